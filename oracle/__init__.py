"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement + rebuilt reference kernels).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``flownet2-pytorch_b200``) never does.
"""
