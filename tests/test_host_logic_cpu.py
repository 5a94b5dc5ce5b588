"""Host-side logic that needs no GPU: the product refuses to run without CUDA (no CPU fallback anywhere)."""
import pytest
import torch

import flownet2_b200 as f


def test_host_pipeline_refuses_cpu_device():
    with pytest.raises(RuntimeError, match="CUDA"):
        f.hostpipe.HostPipeline([(1, 2)], [(1, 2)], "cpu")
    with pytest.raises(ValueError):
        f.hostpipe.HostPipeline([(1, 2)], [(1, 2)], "cuda:0", depth=0)


def test_layers_refuse_cpu_tensors():
    a = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        f.functional.correlation_forward(a, a, 4, 1, 4, 1, 2)
    with pytest.raises(RuntimeError):
        f.functional.channelnorm_forward(a)
    with pytest.raises(RuntimeError):
        f.functional.resample2d_forward(a, torch.zeros(1, 2, 8, 8))


def test_correlation_out_shape_matches_reference_arithmetic():
    # correlation_cuda.cc:19-34: padded = H + 2 pad, border = md + (k - 1) / 2, out = ceil((padded - 2 border) / s1),
    # D = (2 (md / s2) + 1)^2
    assert f.functional.correlation_out_shape(256, 48, 64, 20, 1, 20, 1, 2) == (441, 48, 64)
    assert f.functional.correlation_out_shape(4, 12, 13, 2, 1, 4, 2, 2) == (25, 4, 5)
    assert f.functional.correlation_out_shape(8, 16, 16, 4, 3, 4, 1, 1) == (81, 14, 14)


def test_numa_helpers(tmp_path):
    """flownet2_b200.numa: sysfs parsing and binding are pure host logic (no GPU, no CUDA library call)."""
    import os
    from flownet2_b200 import numa
    assert numa.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa.parse_cpulist("") == []
    assert numa.pci_address(0, 0x1B, 0) == "0000:1b:00.0"
    dev = tmp_path / "0000:9c:00.0"
    dev.mkdir()
    (dev / "numa_node").write_text("1\n")
    assert numa.node_of_pci("0000:9c:00.0", sysfs=str(tmp_path)) == 1
    (dev / "numa_node").write_text("-1\n")                   # single-node boxes report -1
    assert numa.node_of_pci("0000:9c:00.0", sysfs=str(tmp_path)) is None
    assert numa.node_of_pci("0000:00:00.0", sysfs=str(tmp_path)) is None
    before = os.sched_getaffinity(0)
    info = numa.bind_to_node(None)
    assert info["node"] is None and not info["affinity"]
    nodes = numa.online_nodes()
    if nodes:
        info = numa.bind_to_node(nodes[0])
        assert info["node"] == nodes[0] and info["affinity"] and info["cpus"] > 0
        assert os.sched_getaffinity(0) <= set(numa.node_cpus(nodes[0]))
    os.sched_setaffinity(0, before)
    os.environ["FN2B200_NUMA"] = "0"
    try:
        assert numa.bind_to_device_node(0).get("disabled")
    finally:
        del os.environ["FN2B200_NUMA"]


def test_test_library_is_separate_from_the_product():
    """The hardware self-tests / micro-benchmarks live in libfn2b200_test.so; the product library exports none of them."""
    import ctypes
    import flownet2_b200
    import testlib
    prod = ctypes.CDLL(flownet2_b200._lib.LIB_PATH)
    for sym in testlib.SYMBOLS:
        assert not hasattr(prod, sym), sym
    assert not hasattr(prod, "fn2b200_debug_umma_gemm") and not hasattr(prod, "fn2b200_debug_tma_feed")
    tl = ctypes.CDLL(testlib.LIB_PATH)
    for sym in testlib.SYMBOLS:
        assert hasattr(tl, sym), sym
    import re
    hdr = re.sub(r"/\*.*?\*/", "", open(testlib.HEADER).read(), flags=re.S)
    assert sorted(set(re.findall(r"\b(fn2b200_test_\w+)\s*\(", hdr))) == sorted(testlib.SYMBOLS)
