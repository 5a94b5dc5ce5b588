"""Host-resident inputs and outputs: overlap the H2D copy, the kernels and the D2H copy of consecutive steps.

The layers themselves take device tensors (as the reference's do).  When the data lives in (pinned) host
memory -- bench.py's `e2e` leg, or a data loader feeding feature maps -- a step is H2D + kernels + D2H, and
at cfg2 the two PCIe transfers (874 MB each way) dwarf the 1.2 ms of kernels.  PCIe is full duplex, so the
pipeline keeps `depth` sets of device buffers and three streams: while step i's results travel to the host,
step i+1's inputs travel to the device.  Every step still copies all of its inputs and all of its outputs.
"""
import torch


class HostPipeline(object):
    """pipe = HostPipeline(in_shapes, out_shapes, device); pipe.submit(compute, host_in, host_out); pipe.drain()

    compute(dev_in, dev_out) is called on the pipeline's compute stream with lists of device tensors; it must
    only enqueue work on the current stream (our layers and the reference's do).  host_in / host_out are
    sequences of pinned CPU tensors of the declared shapes.  submit() returns immediately.

    Host-buffer lifetime: a host_in tensor must not be overwritten, and a host_out tensor must not be read, until
    the step that uses it has left the pipeline -- i.e. after drain() (which by default blocks the calling thread until
    the last D2H copy has landed) or after ``depth`` further submit() calls followed by drain(host_sync=False) plus
    a synchronisation of the caller's stream.  Pass ``numa_bind=True`` (one process per GPU) to pin the process to the
    GPU's NUMA node first, so that pinned buffers allocated AFTERWARDS sit behind the GPU's own PCIe root."""

    def __init__(self, in_shapes, out_shapes, device, depth=2, dtype=torch.float32, numa_bind=False):
        if depth < 1:
            raise ValueError("HostPipeline: depth must be >= 1")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device (there is no CPU path)")
        self.numa = None
        if numa_bind:
            from . import numa
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.numa = numa.bind_to_device_node(idx)
        self.depth = depth
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.dev_in = [[torch.empty(tuple(s), device=self.device, dtype=dtype) for s in in_shapes] for _ in range(depth)]
        self.dev_out = [[torch.empty(tuple(s), device=self.device, dtype=dtype) for s in out_shapes] for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]      # inputs of the slot have arrived
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]    # kernels of the slot are done
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]     # outputs of the slot have left
        self.count = 0

    def submit(self, compute, host_in, host_out):
        s = self.count % self.depth
        first_use = self.count < self.depth
        self.count += 1
        cur = torch.cuda.current_stream(self.device)
        for t in list(host_in) + list(host_out):
            if t.device.type != "cpu" or not t.is_pinned():
                raise RuntimeError("HostPipeline: host tensors must be pinned CPU tensors")
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_stream(cur)                      # whatever the caller enqueued before (e.g. a timing event)
            if not first_use:
                self.s_in.wait_event(self.ev_comp[s])       # the slot's previous kernels no longer read the inputs
            for d, h in zip(self.dev_in[s], host_in):
                d.copy_(h, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.ev_in[s])
            if not first_use:
                self.s_comp.wait_event(self.ev_out[s])      # the slot's previous outputs have left
            compute(self.dev_in[s], self.dev_out[s])
            self.ev_comp[s].record(self.s_comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[s])
            for h, d in zip(host_out, self.dev_out[s]):
                h.copy_(d, non_blocking=True)
            self.ev_out[s].record(self.s_out)

    def drain(self, host_sync=True):
        """Make the caller's current stream wait for everything submitted so far; with host_sync (default) also block
        the calling thread until the last D2H copy has completed, so host_out can be read and host_in reused."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_comp)
        cur.wait_stream(self.s_out)
        if host_sync and self.count:
            self.ev_out[(self.count - 1) % self.depth].synchronize()
