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
