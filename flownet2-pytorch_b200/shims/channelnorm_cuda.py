"""Module named like the reference's extension ``channelnorm_cuda`` (channelnorm_cuda.cc:27-30)."""
from flownet2_b200 import functional as _F


def forward(input1, output, norm_deg):
    _F.channelnorm_forward(input1, norm_deg, out=output)
    return 1


def backward(input1, output, gradOutput, gradInput1, norm_deg):
    _F.channelnorm_backward(input1, output, gradOutput, norm_deg, out=gradInput1)
    return 1
