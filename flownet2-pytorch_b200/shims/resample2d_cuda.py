"""Module named like the reference's extension ``resample2d_cuda`` (resample2d_cuda.cc:28-31).

The caller pre-allocates (and pre-zeroes) the outputs at their final shape (resample2d.py:18,31-32);
they are written in place.  ``backward`` accumulates into the caller-zeroed ``gradInput1`` exactly
like the reference's atomicAdd kernel.
"""
from flownet2_b200 import functional as _F


def forward(input1, input2, output, kernel_size, bilinear):
    _F.resample2d_forward(input1, input2, kernel_size, bilinear, out=output)
    return 1


def backward(input1, input2, gradOutput, gradInput1, gradInput2, kernel_size, bilinear):
    _F.resample2d_backward(input1, input2, gradOutput, kernel_size, bilinear, out1=gradInput1, out2=gradInput2,
                           zero_out1=False)
    return 1
