"""Module named like the reference's extension ``correlation_cuda`` (correlation_cuda.cc:169-172).

Same two entry points, same out-parameter convention: the caller passes empty tensors
(``input1.new()``), ``forward``/``backward`` resize and overwrite them and return 1.  ``rInput1`` /
``rInput2`` (the reference's padded NHWC scratch, correlation_cuda.cc:36-41) are left untouched:
this implementation needs no scratch.
"""
from flownet2_b200 import functional as _F


def forward(input1, input2, rInput1, rInput2, output, pad_size, kernel_size, max_displacement, stride1, stride2,
            corr_type_multiply):
    B, C, H, W = input1.shape
    D, oH, oW = _F.correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    output.resize_(B, D, oH, oW)
    _F.correlation_forward(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2,
                           corr_type_multiply, out=output)
    return 1


def backward(input1, input2, rInput1, rInput2, gradOutput, gradInput1, gradInput2, pad_size, kernel_size,
             max_displacement, stride1, stride2, corr_type_multiply):
    gradInput1.resize_(input1.shape)
    gradInput2.resize_(input2.shape)
    _F.correlation_backward(input1, input2, gradOutput, pad_size, kernel_size, max_displacement, stride1, stride2,
                            corr_type_multiply, out1=gradInput1, out2=gradInput2)
    return 1
