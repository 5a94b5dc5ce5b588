// correlation_cuda.cc -- the reference's extension module `correlation_cuda` (correlation_package/correlation_cuda.cc:
// 10-16, 89-96, 169-172: same function names, argument order, out-parameter convention, return value) on libfn2b200.
#include "ext_common.h"
using namespace fn2ext;

int correlation_forward_cuda(at::Tensor &input1, at::Tensor &input2, at::Tensor &rInput1, at::Tensor &rInput2,
                             at::Tensor &output, int pad_size, int kernel_size, int max_displacement, int stride1,
                             int stride2, int corr_type_multiply) {
    (void)rInput1; (void)rInput2;        // the reference's padded NHWC scratch (:36-41): not needed, left untouched
    need_cuda_f32(input1, "input1"); need_cuda_f32(input2, "input2"); need_out(output, "output");
    TORCH_CHECK(input1.dim() == 4 && input1.sizes() == input2.sizes(), "input1 / input2 must be 4-D with equal shapes");
    int rc;
    {   // the status is checked (and the exception thrown) after the guard and the temporaries are gone
        c10::cuda::CUDAGuard guard(input1.device());
        at::Tensor a = input1.contiguous(), b = input2.contiguous();
        const int B = a.size(0), C = a.size(1), H = a.size(2), W = a.size(3);
        int D = 0, oH = 0, oW = 0;
        rc = fn2b200_correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, &D, &oH, &oW);
        if (rc == 0) {
            output.resize_({B, D, oH, oW});      // :38 (no fill_(0): every element is written)
            const size_t ws = fn2b200_correlation_forward_workspace(B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
            at::Tensor wsb = scratch(ws, a);
            rc = fn2b200_correlation_forward_ws(a.data_ptr<float>(), b.data_ptr<float>(), output.data_ptr<float>(), B, C, H, W,
                                                pad_size, kernel_size, max_displacement, stride1, stride2, corr_type_multiply,
                                                ws ? wsb.data_ptr() : nullptr, ws, stream_of(a));
        }
    }
    check(rc, "correlation_forward");
    return 1;
}

int correlation_backward_cuda(at::Tensor &input1, at::Tensor &input2, at::Tensor &rInput1, at::Tensor &rInput2,
                              at::Tensor &gradOutput, at::Tensor &gradInput1, at::Tensor &gradInput2, int pad_size,
                              int kernel_size, int max_displacement, int stride1, int stride2, int corr_type_multiply) {
    (void)rInput1; (void)rInput2;
    need_cuda_f32(input1, "input1"); need_cuda_f32(input2, "input2"); need_cuda_f32(gradOutput, "gradOutput");
    need_out(gradInput1, "gradInput1"); need_out(gradInput2, "gradInput2");
    int rc;
    {
        c10::cuda::CUDAGuard guard(input1.device());
        at::Tensor a = input1.contiguous(), b = input2.contiguous(), g = gradOutput.contiguous();
        const int B = a.size(0), C = a.size(1), H = a.size(2), W = a.size(3);
        gradInput1.resize_(a.sizes());       // :109-110
        gradInput2.resize_(b.sizes());
        const size_t ws = fn2b200_correlation_backward_workspace(B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
        at::Tensor wsb = scratch(ws, a);
        rc = fn2b200_correlation_backward_ws(a.data_ptr<float>(), b.data_ptr<float>(), g.data_ptr<float>(),
                                             gradInput1.data_ptr<float>(), gradInput2.data_ptr<float>(), B, C, H, W, pad_size,
                                             kernel_size, max_displacement, stride1, stride2, corr_type_multiply,
                                             ws ? wsb.data_ptr() : nullptr, ws, 0, stream_of(a));
    }
    check(rc, "correlation_backward");
    return 1;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward", &correlation_forward_cuda, "Correlation forward (CUDA)");
    m.def("backward", &correlation_backward_cuda, "Correlation backward (CUDA)");
}
