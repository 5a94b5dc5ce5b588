// capi.cu -- extern "C" boundary of libfn2b200 (see include/fn2b200.h for the contract).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fn2 {

static int fill_corr_params(CorrParams &p, int B, int C, int H, int W, int pad, int k, int md,
                            int s1, int s2) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0)
        return fail(FN2B200_EINVAL, "correlation: bad input shape [%d,%d,%d,%d]", B, C, H, W);
    if (pad < 0 || k < 1 || md < 0 || s1 < 1 || s2 < 1)
        return fail(FN2B200_EINVAL,
                    "correlation: bad parameters pad=%d kernel_size=%d max_displacement=%d "
                    "stride1=%d stride2=%d", pad, k, md, s1, s2);
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.pad = pad; p.k = k; p.md = md; p.s1 = s1; p.s2 = s2;
    p.kr = (k - 1) / 2;
    p.dr = md / s2;
    p.ds = 2 * p.dr + 1;
    p.D = p.ds * p.ds;
    int br = p.kr + md;
    // correlation_cuda.cc:33-34 -- float ceil of a float division
    p.oH = (int)ceilf((float)(H + 2 * pad - 2 * br) / (float)s1);
    p.oW = (int)ceilf((float)(W + 2 * pad - 2 * br) / (float)s1);
    if (p.oH <= 0 || p.oW <= 0)
        return fail(FN2B200_EINVAL, "correlation: empty output (%d x %d); pad_size too small for "
                    "max_displacement/kernel_size", p.oH, p.oW);
    if ((int64_t)B * C * H * W >= (1LL << 31) || (int64_t)B * p.D * p.oH * p.oW >= (1LL << 31))
        return fail(FN2B200_EINVAL, "correlation: tensor exceeds 2^31 elements");
    return 0;
}

// FN2B200_CORR_FWD = "fma" forces the FP32-FMA kernels, "tc" (default) uses tensor cores when the
// configuration supports them and a workspace is supplied.  Read per call: no cached state.
static bool tc_enabled() {
    const char *e = getenv("FN2B200_CORR_FWD");
    return !(e && (e[0] == 'f' || e[0] == 'F'));
}
static bool tc_bwd_enabled() {
    const char *e = getenv("FN2B200_CORR_BWD");
    return !(e && (e[0] == 'f' || e[0] == 'F'));
}

}  // namespace fn2

using namespace fn2;

extern "C" {

int fn2b200_version(void) { return FN2B200_VERSION; }

const char *fn2b200_last_error(void) { return last_error_text(); }
uint64_t fn2b200_launch_count(void) { return launches_so_far(); }

int fn2b200_correlation_out_shape(int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                  int *D, int *oH, int *oW) {
    CorrParams p;
    // shape query must work for empty outputs too -> replicate arithmetic without the checks
    if (C <= 0 || H <= 0 || W <= 0 || pad < 0 || k < 1 || md < 0 || s1 < 1 || s2 < 1)
        return fail(FN2B200_EINVAL, "correlation_out_shape: bad parameters");
    int kr = (k - 1) / 2, br = kr + md, dr = md / s2;
    if (D) *D = (2 * dr + 1) * (2 * dr + 1);
    if (oH) *oH = (int)ceilf((float)(H + 2 * pad - 2 * br) / (float)s1);
    if (oW) *oW = (int)ceilf((float)(W + 2 * pad - 2 * br) / (float)s1);
    (void)p;
    return 0;
}

int fn2b200_correlation_path(int C, int H, int W, int pad, int k, int md, int s1, int s2) {
    CorrParams p;
    if (fill_corr_params(p, 1, C, H, W, pad, k, md, s1, s2)) return -1;
    if (tc_enabled() && corr_tc_supported(p)) return 2;
    return corr_tiled_supported(p) ? 1 : 0;
}

int fn2b200_correlation_forward(const float *in1, const float *in2, float *out, int B, int C,
                                int H, int W, int pad, int k, int md, int s1, int s2,
                                int corr_type_multiply, void *stream) {
    (void)corr_type_multiply;  // accepted and ignored: correlation_cuda_kernel.cu:369
    CorrParams p;
    int rc = fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!in1 || !in2 || !out) return fail(FN2B200_ENULL, "correlation_forward: null pointer");
    if ((rc = bind_device_of(in1))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tiled_supported(p)) return corr_forward_tiled(in1, in2, out, p, st);
    return corr_forward_generic(in1, in2, out, p, st);
}

size_t fn2b200_correlation_forward_workspace(int B, int C, int H, int W, int pad, int k, int md, int s1,
                                             int s2) {
    CorrParams p;
    if (fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2)) return 0;
    if (!tc_enabled()) return 0;
    return corr_tc_workspace_bytes(p);
}

int fn2b200_correlation_forward_ws(const float *in1, const float *in2, float *out, int B, int C, int H,
                                   int W, int pad, int k, int md, int s1, int s2, int corr_type_multiply,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    CorrParams p;
    int rc = fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!in1 || !in2 || !out) return fail(FN2B200_ENULL, "correlation_forward: null pointer");
    if (workspace && tc_enabled() && corr_tc_supported(p) && workspace_bytes >= corr_tc_workspace_bytes(p)) {
        if ((rc = bind_device_of(in1))) return rc;
        return corr_forward_tc(in1, in2, out, p, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    return fn2b200_correlation_forward(in1, in2, out, B, C, H, W, pad, k, md, s1, s2, corr_type_multiply, stream);
}

int fn2b200_correlation_backward(const float *in1, const float *in2, const float *gout,
                                 float *gin1, float *gin2, int B, int C, int H, int W, int pad,
                                 int k, int md, int s1, int s2, int corr_type_multiply,
                                 void *stream) {
    (void)corr_type_multiply;
    CorrParams p;
    int rc = fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    if (s1 != 1)
        return fail(FN2B200_EUNSUPPORTED,
                    "correlation_backward: stride1=%d unsupported (the reference kernels index out "
                    "of bounds for stride1 != 1, correlation_cuda_kernel.cu:163-164 vs :520)", s1);
    if (B == 0) return 0;
    if (!in1 || !in2 || !gout) return fail(FN2B200_ENULL, "correlation_backward: null pointer");
    if (!gin1 && !gin2) return 0;
    if ((rc = bind_device_of(in1))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tiled_supported(p)) return corr_backward_tiled(in1, in2, gout, gin1, gin2, p, st);
    return corr_backward_generic(in1, in2, gout, gin1, gin2, p, st);
}

size_t fn2b200_correlation_backward_workspace(int B, int C, int H, int W, int pad, int k, int md, int s1,
                                              int s2) {
    CorrParams p;
    if (fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2)) return 0;
    if (!tc_bwd_enabled()) return 0;
    return corr_tc_workspace_bytes(p);
}

int fn2b200_correlation_backward_ws(const float *in1, const float *in2, const float *gout, float *gin1,
                                    float *gin2, int B, int C, int H, int W, int pad, int k, int md, int s1,
                                    int s2, int corr_type_multiply, void *workspace, size_t workspace_bytes,
                                    int workspace_has_split, void *stream) {
    CorrParams p;
    int rc = fill_corr_params(p, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    if (workspace && tc_bwd_enabled() && corr_tc_supported(p) && workspace_bytes >= corr_tc_workspace_bytes(p)) {
        if (B == 0) return 0;
        if (!in1 || !in2 || !gout) return fail(FN2B200_ENULL, "correlation_backward: null pointer");
        if (!gin1 && !gin2) return 0;
        if ((rc = bind_device_of(in1))) return rc;
        return corr_backward_tc(in1, in2, gout, gin1, gin2, p, workspace, workspace_bytes, workspace_has_split,
                                (cudaStream_t)stream);
    }
    return fn2b200_correlation_backward(in1, in2, gout, gin1, gin2, B, C, H, W, pad, k, md, s1, s2,
                                        corr_type_multiply, stream);
}

static int check_resample(const char *who, const int64_t *istride, int B, int C, int iH, int iW,
                          int H, int W, int kernel_size) {
    if (B < 0 || C <= 0 || iH <= 0 || iW <= 0 || H <= 0 || W <= 0)
        return fail(FN2B200_EINVAL, "%s: bad shape B=%d C=%d iH=%d iW=%d H=%d W=%d", who, B, C, iH,
                    iW, H, W);
    if (!istride) return fail(FN2B200_ENULL, "%s: null input1 stride array", who);
    if (kernel_size != 1)
        return fail(FN2B200_EUNSUPPORTED,
                    "%s: kernel_size=%d unsupported (the reference's kernel_size > 1 taps are "
                    "unclamped and read out of bounds, resample2d_kernel.cu:54-61)", who, kernel_size);
    if (iH < H || iW < W)
        return fail(FN2B200_EUNSUPPORTED,
                    "%s: input1 spatial dims (%d,%d) smaller than the flow's (%d,%d): the reference "
                    "clamps with the flow dims and would read out of bounds", who, iH, iW, H, W);
    if ((int64_t)B * C * H * W >= (1LL << 31) || (int64_t)B * C * iH * iW >= (1LL << 31))
        return fail(FN2B200_EINVAL, "%s: tensor exceeds 2^31 elements", who);
    return 0;
}

int fn2b200_resample2d_forward(const float *img, const int64_t *istride, const float *flow,
                               float *out, int B, int C, int iH, int iW, int H, int W,
                               int kernel_size, int bilinear, void *stream) {
    int rc = check_resample("resample2d_forward", istride, B, C, iH, iW, H, W, kernel_size);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!img || !flow || !out) return fail(FN2B200_ENULL, "resample2d_forward: null pointer");
    if ((rc = bind_device_of(flow))) return rc;
    return resample2d_forward(img, istride, flow, out, B, C, iH, iW, H, W, bilinear,
                              (cudaStream_t)stream);
}

int fn2b200_resample2d_backward(const float *img, const int64_t *istride, const float *flow,
                                const float *gout, float *gimg, float *gflow, int B, int C,
                                int iH, int iW, int H, int W, int kernel_size, int bilinear,
                                int zero_grad_input1, void *stream) {
    (void)bilinear;  // the reference's backward ignores it (resample2d_kernel.cu:75-198)
    int rc = check_resample("resample2d_backward", istride, B, C, iH, iW, H, W, kernel_size);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!img || !flow || !gout) return fail(FN2B200_ENULL, "resample2d_backward: null pointer");
    if ((rc = bind_device_of(flow))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (gimg && zero_grad_input1) {
        cudaError_t e = cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)B * C * iH * iW, st);
        if (e != cudaSuccess)
            return fail((int)e, "resample2d_backward: memset failed (%s)", cudaGetErrorString(e));
    }
    if (!gimg && !gflow) return 0;
    return resample2d_backward(img, istride, flow, gout, gimg, gflow, B, C, iH, iW, H, W, st);
}

int fn2b200_channelnorm_forward(const float *in, float *out, int B, int C, int H, int W,
                                int norm_deg, void *stream) {
    (void)norm_deg;  // ignored by the reference kernels too (channelnorm_kernel.cu:18-60)
    if (B < 0 || C <= 0 || H <= 0 || W <= 0)
        return fail(FN2B200_EINVAL, "channelnorm_forward: bad shape [%d,%d,%d,%d]", B, C, H, W);
    if ((int64_t)B * C * H * W >= (1LL << 31))
        return fail(FN2B200_EINVAL, "channelnorm_forward: tensor exceeds 2^31 elements");
    if (B == 0) return 0;
    if (!in || !out) return fail(FN2B200_ENULL, "channelnorm_forward: null pointer");
    if (int rc = bind_device_of(in)) return rc;
    return channelnorm_forward(in, out, B, C, H, W, (cudaStream_t)stream);
}

int fn2b200_channelnorm_backward(const float *in, const float *out, const float *gout, float *gin,
                                 int B, int C, int H, int W, int norm_deg, void *stream) {
    (void)norm_deg;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0)
        return fail(FN2B200_EINVAL, "channelnorm_backward: bad shape [%d,%d,%d,%d]", B, C, H, W);
    if ((int64_t)B * C * H * W >= (1LL << 31))
        return fail(FN2B200_EINVAL, "channelnorm_backward: tensor exceeds 2^31 elements");
    if (B == 0) return 0;
    if (!in || !out || !gout || !gin)
        return fail(FN2B200_ENULL, "channelnorm_backward: null pointer");
    if (int rc = bind_device_of(in)) return rc;
    return channelnorm_backward(in, out, gout, gin, B, C, H, W, (cudaStream_t)stream);
}

int fn2b200_channelnorm_forward_16(const void *in, void *out, int B, int C, int H, int W, int norm_deg,
                                   int dtype, void *stream) {
    (void)norm_deg;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || (dtype != 1 && dtype != 2))
        return fail(FN2B200_EINVAL, "channelnorm_forward_16: bad shape [%d,%d,%d,%d] or dtype %d", B, C, H, W, dtype);
    if ((int64_t)B * C * H * W >= (1LL << 31)) return fail(FN2B200_EINVAL, "channelnorm_forward_16: tensor too large");
    if (B == 0) return 0;
    if (!in || !out) return fail(FN2B200_ENULL, "channelnorm_forward_16: null pointer");
    if (int rc = bind_device_of(in)) return rc;
    return channelnorm_forward_half(in, out, B, C, H, W, dtype, (cudaStream_t)stream);
}

int fn2b200_channelnorm_backward_16(const void *in, const void *out, const void *gout, void *gin, int B, int C,
                                    int H, int W, int norm_deg, int dtype, void *stream) {
    (void)norm_deg;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || (dtype != 1 && dtype != 2))
        return fail(FN2B200_EINVAL, "channelnorm_backward_16: bad shape [%d,%d,%d,%d] or dtype %d", B, C, H, W, dtype);
    if ((int64_t)B * C * H * W >= (1LL << 31)) return fail(FN2B200_EINVAL, "channelnorm_backward_16: tensor too large");
    if (B == 0) return 0;
    if (!in || !out || !gout || !gin) return fail(FN2B200_ENULL, "channelnorm_backward_16: null pointer");
    if (int rc = bind_device_of(in)) return rc;
    return channelnorm_backward_half(in, out, gout, gin, B, C, H, W, dtype, (cudaStream_t)stream);
}

}  // extern "C"
