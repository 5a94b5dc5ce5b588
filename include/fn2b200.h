/*
 * fn2b200.h -- C ABI of libfn2b200.so: B200 (sm_100a) kernels for flownet2-pytorch's three
 * custom layers.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * Each entry point replaces one function of the reference's pybind extensions
 * (paths relative to the reference's networks/ directory):
 *
 *   fn2b200_correlation_forward    <- correlation_package/correlation_cuda.cc:10-87
 *                                     (+ correlation_cuda_kernel.cu:336-427: K1 x2, K2)
 *   fn2b200_correlation_backward   <- correlation_package/correlation_cuda.cc:89-167
 *                                     (+ correlation_cuda_kernel.cu:430-564: K1 x2, K3 xB, K4 xB)
 *   fn2b200_correlation_out_shape  <- correlation_package/correlation_cuda.cc:19-34
 *   fn2b200_resample2d_forward     <- resample2d_package/resample2d_cuda.cc:6-13
 *                                     (+ resample2d_kernel.cu:200-242: K5)
 *   fn2b200_resample2d_backward    <- resample2d_package/resample2d_cuda.cc:15-24
 *                                     (+ resample2d_kernel.cu:244-323: K6, K7)
 *   fn2b200_channelnorm_forward    <- channelnorm_package/channelnorm_cuda.cc:6-13
 *                                     (+ channelnorm_kernel.cu:98-129: K8)
 *   fn2b200_channelnorm_backward   <- channelnorm_package/channelnorm_cuda.cc:16-25
 *                                     (+ channelnorm_kernel.cu:131-177: K9)
 *
 * Conventions
 *   - All data pointers are DEVICE pointers to fp32, on the device that is current when the call
 *     is made.  Tensors are contiguous NCHW unless a stride array is passed.
 *   - `stream` is a cudaStream_t / CUstream handle (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host, like the reference (it never synchronises).
 *   - Return value: 0 on success; a negative FN2B200_E* code for argument errors; a positive
 *     cudaError_t value when the CUDA runtime reported a failure.  fn2b200_last_error() returns a
 *     thread-local human-readable message for the last failing call on this thread.  (The
 *     reference raises AT_ERROR("CUDA call failed"), correlation_cuda.cc:81-83; the Python host
 *     layer turns a non-zero return into RuntimeError.)
 *   - The library keeps no mutable global state (bar a launch counter): every call is re-entrant and thread-safe
 *     (nn.DataParallel calls these from one Python thread per GPU, main.py:200).
 *   - Outputs are fully overwritten; no pre-zeroing is required except where noted.
 */
#ifndef FN2B200_H_
#define FN2B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FN2B200_VERSION 100

#define FN2B200_EINVAL (-1)      /* bad shape / parameter */
#define FN2B200_EUNSUPPORTED (-2) /* parameter combination the reference itself cannot run */
#define FN2B200_ENULL (-3)       /* null pointer with a non-empty tensor */

int fn2b200_version(void);
const char *fn2b200_last_error(void);

/* D = (2*(md/s2)+1)^2, oH/oW = ceil((H + 2*pad - 2*((k-1)/2 + md)) / s1)  (float ceil). */
int fn2b200_correlation_out_shape(int C, int H, int W, int pad_size, int kernel_size,
                                  int max_displacement, int stride1, int stride2, int *D,
                                  int *oH, int *oW);

/*
 * input1, input2: [B,C,H,W]; output: [B,D,oH,oW].
 * corr_type_multiply is accepted and ignored, exactly as in the reference
 * (correlation_cuda_kernel.cu:369).  No padded/transposed scratch tensors are needed (the
 * reference's rInput1/rInput2 have no counterpart here).
 */
int fn2b200_correlation_forward(const float *input1, const float *input2, float *output, int B,
                                int C, int H, int W, int pad_size, int kernel_size,
                                int max_displacement, int stride1, int stride2,
                                int corr_type_multiply, void *stream);

/*
 * Tensor-core forward (tcgen05, bf16 hi/lo operand split, fp32 accumulation in TMEM; error ~1e-5
 * relative, inside the 1e-4 contract).  It needs scratch for the split operands:
 * fn2b200_correlation_forward_workspace() returns the bytes required (0 = this configuration runs
 * on the FP32-FMA kernels; set the environment variable FN2B200_CORR_FWD=fma to force that), and
 * fn2b200_correlation_forward_ws() is fn2b200_correlation_forward() plus a 128-byte-aligned device
 * workspace of at least that size (NULL / too small -> the FMA kernels).  The workspace plays the
 * role of the reference's rInput1/rInput2 scratch tensors (correlation_cuda.cc:36-41) and may be
 * reused or freed (stream-ordered) as soon as the call returns.
 */
size_t fn2b200_correlation_forward_workspace(int B, int C, int H, int W, int pad_size, int kernel_size,
                                             int max_displacement, int stride1, int stride2);
int fn2b200_correlation_forward_ws(const float *input1, const float *input2, float *output, int B, int C,
                                   int H, int W, int pad_size, int kernel_size, int max_displacement,
                                   int stride1, int stride2, int corr_type_multiply, void *workspace,
                                   size_t workspace_bytes, void *stream);

/*
 * Tensor-core backward, same workspace size and layout as the forward's.  workspace_has_split != 0
 * promises that the workspace still holds what fn2b200_correlation_forward_ws wrote for the SAME
 * input1/input2 (the split pass is then skipped).  FN2B200_CORR_BWD=fma forces the FMA kernels.
 */
size_t fn2b200_correlation_backward_workspace(int B, int C, int H, int W, int pad_size, int kernel_size,
                                              int max_displacement, int stride1, int stride2);
int fn2b200_correlation_backward_ws(const float *input1, const float *input2, const float *grad_output,
                                    float *grad_input1, float *grad_input2, int B, int C, int H, int W,
                                    int pad_size, int kernel_size, int max_displacement, int stride1,
                                    int stride2, int corr_type_multiply, void *workspace,
                                    size_t workspace_bytes, int workspace_has_split, void *stream);

/*
 * grad_output: [B,D,oH,oW]; grad_input1, grad_input2: [B,C,H,W] (either may be NULL to skip it).
 * stride1 must be 1 (the reference's backward indexes out of bounds otherwise,
 * correlation_cuda_kernel.cu:163-164 vs :520) -> FN2B200_EUNSUPPORTED.
 */
int fn2b200_correlation_backward(const float *input1, const float *input2,
                                 const float *grad_output, float *grad_input1,
                                 float *grad_input2, int B, int C, int H, int W, int pad_size,
                                 int kernel_size, int max_displacement, int stride1, int stride2,
                                 int corr_type_multiply, void *stream);

/*
 * input1 (image): [*,C,iH,iW] with element strides istride[4] = {b, c, h, w} (so the
 * non-contiguous channel slice models.py:133 passes needs no .contiguous() copy);
 * input2 (flow): [B,2,H,W] contiguous; output: [B,C,H,W] contiguous.
 * kernel_size must be 1 (the reference's kernel_size > 1 taps are unclamped and read out of
 * bounds, resample2d_kernel.cu:54-61) -> FN2B200_EUNSUPPORTED.  bilinear == 0 -> nearest.
 */
int fn2b200_resample2d_forward(const float *input1, const int64_t *istride, const float *input2,
                               float *output, int B, int C, int iH, int iW, int H, int W,
                               int kernel_size, int bilinear, void *stream);

/*
 * grad_output: [B,C,H,W]; grad_input1: [B,C,iH,iW] contiguous; grad_input2: [B,2,H,W].
 * grad_input1 is accumulated with fp32 atomics (order non-deterministic, like the reference's
 * atomicAdd, resample2d_kernel.cu:118-121).  zero_grad_input1 != 0 -> the library zero-fills it
 * on `stream` first; 0 -> the caller already did (resample2d.py:31).  Either grad may be NULL.
 * The backward is always the bilinear one (`bilinear` is ignored, as in the reference).
 */
int fn2b200_resample2d_backward(const float *input1, const int64_t *istride, const float *input2,
                                const float *grad_output, float *grad_input1, float *grad_input2,
                                int B, int C, int iH, int iW, int H, int W, int kernel_size,
                                int bilinear, int zero_grad_input1, void *stream);

/* input1: [B,C,H,W]; output: [B,1,H,W] = sqrt(sum_c x^2).  norm_deg is ignored (reference: same). */
int fn2b200_channelnorm_forward(const float *input1, float *output, int B, int C, int H, int W,
                                int norm_deg, void *stream);

/* grad_input1[b,c,p] = grad_output[b,p] * input1[b,c,p] / (output[b,p] + 1e-9). */
int fn2b200_channelnorm_backward(const float *input1, const float *output,
                                 const float *grad_output, float *grad_input1, int B, int C,
                                 int H, int W, int norm_deg, void *stream);

/*
 * 16-bit storage variants of ChannelNorm (dtype: 1 = fp16, 2 = bf16; tensors of that type, fp32
 * arithmetic inside).  The reference dispatches these kernels on half as well
 * (channelnorm_kernel.cu:111,152); ChannelNorm is the one custom layer fed fp16 in --fp16 mode.
 */
int fn2b200_channelnorm_forward_16(const void *input1, void *output, int B, int C, int H, int W, int norm_deg,
                                   int dtype, void *stream);
int fn2b200_channelnorm_backward_16(const void *input1, const void *output, const void *grad_output,
                                    void *grad_input1, int B, int C, int H, int W, int norm_deg, int dtype,
                                    void *stream);

/*
 * Introspection for benchmarks/tests: which kernel family the correlation entry points would
 * dispatch to for these parameters.  0 = generic gather kernels, 1 = TMA-tiled FMA kernels,
 * 2 = forward on tensor cores (when a workspace is supplied) + TMA-tiled FMA backward.
 */
int fn2b200_correlation_path(int C, int H, int W, int pad_size, int kernel_size,
                             int max_displacement, int stride1, int stride2);

/*
 * Hardware self-test of the tcgen05 / TMEM / TMA-swizzle plumbing the tensor-core correlation path
 * is built on: D[128 x 144] (fp32, row-major) = A[128 x K] * B[144 x K]^T, A and B bf16 row-major
 * device buffers, K a multiple of 64.  K == -144 selects the second form (the backward kernel's
 * operand layouts): D[128 x 64] = A[128 x 144] * Bt[144 x 64], A written to shared memory by the
 * threads in the no-swizzle core-matrix layout, Bt ([K][N] row-major) loaded as an MN-major
 * SW128 operand; K == -145 is the same with A in the 32-byte-swizzle K-major layout.
 * Test hook only (tests/test_gpu_umma.py).
  * -1000 < K <= -500: the K > 0 product with K' = -K - 500 (64...256) and A read from tensor memory (written there by
 * the threads with tcgen05.st; groundwork for round 2, tools/umma_ts_check.py).
 * K <= -1000: MMA issue-rate benchmark instead (A_bf16 / B_bf16 unused but non-null): K = -(1000 + 1000 * mode + N),
 * mode 0 = A, B from shared memory (K-major), 1 = A from tensor memory, 2 = A, B MN-major; D[sm] = cycles per
 * M128 x N x K16 MMA on that SM (D holds >= #SM floats).
 */
int fn2b200_debug_umma_gemm(const void *A_bf16, const void *B_bf16, float *D, int K, void *stream);

/*
 * TMA feed micro-benchmark (tools/tma_feed.py): persistent CTAs pull halo-style SW128 boxes
 * (64 channels x box_w x box_h) of a [nimg][Hc][Wc][C] bf16 tensor into a `stages`-deep ring with no
 * consumer; out[2*cta] = cycles, out[2*cta+1] = bytes.  cluster > 1: thread-block clusters of that size, rank 0
 * issues every box with TMA multicast to the whole cluster (cluster < -1: every rank issues its share).
 * producer_warps (1..8, unicast only): that many warps each drive a private ring.  Test / tuning hook only.
 */
int fn2b200_debug_tma_feed(const void *base_bf16, long long *out, int nimg, int C, int Hc, int Wc, int box_w,
                           int box_h, int stages, int boxes_per_stage, int iters, int grid, int cluster,
                           int producer_warps, void *stream);

/* Number of kernel launches (ours) issued by this library in this process so far (statistics). */
uint64_t fn2b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FN2B200_H_ */
