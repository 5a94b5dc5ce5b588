// common.cuh -- shared host/device helpers for libfn2b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fn2b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfn2b200 is written for sm_100a (B200) only"
#endif

namespace fn2 {

// ---- error plumbing (thread-local, no global mutable state) -------------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
void count_launch(int n = 1);
uint64_t launches_so_far();
const char *last_error_text();
int check_launch(const char *what);
int bind_device_of(const void *ptr);  // make the device owning ptr current on this thread  // cudaGetLastError -> return code (+ message)

// ---- TMA descriptor creation (driver entry point resolved through the runtime; no -lcuda) --
// Encodes a tiled tensor map over an fp32 tensor of `rank` dims (dims[0] fastest).
// strides_bytes[i] is the byte stride of dim i+1 (rank-1 entries).  Returns 0 / error code.
int make_tensor_map_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                        const uint64_t *strides_bytes, const uint32_t *box);

// ---- device-side PTX wrappers ---------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 3D / 4D tiled TMA loads global -> shared, completion on an mbarrier (transaction bytes).
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// streaming (read-once) 128-bit / 32-bit global loads that do not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// write-once streaming stores (evict-first in L2 is left to hardware; bypass L1 allocation)
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#endif  // __CUDACC__

// ---- kernel launchers (one per .cu) -----------------------------------------------------------
int channelnorm_forward(const float *in, float *out, int B, int C, int H, int W, cudaStream_t st);
int channelnorm_backward(const float *in, const float *out, const float *gout, float *gin, int B,
                         int C, int H, int W, cudaStream_t st);
int channelnorm_forward_half(const void *in, void *out, int B, int C, int H, int W, int dtype, cudaStream_t st);
int channelnorm_backward_half(const void *in, const void *out, const void *gout, void *gin, int B, int C, int H,
                              int W, int dtype, cudaStream_t st);
int resample2d_forward(const float *img, const int64_t *istride, const float *flow, float *out,
                       int B, int C, int iH, int iW, int H, int W, int bilinear, cudaStream_t st);
int resample2d_backward(const float *img, const int64_t *istride, const float *flow,
                        const float *gout, float *gimg, float *gflow, int B, int C, int iH, int iW,
                        int H, int W, cudaStream_t st);

struct CorrParams {
    int B, C, H, W;        // inputs [B,C,H,W]
    int pad, k, md, s1, s2;
    int kr, dr, ds, D;     // kernel radius, displacement radius / size / count
    int oH, oW;            // output spatial dims
};
int corr_forward_generic(const float *in1, const float *in2, float *out, const CorrParams &p,
                         cudaStream_t st);
int corr_backward_generic(const float *in1, const float *in2, const float *gout, float *gin1,
                          float *gin2, const CorrParams &p, cudaStream_t st);
// TMA-tiled FMA kernels: kernel_size == 1, stride1 == 1, (s2, dr) in the instantiated set, W % 4 == 0.
bool corr_tiled_supported(const CorrParams &p);
int corr_forward_tiled(const float *in1, const float *in2, float *out, const CorrParams &p,
                       cudaStream_t st);
int corr_backward_tiled(const float *in1, const float *in2, const float *gout, float *gin1,
                        float *gin2, const CorrParams &p, cudaStream_t st);

// Tensor-core (tcgen05) forward for FlowNetC's configuration; needs a caller-provided workspace.
bool corr_tc_supported(const CorrParams &p);
size_t corr_tc_workspace_bytes(const CorrParams &p);
int corr_backward_tc(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                     const CorrParams &p, void *workspace, size_t workspace_bytes, int have_split, cudaStream_t st);
int corr_forward_tc(const float *in1, const float *in2, float *out, const CorrParams &p, void *workspace,
                    size_t workspace_bytes, cudaStream_t st);

}  // namespace fn2
