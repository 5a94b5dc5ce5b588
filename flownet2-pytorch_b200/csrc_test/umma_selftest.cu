// umma_selftest.cu -- hardware self-test of the tcgen05 / TMEM / TMA-swizzle plumbing in umma.cuh.
// D[128 x N] (fp32) = A[128 x K] * B[N x K]^T with bf16 operands, K a multiple of 64, N = 144.
// Exposed through libfn2b200_test.so (csrc_test/fn2b200_test.h; tests/test_gpu_umma.py compares with a CPU product).
#include "../csrc/umma.cuh"
#include <cuda_bf16.h>

namespace fn2 {

constexpr int ST_M = 128, ST_N = 144, ST_KB = 64;

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     float *__restrict__ D, int K) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = smem;                         // [128 rows][128 B], SW128
    unsigned char *sB = smem + ST_M * 128;            // [144 rows][128 B], SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + ST_N * 128);   // [0] full, [1] mma done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_bf16_f32(ST_M, ST_N);

    const int nkb = K / ST_KB;
    for (int kb = 0; kb < nkb; ++kb) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[0], (ST_M + ST_N) * 128);
            tma_load_2d(sA, &mapA, &bars[0], kb * ST_KB, 0);
            tma_load_2d(sB, &mapB, &bars[0], kb * ST_KB, 0);
            mbar_wait(&bars[0], kb & 1);
            tcgen05_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA));
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
#pragma unroll
            for (int ks = 0; ks < ST_KB / 16; ++ks)   // advance 16 bf16 = 32 B = 2 x 16-byte units
                umma_bf16_ss(tmem_base, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) != 0);
            umma_commit(&bars[1]);
            mbar_wait(&bars[1], kb & 1);              // smem may be overwritten by the next block
        }
        __syncthreads();
    }
    tcgen05_fence_after();
    // epilogue: warp w reads TMEM lanes [32w, 32w+32), all 144 columns
    const int row = tid;
    for (int c0 = 0; c0 < ST_N; c0 += 16) {
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[row * ST_N + c0 + j] = r[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

int umma_selftest(const void *A, const void *B, float *D, int K, cudaStream_t st) {
    if (K <= 0 || K % ST_KB) return fail(FN2B200_EINVAL, "umma_selftest: K must be a positive multiple of 64");
    CUtensorMap ma, mb;
    uint64_t dimsA[2] = {(uint64_t)K, ST_M}, dimsB[2] = {(uint64_t)K, ST_N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t boxA[2] = {ST_KB, ST_M}, boxB[2] = {ST_KB, ST_N};
    int rc = make_tensor_map_bf16_sw128(&ma, A, 2, dimsA, str, boxA);
    if (rc) return rc;
    rc = make_tensor_map_bf16_sw128(&mb, B, 2, dimsB, str, boxB);
    if (rc) return rc;
    const int smem = (ST_M + ST_N) * 128 + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "umma_selftest: smem attribute (%s)", cudaGetErrorString(e));
    umma_selftest_kernel<<<1, 128, smem, st>>>(ma, mb, D, K);
    count_launch();
    return check_launch("umma_selftest");
}

}  // namespace fn2

// ------------------------------------------------------------------------------------------------
// Variant TS (round-2 groundwork, tools/umma_ts_check.py; not yet used by a kernel): the same product as
// variant 1 with A read from TENSOR MEMORY.  Thread = row m writes its K bf16 values with tcgen05.st, two per
// 32-bit column (ASSUMED layout: column k/2 of lane m, even k in the low half); B comes through TMA as before.
// ------------------------------------------------------------------------------------------------
namespace fn2 {

__global__ void __launch_bounds__(128, 1)
umma_selftest_ts_kernel(const __nv_bfloat16_raw *__restrict__ A, const __grid_constant__ CUtensorMap mapB,
                        float *__restrict__ D, int K) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sB = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // [144 rows][128 B], SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + ST_N * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_tmem = tmem_base + 256;                      // columns [256, 256 + K/2)
    {   // A row `tid` -> TMEM lane `tid` (warp w owns lanes [32w, 32w+32))
        const unsigned short *arow = reinterpret_cast<const unsigned short *>(A) + (size_t)tid * K;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                r[j] = (uint32_t)arow[2 * (c0 + j)] | ((uint32_t)arow[2 * (c0 + j) + 1] << 16);
            tmem_st8(a_tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        }
        tmem_st_wait();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t idesc = umma_idesc_bf16_f32(ST_M, ST_N);
    const int nkb = K / ST_KB;
    for (int kb = 0; kb < nkb; ++kb) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[0], ST_N * 128);
            tma_load_2d(sB, &mapB, &bars[0], kb * ST_KB, 0);
            mbar_wait(&bars[0], kb & 1);
            tcgen05_fence_after();
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
#pragma unroll
            for (int ks = 0; ks < ST_KB / 16; ++ks)
                umma_bf16_ts(tmem_base, a_tmem + (kb * ST_KB + ks * 16) / 2, db + 2 * ks, idesc, (kb | ks) != 0);
            umma_commit(&bars[1]);
            mbar_wait(&bars[1], kb & 1);
        }
        __syncthreads();
    }
    tcgen05_fence_after();
    for (int c0 = 0; c0 < ST_N; c0 += 16) {
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * ST_N + c0 + j] = r[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// Variant TS-cp: A arrives in shared memory through TMA exactly as in variant 1 (K-major, 128-byte swizzle) and is copied
// into tensor memory with tcgen05.cp.128x256b (one K = 16 step per copy) by the MMA-issuing thread, straight before the
// TS-mode MMAs that read it -- the route the forward correlation kernel uses for its A_hi operand.
__global__ void __launch_bounds__(128, 1)
umma_selftest_tscp_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                          float *__restrict__ D, int K) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // [128 rows][128 B], SW128
    unsigned char *sB = sA + ST_M * 128;                                                // [144 rows][128 B], SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + ST_N * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_tmem = tmem_base + 288;                      // columns [288, 288 + K/2): next to two 144-column accumulators
    const uint32_t idesc = umma_idesc_bf16_f32(ST_M, ST_N);
    const int nkb = K / ST_KB;
    for (int kb = 0; kb < nkb; ++kb) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[0], (ST_M + ST_N) * 128);
            tma_load_2d(sA, &mapA, &bars[0], kb * ST_KB, 0);
            tma_load_2d(sB, &mapB, &bars[0], kb * ST_KB, 0);
            mbar_wait(&bars[0], kb & 1);
            tcgen05_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
#pragma unroll
            for (int ks = 0; ks < ST_KB / 16; ++ks) umma_cp_128x256b(a_tmem + (kb * ST_KB + ks * 16) / 2, da + 2 * ks);
#pragma unroll
            for (int ks = 0; ks < ST_KB / 16; ++ks)
                umma_bf16_ts(tmem_base, a_tmem + (kb * ST_KB + ks * 16) / 2, db + 2 * ks, idesc, (kb | ks) != 0);
            umma_commit(&bars[1]);
            mbar_wait(&bars[1], kb & 1);
        }
        __syncthreads();
    }
    tcgen05_fence_after();
    for (int c0 = 0; c0 < ST_N; c0 += 16) {
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * ST_N + c0 + j] = r[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int umma_selftest_tscp(const void *A, const void *B, float *D, int K, cudaStream_t st) {
    if (K <= 0 || K % ST_KB || K > 256) return fail(FN2B200_EINVAL, "umma_selftest_tscp: K must be 64, 128, 192 or 256");
    CUtensorMap ma, mb;
    uint64_t dimsA[2] = {(uint64_t)K, ST_M}, dimsB[2] = {(uint64_t)K, ST_N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t boxA[2] = {ST_KB, ST_M}, boxB[2] = {ST_KB, ST_N};
    int rc = make_tensor_map_bf16_sw128(&ma, A, 2, dimsA, str, boxA);
    if (rc) return rc;
    rc = make_tensor_map_bf16_sw128(&mb, B, 2, dimsB, str, boxB);
    if (rc) return rc;
    const int smem = (ST_M + ST_N) * 128 + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_tscp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "umma_selftest_tscp: smem attribute (%s)", cudaGetErrorString(e));
    umma_selftest_tscp_kernel<<<1, 128, smem, st>>>(ma, mb, D, K);
    count_launch();
    return check_launch("umma_selftest_tscp");
}

int umma_selftest_ts(const void *A, const void *B, float *D, int K, cudaStream_t st) {
    if (K <= 0 || K % ST_KB || K > 256) return fail(FN2B200_EINVAL, "umma_selftest_ts: K must be 64, 128, 192 or 256");
    CUtensorMap mb;
    uint64_t dimsB[2] = {(uint64_t)K, ST_N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t boxB[2] = {ST_KB, ST_N};
    int rc = make_tensor_map_bf16_sw128(&mb, B, 2, dimsB, str, boxB);
    if (rc) return rc;
    const int smem = ST_N * 128 + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "umma_selftest_ts: smem attribute (%s)", cudaGetErrorString(e));
    umma_selftest_ts_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16_raw *>(A), mb, D, K);
    count_launch();
    return check_launch("umma_selftest_ts");
}

}  // namespace fn2

// ------------------------------------------------------------------------------------------------
// Variant 2 (the backward kernel's operand forms): D[128 x 64] = A[128 x K] * Bt[K x 64] with
//   A : K-major, NO swizzle, written to shared memory by the threads themselves (core-matrix layout)
//   Bt: MN-major (N contiguous), SW128, loaded by TMA from a row-major [K][64] bf16 matrix.
// K = 144 (9 k-steps), like one unit of the backward kernel.
// ------------------------------------------------------------------------------------------------
namespace fn2 {

constexpr int ST2_K = 144, ST2_N = 64;

__global__ void __launch_bounds__(128, 1)
umma_selftest2_kernel(const __nv_bfloat16_raw *__restrict__ A, const __grid_constant__ CUtensorMap mapB,
                      float *__restrict__ D, int a_sw32) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sB = smem;                              // [144 rows(K)][128 B], SW128 (MN-major)
    unsigned char *sA = smem + ST2_K * 128;                // 9 k-steps x [2 chunks][16 groups][8 rows x 16 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + 9 * 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    // every thread writes its own row of A into the no-swizzle core-matrix layout
    {
        const int p = tid;
        const unsigned short *arow = reinterpret_cast<const unsigned short *>(A) + (size_t)p * ST2_K;
        for (int k = 0; k < ST2_K; ++k) {
            const int ks = k >> 4, c = (k >> 3) & 1, e = k & 7;
            const uint32_t off = a_sw32 ? ks * 4096 + sw32_offset(p, k & 15)
                                        : ks * 4096 + c * 2048 + (p >> 3) * 128 + (p & 7) * 16 + e * 2;
            *reinterpret_cast<unsigned short *>(sA + off) = arow[k];
        }
    }
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], ST2_K * 128);
        tma_load_2d(sB, &mapB, &bars[0], 0, 0);
        mbar_wait(&bars[0], 0);
        tcgen05_fence_after();
        const uint32_t idesc = umma_idesc_bf16_f32(128, ST2_N, 1);
        const uint64_t db = umma_desc_mn_sw128(smem_u32(sB), ST2_K * 128);
#pragma unroll
        for (int ks = 0; ks < ST2_K / 16; ++ks) {
            const uint64_t da = a_sw32 ? umma_desc_k_sw32(smem_u32(sA + ks * 4096))
                                       : umma_desc_k_noswz(smem_u32(sA + ks * 4096), 2048, 128);
            umma_bf16_ss(tmem_base, da, db + (uint64_t)((ks * 16 * 128) >> 4), idesc, ks != 0);
        }
        umma_commit(&bars[1]);
        mbar_wait(&bars[1], 0);
    }
    __syncthreads();
    tcgen05_fence_after();
    for (int c0 = 0; c0 < ST2_N; c0 += 16) {
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[tid * ST2_N + c0 + j] = r[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

int umma_selftest2(const void *A, const void *Bt, float *D, int a_sw32, cudaStream_t st) {
    CUtensorMap mb;
    uint64_t dims[2] = {ST2_N, ST2_K};
    uint64_t str[1] = {ST2_N * 2};
    uint32_t box[2] = {ST2_N, ST2_K};
    int rc = make_tensor_map_bf16_sw128(&mb, Bt, 2, dims, str, box);
    if (rc) return rc;
    const int smem = ST2_K * 128 + 9 * 4096 + 1024 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "umma_selftest2: smem attribute (%s)", cudaGetErrorString(e));
    umma_selftest2_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16_raw *>(A), mb, D, a_sw32);
    count_launch();
    return check_launch("umma_selftest2");
}

}  // namespace fn2

// ------------------------------------------------------------------------------------------------
// MMA issue-rate micro-benchmark: every SM issues `iters` back-to-back M128 x N x K16 bf16 MMAs on fixed
// (zeroed) operands and reports cycles per MMA.  mode 0: A and B from shared memory, K-major SW128 (the
// forward's operands); mode 1: A from tensor memory, B as mode 0; mode 2: A and B MN-major SW128 (the
// backward's).  extra_smem_kb: a second warp streams that many KB of generic shared-memory writes per
// MMA batch to emulate TMA / epilogue traffic (0 = none).
// ------------------------------------------------------------------------------------------------
namespace fn2 {

__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(float *__restrict__ out, int mode, int N, int iters) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);      // A: 128 x 64, B: up to 256 x 64
        const uint32_t idesc = umma_idesc_bf16_f32(128, N, mode == 2, mode == 2);
        const uint64_t ad = mode == 2 ? umma_desc_mn_sw128(a_addr, 8192) : umma_desc_k_sw128(a_addr);
        const uint64_t bd = mode == 2 ? umma_desc_mn_sw128(b_addr, 8192) : umma_desc_k_sw128(b_addr);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = mode == 2 ? (uint64_t)((ks * 16 * 128) >> 4) : (uint64_t)(2 * ks);
                    if (mode == 1) umma_bf16_ts(tmem_base, tmem_base + 256 + 8 * ks, bd + adv, idesc, 1);
                    else umma_bf16_ss(tmem_base, ad + adv, bd + adv, idesc, 1);
                }
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (tid == 0) out[blockIdx.x] = (float)(t1 - t0) / (4.0f * iters);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int umma_rate_bench(float *out, int mode, int N, int iters, cudaStream_t st) {
    if (mode < 0 || mode > 2 || N < 16 || N > 256 || (N & 15) || iters < 1)
        return fail(FN2B200_EINVAL, "umma_rate_bench: mode %d N %d iters %d", mode, N, iters);
    const int smem = 64 * 1024 + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "umma_rate_bench: smem attribute (%s)", cudaGetErrorString(e));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    umma_rate_kernel<<<sms, 128, smem, st>>>(out, mode, N, iters);
    count_launch();
    return check_launch("umma_rate_bench");
}

}  // namespace fn2

// ------------------------------------------------------------------------------------------------
// TMA feed micro-benchmark (tools/tma_feed.py): how fast can one SM pull halo-style boxes
// (64 channels x bw x bh class pixels of a [img][Hc][Wc][C] bf16 tensor, SW128) when nothing consumes
// them?  Persistent CTAs, `stages`-deep ring, `per_stage` boxes per stage, access pattern of the
// correlation kernels (tiles of 8x16, 7 units, C/64 k-blocks).  Writes cycles and bytes per CTA.
// ------------------------------------------------------------------------------------------------
namespace fn2 {

__global__ void __launch_bounds__(256, 1)
tma_feed_kernel(const __grid_constant__ CUtensorMap map, long long *__restrict__ out, int C, int Hc, int Wc,
                int nimg, int bw, int bh, int stages, int per_stage, int iters, int warps) {
    // `warps` producer warps, each with a private `stages`-deep ring (own slots, own barriers)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int box_bytes = bw * bh * 128;
    const int tid = threadIdx.x, wid = tid >> 5;
    uint64_t *bars0 = reinterpret_cast<uint64_t *>(smem0 + (size_t)warps * stages * per_stage * box_bytes);
    if (tid == 0) {
        prefetch_tensormap(&map);
        for (int i = 0; i < warps * stages; ++i) mbar_init(&bars0[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (wid >= warps) return;
    unsigned char *smem = smem0 + (size_t)wid * stages * per_stage * box_bytes;
    uint64_t *bars = bars0 + wid * stages;
    const int nkb = C / 64, nxt = (Wc + 15) / 16, nyt = (Hc + 7) / 8;
    const int ntiles = nimg * nxt * nyt;
    long long t0 = 0, bytes = 0;
    int issued = 0, waited = 0;
    auto issue = [&](int n0) {
        // n-th stage of this CTA: tile, unit, k-block as the correlation kernels walk them
        const int n = n0 * warps + wid;
        const int per_tile = 7 * nkb;
        const int tile = (blockIdx.x + (n / per_tile) * gridDim.x) % ntiles;
        const int u = (n % per_tile) / nkb, kb = n % nkb;
        const int img = tile / (nxt * nyt), yc0 = ((tile / nxt) % nyt) * 8, xc0 = (tile % nxt) * 16;
        const int s = n0 % stages;
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(&bars[s], (uint32_t)(per_stage * box_bytes));
            for (int b = 0; b < per_stage; ++b)
                tma_load_4d(smem + ((size_t)s * per_stage + b) * box_bytes, &map, &bars[s], kb * 64,
                            xc0 - 10 + (b * bw) % 36, yc0 - 10 + u * 4, (img + b) % nimg);
        }
        __syncwarp();
    };
    for (; issued < stages && issued < iters; ++issued) issue(issued);
    t0 = clock64();
    for (; waited < iters; ++waited) {
        mbar_wait(&bars[waited % stages], (waited / stages) & 1);
        bytes += (long long)per_stage * box_bytes;
        if (issued < iters) { issue(issued); ++issued; }
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) {
        atomicMax((unsigned long long *)&out[2 * blockIdx.x], (unsigned long long)(t1 - t0));
        atomicAdd((unsigned long long *)&out[2 * blockIdx.x + 1], (unsigned long long)bytes);
    }
}

// Cluster variant: CTA rank 0 of each cluster issues every box with .multicast::cluster to all `cs`
// CTAs (same smem offset, same barrier offset in each); a slot is re-armed once every CTA of the
// cluster has seen it complete (remote arrivals on rank 0's `empty` barriers).  out[] as above, bytes =
// what landed in that CTA.  Answers: is the ~28 B/clk/SM unicast ceiling an SM-ingest or an L2-side limit?
__device__ __forceinline__ void tma_load_4d_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}

__global__ void __launch_bounds__(64, 1)
tma_feed_mc_kernel(const __grid_constant__ CUtensorMap map, long long *__restrict__ out, int C, int Hc, int Wc,
                   int nimg, int bw, int bh, int stages, int per_stage, int iters, int cs_signed) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int cs = cs_signed < 0 ? -cs_signed : cs_signed, cs_mode = cs_signed < 0;
    const int box_bytes = bw * bh * 128;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * per_stage * box_bytes);
    uint64_t *empty = full + stages;
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_ctarank();
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (uint32_t)cs); }
        fence_barrier_init();
    }
    cluster_sync_all();
    if (tid < 32) {
        const int nkb = C / 64, nxt = (Wc + 15) / 16, nyt = (Hc + 7) / 8;
        const int ntiles = nimg * nxt * nyt;
        const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
        const uint16_t mask = (uint16_t)((1u << cs) - 1u);
        long long bytes = 0;
        auto arm = [&](int n) {
            if (elect_one_sync()) mbar_arrive_expect_tx(&full[n % stages], (uint32_t)(per_stage * box_bytes));
            __syncwarp();
        };
        const bool all_issue = cs_mode != 0;       // every rank issues the boxes b % cs == rank
        auto issue = [&](int n) {
            const int per_tile = 7 * nkb;
            const int tile = (cid + (n / per_tile) * ncl) % ntiles;
            const int u = (n % per_tile) / nkb, kb = n % nkb;
            const int img = tile / (nxt * nyt), yc0 = ((tile / nxt) % nyt) * 8, xc0 = (tile % nxt) * 16;
            const int s = n % stages;
            if (elect_one_sync())
                for (int b = 0; b < per_stage; ++b)
                    if (!all_issue || (uint32_t)(b % cs) == rank)
                    tma_load_4d_mc(smem + ((size_t)s * per_stage + b) * box_bytes, &map, &full[s], kb * 64,
                                   xc0 - 10 + (b * bw) % 36, yc0 - 10 + u * 4, (img + b) % nimg, mask);
            __syncwarp();
        };
        int n = 0;
        for (; n < stages && n < iters; ++n) { arm(n); if (rank == 0 || all_issue) issue(n); }
        const long long t0 = clock64();
        for (int w = 0; w < iters; ++w) {
            const int s = w % stages;
            mbar_wait(&full[s], (w / stages) & 1);
            bytes += (long long)per_stage * box_bytes;
            if (elect_one_sync()) {
                if (all_issue) for (int r = 0; r < cs; ++r) mbar_arrive_remote(&empty[s], (uint32_t)r);
                else mbar_arrive_remote(&empty[s], 0);
            }
            __syncwarp();
            if (n < iters) {
                arm(n);
                if (rank == 0 || all_issue) { mbar_wait(&empty[s], (w / stages) & 1); issue(n); }
                ++n;
            }
        }
        const long long t1 = clock64();
        if (tid == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = bytes; }
    }
    cluster_sync_all();     // nobody leaves while a peer may still signal its barriers
}

int tma_feed_bench(const void *base, long long *out, int nimg, int C, int Hc, int Wc, int bw, int bh, int stages,
                   int per_stage, int iters, int grid, int cluster, int warps, cudaStream_t st) {
    CUtensorMap m;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wc, (uint64_t)Hc, (uint64_t)nimg};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)Wc * C * 2, (uint64_t)Hc * Wc * C * 2};
    uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
    int rc = make_tensor_map_bf16_sw128(&m, base, 4, dims, strides, box);
    if (rc) return rc;
    if (warps < 1 || warps > 8) return fail(FN2B200_EINVAL, "tma_feed_bench: producer warps %d", warps);
    const int smem = warps * (stages * per_stage * bw * bh * 128 + 2 * stages * 8) + 1024 + 64;
    if (smem > 232448) return fail(FN2B200_EINVAL, "tma_feed_bench: %d bytes of shared memory requested", smem);
    const int cmode = cluster;                 // negative: every rank issues its share of the boxes
    if (cluster < -1) cluster = -cluster;
    if (cluster <= 1) {
        cudaError_t e = cudaFuncSetAttribute(tma_feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail((int)e, "tma_feed_bench: smem attribute (%s)", cudaGetErrorString(e));
        cudaMemsetAsync(out, 0, sizeof(long long) * 2 * grid, st);
        tma_feed_kernel<<<grid, 256, smem, st>>>(m, out, C, Hc, Wc, nimg, bw, bh, stages, per_stage, iters, warps);
    } else {
        if (cluster > 8 || grid % cluster) return fail(FN2B200_EINVAL, "tma_feed_bench: cluster %d, grid %d", cluster, grid);
        cudaError_t e = cudaFuncSetAttribute(tma_feed_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail((int)e, "tma_feed_bench: smem attribute (%s)", cudaGetErrorString(e));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, tma_feed_mc_kernel, m, out, C, Hc, Wc, nimg, bw, bh, stages, per_stage, iters, cmode);
        if (e != cudaSuccess) return fail((int)e, "tma_feed_bench: cluster launch (%s)", cudaGetErrorString(e));
    }
    count_launch();
    return check_launch("tma_feed_bench");
}

}  // namespace fn2
