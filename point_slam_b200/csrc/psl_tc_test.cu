// Self-test of the tensor-core building blocks (psl_tc.cuh): D (128 x N) = A (128 x K) * W (N x K)^T with 3xTF32,
// A either resident in TMEM (TS form, mode 0) or in shared memory (SS form, mode 1), W copied with cp.async.bulk from
// a pre-packed canonical image.  Used by tests/test_gpu_tc.py; not on the product path.
#include "psl_common.cuh"
#include "psl_tc.cuh"

namespace psl {

// pack W (N,K) row-major -> [hi plane | lo plane], each in the canonical K-major layout (floats)
__global__ void k_tc_pack_w(const float* __restrict__ W, int N, int K, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * K) return;
    const int n = e / K, k = e - n * K;
    float hi, lo;
    tc::split_tf32(W[e], hi, lo);
    out[tc::canon_off_floats(n, k, N)] = hi;
    out[N * K + tc::canon_off_floats(n, k, N)] = lo;
}

__global__ void __launch_bounds__(160, 1) k_tc_gemm_test(const float* __restrict__ A, const float* __restrict__ Wp,
                                                         float* __restrict__ D, int K, int N, int mode) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* sW = reinterpret_cast<float*>(smem_raw);                 // 2 * N * K floats
    float* sA = sW + 2 * N * K;                                     // 2 * 128 * K floats (SS mode)
    __shared__ __align__(8) uint64_t bar_w, bar_done;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t A_HI = 0, A_LO = 192, D_COL = 384;
    if (warp == 4 && lane == 0) {                                    // producer: one bulk copy for both planes
        const uint32_t bytes = 2u * N * K * sizeof(float);
        tc::mbar_expect_tx(&bar_w, bytes);
        tc::bulk_g2s(sW, Wp, bytes, &bar_w);
    }
    if (warp < 4) {                                                  // row r of A -> TMEM lanes or shared memory
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float a = (k0 + j < K) ? A[r * K + k0 + j] : 0.f;
                tc::split_tf32(a, hi[j], lo[j]);
            }
            if (mode == 0) {
                tc::tmem_st32(lane_addr + A_HI + k0, hi);
                tc::tmem_st32(lane_addr + A_LO + k0, lo);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (k0 + j < K) {
                        sA[tc::canon_off_floats(r, k0 + j, 128)] = hi[j];
                        sA[128 * K + tc::canon_off_floats(r, k0 + j, 128)] = lo[j];
                    }
            }
        }
        if (mode == 0) tc::tmem_st_wait();
        else tc::fence_proxy_async();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 4 && lane == 0) {                                    // MMA issuer
        tc::mbar_wait(&bar_w, 0);
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint32_t lbo_w = (uint32_t)N * 16u, lbo_a = 128u * 16u;
        const uint32_t w_hi = tc::smem_u32(sW), w_lo = w_hi + (uint32_t)N * K * 4u;
        const uint32_t a_hi_s = tc::smem_u32(sA), a_lo_s = a_hi_s + 128u * K * 4u;
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bh = tc::make_smem_desc(w_hi + ks * 2 * lbo_w, lbo_w, 128);
            const uint64_t bl = tc::make_smem_desc(w_lo + ks * 2 * lbo_w, lbo_w, 128);
            if (mode == 0) {
                tc::mma_tf32_ts(tmem + D_COL, tmem + A_HI + ks * 8, bh, idesc, acc); acc = 1;
                tc::mma_tf32_ts(tmem + D_COL, tmem + A_LO + ks * 8, bh, idesc, 1);
                tc::mma_tf32_ts(tmem + D_COL, tmem + A_HI + ks * 8, bl, idesc, 1);
            } else {
                const uint64_t ah = tc::make_smem_desc(a_hi_s + ks * 2 * lbo_a, lbo_a, 128);
                const uint64_t al = tc::make_smem_desc(a_lo_s + ks * 2 * lbo_a, lbo_a, 128);
                tc::mma_tf32_ss(tmem + D_COL, ah, bh, idesc, acc); acc = 1;
                tc::mma_tf32_ss(tmem + D_COL, al, bh, idesc, 1);
                tc::mma_tf32_ss(tmem + D_COL, ah, bl, idesc, 1);
            }
        }
        tc::mma_commit(&bar_done);
    }
    if (warp < 4) {
        tc::mbar_wait(&bar_done, 0);
        tc::fence_after_sync();
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 32) {
            float v[32];
            tc::tmem_ld32(lane_addr + D_COL + n0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) D[r * N + n0 + j] = v[j];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}


// ---- 16-bit planes (kind::f16, A/B formats mixed): D = Ah Wh^T + Al Wh^T + Ah Wl^T, hi = f16, lo = bf16 --------------------
// pack W (N,K) row-major -> [hi plane (f16) | lo plane (bf16)], canonical 16-bit K-major layout (2-byte elements)
__global__ void k_tc_pack_w_h(const float* __restrict__ W, int N, int K, uint16_t* __restrict__ out, int scheme) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per K-pair
    if (e >= N * K / 2) return;
    const int n = e / (K / 2), k = 2 * (e - n * (K / 2));
    uint32_t hi, lo;
    if (scheme == 1) tc::split_h2_f16(W[n * K + k], W[n * K + k + 1], hi, lo);
    else tc::split_h2(W[n * K + k], W[n * K + k + 1], hi, lo);
    const uint32_t o = tc::canon_off_h(n, k, N);
    *reinterpret_cast<uint32_t*>(out + o) = hi;
    *reinterpret_cast<uint32_t*>(out + N * K + o) = lo;
}

// mode 2: A planes in TMEM (TS), mode 3: A planes in shared memory (SS); variant 1 swaps the halves inside a TMEM word
__global__ void __launch_bounds__(160, 1) k_tc_gemm_test_h(const float* __restrict__ A, const uint16_t* __restrict__ Wp,
                                                           float* __restrict__ D, int K, int N, int mode, int variant_) {
    const int variant = variant_ & 1, scheme = (variant_ >> 1) & 3, prods = (variant_ >> 3) ? (variant_ >> 3) : 7;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint16_t* sW = reinterpret_cast<uint16_t*>(smem_raw);           // 2 * N * K halves
    uint16_t* sA = sW + 2 * N * K;                                  // 2 * 128 * K halves (SS mode)
    __shared__ __align__(8) uint64_t bar_w, bar_done;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t A_HI = 128, A_LO = 192, D_COL = 0;                // the layout of the two-tile kernels: D | A hi | A lo
    if (warp == 4 && lane == 0) {
        const uint32_t bytes = 2u * N * K * 2u;
        tc::mbar_expect_tx(&bar_w, bytes);
        tc::bulk_g2s(sW, Wp, bytes, &bar_w);
    }
    if (warp < 4) {
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 16) {                        // 16 k values = 8 packed words per plane
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t h, l;
                float a0 = A[r * K + k0 + 2 * j], a1 = A[r * K + k0 + 2 * j + 1];
                if (variant) { const float t = a0; a0 = a1; a1 = t; }
                if (scheme == 1) tc::split_h2_f16(a0, a1, h, l);
                else tc::split_h2(a0, a1, h, l);
                hi[j] = __uint_as_float(h); lo[j] = __uint_as_float(l);
            }
            if (mode == 2) {
                tc::tmem_st8(lane_addr + A_HI + k0 / 2, hi);
                tc::tmem_st8(lane_addr + A_LO + k0 / 2, lo);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t o = tc::canon_off_h(r, k0 + 2 * j, 128);
                    *reinterpret_cast<uint32_t*>(sA + o) = __float_as_uint(hi[j]);
                    *reinterpret_cast<uint32_t*>(sA + 128 * K + o) = __float_as_uint(lo[j]);
                }
            }
        }
        if (mode == 2) tc::tmem_st_wait();
        else tc::fence_proxy_async();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 4 && lane == 0) {
        tc::mbar_wait_wd(&bar_w, 0);
        tc::fence_after_sync();
        const uint32_t id_hh = tc::make_idesc_f16(128, N, tc::FMT_F16, tc::FMT_F16);
        const uint32_t lofmt = scheme == 1 ? tc::FMT_F16 : tc::FMT_BF16;
        const uint32_t id_lh = tc::make_idesc_f16(128, N, lofmt, tc::FMT_F16);
        const uint32_t id_hl = tc::make_idesc_f16(128, N, tc::FMT_F16, lofmt);
        const uint32_t lbo_w = (uint32_t)N * 16u, lbo_a = 128u * 16u;      // bytes between K-adjacent core matrices
        const uint32_t w_hi = tc::smem_u32(sW), w_lo = w_hi + (uint32_t)N * K * 2u;
        const uint32_t a_hi_s = tc::smem_u32(sA), a_lo_s = a_hi_s + 128u * K * 2u;
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 16; ++ks) {                       // one MMA covers K = 16 = two core matrices
            const uint64_t bh = tc::make_smem_desc(w_hi + ks * 2 * lbo_w, lbo_w, 128);
            const uint64_t bl = tc::make_smem_desc(w_lo + ks * 2 * lbo_w, lbo_w, 128);
            if (mode == 2) {
                if (prods & 1) { tc::mma_f16_ts(tmem + D_COL, tmem + A_HI + ks * 8, bh, id_hh, acc); acc = 1; }
                if (prods & 2) { tc::mma_f16_ts(tmem + D_COL, tmem + A_LO + ks * 8, bh, id_lh, acc); acc = 1; }
                if (prods & 4) { tc::mma_f16_ts(tmem + D_COL, tmem + A_HI + ks * 8, bl, id_hl, acc); acc = 1; }
            } else {
                const uint64_t ah = tc::make_smem_desc(a_hi_s + ks * 2 * lbo_a, lbo_a, 128);
                const uint64_t al = tc::make_smem_desc(a_lo_s + ks * 2 * lbo_a, lbo_a, 128);
                if (prods & 1) { tc::mma_f16_ss(tmem + D_COL, ah, bh, id_hh, acc); acc = 1; }
                if (prods & 2) { tc::mma_f16_ss(tmem + D_COL, al, bh, id_lh, acc); acc = 1; }
                if (prods & 4) { tc::mma_f16_ss(tmem + D_COL, ah, bl, id_hl, acc); acc = 1; }
            }
        }
        tc::mma_commit(&bar_done);
    }
    if (warp < 4) {
        tc::mbar_wait_wd(&bar_done, 0);
        tc::fence_after_sync();
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 16) {
            float v[16];
            tc::tmem_ld16(lane_addr + D_COL + n0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + j < N) D[r * N + n0 + j] = v[j];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

}  // namespace psl

using namespace psl;
extern "C" int psl_tc_gemm_test(const float* A, const float* W, float* D, float* scratch, int K, int N, int mode, psl_stream_t stream);
extern "C" int psl_tc_gemm_test_h(const float* A, const float* W, float* D, float* scratch, int K, int N, int mode, int variant, psl_stream_t stream);
// A (128,K), W (N,K), D (128,N) device fp32; scratch: 2*N*K floats.  K multiple of 8 (<= 160), N multiple of 16 (<= 128).
extern "C" int psl_tc_gemm_test(const float* A, const float* W, float* D, float* scratch, int K, int N, int mode,
                                psl_stream_t stream) {
    PSL_REQUIRE(A && W && D && scratch, "NULL argument");
    PSL_REQUIRE(K % 8 == 0 && K >= 8 && K <= 160 && N % 16 == 0 && N >= 16 && N <= 128, "unsupported K/N");
    cudaStream_t st = as_stream(stream);
    k_tc_pack_w<<<(N * K + 255) / 256, 256, 0, st>>>(W, N, K, scratch);
    const size_t smem = sizeof(float) * (2 * (size_t)N * K + (mode == 1 ? 2 * 128 * (size_t)K : 0)) + 1024;
    PSL_REQUIRE(smem <= 227 * 1024, "test shape does not fit in shared memory");
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_gemm_test<<<1, 160, smem, st>>>(A, scratch, D, K, N, mode);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// 16-bit planes: A (128,K), W (N,K), D (128,N) device fp32; scratch: N*K floats.  K multiple of 16 (<= 208), N multiple of 16 (<= 128).
// mode 2 = TS form (A planes in TMEM, two k per 32-bit column), mode 3 = SS form; variant: bit 0 swaps the halves of a TMEM word,
// bits 1-2 scheme (0: lo = bf16, formats mixed; 1: lo = f16), bits 3-5 which products to issue (0 = all; 1 hi*hi, 2 lo*hi, 4 hi*lo).
extern "C" int psl_tc_gemm_test_h(const float* A, const float* W, float* D, float* scratch, int K, int N, int mode, int variant,
                                  psl_stream_t stream) {
    PSL_REQUIRE(A && W && D && scratch, "NULL argument");
    PSL_REQUIRE(K % 16 == 0 && K >= 16 && K <= (mode == 2 ? 128 : 208) && N % 16 == 0 && N >= 16 && N <= 128 && (mode == 2 || mode == 3), "unsupported K/N/mode");
    cudaStream_t st = as_stream(stream);
    k_tc_pack_w_h<<<(N * K / 2 + 255) / 256, 256, 0, st>>>(W, N, K, reinterpret_cast<uint16_t*>(scratch), (variant >> 1) & 3);
    const size_t smem = 2 * (2 * (size_t)N * K + (mode == 3 ? 2 * 128 * (size_t)K : 0)) + 1024;
    PSL_REQUIRE(smem <= 227 * 1024, "test shape does not fit in shared memory");
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_tc_gemm_test_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_gemm_test_h<<<1, 160, smem, st>>>(A, reinterpret_cast<const uint16_t*>(scratch), D, K, N, mode, variant);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
