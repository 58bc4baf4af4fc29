// Shared device/host helpers for libpointslam_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pointslam_b200.h"

namespace psl {

void set_error(const char* fmt, ...);

#define PSL_CHECK_CUDA(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            psl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                           __FILE__, __LINE__);                                       \
            return -2;                                                                \
        }                                                                             \
    } while (0)

#define PSL_REQUIRE(cond, msg)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            psl::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);  \
            return -1;                                                                \
        }                                                                             \
    } while (0)

inline cudaStream_t as_stream(psl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// optional per-kernel device timing (psl_timing_enable): CUDA events recorded on the launching stream
enum { T_KNN = 0, T_DECODE_FWD = 1, T_DECODE_BWD = 2, T_COMPOSITE = 3, T_SCATTER = 4, T_PACK = 5, T_REDUCE = 6, T_COLOR_FWD_TC = 7,
       T_COLOR_BWD_TC = 8, T_WGRAD_TC = 9, T_SHELL = 10, T_MAP = 11, T_COUNT = 12 };
struct TimingScope {
    int slot;
    cudaStream_t st;
    TimingScope(int id, cudaStream_t s, int n_kernels = 1);
    ~TimingScope();
};

constexpr float kTwoPi = 6.283185307179586f;    // float32(2*math.pi), decoder.py:33
constexpr uint64_t kEmptyKey = ~0ull;

// ---- spatial hash ------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ int cell_coord(float v, float inv_cell) {
    return (int)floorf(__fmul_rn(v, inv_cell));
}
__device__ __forceinline__ uint64_t cell_key(int cx, int cy, int cz) {
    // 21 bits per axis, biased; z major so that x-neighbours are adjacent after the sort
    const uint64_t B = 1u << 20;
    return ((uint64_t)(cz + B) << 42) | ((uint64_t)(cy + B) << 21) | (uint64_t)(cx + B);
}

// canonical squared distance: ((dx*dx + dy*dy) + dz*dz), dx = c - q, no FMA contraction
__device__ __forceinline__ float sqdist_canonical(float cx, float cy, float cz, float qx, float qy, float qz) {
    const float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// float thresholds equivalent to the reference's float64 comparisons of a float32 D against r^2:
//   (double)D <= r2  <=>  D <= thr_le      (double)D < r2  <=>  D < thr_lt
__device__ __forceinline__ float thr_le_of(double r2) { return __double2float_rd(r2); }
__device__ __forceinline__ float thr_lt_of(double r2) { return __double2float_ru(r2); }

__device__ __forceinline__ float softplus100(float x) {          // nn.Softplus(beta=100, threshold=20)
    const float bx = x * 100.0f;
    return bx > 20.0f ? x : log1pf(expf(bx)) / 100.0f;
}
__device__ __forceinline__ float softplus100_grad(float x) {     // d/dx: z/(z+1), z = exp(beta x); 1 above threshold
    const float bx = x * 100.0f;
    if (bx > 20.0f) return 1.0f;
    const float z = expf(bx);
    return z / (z + 1.0f);
}
// Softplus(beta=100) with hardware ex2/lg2 approximations: |abs error| <~ 2e-9 (the output is >= 6.9e-3 wherever it
// matters), used by the tensor-core epilogues where the transcendental, not the MMA, is the critical path.
__device__ __forceinline__ float softplus100_fast(float x) {
    const float bx = x * 100.0f;
    const float e = __expf(fminf(bx, 20.0f));
    const float y = __logf(1.0f + e) * 0.01f;
    return bx > 20.0f ? x : y;
}
// sin / cos of a Fourier-embedding argument (|x| up to a few 1e3 rad) for the tensor-core kernels: three-term Cody-Waite
// reduction by 2*pi (k*C1 exact for |k| < 2^15) and the SFU sin/cos on [-pi, pi] -- absolute error <~ 7e-7 (MUFU 2^-21.4 +
// 2.5e-7 from the reduction) where sincosf() is ~1 ulp but costs ~35 instructions and a Payne-Hanek slow path in the
// instruction stream.  fp32 spacing of the argument itself is 2.4e-4 at 4e3 rad, so this is far below the input noise.
// -DPSL_PRECISE_TRIG restores sincosf (A/B builds, tests/test_gpu_tc.py checks both against the fp64 oracle).
__device__ __forceinline__ void sincos_embed(float x, float* s, float* c) {
#ifdef PSL_PRECISE_TRIG
    sincosf(x, s, c);
#else
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.28125f, x);
    r = fmaf(k, -0.0019353071693331003f, r);
    r = fmaf(k, -1.0253131677018246e-11f, r);
    *s = __sinf(r);
    *c = __cosf(r);
#endif
}
__device__ __forceinline__ float sin_embed(float x) { float s, c; sincos_embed(x, &s, &c); return s; }
__device__ __forceinline__ float cos_embed(float x) { float s, c; sincos_embed(x, &s, &c); return c; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace psl
