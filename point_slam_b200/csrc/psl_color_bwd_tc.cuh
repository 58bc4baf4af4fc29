// Layout constants of the tensor-core colour backward (transposed operand blob, TMEM column regions, shared memory map) and its
// launch arguments, shared by the production kernel (psl_color_bwd_tc.cu) and the 16-worker-warp experiment
// (psl_color_bwd_tc_w16.cu).
#pragma once
#include "psl_decode.cuh"
#include "psl_tc.cuh"
#include "psl_tc_layout.cuh"

namespace psl {
namespace cbt {

constexpr int TM = 128, NWORK = 256, NTHR = 320;
constexpr uint32_t TP = 0, TQ = 128, TR = 256, TDC = 384, TDE = 416, TDX = 416;    // TMEM column regions (DX reuses DE)

// ---- backward blob (float offsets): every matrix transposed ([k_in][n_out]) as canonical K-major hi|lo chunk images --
constexpr int BB_N2T = 0;                              // N2^T  (128 rows x 32 k)
constexpr int BB_N1T = BB_N2T + 2 * 128 * 32;          // N1^T  (64 rows x 128 k), rows >= 52 zero
constexpr int BB_VEC = BB_N1T + 2 * 64 * 128;          // Brel [3][12] (48) | Bc [3][20] (64)
constexpr int BV_BREL = 0, BV_BC = 48, BV_SIZE = 112;
constexpr int BB_TRUNK = BB_VEC + BV_SIZE;
constexpr int NMAT = 12;   // WoT GoT | La4 G4 | La3 G3 Le3 | La2 G2 | La1 G1 | Le0
__host__ __device__ constexpr int mat_n(int q) { return (q == 0 || q == 2 || q == 4 || q == 7 || q == 9) ? 128 : ((q == 6 || q == 11) ? 48 : 32); }
__host__ __device__ constexpr int mat_k(int q) { return q < 2 ? 16 : 128; }
__host__ __device__ constexpr int mat_off(int q) {
    int o = 0;
    for (int i = 0; i < q; ++i) o += 2 * mat_n(i) * mat_k(i);
    return o;
}
constexpr int BB_TOTAL = BB_TRUNK + mat_off(NMAT);
static_assert(BB_TRUNK % 4 == 0, "alignment");

// shared memory (bytes)
constexpr int SB_NBRW = 0;                             // 98304 resident: N2T | N1T
constexpr int SB_RING = 98304;                         // 2 x 32768
constexpr int SB_VEC = SB_RING + 65536;                // BV_SIZE floats
constexpr int SB_AFF = SB_VEC + BV_SIZE * 4;           // 12 floats (+4 pad)
constexpr int SB_RED = SB_AFF + 64;                    // [4 warps][32] dBrel accumulators
constexpr int SB_BAR = SB_RED + 4 * 32 * 4;
constexpr int SB_TOTAL = SB_BAR + 16 * 8;

struct Args {
    psl_decode_cfg cfg;
    const float* blob;
    const float* pos; long long m;
    const int* I; const float* D; const int* nnum; const double* r2;
    const float* cloud_pos; const float* col_feats; const float* affine;
    const float* raw; const float* d_raw; const float* tsave;
    float* tbwd;                 // TBwd buffer (dhT, doutT, dz1T, dccT, aff)
    float* d_colpair;            // (m,8,32) rel / (m,32) = d_cc otherwise
    float* wn_out;               // (m,8) weights for the feature scatter (0 where masked)
    float* dwn_col;              // (m,8) dL/d(normalised weights), colour part
    float* dpos_col;             // (m,3) dL/dpos, colour part (embedding + rel-pos), NULL when not wanted
    float* part_brel;            // [grid][32] per-CTA partial of dL/dBrel
    int want_wgrad;              // store dhT / dz1T / doutT for the weight-gradient kernel
};

__device__ __forceinline__ float sp_grad_fast(float z) {          // d softplus_100 / dz = sigmoid(100 z)
    return __fdividef(1.0f, 1.0f + __expf(-fminf(100.0f * z, 30.0f)));
}

}  // namespace cbt
}  // namespace psl
