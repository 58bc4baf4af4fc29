// Layout constants of the tensor-core colour forward (operand blob, TMEM column regions, shared memory map) and its launch
// arguments, shared by the production kernel (psl_color_tc.cu) and the 16-worker-warp experiment (psl_color_tc_w16.cu).
#pragma once
#include "psl_decode.cuh"
#include "psl_tc.cuh"
#include "psl_tc_layout.cuh"

namespace psl {
namespace ctc {

constexpr int TM = 128;                 // sample rows per tile (= TMEM lanes)
constexpr int NWORK = 256, NTHR = 320;
constexpr uint32_t TP = 0, TQ = 128, TR = 256, TCC = 384, TSP = 448;   // TMEM column regions

// ---- packed tensor-core blob (float offsets) -----------------------------------------------------------------------
constexpr int TB_N1 = 0;                              // N1 hi | lo, canonical (128 rows x 64 k), cols >= 52 zero
constexpr int TB_N2 = TB_N1 + 2 * 128 * 64;           // N2 hi | lo, canonical (32 rows x 128 k)
constexpr int TB_VEC = TB_N2 + 2 * 32 * 128;          // small vectors, copied to shared memory by the workers:
constexpr int V_B1 = 0, V_B2 = 128, V_BIAS = 160, V_BOUT = 800, V_BC = 816, V_BREL = 880, V_SIZE = 928;
constexpr int TB_TRUNK = TB_VEC + V_SIZE;             // chunk stream of the 6 trunk layers
constexpr int NLAYER = 6;                             // 5 trunk layers + output layer
__host__ __device__ constexpr int l_ne(int l) { return (l == 0 || l == 3) ? 5 : 0; }      // k-steps fed by the embedding (SS)
__host__ __device__ constexpr int l_na(int l) { return l == 0 ? 0 : 16; }                  // k-steps fed by act(z) (TS)
__host__ __device__ constexpr int l_nc(int l) { return l == 0 ? 0 : 4; }                   // k-steps fed by c (TS)
__host__ __device__ constexpr int l_ks(int l) { return l_ne(l) + l_na(l) + l_nc(l); }
__host__ __device__ constexpr int l_n(int l) { return l == 5 ? 16 : 128; }
__host__ __device__ constexpr int l_off(int l) {       // float offset of layer l inside the trunk stream
    int o = 0;
    for (int i = 0; i < l; ++i) o += l_ks(i) * 16 * l_n(i);
    return o;
}
constexpr int TB_TOTAL = TB_TRUNK + l_off(NLAYER);
constexpr int FOLD_LD = 200;                          // row stride of the fp32 folded matrices (scratch)
constexpr int FOLD_FLOATS = NLAYER * 128 * FOLD_LD;

// ---- shared memory (bytes) ----------------------------------------------------------------------------------------------
constexpr int SB_NBRW = 0;                            // 98304: resident neighbour-MLP weights
constexpr int SB_RING = 98304;                        // 2 x 32768
constexpr int SB_E = SB_RING + 65536;                 // embedding A operand, hi | lo, canonical (128 x 40): 40960
constexpr int SB_VEC = SB_E + 40960;                  // V_SIZE floats
constexpr int SB_RAND = SB_VEC + V_SIZE * 4;          // 32 floats rand_col + 12 affine (+pad)
constexpr int SB_BAR = SB_RAND + 48 * 4;              // mbarriers
constexpr int SB_TOTAL = SB_BAR + 16 * 8;
static_assert(SB_TOTAL <= 227 * 1024, "shared memory over budget");

struct Args {
    psl_decode_cfg cfg;
    const float* blob;                  // packed tensor-core blob
    const float* pos; long long m;
    const int* I; const float* D; const int* nnum; const double* r2;
    const float* cloud_pos; const float* col_feats; const float* rand_col; const float* affine;
    float* raw;                         // (m,4): xyz written here, w (occupancy) untouched
    float* save;                        // SAVE == 1: psl_decode.cuh SaveLayout (consumed by the FFMA backward)
    float* tsave;                       // SAVE == 2: psl_tc_layout.cuh TSave (consumed by the tensor-core backward)
};

}  // namespace ctc
}  // namespace psl
