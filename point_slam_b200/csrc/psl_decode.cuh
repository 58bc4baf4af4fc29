// Shared definitions of the decode kernels: packed-parameter blob layout, tile geometry, save layout.
#pragma once
#include "psl_common.cuh"

namespace psl {

constexpr int TS = 64;     // samples per tile (one CTA iteration)
constexpr int LD = 68;     // row stride (floats) of the [channel][sample] activation arrays in shared memory
constexpr int NWARP = 8;   // warps per CTA
constexpr int SPW = 8;     // samples per warp

// ---- packed parameter blob (float offsets).  All matrices are stored TRANSPOSED, [K][N] with N contiguous,
// ---- so that lane l reads output channels l, l+32, ... without bank conflicts.
// geometry stage (staged as one block)
constexpr int G_B = 0;                       // [3][96]   embedder._B, cols 93..95 zero
constexpr int G_L0 = G_B + 3 * 96;           // [96][32]  pts_linears.0 (rows 93..95 zero)
constexpr int G_L1 = G_L0 + 96 * 32;         // [32][32]
constexpr int G_L2 = G_L1 + 32 * 32;
constexpr int G_L3 = G_L2 + 32 * 32;         // [128][32] rows 0..95: embedding part, rows 96..127: hidden part
constexpr int G_L4 = G_L3 + 128 * 32;
constexpr int G_FC = G_L4 + 32 * 32;         // 5 x [32][32]  fc_c
constexpr int G_BIAS = G_FC + 5 * 32 * 32;   // 5 x [32]
constexpr int G_BIASC = G_BIAS + 5 * 32;     // 5 x [32]
constexpr int G_WO = G_BIASC + 5 * 32;       // [32]
constexpr int G_BO = G_WO + 32;              // [1] (+3 pad)
constexpr int G_SIZE = G_BO + 4;
// neighbour-MLP stage
constexpr int N_BREL = 0;                    // [3][12]  embedder_rel_pos._B, cols 10..11 zero
constexpr int N_W1 = N_BREL + 36;            // [52][128] rows 0..19: rel-pos embedding, rows 20..51: feature
constexpr int N_W2 = N_W1 + 52 * 128;        // [128][32]
constexpr int N_B1 = N_W2 + 128 * 32;        // [128]
constexpr int N_B2 = N_B1 + 128;             // [32]
constexpr int N_SIZE = N_B2 + 32;
// colour trunk layer i:  [K_i][128] weights | [32][128] fc_c | [128] b | [128] bc | extra
__host__ __device__ constexpr int col_k(int i) { return i == 0 ? 40 : (i == 3 ? 168 : 128); }
__host__ __device__ constexpr int CL_FC(int i) { return col_k(i) * 128; }
__host__ __device__ constexpr int CL_B(int i) { return CL_FC(i) + 32 * 128; }
__host__ __device__ constexpr int CL_BC(int i) { return CL_B(i) + 128; }
__host__ __device__ constexpr int CL_X(int i) { return CL_BC(i) + 128; }                 // i==0: Bc [3][20]; i==4: Wo^T [128][4], bo[4]
__host__ __device__ constexpr int CL_SIZE(int i) { return CL_X(i) + (i == 0 ? 60 : (i == 4 ? 516 : 0)); }
constexpr int OFF_GEO = 0;
constexpr int OFF_NBR = OFF_GEO + G_SIZE;
__host__ __device__ constexpr int OFF_COL(int i) {
    int o = OFF_NBR + N_SIZE;
    for (int j = 0; j < i; ++j) o += CL_SIZE(j);
    return o;
}
constexpr int PACKED_FLOATS = OFF_COL(5);
constexpr int SW_FLOATS = CL_SIZE(3);        // largest stage (25856 floats)
static_assert(G_SIZE <= SW_FLOATS && N_SIZE <= SW_FLOATS, "stage does not fit");
static_assert(G_SIZE % 4 == 0 && N_SIZE % 4 == 0 && CL_SIZE(0) % 4 == 0 && CL_SIZE(4) % 4 == 0, "float4 staging");

// ---- activations saved by the forward for the backward: section-major, each section [M][width] ----------------
struct SaveLayout {
    long long cg, gz, gh, cc, cz, ch, nz1, nf, total;     // float offsets per sample (multiply by M for the base)
};
__host__ __device__ inline SaveLayout save_layout(int stage_color, int rel) {
    SaveLayout L{};
    long long o = 0;
    L.cg = o; o += 32;
    L.gz = o; o += 5 * 32;
    L.gh = o; o += 5 * 32;
    L.cc = o; if (stage_color) o += 32;
    L.cz = o; if (stage_color) o += 5 * 128;
    L.ch = o;                                  // (h is recomputed by the backward from z and c: not stored)
    L.nz1 = o; if (stage_color && rel) o += 8 * 128;
    L.nf = o; if (stage_color && rel) o += 8 * 32;
    L.total = o;
    return L;
}

struct DecodeArgs {
    psl_decode_cfg cfg;
    const float* packed;
    const float* pos; long long m;
    const int* I; const float* D; const int* nnum; const double* r2;
    const float* cloud_pos; const float* geo_feats; const float* col_feats;
    const float* rand_geo; const float* rand_col; const float* affine;
    float* raw; unsigned char* has_nb; float* save;
};

// one dense layer for the 8 samples of a warp: acc[j][s] += sum_k Wt[k][lane+32j] * in[k][s]
template <int NJ>
__device__ __forceinline__ void dense8(float (&acc)[NJ][8], const float* __restrict__ in, int K,
                                       const float* __restrict__ Wt, int lane, int ldw = 32 * NJ) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(in + k * LD);
        const float4 a1 = *reinterpret_cast<const float4*>(in + k * LD + 4);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float w = Wt[k * ldw + lane + 32 * j];
            acc[j][0] = fmaf(w, a0.x, acc[j][0]); acc[j][1] = fmaf(w, a0.y, acc[j][1]);
            acc[j][2] = fmaf(w, a0.z, acc[j][2]); acc[j][3] = fmaf(w, a0.w, acc[j][3]);
            acc[j][4] = fmaf(w, a1.x, acc[j][4]); acc[j][5] = fmaf(w, a1.y, acc[j][5]);
            acc[j][6] = fmaf(w, a1.z, acc[j][6]); acc[j][7] = fmaf(w, a1.w, acc[j][7]);
        }
    }
}

__device__ __forceinline__ void stage_weights(float* sW, const float* __restrict__ src, int nfloats) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(sW);
    for (int i = threadIdx.x; i < nfloats / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- geometry branch on the warp-level tensor-core path (psl_geo_mma.cu): data path without parameter gradients -----------------
// selected by bit 1 of psl_decode_cfg.reserved on a geometry-stage call; the fragment images live behind the FFMA blob in `packed`
constexpr int PSL_GEO_MMA_BIT = 2;
constexpr int GEO_MMA_SAVE_WORDS = 5;            // one 32-bit ReLU mask per layer
size_t geo_mma_floats();
int geo_mma_pack(const psl_decoder_params* P, float* packed, cudaStream_t st);
int geo_fwd_mma(const psl_decode_cfg* cfg, const float* packed, const float* pos, long long m, const int* I, const float* D,
                const int* nnum, const double* r2, const float* geo_feats, const float* rand_geo, float* raw, unsigned char* has_nb,
                float* save, cudaStream_t st);
int geo_bwd_mma(const psl_decode_cfg* cfg, const float* packed, const float* pos, long long m, const int* I, const float* D,
                const int* nnum, const double* r2, const float* cloud_pos, const float* geo_feats, const float* save,
                const float* d_raw, float* d_pos, float* d_cg, float* wn, const float* dwn_extra, const float* dpos_extra,
                cudaStream_t st);

int geo_idw_chain(const psl_decode_cfg* cfg, const float* pos, long long m, const int* I, const float* D, const double* r2,
                  const float* cloud_pos, const float* dwn, const float* dpos_add, float* d_pos, cudaStream_t st);

// IDW weight of one (sample, neighbour) slot before normalisation (decoder.py:152-157)
__device__ __forceinline__ float idw_raw(float D, int idx, float thr_le, int weighting) {
    if (idx < 0 || !(D <= thr_le)) return 0.f;
    if (weighting == PSL_WEIGHT_EXPO) return expf(-20.0f * sqrtf(D));
    return __fdiv_rn(1.0f, __fadd_rn(D, 1e-10f));
}

}  // namespace psl
