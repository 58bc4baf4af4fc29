// Geometry branch (decoder.py:130-222) on the tensor cores, data path only: IDW interpolation of the geometry features, the
// 5-block 32-wide MLP with its 93-channel Fourier embedding, occupancy output -- and the matching backward that returns the
// gradients with respect to the interpolated feature, the IDW weights and the sample position (no parameter gradients: the
// geometry decoder is frozen in tracking and, with mapping.fix_geo_decoder, in mapping; psl_decode_bwd keeps the FFMA kernel for
// the configuration that optimises it).
//
// The layers are N = 32 wide: far too narrow for a 128-row tcgen05 tile with TMEM round trips, so the GEMMs run on the warp-level
// tensor-core path (mma.sync m16n8k8, fp32 accumulate) with the 3xTF32 operand split (x = hi + lo, D += lo*hi + hi*lo + hi*hi),
// which keeps fp32 accuracy (|err| ~ 2^-21 per product).  One warp owns 16 samples from the neighbour gather to the output:
//   * activations never leave registers: the accumulator fragment of layer i (rows g, g+8; columns 8n+2t, 8n+2t+1) is fed
//     back as the A fragment of layer i+1 by permuting the reduction index (k-step s, logical k = t / t+4  <->  channel 8s+2t /
//     8s+2t+1), the permutation being applied once to the weights when they are packed;
//   * weights: pre-split hi/lo B fragments, one 16-byte load per (k-step, n-tile, lane), read through L1 (122 KB per direction,
//     shared by every warp of the SM) -- no shared-memory staging, no CTA barrier anywhere in the kernel;
//   * kept for the backward: one 32-bit ReLU mask per sample and layer (20 B / sample instead of 1.4 KB).
#include "psl_decode.cuh"

namespace psl {
namespace gm {

constexpr int WPB = 4;                                  // warps per block (independent: block size only sets the scheduling grain)
constexpr int ROWS = 16;                                // samples per warp tile

// ---- forward blob: [k-step][n-tile 4][lane 32] float4 {b0 hi, b1 hi, b0 lo, b1 lo} --------------------------------------------
constexpr int F_L0 = 0, F_L1 = 12, F_L2 = 16, F_L3E = 20, F_L3H = 32, F_L4 = 36, F_FC = 40, F_KSTEPS = 60;
constexpr int FWD_ITEMS = F_KSTEPS * 4 * 32;            // float4 items
// ---- backward blob: [pair (s, n-tile)][lane] float4; per layer i = 4..0: Fc_i (16 pairs), W_i hidden part (16, i >= 1),
// ---- W_i embedding part (48 = 4 s x 12 n-tiles, i = 3 and 0)
constexpr int B_FC4 = 0, B_H4 = 16, B_FC3 = 32, B_H3 = 48, B_E3 = 64, B_FC2 = 112, B_H2 = 128, B_FC1 = 144, B_H1 = 160, B_FC0 = 176,
              B_E0 = 192, B_PAIRS = 240;
constexpr int BWD_ITEMS = B_PAIRS * 32;
constexpr int MMA_FLOATS = 4 * (FWD_ITEMS + BWD_ITEMS);
static_assert(PACKED_FLOATS % 4 == 0, "the fragment images behind the FFMA blob must stay 16-byte aligned");

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// one k-step (8 reduction indices) against NT n-tiles: acc[n] += x (16x8, A-fragment order, fp32) * B[n] (8x8)
template <int NT>
__device__ __forceinline__ void kstep(float (&acc)[NT][4], float x0, float x1, float x2, float x3, const float4* __restrict__ B, int lane) {
    uint32_t ah[4], al[4];
    split(x0, ah[0], al[0]); split(x1, ah[1], al[1]); split(x2, ah[2], al[2]); split(x3, ah[3], al[3]);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float4 b = __ldg(B + n * 32 + lane);
        const uint32_t b0h = __float_as_uint(b.x), b1h = __float_as_uint(b.y), b0l = __float_as_uint(b.z), b1l = __float_as_uint(b.w);
        mma8(acc[n], al, b0h, b1h);
        mma8(acc[n], ah, b0l, b1l);
        mma8(acc[n], ah, b0h, b1h);
    }
}
// an accumulator block (columns 8s .. 8s+7 of a 16-row tile) as the A operand of k-step s
#define PSL_GM_AFRAG(c) (c)[0], (c)[2], (c)[1], (c)[3]

// ---- packing: reference-layout matrices -> pre-split B fragments ---------------------------------------------------------------
struct PackSrc { const float* W[5]; const float* Wc[5]; };

__device__ __forceinline__ float4 split_pair(float v0, float v1) {
    uint32_t h0, l0, h1, l1;
    split(v0, h0, l0); split(v1, h1, l1);
    return make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
}

__device__ void pack_items(const PackSrc& S, float4* __restrict__ dst, int first, int stride) {
    for (int e = first; e < FWD_ITEMS + BWD_ITEMS; e += stride) {
        const int lane = e & 31, g = lane >> 2, t = lane & 3;
        float v0, v1;
        if (e < FWD_ITEMS) {
            // B(k, n) = W[n][k0 + k]: forward, reduction over the layer INPUT
            const int nt = (e >> 5) & 3, ks = e >> 7;
            const float* src; int ld, koff = 0, valid = 32, s; bool natural = false;
            if (ks < F_L1) { src = S.W[0]; ld = 93; valid = 93; natural = true; s = ks; }
            else if (ks < F_L2) { src = S.W[1]; ld = 32; s = ks - F_L1; }
            else if (ks < F_L3E) { src = S.W[2]; ld = 32; s = ks - F_L2; }
            else if (ks < F_L3H) { src = S.W[3]; ld = 125; valid = 93; natural = true; s = ks - F_L3E; }
            else if (ks < F_L4) { src = S.W[3]; ld = 125; koff = 93; s = ks - F_L3H; }
            else if (ks < F_FC) { src = S.W[4]; ld = 32; s = ks - F_L4; }
            else { src = S.Wc[(ks - F_FC) >> 2]; ld = 32; s = (ks - F_FC) & 3; }
            const int n = 8 * nt + g;
            const int k0 = natural ? 8 * s + t : 8 * s + 2 * t, k1 = natural ? k0 + 4 : k0 + 1;
            v0 = k0 < valid ? src[n * ld + koff + k0] : 0.f;
            v1 = k1 < valid ? src[n * ld + koff + k1] : 0.f;
        } else {
            // B(k, n) = W[k][c0 + n]: backward, reduction over the layer OUTPUT (k-step s <-> output channels 8s+2t, 8s+2t+1)
            const int p = (e - FWD_ITEMS) >> 5;
            const float* src; int ld = 32, coff = 0, valid = 32, q, NT = 4;
            if (p < B_H4) { src = S.Wc[4]; q = p - B_FC4; }
            else if (p < B_FC3) { src = S.W[4]; q = p - B_H4; }
            else if (p < B_H3) { src = S.Wc[3]; q = p - B_FC3; }
            else if (p < B_E3) { src = S.W[3]; ld = 125; coff = 93; q = p - B_H3; }
            else if (p < B_FC2) { src = S.W[3]; ld = 125; valid = 93; NT = 12; q = p - B_E3; }
            else if (p < B_H2) { src = S.Wc[2]; q = p - B_FC2; }
            else if (p < B_FC1) { src = S.W[2]; q = p - B_H2; }
            else if (p < B_H1) { src = S.Wc[1]; q = p - B_FC1; }
            else if (p < B_FC0) { src = S.W[1]; q = p - B_H1; }
            else if (p < B_E0) { src = S.Wc[0]; q = p - B_FC0; }
            else { src = S.W[0]; ld = 93; valid = 93; NT = 12; q = p - B_E0; }
            const int s = q / NT, nt = q - s * NT;
            const int n = 8 * nt + g, j0 = 8 * s + 2 * t;
            v0 = n < valid ? src[j0 * ld + coff + n] : 0.f;
            v1 = n < valid ? src[(j0 + 1) * ld + coff + n] : 0.f;
        }
        dst[e] = split_pair(v0, v1);
    }
}

__global__ void k_geo_mma_pack(PackSrc S, float4* __restrict__ dst) {
    pack_items(S, dst, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---- shared by both directions: sample meta of a 16-row warp tile ---------------------------------------------------------------
struct GeoArgs {
    psl_decode_cfg cfg;
    const float* packed;                 // FFMA blob (biases, embedder matrix, output layer) followed by the fragment images
    const float* pos; long long m;
    const int* I; const float* D; const int* nnum; const double* r2;
    const float* cloud_pos; const float* geo_feats; const float* rand_geo;
    // forward
    float* raw; unsigned char* has_nb; uint32_t* masks_out;
    // backward
    const uint32_t* masks; const float* d_raw;
    float* d_pos; float* d_cg; float* wn_out;
    const float* dwn_extra; const float* dpos_extra;
};

// Fourier embedding argument of channel j for a sample whose position was pre-multiplied by 2 pi (decoder.py:31-34)
__device__ __forceinline__ float emb_arg(const float* __restrict__ gB, int j, float x, float y, float z) {
    return fmaf(z, __ldg(gB + 2 * 96 + j), fmaf(y, __ldg(gB + 96 + j), x * __ldg(gB + j)));
}

template <bool SAVE>
__global__ void __launch_bounds__(WPB * 32, 3) k_geo_fwd_mma(GeoArgs a) {
    __shared__ float sWnAll[WPB][ROWS * 8];
    __shared__ int sIAll[WPB][ROWS * 8];
    __shared__ int sHasAll[WPB][ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long M = a.m;
    const long long m0 = ((long long)blockIdx.x * WPB + warp) * ROWS;
    if (m0 >= M) return;
    float* sWn = sWnAll[warp];
    int* sI = sIAll[warp];
    int* sHas = sHasAll[warp];
    const float* __restrict__ pk = a.packed + OFF_GEO;
    const float4* __restrict__ blob = reinterpret_cast<const float4*>(a.packed + PACKED_FLOATS);

    // ---- sample meta + normalised IDW weights (decoder.py:152-163) -------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int q = lane + 32 * h, s = q >> 3, k = q & 7;
        const long long m = m0 + s;
        int idx = -1;
        float w = 0.f;
        if (m < M) {
            idx = a.I[m * 8 + k];
            const double r2 = a.r2 ? a.r2[m / a.cfg.r2_group] : a.cfg.r2_scalar;
            w = idw_raw(a.D[m * 8 + k], idx, thr_le_of(r2), a.cfg.weighting);
        }
        float sum = fabsf(w);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sWn[q] = __fdiv_rn(w, fmaxf(sum, 1e-12f));
        sI[q] = (w != 0.f) ? idx : -1;
        if (k == 0) {
            const int has = (m < M) && (a.nnum[m] >= a.cfg.min_nn);
            sHas[s] = has;
            if (m < M) a.has_nb[m] = (unsigned char)has;
        }
    }
    __syncwarp();
    const long long mr[2] = {m0 + g, m0 + g + 8};
    const bool ok[2] = {mr[0] < M, mr[1] < M};

    // ---- interpolated geometry feature c_g (decoder.py:164-171) in accumulator-fragment order -----------------------------------
    float cg[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) cg[s][0] = cg[s][1] = cg[s][2] = cg[s][3] = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = g + 8 * r;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = sI[row * 8 + k];
            if (idx >= 0) {
                const float w = sWn[row * 8 + k];
                const float2* f = reinterpret_cast<const float2*>(a.geo_feats + (size_t)idx * 32) + t;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float2 v = __ldg(f + 4 * s);
                    cg[s][2 * r] = fmaf(w, v.x, cg[s][2 * r]);
                    cg[s][2 * r + 1] = fmaf(w, v.y, cg[s][2 * r + 1]);
                }
            }
        }
        if (!sHas[row]) {
#pragma unroll
            for (int s = 0; s < 4; ++s) { cg[s][2 * r] = __ldg(a.rand_geo + 8 * s + 2 * t); cg[s][2 * r + 1] = __ldg(a.rand_geo + 8 * s + 2 * t + 1); }
        }
    }

    // ---- Fourier embedding sin(2 pi p B), 93 channels (+3 zero), A-fragment order of k-step s: channels 8s+t, 8s+t+4 -------------
    float px[2], py[2], pz[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        px[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3]) : 0.f;
        py[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3 + 1]) : 0.f;
        pz[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3 + 2]) : 0.f;
    }
    float e[12][4];
#pragma unroll
    for (int s = 0; s < 12; ++s) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = 8 * s + t + 4 * c;
#pragma unroll
            for (int r = 0; r < 2; ++r)
                e[s][2 * c + r] = j < PSL_GEO_EMB ? sinf(emb_arg(pk + G_B, j, px[r], py[r], pz[r])) : 0.f;
        }
    }

    // ---- trunk ---------------------------------------------------------------------------------------------------------------
    float h[4][4];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float z[4][4], fc[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(pk + G_BIAS + 32 * i + 8 * n) + t);
            const float2 bc = __ldg(reinterpret_cast<const float2*>(pk + G_BIASC + 32 * i + 8 * n) + t);
            z[n][0] = z[n][2] = b.x; z[n][1] = z[n][3] = b.y;
            fc[n][0] = fc[n][2] = bc.x; fc[n][1] = fc[n][3] = bc.y;
        }
        if (i == 0 || i == 3) {
            const float4* B = blob + (i == 0 ? F_L0 : F_L3E) * 128;
#pragma unroll
            for (int s = 0; s < 12; ++s) kstep<4>(z, e[s][0], e[s][1], e[s][2], e[s][3], B + s * 128, lane);
        }
        if (i >= 1) {
            const float4* B = blob + (i == 1 ? F_L1 : i == 2 ? F_L2 : i == 3 ? F_L3H : F_L4) * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep<4>(z, PSL_GM_AFRAG(h[s]), B + s * 128, lane);
        }
        {
            const float4* B = blob + (F_FC + 4 * i) * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep<4>(fc, PSL_GM_AFRAG(cg[s]), B + s * 128, lane);
        }
        uint32_t mk[2] = {0u, 0u};
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                h[n][c] = __fadd_rn(fmaxf(z[n][c], 0.f), fc[n][c]);
                if (SAVE) mk[c >> 1] |= (z[n][c] > 0.f ? 1u : 0u) << (8 * n + 2 * t + (c & 1));
            }
        if (SAVE) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mk[r] |= __shfl_xor_sync(0xffffffffu, mk[r], 1);
                mk[r] |= __shfl_xor_sync(0xffffffffu, mk[r], 2);
                if (t == 0 && ok[r]) a.masks_out[(long long)i * M + mr[r]] = mk[r];
            }
        }
    }
    // ---- occupancy = Wo h4 + bo -------------------------------------------------------------------------------------------------
    float occ[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float2 wo = __ldg(reinterpret_cast<const float2*>(pk + G_WO + 8 * n) + t);
        occ[0] = fmaf(wo.y, h[n][1], fmaf(wo.x, h[n][0], occ[0]));
        occ[1] = fmaf(wo.y, h[n][3], fmaf(wo.x, h[n][2], occ[1]));
    }
    const float bo = __ldg(pk + G_BO);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        occ[r] += __shfl_xor_sync(0xffffffffu, occ[r], 1);
        occ[r] += __shfl_xor_sync(0xffffffffu, occ[r], 2);
        if (t == 0 && ok[r]) {
            if (a.cfg.reserved & 1) a.raw[mr[r] * 4 + 3] = occ[r] + bo;      // occupancy only: rgb belongs to a concurrent colour kernel
            else reinterpret_cast<float4*>(a.raw)[mr[r]] = make_float4(0.f, 0.f, 0.f, occ[r] + bo);
        }
    }
}

// =================================================================================================================================
// backward (data gradients)
// =================================================================================================================================
__global__ void __launch_bounds__(WPB * 32, 3) k_geo_bwd_mma(GeoArgs a) {
    __shared__ float sWnAll[WPB][ROWS * 8], sWrAll[WPB][ROWS * 8], sDWnAll[WPB][ROWS * 8];
    __shared__ int sIAll[WPB][ROWS * 8];
    __shared__ float sPAll[WPB][ROWS * 4], sDPAll[WPB][ROWS * 4], sDenAll[WPB][ROWS];
    __shared__ int sHasAll[WPB][ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long M = a.m;
    const long long m0 = ((long long)blockIdx.x * WPB + warp) * ROWS;
    if (m0 >= M) return;
    float *sWn = sWnAll[warp], *sWr = sWrAll[warp], *sDWn = sDWnAll[warp], *sP = sPAll[warp], *sDP = sDPAll[warp], *sDen = sDenAll[warp];
    int *sI = sIAll[warp], *sHas = sHasAll[warp];
    const float* __restrict__ pk = a.packed + OFF_GEO;
    const float4* __restrict__ blob = reinterpret_cast<const float4*>(a.packed + PACKED_FLOATS) + FWD_ITEMS;
    const bool need_de = a.d_pos != nullptr;          // the embedding gradient only feeds the sample position

#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int q = lane + 32 * h, s = q >> 3, k = q & 7;
        const long long m = m0 + s;
        int idx = -1;
        float w = 0.f;
        if (m < M) {
            idx = a.I[m * 8 + k];
            const double r2 = a.r2 ? a.r2[m / a.cfg.r2_group] : a.cfg.r2_scalar;
            w = idw_raw(a.D[m * 8 + k], idx, thr_le_of(r2), a.cfg.weighting);
        }
        float sum = fabsf(w);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        const float den = fmaxf(sum, 1e-12f);
        const float wn = __fdiv_rn(w, den);
        const int keep = (w != 0.f) ? idx : -1;
        const bool has = (m < M) && (a.nnum[m] >= a.cfg.min_nn);
        sWn[q] = wn; sWr[q] = w; sDWn[q] = (a.dwn_extra && m < M) ? a.dwn_extra[m * 8 + k] : 0.f;
        sI[q] = keep;
        if (k == 0) {
            sHas[s] = has;
            sDen[s] = sum > 1e-12f ? den : 0.f;              // 0 => the clamp is active, no gradient through the norm
        }
        if (a.wn_out && m < M) a.wn_out[m * 8 + k] = (keep >= 0 && has) ? wn : 0.f;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int q = lane + 32 * h, s = q >> 2, c = q & 3;
        const long long m = m0 + s;
        sP[q] = (m < M && c < 3) ? a.pos[m * 3 + c] : 0.f;
        sDP[q] = (a.dpos_extra && m < M && c < 3) ? a.dpos_extra[m * 3 + c] : 0.f;
    }
    __syncwarp();
    const long long mr[2] = {m0 + g, m0 + g + 8};
    const bool ok[2] = {mr[0] < M, mr[1] < M};
    uint32_t mk[5][2];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int r = 0; r < 2; ++r) mk[i][r] = ok[r] ? __ldg(a.masks + (long long)i * M + mr[r]) : 0u;
    const float docc[2] = {ok[0] ? a.d_raw[mr[0] * 4 + 3] : 0.f, ok[1] ? a.d_raw[mr[1] * 4 + 3] : 0.f};

    // dh4 = Wo^T d occ
    float dh[4][4], dcg[4][4], de[12][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float2 wo = __ldg(reinterpret_cast<const float2*>(pk + G_WO + 8 * n) + t);
        dh[n][0] = wo.x * docc[0]; dh[n][1] = wo.y * docc[0]; dh[n][2] = wo.x * docc[1]; dh[n][3] = wo.y * docc[1];
        dcg[n][0] = dcg[n][1] = dcg[n][2] = dcg[n][3] = 0.f;
    }
#pragma unroll
    for (int n = 0; n < 12; ++n) de[n][0] = de[n][1] = de[n][2] = de[n][3] = 0.f;

#pragma unroll
    for (int i = 4; i >= 0; --i) {
        const int pfc = i == 4 ? B_FC4 : i == 3 ? B_FC3 : i == 2 ? B_FC2 : i == 1 ? B_FC1 : B_FC0;
        // (a) d c_g += Fc_i^T dh_i
#pragma unroll
        for (int s = 0; s < 4; ++s) kstep<4>(dcg, PSL_GM_AFRAG(dh[s]), blob + (pfc + 4 * s) * 32, lane);
        // (b) dz_i = dh_i * relu'(z_i)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (!((mk[i][c >> 1] >> (8 * n + 2 * t + (c & 1))) & 1u)) dh[n][c] = 0.f;
        // (c) embedding part of the input (layers 0 and 3)
        if (need_de && (i == 0 || i == 3)) {
            const int pe = i == 3 ? B_E3 : B_E0;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep<12>(de, PSL_GM_AFRAG(dh[s]), blob + (pe + 12 * s) * 32, lane);
        }
        // (d) dh_{i-1} = W_i[:, hidden]^T dz_i
        if (i >= 1) {
            const int ph = i == 4 ? B_H4 : i == 3 ? B_H3 : i == 2 ? B_H2 : B_H1;
            float dn[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) dn[n][0] = dn[n][1] = dn[n][2] = dn[n][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep<4>(dn, PSL_GM_AFRAG(dh[s]), blob + (ph + 4 * s) * 32, lane);
#pragma unroll
            for (int n = 0; n < 4; ++n) { dh[n][0] = dn[n][0]; dh[n][1] = dn[n][1]; dh[n][2] = dn[n][2]; dh[n][3] = dn[n][3]; }
        }
    }

    // ---- embedding gradient: d arg = de cos(arg); d pos += 2 pi B d arg (columns 8n+2t, 8n+2t+1 of rows g, g+8) -------------------
    if (need_de) {
        float gx[2] = {0.f, 0.f}, gy[2] = {0.f, 0.f}, gz[2] = {0.f, 0.f};
        float px[2], py[2], pz[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            px[r] = __fmul_rn(kTwoPi, sP[(g + 8 * r) * 4]); py[r] = __fmul_rn(kTwoPi, sP[(g + 8 * r) * 4 + 1]); pz[r] = __fmul_rn(kTwoPi, sP[(g + 8 * r) * 4 + 2]);
        }
#pragma unroll
        for (int n = 0; n < 12; ++n)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = 8 * n + 2 * t + c;
                if (j < PSL_GEO_EMB) {
                    const float b0 = __ldg(pk + G_B + j), b1 = __ldg(pk + G_B + 96 + j), b2 = __ldg(pk + G_B + 2 * 96 + j);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float da = de[n][2 * r + c] * cosf(fmaf(pz[r], b2, fmaf(py[r], b1, px[r] * b0)));
                        gx[r] = fmaf(da, b0, gx[r]); gy[r] = fmaf(da, b1, gy[r]); gz[r] = fmaf(da, b2, gz[r]);
                    }
                }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            gx[r] += __shfl_xor_sync(0xffffffffu, gx[r], 1); gx[r] += __shfl_xor_sync(0xffffffffu, gx[r], 2);
            gy[r] += __shfl_xor_sync(0xffffffffu, gy[r], 1); gy[r] += __shfl_xor_sync(0xffffffffu, gy[r], 2);
            gz[r] += __shfl_xor_sync(0xffffffffu, gz[r], 1); gz[r] += __shfl_xor_sync(0xffffffffu, gz[r], 2);
            if (t == 0) {
                float* d = sDP + (g + 8 * r) * 4;
                d[0] += kTwoPi * gx[r]; d[1] += kTwoPi * gy[r]; d[2] += kTwoPi * gz[r];
            }
        }
    }

    // ---- d c_g -> output (zero where the sample has no neighbours) and d wn_k += d c_g . geo_feats[I_k] --------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = g + 8 * r;
        const bool has = sHas[row] != 0;
        float2 dc[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) dc[s] = has ? make_float2(dcg[s][2 * r], dcg[s][2 * r + 1]) : make_float2(0.f, 0.f);
        if (a.d_cg && ok[r]) {
#pragma unroll
            for (int s = 0; s < 4; ++s) reinterpret_cast<float2*>(a.d_cg + mr[r] * 32)[4 * s + t] = dc[s];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = sI[row * 8 + k];
            float dot = 0.f;
            if (idx >= 0) {
                const float2* f = reinterpret_cast<const float2*>(a.geo_feats + (size_t)idx * 32) + t;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float2 v = __ldg(f + 4 * s);
                    dot = fmaf(dc[s].y, v.y, fmaf(dc[s].x, v.x, dot));
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            if (t == 0) sDWn[row * 8 + k] += dot;
        }
    }
    __syncwarp();

    // ---- IDW weights -> d_pos (tracker: D is a function of the sample position, decoder.py:143-148) -------------------------------
    if (a.d_pos) {
        if (a.cfg.is_tracker) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int q = lane + 32 * h, s = q >> 3, k = q & 7;
                const float wn = sWn[q], wr = sWr[q], den = sDen[s];
                const float dwn = sDWn[q];
                float dot = dwn * wn;                        // sum_j d wn_j * wn_j over the 8 neighbours
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                float gx = 0.f, gy = 0.f, gz = 0.f;
                const int idx = sI[q];
                if (idx >= 0 && wr != 0.f) {
                    // wn = w / den, den = sum w (w >= 0):  d w = (d wn - dot) / den  (den clamp inactive), else d wn / 1e-12
                    const float dw = den > 0.f ? (dwn - dot) / den : dwn / 1e-12f;
                    float dD;
                    const float Dv = a.D[(m0 + s) * 8 + k];
                    if (a.cfg.weighting == PSL_WEIGHT_EXPO) dD = dw * wr * (-10.0f / sqrtf(Dv));
                    else dD = -dw * wr * wr;
                    // D = sum (c - p)^2  ->  dD/dp = -2 (c - p)
                    const float cx = __ldg(a.cloud_pos + (size_t)idx * 3) - sP[s * 4];
                    const float cy = __ldg(a.cloud_pos + (size_t)idx * 3 + 1) - sP[s * 4 + 1];
                    const float cz = __ldg(a.cloud_pos + (size_t)idx * 3 + 2) - sP[s * 4 + 2];
                    gx = -2.0f * dD * cx; gy = -2.0f * dD * cy; gz = -2.0f * dD * cz;
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    gx += __shfl_xor_sync(0xffffffffu, gx, o);
                    gy += __shfl_xor_sync(0xffffffffu, gy, o);
                    gz += __shfl_xor_sync(0xffffffffu, gz, o);
                }
                if (k == 0) { sDP[s * 4] += gx; sDP[s * 4 + 1] += gy; sDP[s * 4 + 2] += gz; }
            }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = lane + 32 * h;
            if (q < ROWS * 3) {
                const int s = q / 3, c = q - 3 * s;
                if (m0 + s < M) a.d_pos[(m0 + s) * 3 + c] = sDP[s * 4 + c];
            }
        }
    }
}

}  // namespace gm

// ---- host side (called by psl_pack_params / psl_decode_fwd / psl_decode_bwd) ------------------------------------------------------
size_t geo_mma_floats() { return (size_t)gm::MMA_FLOATS; }

int geo_mma_pack(const psl_decoder_params* P, float* packed, cudaStream_t st) {
    gm::PackSrc S;
    for (int i = 0; i < 5; ++i) { S.W[i] = P->g_W[i]; S.Wc[i] = P->g_Wc[i]; }
    gm::k_geo_mma_pack<<<30, 256, 0, st>>>(S, reinterpret_cast<float4*>(packed + PACKED_FLOATS));
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static unsigned geo_blocks(long long m) { return (unsigned)((m + gm::WPB * gm::ROWS - 1) / (gm::WPB * gm::ROWS)); }

int geo_fwd_mma(const psl_decode_cfg* cfg, const float* packed, const float* pos, long long m, const int* I, const float* D,
                const int* nnum, const double* r2, const float* geo_feats, const float* rand_geo, float* raw, unsigned char* has_nb,
                float* save, cudaStream_t st) {
    gm::GeoArgs a{};
    a.cfg = *cfg; a.packed = packed; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.geo_feats = geo_feats; a.rand_geo = rand_geo; a.raw = raw; a.has_nb = has_nb; a.masks_out = reinterpret_cast<uint32_t*>(save);
    TimingScope ts(T_DECODE_FWD, st);
    if (save) gm::k_geo_fwd_mma<true><<<geo_blocks(m), gm::WPB * 32, 0, st>>>(a);
    else gm::k_geo_fwd_mma<false><<<geo_blocks(m), gm::WPB * 32, 0, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int geo_bwd_mma(const psl_decode_cfg* cfg, const float* packed, const float* pos, long long m, const int* I, const float* D,
                const int* nnum, const double* r2, const float* cloud_pos, const float* geo_feats, const float* save,
                const float* d_raw, float* d_pos, float* d_cg, float* wn, const float* dwn_extra, const float* dpos_extra,
                cudaStream_t st) {
    gm::GeoArgs a{};
    a.cfg = *cfg; a.packed = packed; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.geo_feats = geo_feats; a.masks = reinterpret_cast<const uint32_t*>(save); a.d_raw = d_raw;
    a.d_pos = d_pos; a.d_cg = d_cg; a.wn_out = wn; a.dwn_extra = dwn_extra; a.dpos_extra = dpos_extra;
    TimingScope ts(T_DECODE_BWD, st);
    gm::k_geo_bwd_mma<<<geo_blocks(m), gm::WPB * 32, 0, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace psl
