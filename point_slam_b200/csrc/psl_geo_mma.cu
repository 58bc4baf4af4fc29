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
//   * weights: pre-split hi/lo B fragments, one 16-byte shared-memory load per (k-step, n-tile, lane); the 120 KB image of a
//     direction is brought in by one bulk copy per block that flies while the warps gather (the first version read it through L1:
//     38 % of the stall samples sat on those loads, profiles/r02_summary.md); the blocks are persistent over the 16-sample tiles,
//     so the image is loaded once per SM however large the launch; one __syncthreads at block start, none afterwards;
//   * rolled layer loop (a fully unrolled tile is 12.5 k instructions executed once: instruction-fetch bound), the two embedding
//     products (layers 0 and 3) accumulated in ONE pass over the Fourier channels, so sin() is evaluated once and never stored;
//   * kept for the backward: one 32-bit ReLU mask per sample and layer (20 B / sample instead of 1.4 KB).
#include "psl_decode.cuh"
#include "psl_tc.cuh"

namespace psl {
namespace gm {

constexpr int WPB = 12;                                 // warps per block: independent 16-sample tiles sharing one weight image in smem
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

// ---- shared by both directions ----------------------------------------------------------------------------------------------------
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

// Block = WPB independent warps sharing one copy of the direction's fragment image in shared memory (one bulk copy per block,
// in flight while the warps gather).  Per-warp scratch follows the image.
constexpr int BLOB_BYTES = FWD_ITEMS * 16;               // both directions: 122 880 B
static_assert(FWD_ITEMS == BWD_ITEMS, "one shared-memory plan for both directions");
constexpr int S_GB = BLOB_BYTES;                         // [3][96] embedder matrix
constexpr int S_BAR = S_GB + 3 * 96 * 4;                 // mbarrier
constexpr int S_WARP = S_BAR + 16;                       // per-warp scratch from here
constexpr int FW_WN = 0, FW_I = 128, FW_HAS = 256, FW_WORDS = 272;                                     // forward: 1088 B per warp
constexpr int BW_WN = 0, BW_WR = 128, BW_DWN = 256, BW_I = 384, BW_P = 512, BW_DP = 576, BW_DEN = 640, BW_HAS = 656,
              BW_MASK = 672, BW_WORDS = 752;                                                           // backward: 3008 B per warp
constexpr int SM_GEO_FWD = S_WARP + WPB * FW_WORDS * 4;
constexpr int SM_GEO_BWD = S_WARP + WPB * BW_WORDS * 4;
static_assert(SM_GEO_BWD <= 227 * 1024, "shared memory over budget");

// block prologue: thread 0 starts the bulk copy of the fragment image; everybody stages the embedder matrix
__device__ __forceinline__ void start_image(unsigned char* smem, const float4* __restrict__ image, const float* __restrict__ gB) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S_BAR);
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
        tc::mbar_expect_tx(bar, BLOB_BYTES);
        constexpr int CH = BLOB_BYTES / 8;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) tc::bulk_g2s(smem + c * CH, reinterpret_cast<const unsigned char*>(image) + c * CH, CH, bar);
    }
    float* sGB = reinterpret_cast<float*>(smem + S_GB);
    for (int i = threadIdx.x; i < 3 * 96; i += blockDim.x) sGB[i] = __ldg(gB + i);
    __syncthreads();                                     // the only block-wide barrier: barrier init + sGB visible
}
__device__ __forceinline__ void wait_image(unsigned char* smem) { tc::mbar_wait_wd(reinterpret_cast<uint64_t*>(smem + S_BAR), 0); }

// Fourier embedding argument of channel j for a sample whose position was pre-multiplied by 2 pi (decoder.py:31-34)
__device__ __forceinline__ float emb_arg(const float* sGB, int j, float x, float y, float z) {
    return fmaf(z, sGB[2 * 96 + j], fmaf(y, sGB[96 + j], x * sGB[j]));
}
// one k-step (8 reduction indices) against NT n-tiles of the shared-memory image: acc[n] += x (16x8, A-fragment order, fp32) * B[n]
template <int NT>
__device__ __forceinline__ void kstep_s(float (&acc)[NT][4], float x0, float x1, float x2, float x3, const float4* B, int lane) {
    uint32_t ah[4], al[4];
    split(x0, ah[0], al[0]); split(x1, ah[1], al[1]); split(x2, ah[2], al[2]); split(x3, ah[3], al[3]);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float4 b = B[n * 32 + lane];
        const uint32_t b0h = __float_as_uint(b.x), b1h = __float_as_uint(b.y), b0l = __float_as_uint(b.z), b1l = __float_as_uint(b.w);
        mma8(acc[n], al, b0h, b1h);
        mma8(acc[n], ah, b0l, b1l);
        mma8(acc[n], ah, b0h, b1h);
    }
}

template <bool SAVE>
__global__ void __launch_bounds__(WPB * 32, 1) k_geo_fwd_mma(GeoArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* __restrict__ pk = a.packed + OFF_GEO;
    start_image(smem, reinterpret_cast<const float4*>(a.packed + PACKED_FLOATS), pk + G_B);
    const long long M = a.m;
    const float4* sB = reinterpret_cast<const float4*>(smem);
    const float* sGB = reinterpret_cast<const float*>(smem + S_GB);
    float* sWn = reinterpret_cast<float*>(smem + S_WARP) + warp * FW_WORDS + FW_WN;
    int* sI = reinterpret_cast<int*>(smem + S_WARP) + warp * FW_WORDS + FW_I;
    int* sHas = reinterpret_cast<int*>(smem + S_WARP) + warp * FW_WORDS + FW_HAS;
    // persistent over the 16-sample tiles: the block's copy of the weight image is loaded once, however large the launch
    const long long tstep = (long long)gridDim.x * (blockDim.x >> 5);
#pragma unroll 1
    for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + warp; tile * ROWS < M; tile += tstep) {
    const long long m0 = tile * ROWS;
    __syncwarp();                                        // the previous tile's readers of the per-warp scratch are done

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
    float px[2], py[2], pz[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        px[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3]) : 0.f;
        py[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3 + 1]) : 0.f;
        pz[r] = ok[r] ? __fmul_rn(kTwoPi, a.pos[mr[r] * 3 + 2]) : 0.f;
    }

    // ---- embedding products of layers 0 and 3 in ONE pass over the 93 (+3 zero) Fourier channels: sin(2 pi p B) is evaluated once,
    // ---- in the A-fragment order of k-step s (channels 8s+t, 8s+t+4), and never stored ---------------------------------------------
    float z0[4][4], z3[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float2 b0 = __ldg(reinterpret_cast<const float2*>(pk + G_BIAS + 8 * n) + t);
        const float2 b3 = __ldg(reinterpret_cast<const float2*>(pk + G_BIAS + 32 * 3 + 8 * n) + t);
        z0[n][0] = z0[n][2] = b0.x; z0[n][1] = z0[n][3] = b0.y;
        z3[n][0] = z3[n][2] = b3.x; z3[n][1] = z3[n][3] = b3.y;
    }
    wait_image(smem);
#pragma unroll 1
    for (int s = 0; s < 12; ++s) {
        float e[4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = 8 * s + t + 4 * c;
#pragma unroll
            for (int r = 0; r < 2; ++r) e[2 * c + r] = j < PSL_GEO_EMB ? sinf(emb_arg(sGB, j, px[r], py[r], pz[r])) : 0.f;
        }
        uint32_t ah[4], al[4];
        split(e[0], ah[0], al[0]); split(e[1], ah[1], al[1]); split(e[2], ah[2], al[2]); split(e[3], ah[3], al[3]);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float4 b = sB[(F_L0 + s) * 128 + n * 32 + lane];
            const float4 d = sB[(F_L3E + s) * 128 + n * 32 + lane];
            mma8(z0[n], al, __float_as_uint(b.x), __float_as_uint(b.y));
            mma8(z3[n], al, __float_as_uint(d.x), __float_as_uint(d.y));
            mma8(z0[n], ah, __float_as_uint(b.z), __float_as_uint(b.w));
            mma8(z3[n], ah, __float_as_uint(d.z), __float_as_uint(d.w));
            mma8(z0[n], ah, __float_as_uint(b.x), __float_as_uint(b.y));
            mma8(z3[n], ah, __float_as_uint(d.x), __float_as_uint(d.y));
        }
    }

    // ---- trunk (one rolled loop: the code of a layer fits the instruction cache) -----------------------------------------------------
    float h[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) h[n][0] = h[n][1] = h[n][2] = h[n][3] = 0.f;
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
        float z[4][4], fc[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float2 bc = __ldg(reinterpret_cast<const float2*>(pk + G_BIASC + 32 * i + 8 * n) + t);
            fc[n][0] = fc[n][2] = bc.x; fc[n][1] = fc[n][3] = bc.y;
            if (i == 0) { z[n][0] = z0[n][0]; z[n][1] = z0[n][1]; z[n][2] = z0[n][2]; z[n][3] = z0[n][3]; }
            else if (i == 3) { z[n][0] = z3[n][0]; z[n][1] = z3[n][1]; z[n][2] = z3[n][2]; z[n][3] = z3[n][3]; }
            else {
                const float2 b = __ldg(reinterpret_cast<const float2*>(pk + G_BIAS + 32 * i + 8 * n) + t);
                z[n][0] = z[n][2] = b.x; z[n][1] = z[n][3] = b.y;
            }
        }
        if (i >= 1) {
            const float4* B = sB + (i == 1 ? F_L1 : i == 2 ? F_L2 : i == 3 ? F_L3H : F_L4) * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep_s<4>(z, PSL_GM_AFRAG(h[s]), B + s * 128, lane);
        }
        {
            const float4* B = sB + (F_FC + 4 * i) * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep_s<4>(fc, PSL_GM_AFRAG(cg[s]), B + s * 128, lane);
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
    }   // tile loop
}

// =================================================================================================================================
// backward (data gradients)
// =================================================================================================================================
__global__ void __launch_bounds__(WPB * 32, 1) k_geo_bwd_mma(GeoArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* __restrict__ pk = a.packed + OFF_GEO;
    start_image(smem, reinterpret_cast<const float4*>(a.packed + PACKED_FLOATS) + FWD_ITEMS, pk + G_B);
    const long long M = a.m;
    const float4* sB = reinterpret_cast<const float4*>(smem);
    const float* sGB = reinterpret_cast<const float*>(smem + S_GB);
    float* sw = reinterpret_cast<float*>(smem + S_WARP) + warp * BW_WORDS;
    float *sWn = sw + BW_WN, *sWr = sw + BW_WR, *sDWn = sw + BW_DWN, *sP = sw + BW_P, *sDP = sw + BW_DP, *sDen = sw + BW_DEN;
    int *sI = reinterpret_cast<int*>(sw) + BW_I, *sHas = reinterpret_cast<int*>(sw) + BW_HAS;
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sw) + BW_MASK;            // [5][16]
    const bool need_de = a.d_pos != nullptr;          // the embedding gradient only feeds the sample position
    const long long tstep = (long long)gridDim.x * (blockDim.x >> 5);
#pragma unroll 1
    for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + warp; tile * ROWS < M; tile += tstep) {
    const long long m0 = tile * ROWS;
    __syncwarp();

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
#pragma unroll
    for (int h = 0; h < 3; ++h) {                          // ReLU masks of the 5 layers x 16 rows
        const int q = lane + 32 * h;
        if (q < 80) { const long long m = m0 + (q & 15); sMask[q] = m < M ? __ldg(a.masks + (long long)(q >> 4) * M + m) : 0u; }
    }
    __syncwarp();
    const long long mr[2] = {m0 + g, m0 + g + 8};
    const bool ok[2] = {mr[0] < M, mr[1] < M};
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

    wait_image(smem);
#pragma unroll 1
    for (int i = 4; i >= 0; --i) {
        const int pfc = i == 4 ? B_FC4 : i == 3 ? B_FC3 : i == 2 ? B_FC2 : i == 1 ? B_FC1 : B_FC0;
        // (a) d c_g += Fc_i^T dh_i
#pragma unroll
        for (int s = 0; s < 4; ++s) kstep_s<4>(dcg, PSL_GM_AFRAG(dh[s]), sB + (pfc + 4 * s) * 32, lane);
        // (b) dz_i = dh_i * relu'(z_i)
        const uint32_t mk[2] = {sMask[i * 16 + g], sMask[i * 16 + g + 8]};
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (!((mk[c >> 1] >> (8 * n + 2 * t + (c & 1))) & 1u)) dh[n][c] = 0.f;
        // (c) embedding part of the input (layers 0 and 3)
        if (need_de && (i == 0 || i == 3)) {
            const int pe = i == 3 ? B_E3 : B_E0;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep_s<12>(de, PSL_GM_AFRAG(dh[s]), sB + (pe + 12 * s) * 32, lane);
        }
        // (d) dh_{i-1} = W_i[:, hidden]^T dz_i
        if (i >= 1) {
            const int ph = i == 4 ? B_H4 : i == 3 ? B_H3 : i == 2 ? B_H2 : B_H1;
            float dn[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) dn[n][0] = dn[n][1] = dn[n][2] = dn[n][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) kstep_s<4>(dn, PSL_GM_AFRAG(dh[s]), sB + (ph + 4 * s) * 32, lane);
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
                    const float b0 = sGB[j], b1 = sGB[96 + j], b2 = sGB[2 * 96 + j];
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
    }   // tile loop
}

// ---- IDW-weight chain rule alone: d_pos += d_pos_add + (d wn -> d D -> d pos), for the gradient the COLOUR branch put on the
// normalised weights (tracker, decoder.py:143-163).  The chain is linear in d wn, so the geometry backward applies it to its own
// d wn while the colour backward runs, and this kernel adds the colour branch's share afterwards (8 lanes per sample).
__global__ void k_idw_chain(GeoArgs a) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long m = q >> 3;
    const int k = (int)(q & 7);
    const bool ok = m < a.m;
    int idx = -1;
    float w = 0.f, Dv = 0.f;
    if (ok) {
        idx = a.I[m * 8 + k];
        Dv = a.D[m * 8 + k];
        const double r2 = a.r2 ? a.r2[m / a.cfg.r2_group] : a.cfg.r2_scalar;
        w = idw_raw(Dv, idx, thr_le_of(r2), a.cfg.weighting);
    }
    float sum = fabsf(w);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float den0 = fmaxf(sum, 1e-12f);
    const float wn = __fdiv_rn(w, den0);
    const float den = sum > 1e-12f ? den0 : 0.f;
    const float dwn = ok ? a.dwn_extra[m * 8 + k] : 0.f;
    float dot = dwn * wn;
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (a.cfg.is_tracker && ok && idx >= 0 && w != 0.f) {      // mapping: D does not depend on the sample position in the graph
        const float dw = den > 0.f ? (dwn - dot) / den : dwn / 1e-12f;
        float dD;
        if (a.cfg.weighting == PSL_WEIGHT_EXPO) dD = dw * w * (-10.0f / sqrtf(Dv));
        else dD = -dw * w * w;
        const float cx = __ldg(a.cloud_pos + (size_t)idx * 3) - a.pos[m * 3];
        const float cy = __ldg(a.cloud_pos + (size_t)idx * 3 + 1) - a.pos[m * 3 + 1];
        const float cz = __ldg(a.cloud_pos + (size_t)idx * 3 + 2) - a.pos[m * 3 + 2];
        gx = -2.0f * dD * cx; gy = -2.0f * dD * cy; gz = -2.0f * dD * cz;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gz += __shfl_xor_sync(0xffffffffu, gz, o);
    }
    if (ok && k < 3) {
        const float g = k == 0 ? gx : (k == 1 ? gy : gz);
        a.d_pos[m * 3 + k] += g + (a.dpos_extra ? a.dpos_extra[m * 3 + k] : 0.f);
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

// launch shape: 16-sample warp tiles spread over all SMs first (the work of a launch is less than one wave: a block's latency is
// what is timed, and the legacy tensor path is shared by the warps of an SM), at most WPB warps per block
static void geo_shape(long long m, unsigned* blocks, unsigned* threads) {
    const long long tiles = (m + gm::ROWS - 1) / gm::ROWS;
    long long nw = (tiles + sm_count() - 1) / sm_count();
    nw = nw < 1 ? 1 : (nw > gm::WPB ? gm::WPB : nw);
    long long nb = (tiles + nw - 1) / nw;
    if (nb > sm_count()) nb = sm_count();                 // one block per SM (the image takes most of its shared memory): persistent
    *blocks = (unsigned)nb;
    *threads = (unsigned)(32 * nw);
}

int geo_fwd_mma(const psl_decode_cfg* cfg, const float* packed, const float* pos, long long m, const int* I, const float* D,
                const int* nnum, const double* r2, const float* geo_feats, const float* rand_geo, float* raw, unsigned char* has_nb,
                float* save, cudaStream_t st) {
    gm::GeoArgs a{};
    a.cfg = *cfg; a.packed = packed; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.geo_feats = geo_feats; a.rand_geo = rand_geo; a.raw = raw; a.has_nb = has_nb; a.masks_out = reinterpret_cast<uint32_t*>(save);
    PSL_CHECK_CUDA(cudaFuncSetAttribute(gm::k_geo_fwd_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gm::SM_GEO_FWD));
    PSL_CHECK_CUDA(cudaFuncSetAttribute(gm::k_geo_fwd_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gm::SM_GEO_FWD));
    unsigned nb, nt;
    geo_shape(m, &nb, &nt);
    TimingScope ts(T_DECODE_FWD, st);
    if (save) gm::k_geo_fwd_mma<true><<<nb, nt, gm::SM_GEO_FWD, st>>>(a);
    else gm::k_geo_fwd_mma<false><<<nb, nt, gm::SM_GEO_FWD, st>>>(a);
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
    PSL_CHECK_CUDA(cudaFuncSetAttribute(gm::k_geo_bwd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, gm::SM_GEO_BWD));
    unsigned nb, nt;
    geo_shape(m, &nb, &nt);
    TimingScope ts(T_DECODE_BWD, st);
    gm::k_geo_bwd_mma<<<nb, nt, gm::SM_GEO_BWD, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int geo_idw_chain(const psl_decode_cfg* cfg, const float* pos, long long m, const int* I, const float* D, const double* r2,
                  const float* cloud_pos, const float* dwn, const float* dpos_add, float* d_pos, cudaStream_t st) {
    gm::GeoArgs a{};
    a.cfg = *cfg; a.pos = pos; a.m = m; a.I = I; a.D = D; a.r2 = r2; a.cloud_pos = cloud_pos; a.dwn_extra = dwn; a.dpos_extra = dpos_add;
    a.d_pos = d_pos;
    TimingScope ts(T_DECODE_BWD, st);
    gm::k_idw_chain<<<(unsigned)((m * 8 + 255) / 256), 256, 0, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace psl
