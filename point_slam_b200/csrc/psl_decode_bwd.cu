// K2+K3 backward (fp32 FFMA path).  One CTA walks 64-sample tiles; per tile the gradient is pushed top-down through
// the colour trunk, the per-neighbour colour MLP, the geometry trunk and the IDW weights.  Data gradients are
// per-warp (8 samples per warp, lane <-> channel); weight gradients are CTA-cooperative outer products over the
// 64 samples of the tile, accumulated into a per-CTA partial buffer (no atomics) that a second kernel reduces in a
// fixed order -> bit-deterministic (the reference's only test is run-to-run bit identity, test_deterministic.py).
//
// Autograd semantics reproduced: src/conv_onet/models/decoder.py:130-222, :341-449.
#include "psl_decode.cuh"

namespace psl {

// ---- flat gradient blob (reference layouts, row-major (out,in)) ------------------------------------------------
constexpr int GR_gB = 0;                        // (3,93)
constexpr int GR_gW0 = GR_gB + 279;             // (32,93)
constexpr int GR_gW1 = GR_gW0 + 32 * 93;
constexpr int GR_gW2 = GR_gW1 + 1024;
constexpr int GR_gW3 = GR_gW2 + 1024;           // (32,125)
constexpr int GR_gW4 = GR_gW3 + 32 * 125;
constexpr int GR_gb = GR_gW4 + 1024;            // 5 x 32
constexpr int GR_gWc = GR_gb + 160;             // 5 x (32,32)
constexpr int GR_gbc = GR_gWc + 5120;           // 5 x 32
constexpr int GR_gWo = GR_gbc + 160;            // 32
constexpr int GR_gbo = GR_gWo + 32;             // 1
constexpr int GR_cBrel = GR_gbo + 1;            // (3,10)
constexpr int GR_cN1 = GR_cBrel + 30;           // (128,52)
constexpr int GR_cn1b = GR_cN1 + 128 * 52;
constexpr int GR_cN2 = GR_cn1b + 128;           // (32,128)
constexpr int GR_cn2b = GR_cN2 + 4096;
__host__ __device__ constexpr int GR_cW(int i) {
    int o = GR_cn2b + 32;
    for (int j = 0; j < i; ++j) o += 128 * col_k(j);
    return o;
}
constexpr int GR_cb = GR_cW(5);                 // 5 x 128
constexpr int GR_cWc = GR_cb + 640;             // 5 x (128,32)
constexpr int GR_cbc = GR_cWc + 5 * 4096;       // 5 x 128
constexpr int GR_cWo = GR_cbc + 640;            // (3,128)
constexpr int GR_cbo = GR_cWo + 384;            // 3
constexpr int GR_aff = GR_cbo + 3;              // 12
constexpr int GR_TOTAL = ((GR_aff + 12 + 3) / 4) * 4;

__host__ __device__ constexpr int GR_gW(int i) { return i == 0 ? GR_gW0 : i == 1 ? GR_gW1 : i == 2 ? GR_gW2 : i == 3 ? GR_gW3 : GR_gW4; }
__host__ __device__ constexpr int geo_k(int i) { return i == 0 ? 93 : (i == 3 ? 125 : 32); }

struct BwdArgs {
    DecodeArgs f;                       // forward inputs (raw = forward output, save = activations)
    psl_decoder_params P;               // original-layout parameter pointers (device)
    const float* d_raw;
    float* d_pos; float* d_cg; float* wn_out; float* d_colpair; float* d_cc;
    float* partial;                     // [gridDim.x][GR_TOTAL]
    const float* dwn_extra;             // (m,8)  dL/d(normalised weights) from another kernel (tensor-core colour branch)
    const float* dpos_extra;            // (m,3)  dL/dpos from another kernel
    int want_geo_params, want_col_params;
};

// ---- shared memory plan (floats) --------------------------------------------------------------------------------
constexpr int B_DH = 0;                          // [128][LD]
constexpr int B_DH2 = B_DH + 128 * LD;           // [128][LD]
constexpr int B_U = B_DH2 + 128 * LD;            // union: staged weights (<= 128*168) / layer inputs / nbr scratch
constexpr int B_U_FLOATS = 128 * 168;
constexpr int B_C = B_U + B_U_FLOATS;            // [32][LD] interpolated feature c (colour, then geometry)
constexpr int B_DE = B_C + 32 * LD;              // [96][LD] gradient w.r.t. the Fourier embedding (colour 40 / geo 93)
constexpr int B_DC = B_DE + 96 * LD;             // [32][LD] gradient w.r.t. c
constexpr int B_WN = B_DC + 32 * LD;             // per warp 64: normalised weights
constexpr int B_WR = B_WN + NWARP * 64;          // per warp 64: raw weights
constexpr int B_DWN = B_WR + NWARP * 64;         // per warp 64: gradient w.r.t. normalised weights
constexpr int B_I = B_DWN + NWARP * 64;          // per warp 64 int
constexpr int B_P = B_I + NWARP * 64;            // per warp 32: pos (8x4)
constexpr int B_DP = B_P + NWARP * 32;           // per warp 32: d_pos accumulators (8x4)
constexpr int B_DRAW = B_DP + NWARP * 32;        // per warp 32: d_raw (8x4)
constexpr int B_DOUT = B_DRAW + NWARP * 32;      // [4][LD] gradient at the 3 colour outputs (block-wide)
constexpr int B_OUTV = B_DOUT + 4 * LD;          // [4][LD] pre-affine colour outputs
constexpr int B_MISC = B_OUTV + 4 * LD;          // per warp 16: has flags (8) + denominators (8)
constexpr int B_FLOATS = B_MISC + NWARP * 16;
constexpr size_t SM_BWD_BYTES = sizeof(float) * B_FLOATS;
static_assert(SM_BWD_BYTES <= 227 * 1024, "backward shared memory over budget");

// acc[j][s] += sum_n W[n][k0 + lane + 32 j] * g[n][s]   (W row-major (N, ldw) in shared memory; k < K guarded)
template <int NJ>
__device__ __forceinline__ void dense8T(float (&acc)[NJ][8], const float* __restrict__ g, int N,
                                        const float* __restrict__ W, int ldw, int k0, int K, int lane) {
    bool ok[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) ok[j] = (lane + 32 * j) < K;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float4 a0 = *reinterpret_cast<const float4*>(g + n * LD);
        const float4 a1 = *reinterpret_cast<const float4*>(g + n * LD + 4);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float w = ok[j] ? W[n * ldw + k0 + lane + 32 * j] : 0.f;
            acc[j][0] = fmaf(w, a0.x, acc[j][0]); acc[j][1] = fmaf(w, a0.y, acc[j][1]);
            acc[j][2] = fmaf(w, a0.z, acc[j][2]); acc[j][3] = fmaf(w, a0.w, acc[j][3]);
            acc[j][4] = fmaf(w, a1.x, acc[j][4]); acc[j][5] = fmaf(w, a1.y, acc[j][5]);
            acc[j][6] = fmaf(w, a1.z, acc[j][6]); acc[j][7] = fmaf(w, a1.w, acc[j][7]);
        }
    }
}

// CTA-cooperative outer product over the 64 samples of the tile:
//   out[a * ldo + col0 + b] += sum_s A[a][s] * B[b][s],  a < 16*RA (a = ta + 16 i), b < NB <= 16*RB (b = tb + 16 j)
template <int RA, int RB>
__device__ __forceinline__ void outer_acc(const float* __restrict__ A, const float* __restrict__ Bm, int NB,
                                          float* __restrict__ out, int ldo, int col0o) {
    const int tb = threadIdx.x & 15, ta = threadIdx.x >> 4;
    float acc[RA][RB];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int s4 = 0; s4 < TS / 4; ++s4) {
        float4 av[RA], bv[RB];
#pragma unroll
        for (int i = 0; i < RA; ++i) av[i] = *reinterpret_cast<const float4*>(A + (ta + 16 * i) * LD + 4 * s4);
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int b = tb + 16 * j;
            bv[j] = b < NB ? *reinterpret_cast<const float4*>(Bm + b * LD + 4 * s4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < RA; ++i)
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]); acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]); acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int b = tb + 16 * j;
            if (b < NB) out[(ta + 16 * i) * ldo + col0o + b] += acc[i][j];
        }
}

// out[a] += sum_s A[a][s]   (bias gradients)
__device__ __forceinline__ void rowsum_acc(const float* __restrict__ A, int NA, float* __restrict__ out) {
    for (int a = threadIdx.x; a < NA; a += blockDim.x) {
        float s = 0.f;
#pragma unroll 4
        for (int c = 0; c < TS; ++c) s += A[a * LD + c];
        out[a] += s;
    }
}

__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int nfloats) {
    // plain copy global -> shared (float granularity: reference-layout matrices are not always 16 B multiples)
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = __ldg(src + i);
}

constexpr int FCT_LD = 132;                      // padded row stride of the transposed fc_c matrix in shared memory
constexpr int U_FCT = (128 + 40) * LD;           // offsets inside the union buffer sU
constexpr int U_BC = U_FCT + 32 * FCT_LD;
static_assert(U_BC + 128 <= B_U_FLOATS, "fc_c staging does not fit");

// stage Fc (128,32) transposed as [32][FCT_LD] and its bias (all threads of the CTA)
__device__ __forceinline__ void stage_fct(float* sU, const float* __restrict__ Fc, const float* __restrict__ bc) {
    for (int e = threadIdx.x; e < 128 * 32; e += blockDim.x) sU[U_FCT + (e & 31) * FCT_LD + (e >> 5)] = __ldg(Fc + e);
    for (int e = threadIdx.x; e < 128; e += blockDim.x) sU[U_BC + e] = __ldg(bc + e);
}

// h_i = softplus(z_i) + Fc_i c + bc_i for the 8 samples of this warp (the forward keeps only z): h[j][s], channel lane+32j
__device__ __forceinline__ void recompute_h(float (&h)[4][8], const float* sU, const float* sC_cols, const float* __restrict__ zsave,
                                            long long M, long long m0, int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int s = 0; s < 8; ++s) h[j][s] = sU[U_BC + lane + 32 * j];
    dense8<4>(h, sC_cols, 32, sU + U_FCT, lane, FCT_LD);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float z = (m0 + s < M) ? zsave[(m0 + s) * 128 + lane + 32 * j] : 0.f;
            h[j][s] = (m0 + s < M) ? __fadd_rn(softplus100(z), h[j][s]) : 0.f;
        }
}

__global__ void __launch_bounds__(NWARP * 32, 1) k_decode_bwd(BwdArgs a, long long n_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* sDH = smem + B_DH;
    float* sDH2 = smem + B_DH2;
    float* sU = smem + B_U;
    float* sC = smem + B_C;
    float* sDE = smem + B_DE;
    float* sDC = smem + B_DC;
    float* sDOut = smem + B_DOUT;
    float* sOutV = smem + B_OUTV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sWn = smem + B_WN + warp * 64;
    float* sWr = smem + B_WR + warp * 64;
    float* sDWn = smem + B_DWN + warp * 64;
    int* sI = reinterpret_cast<int*>(smem + B_I) + warp * 64;
    float* sP = smem + B_P + warp * 32;
    float* sDP = smem + B_DP + warp * 32;
    float* sDRaw = smem + B_DRAW + warp * 32;
    int* sHas = reinterpret_cast<int*>(smem + B_MISC) + warp * 16;
    float* sDen = smem + B_MISC + warp * 16 + 8;
    const int col0 = warp * SPW;
    const DecodeArgs& f = a.f;
    const long long M = f.m;
    const bool color = f.cfg.stage == PSL_STAGE_COLOR;
    const bool rel = color && f.cfg.encode_rel_pos;
    const SaveLayout SL = save_layout(color, f.cfg.encode_rel_pos);
    float* part = a.partial + (size_t)blockIdx.x * GR_TOTAL;
    const bool wg = a.want_geo_params, wc = a.want_col_params;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long m0 = tile * TS + col0;
        // ================= B0: meta, weights, incoming gradient ============================================
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = lane + 32 * h, s = q >> 3, k = q & 7;
            const long long m = m0 + s;
            int idx = -1;
            float w = 0.f;
            if (m < M) {
                idx = f.I[m * 8 + k];
                const double r2 = f.r2 ? f.r2[m / f.cfg.r2_group] : f.cfg.r2_scalar;
                w = idw_raw(f.D[m * 8 + k], idx, thr_le_of(r2), f.cfg.weighting);
            }
            float sum = fabsf(w);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            const float den = fmaxf(sum, 1e-12f);
            const float wn = __fdiv_rn(w, den);
            sWn[q] = wn; sWr[q] = w; sDWn[q] = (a.dwn_extra && m < M) ? a.dwn_extra[m * 8 + k] : 0.f;
            sI[q] = (w != 0.f) ? idx : -1;
            if (k == 0) {
                sHas[s] = (m < M) && (f.nnum[m] >= f.cfg.min_nn);
                sDen[s] = sum > 1e-12f ? den : 0.f;          // 0 => the clamp is active, no gradient through the norm
            }
            if (a.wn_out && m < M) a.wn_out[m * 8 + k] = (sI[q] >= 0 && (f.nnum[m] >= f.cfg.min_nn)) ? wn : 0.f;
        }
        {
            const int s = lane >> 2, c = lane & 3;
            const long long m = m0 + s;
            sP[lane] = (m < M && c < 3) ? f.pos[m * 3 + c] : 0.f;
            sDP[lane] = (a.dpos_extra && m < M && c < 3) ? a.dpos_extra[m * 3 + c] : 0.f;
            sDRaw[lane] = (m < M) ? a.d_raw[m * 4 + c] : 0.f;
        }
        __syncwarp();

        // ================= colour branch ======================================================================
        if (color) {
            // ---- B1: output layer ------------------------------------------------------------------------------
            // h4 -> sU (as [128][LD]) for dWo; out = Wo h4 + bo (needed only for the affine mode)
            // c (interpolated colour feature): needed to rebuild h_i = softplus(z_i) + Fc_i c + bc_i and for the fc_c gradients
#pragma unroll
            for (int s = 0; s < 8; ++s)
                sC[lane * LD + col0 + s] = (m0 + s < M) ? f.save[SL.cc * M + (m0 + s) * 32 + lane] : 0.f;
            __syncthreads();
            stage_fct(sU, a.P.c_Wc[4], a.P.c_bc[4]);
            __syncthreads();
            float h4[4][8];
            recompute_h(h4, sU, sC + col0, f.save + SL.cz * M + 4ll * M * 128, M, m0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int s = 0; s < 8; ++s) sU[(lane + 32 * j) * LD + col0 + s] = h4[j][s];
            float wo[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) wo[c][j] = __ldg(a.P.c_Wo + c * 128 + lane + 32 * j);
            if (f.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        float p = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) p = fmaf(wo[c][j], h4[j][s], p);
                        p = warp_sum(p) + __ldg(a.P.c_bo + c);
                        if (lane == s) sOutV[c * LD + col0 + s] = p;
                    }
            }
            __syncwarp();
            if (lane < SPW) {
                const int s = lane;
                const long long m = m0 + s;
                float g0 = sDRaw[s * 4], g1 = sDRaw[s * 4 + 1], g2 = sDRaw[s * 4 + 2];
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
                if (m < M && f.cfg.rgb_mode != PSL_RGB_RAW) {
                    const float4 rv = reinterpret_cast<const float4*>(f.raw)[m];
                    g0 *= rv.x * (1.0f - rv.x); g1 *= rv.y * (1.0f - rv.y); g2 *= rv.z * (1.0f - rv.z);
                }
                if (f.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                    const float* A = f.affine;
                    o0 = A[0] * g0 + A[1] * g1 + A[2] * g2;
                    o1 = A[3] * g0 + A[4] * g1 + A[5] * g2;
                    o2 = A[6] * g0 + A[7] * g1 + A[8] * g2;
                    // post-affine gradients parked in sDRaw for the rot/trans reduction below
                    sDRaw[s * 4] = g0; sDRaw[s * 4 + 1] = g1; sDRaw[s * 4 + 2] = g2;
                } else {
                    o0 = g0; o1 = g1; o2 = g2;
                }
                sDOut[0 * LD + col0 + s] = o0; sDOut[1 * LD + col0 + s] = o1; sDOut[2 * LD + col0 + s] = o2;
            }
            __syncwarp();
            {   // dh4 = Wo^T dout
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const float v = wo[0][j] * sDOut[0 * LD + col0 + s] + wo[1][j] * sDOut[1 * LD + col0 + s] +
                                        wo[2][j] * sDOut[2 * LD + col0 + s];
                        sDH[(lane + 32 * j) * LD + col0 + s] = v;
                    }
            }
            __syncthreads();
            if (wc) {   // dWo (3,128), dbo (3), affine (12)
                if (threadIdx.x < 128) {
                    const int n = threadIdx.x;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                    for (int c = 0; c < TS; ++c) {
                        const float hv = sU[n * LD + c];
                        s0 = fmaf(sDOut[0 * LD + c], hv, s0); s1 = fmaf(sDOut[1 * LD + c], hv, s1); s2 = fmaf(sDOut[2 * LD + c], hv, s2);
                    }
                    part[GR_cWo + n] += s0; part[GR_cWo + 128 + n] += s1; part[GR_cWo + 256 + n] += s2;
                } else if (threadIdx.x < 131) {
                    const int c = threadIdx.x - 128;
                    float s = 0.f;
                    for (int k = 0; k < TS; ++k) s += sDOut[c * LD + k];
                    part[GR_cbo + c] += s;
                }
            }
            if (f.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID && threadIdx.x >= 160 && threadIdx.x < 172) {
                // d rot[a][b] = sum_s out_a * g_b ; d trans[b] = sum_s g_b   (g parked in each warp's sDRaw)
                const int e = threadIdx.x - 160;
                float s = 0.f;
                for (int k = 0; k < TS; ++k) {
                    const float* dr = smem + B_DRAW + (k >> 3) * 32 + (k & 7) * 4;
                    if (e < 9) s = fmaf(sOutV[(e / 3) * LD + k], dr[e % 3], s);
                    else s += dr[e - 9];
                }
                part[GR_aff + e] += s;
            }
            __syncthreads();

            // ---- B2: trunk layers 4..0 ----------------------------------------------------------------------------
            float dcc[1][8];
#pragma unroll
            for (int s = 0; s < 8; ++s) dcc[0][s] = 0.f;
            float* dh = sDH;
            float* dhn = sDH2;
#pragma unroll 1
            for (int i = 4; i >= 0; --i) {
                const int K = col_k(i);
                // (a) dc += Fc_i^T dh      (Fc_i (128,32) staged in sU)
                __syncthreads();
                stage_rows(sU, a.P.c_Wc[i], 128 * 32);
                __syncthreads();
                dense8T<1>(dcc, dh + col0, 128, sU, 32, 0, 32, lane);
                // (b) dFc_i += dh (x) c ; dbc_i += sum dh
                if (wc) {
                    outer_acc<8, 2>(dh, sC, 32, part + GR_cWc + 4096 * i, 32, 0);
                    rowsum_acc(dh, 128, part + GR_cbc + 128 * i);
                }
                __syncthreads();
                // (c) dz = dh * softplus'(z_i)    (in place, own columns)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const float z = (m0 + s < M) ? f.save[SL.cz * M + ((long long)i * M + m0 + s) * 128 + lane + 32 * j] : 0.f;
                        dh[(lane + 32 * j) * LD + col0 + s] *= softplus100_grad(z);
                    }
                // (d) layer input h_{i-1} -> sU rows [0,128) ; embedding rows recomputed -> sU rows [128,168)
                float* sIn = sU;
                float* sEmb = sU + 128 * LD;
                if (i >= 1) {
                    __syncthreads();                         // everyone is done with the fc_c stage of step (a)
                    stage_fct(sU, a.P.c_Wc[i - 1], a.P.c_bc[i - 1]);
                    __syncthreads();
                    float hp[4][8];
                    recompute_h(hp, sU, sC + col0, f.save + SL.cz * M + (long long)(i - 1) * M * 128, M, m0, lane);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int s = 0; s < 8; ++s) sIn[(lane + 32 * j) * LD + col0 + s] = hp[j][s];
                }
                if ((i == 0 || i == 3) && lane < PSL_COL_EMB) {
                    const float b0 = __ldg(a.P.c_B + lane), b1 = __ldg(a.P.c_B + 20 + lane), b2 = __ldg(a.P.c_B + 40 + lane);
                    for (int s = 0; s < SPW; ++s) {
                        const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]), z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
                        float sn, cs;
                        sincosf(fmaf(z, b2, fmaf(y, b1, x * b0)), &sn, &cs);
                        sEmb[lane * LD + col0 + s] = sn;
                        sEmb[(20 + lane) * LD + col0 + s] = cs;
                    }
                }
                __syncthreads();
                // (e) dW_i += dz (x) in ; db_i += sum dz
                if (wc) {
                    if (i == 0) outer_acc<8, 3>(dh, sEmb, 40, part + GR_cW(0), 40, 0);
                    else if (i == 3) {
                        outer_acc<8, 3>(dh, sEmb, 40, part + GR_cW(3), 168, 0);
                        outer_acc<8, 8>(dh, sIn, 128, part + GR_cW(3), 168, 40);
                    } else outer_acc<8, 8>(dh, sIn, 128, part + GR_cW(i), 128, 0);
                    rowsum_acc(dh, 128, part + GR_cb + 128 * i);
                }
                // (f) d_in = W_i^T dz       (W_i (128,K) staged in sU)
                __syncthreads();
                stage_rows(sU, a.P.c_W[i], 128 * K);
                __syncthreads();
                if (i == 0 || i == 3) {
                    float de[2][8];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int s = 0; s < 8; ++s) de[j][s] = 0.f;
                    dense8T<2>(de, dh + col0, 128, sU, K, 0, 40, lane);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (lane + 32 * j < 40)
#pragma unroll
                            for (int s = 0; s < 8; ++s) {
                                float* d = sDE + (lane + 32 * j) * LD + col0 + s;
                                *d = (i == 3) ? de[j][s] : (*d + de[j][s]);
                            }
                }
                if (i >= 1) {
                    float din[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int s = 0; s < 8; ++s) din[j][s] = 0.f;
                    dense8T<4>(din, dh + col0, 128, sU, K, i == 3 ? 40 : 0, 128, lane);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int s = 0; s < 8; ++s) dhn[(lane + 32 * j) * LD + col0 + s] = din[j][s];
                }
                float* t = dh; dh = dhn; dhn = t;
            }
            __syncwarp();
            // ---- B3: colour embedding gradient -> d_pos (colour B is not a parameter, decoder.py:27-28) ----------
            if (a.d_pos) {
                const bool lv = lane < PSL_COL_EMB;
                const float b0 = lv ? __ldg(a.P.c_B + lane) : 0.f, b1 = lv ? __ldg(a.P.c_B + 20 + lane) : 0.f,
                            b2 = lv ? __ldg(a.P.c_B + 40 + lane) : 0.f;
                for (int s = 0; s < SPW; ++s) {
                    const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]), z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
                    float da = 0.f;
                    if (lv) {
                        float sn, cs;
                        sincosf(fmaf(z, b2, fmaf(y, b1, x * b0)), &sn, &cs);
                        da = sDE[lane * LD + col0 + s] * cs - sDE[(20 + lane) * LD + col0 + s] * sn;
                    }
                    const float gx = warp_sum(da * b0), gy = warp_sum(da * b1), gz = warp_sum(da * b2);
                    if (lane == 0) { sDP[s * 4] += kTwoPi * gx; sDP[s * 4 + 1] += kTwoPi * gy; sDP[s * 4 + 2] += kTwoPi * gz; }
                }
            }
            __syncwarp();
            // park dc_c in sDC (zero where the sample had no neighbours: c was the random vector)
#pragma unroll
            for (int s = 0; s < 8; ++s) sDC[lane * LD + col0 + s] = sHas[s] ? dcc[0][s] : 0.f;
            __syncwarp();

            // ---- B4: per-neighbour colour MLP ---------------------------------------------------------------------
            if (rel) {
                // scratch inside sU: N1 (128,52) | N2 (32,128) | X [52][LD] | DF [32][LD]
                float* sN1 = sU;
                float* sN2 = sU + 128 * 52;
                float* sX = sN2 + 32 * 128;
                float* sDF = sX + 52 * LD;
                __syncthreads();
                stage_rows(sN1, a.P.c_N1, 128 * 52);
                stage_rows(sN2, a.P.c_N2, 32 * 128);
                __syncthreads();
                float accN1[8][4], accN2[2][8], accB1 = 0.f, accB2 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) accN1[i][j] = 0.f;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) accN2[i][j] = 0.f;
                const float brx = lane < 30 ? __ldg(a.P.c_Brel + lane) : 0.f;   // lane j<10: B[0][j], 10..19: B[1][j-10], 20..29: B[2][j-20]
                float dBrel = 0.f;                                              // lane e<30 accumulates d Brel[e/10][e%10]
                for (int s = 0; s < SPW; ++s) {
                    const long long m = m0 + s;
                    const bool act = sHas[s];
                    // rebuild x (same arithmetic as the forward)
                    const int r = lane & 7;
                    const int idx = sI[s * 8 + r];
                    float rx = 0.f, ry = 0.f, rz = 0.f;
                    if (idx >= 0) {
                        rx = __fmul_rn(kTwoPi, __fsub_rn(__ldg(f.cloud_pos + (size_t)idx * 3), sP[s * 4]));
                        ry = __fmul_rn(kTwoPi, __fsub_rn(__ldg(f.cloud_pos + (size_t)idx * 3 + 1), sP[s * 4 + 1]));
                        rz = __fmul_rn(kTwoPi, __fsub_rn(__ldg(f.cloud_pos + (size_t)idx * 3 + 2), sP[s * 4 + 2]));
                    }
                    float sn3[3], cs3[3];
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int jj = (lane >> 3) + 4 * t;
                        const int jc = jj < PSL_REL_EMB ? jj : 0;
                        sn3[t] = 0.f; cs3[t] = 0.f;
                        const float B0 = __shfl_sync(0xffffffffu, brx, jc), B1 = __shfl_sync(0xffffffffu, brx, 10 + jc),
                                    B2 = __shfl_sync(0xffffffffu, brx, 20 + jc);
                        if (jj < PSL_REL_EMB) {
                            if (idx >= 0) sincosf(fmaf(rz, B2, fmaf(ry, B1, rx * B0)), &sn3[t], &cs3[t]);
                            sX[jj * LD + col0 + r] = sn3[t];
                            sX[(10 + jj) * LD + col0 + r] = cs3[t];
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int id2 = sI[s * 8 + rr];
                        sX[(20 + lane) * LD + col0 + rr] = id2 >= 0 ? __ldg(f.col_feats + (size_t)id2 * 32 + lane) : 0.f;
                    }
                    // df = wn * dc_c ; d wn += dc_c . f
                    const float dcl = sDC[lane * LD + col0 + s];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const float fv = (m < M) ? f.save[SL.nf * M + (m * 8 + rr) * 32 + lane] : 0.f;
                        const float dot = warp_sum(dcl * fv);
                        if (lane == 0 && act) sDWn[s * 8 + rr] += dot;
                        sDF[lane * LD + col0 + rr] = sWn[s * 8 + rr] * dcl;
                    }
                    __syncwarp();
                    // dh1 = N2^T df ; dz1 = dh1 * softplus'(z1) -> sDH ; h1 = softplus(z1) -> sDH2
                    float dh1[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) dh1[j][rr] = 0.f;
                    dense8T<4>(dh1, sDF + col0, 32, sN2, 128, 0, 128, lane);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const float z1 = (m < M) ? f.save[SL.nz1 * M + (m * 8 + rr) * 128 + lane + 32 * j] : 0.f;
                            sDH[(lane + 32 * j) * LD + col0 + rr] = dh1[j][rr] * softplus100_grad(z1);
                            sDH2[(lane + 32 * j) * LD + col0 + rr] = softplus100(z1);
                        }
                    __syncwarp();
                    // dx = N1^T dz1 (52 rows: 20 embedding + 32 feature)
                    float dx[2][8];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) dx[j][rr] = 0.f;
                    dense8T<2>(dx, sDH + col0, 128, sN1, 52, 0, 52, lane);
                    // feature part: rows 20..51  -> lanes 20..31 hold rows 20..31 (j=0), lanes 0..19 hold rows 32..51 (j=1)
                    if (a.d_colpair && m < M) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            float* dst = a.d_colpair + ((size_t)m * 8 + rr) * 32;
                            const bool live = act && sI[s * 8 + rr] >= 0;
                            if (lane >= 20) dst[lane - 20] = live ? dx[0][rr] : 0.f;
                            if (lane < 20) dst[12 + lane] = live ? dx[1][rr] : 0.f;
                        }
                    }
                    // embedding part: rows 0..19 in lanes 0..19 (j=0): park in sDE rows 0..19 (own columns) for the
                    // rel-pos gradient (needs the [row = neighbour] view)
                    if (lane < 20) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) sDE[lane * LD + col0 + rr] = act ? dx[0][rr] : 0.f;
                    }
                    __syncwarp();
                    {   // d arg_jj = dsin*cos - dcos*sin ; d rel = 2 pi * sum_jj d arg * B[:,jj] ; d pos -= d rel
                        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const int jj = (lane >> 3) + 4 * t;
                            const int jc = jj < PSL_REL_EMB ? jj : 0;
                            const float B0 = __shfl_sync(0xffffffffu, brx, jc), B1 = __shfl_sync(0xffffffffu, brx, 10 + jc),
                                        B2 = __shfl_sync(0xffffffffu, brx, 20 + jc);
                            float da = 0.f;
                            if (jj < PSL_REL_EMB && idx >= 0)
                                da = sDE[jj * LD + col0 + r] * cs3[t] - sDE[(10 + jj) * LD + col0 + r] * sn3[t];
                            gx = fmaf(da, B0, gx); gy = fmaf(da, B1, gy); gz = fmaf(da, B2, gz);
                            // d Brel[c][jj] += (2 pi rel_c) * d arg : reduce over the 8 neighbours (lanes with equal lane>>3)
                            float t0 = da * rx, t1 = da * ry, t2 = da * rz;
                            t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t0 += __shfl_xor_sync(0xffffffffu, t0, 2); t0 += __shfl_xor_sync(0xffffffffu, t0, 4);
                            t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 2); t1 += __shfl_xor_sync(0xffffffffu, t1, 4);
                            t2 += __shfl_xor_sync(0xffffffffu, t2, 1); t2 += __shfl_xor_sync(0xffffffffu, t2, 2); t2 += __shfl_xor_sync(0xffffffffu, t2, 4);
                            // lane e (<30) owns Brel[e/10][e%10]; the value for jj sits in lanes 8*(jj%4).. of pass t=jj/4
                            const int e = lane, ej = e % 10, ec = e / 10;
                            const float v0 = __shfl_sync(0xffffffffu, t0, (ej & 3) * 8), v1 = __shfl_sync(0xffffffffu, t1, (ej & 3) * 8),
                                        v2 = __shfl_sync(0xffffffffu, t2, (ej & 3) * 8);
                            if (e < 30 && (ej >> 2) == t) dBrel += ec == 0 ? v0 : (ec == 1 ? v1 : v2);
                        }
                        // sum over the 4 jj-groups sharing a neighbour (lanes r, r+8, r+16, r+24)
                        gx += __shfl_xor_sync(0xffffffffu, gx, 8); gx += __shfl_xor_sync(0xffffffffu, gx, 16);
                        gy += __shfl_xor_sync(0xffffffffu, gy, 8); gy += __shfl_xor_sync(0xffffffffu, gy, 16);
                        gz += __shfl_xor_sync(0xffffffffu, gz, 8); gz += __shfl_xor_sync(0xffffffffu, gz, 16);
                        // and over the 8 neighbours
                        gx += __shfl_xor_sync(0xffffffffu, gx, 1); gx += __shfl_xor_sync(0xffffffffu, gx, 2); gx += __shfl_xor_sync(0xffffffffu, gx, 4);
                        gy += __shfl_xor_sync(0xffffffffu, gy, 1); gy += __shfl_xor_sync(0xffffffffu, gy, 2); gy += __shfl_xor_sync(0xffffffffu, gy, 4);
                        gz += __shfl_xor_sync(0xffffffffu, gz, 1); gz += __shfl_xor_sync(0xffffffffu, gz, 2); gz += __shfl_xor_sync(0xffffffffu, gz, 4);
                        if (lane == 0) { sDP[s * 4] -= kTwoPi * gx; sDP[s * 4 + 1] -= kTwoPi * gy; sDP[s * 4 + 2] -= kTwoPi * gz; }
                    }
                    // CTA-cooperative weight gradients of this round (64 neighbour rows: 8 warps x 8)
                    __syncthreads();
                    if (wc) {
                        const int tb = threadIdx.x & 15, ta = threadIdx.x >> 4;
                        for (int s4 = 0; s4 < TS / 4; ++s4) {
                            float4 av[8], bv[4];
#pragma unroll
                            for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(sDH + (ta + 16 * i) * LD + 4 * s4);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int b = tb + 16 * j;
                                bv[j] = b < 52 ? *reinterpret_cast<const float4*>(sX + b * LD + 4 * s4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    accN1[i][j] = fmaf(av[i].x, bv[j].x, accN1[i][j]); accN1[i][j] = fmaf(av[i].y, bv[j].y, accN1[i][j]);
                                    accN1[i][j] = fmaf(av[i].z, bv[j].z, accN1[i][j]); accN1[i][j] = fmaf(av[i].w, bv[j].w, accN1[i][j]);
                                }
                            float4 dv[2], hv[8];
#pragma unroll
                            for (int i = 0; i < 2; ++i) dv[i] = *reinterpret_cast<const float4*>(sDF + (ta + 16 * i) * LD + 4 * s4);
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = *reinterpret_cast<const float4*>(sDH2 + (tb + 16 * j) * LD + 4 * s4);
#pragma unroll
                            for (int i = 0; i < 2; ++i)
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    accN2[i][j] = fmaf(dv[i].x, hv[j].x, accN2[i][j]); accN2[i][j] = fmaf(dv[i].y, hv[j].y, accN2[i][j]);
                                    accN2[i][j] = fmaf(dv[i].z, hv[j].z, accN2[i][j]); accN2[i][j] = fmaf(dv[i].w, hv[j].w, accN2[i][j]);
                                }
                        }
                        if (threadIdx.x < 128) { float t = 0.f; for (int c = 0; c < TS; ++c) t += sDH[threadIdx.x * LD + c]; accB1 += t; }
                        else if (threadIdx.x < 160) { float t = 0.f; for (int c = 0; c < TS; ++c) t += sDF[(threadIdx.x - 128) * LD + c]; accB2 += t; }
                    }
                    __syncthreads();
                }
                if (wc) {
                    const int tb = threadIdx.x & 15, ta = threadIdx.x >> 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const int b = tb + 16 * j; if (b < 52) part[GR_cN1 + (ta + 16 * i) * 52 + b] += accN1[i][j]; }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) part[GR_cN2 + (ta + 16 * i) * 128 + tb + 16 * j] += accN2[i][j];
                    if (threadIdx.x < 128) part[GR_cn1b + threadIdx.x] += accB1;
                    else if (threadIdx.x < 160) part[GR_cn2b + threadIdx.x - 128] += accB2;
                    // d Brel: one value per (warp, lane<30); reduce the 8 warps in a fixed order through shared memory
                    __syncthreads();
                    if (lane < 30) sU[warp * 32 + lane] = dBrel;
                    __syncthreads();
                    if (threadIdx.x < 30) {
                        float t = 0.f;
                        for (int w = 0; w < NWARP; ++w) t += sU[w * 32 + threadIdx.x];
                        part[GR_cBrel + threadIdx.x] += t;                   // rx/ry/rz already carry the 2 pi factor
                    }
                }
                __syncthreads();
            } else {
                // plain IDW of the colour features: d wn_k += dc_c . col_feats[I_k] ; feature grads via wn * d_cc
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float dcl = sDC[lane * LD + col0 + s];
                    if (a.d_cc && m0 + s < M) a.d_cc[(m0 + s) * 32 + lane] = dcl;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int idx = sI[s * 8 + k];
                        const float fv = idx >= 0 ? __ldg(f.col_feats + (size_t)idx * 32 + lane) : 0.f;
                        const float dot = warp_sum(dcl * fv);
                        if (lane == 0 && sHas[s]) sDWn[s * 8 + k] += dot;
                    }
                }
            }
        }

        // ================= geometry branch ====================================================================
        {
            __syncthreads();
            // stage all geometry matrices (reference layouts) into sU:  W0 | W1 | W2 | W3 | W4 | Wc[5]
            float* gW0 = sU;                    // (32,93)
            float* gW1 = gW0 + 32 * 93;
            float* gW2 = gW1 + 1024;
            float* gW3 = gW2 + 1024;            // (32,125)
            float* gW4 = gW3 + 32 * 125;
            float* gWc = gW4 + 1024;            // 5 x (32,32)
            float* sIn = gWc + 5 * 1024;        // [32][LD] layer input (16 B aligned: offset is a multiple of 4)
            float* sEmb = sDH2;                 // [96][LD] Fourier embedding (sDH2 is free in this phase)
            static_assert((32 * 93 + 3 * 1024 + 32 * 125 + 5 * 1024) % 4 == 0, "alignment");
            if (color || wg || tile == (long long)blockIdx.x) {   // geometry-only data-gradient launches keep the matrices resident
                stage_rows(gW0, a.P.g_W[0], 32 * 93);
                stage_rows(gW1, a.P.g_W[1], 1024);
                stage_rows(gW2, a.P.g_W[2], 1024);
                stage_rows(gW3, a.P.g_W[3], 32 * 125);
                stage_rows(gW4, a.P.g_W[4], 1024);
                for (int i = 0; i < 5; ++i) stage_rows(gWc + 1024 * i, a.P.g_Wc[i], 1024);
            }
            // ReLU masks of all five layers up front (one coalesced 128-byte row per sample and layer): the loads are in flight
            // while the weights are staged instead of being waited for layer by layer
            unsigned zmask[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                unsigned mk = 0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float z = (m0 + s < M) ? f.save[SL.gz * M + ((long long)i * M + m0 + s) * 32 + lane] : 0.f;
                    mk |= (z > 0.f ? 1u : 0u) << s;
                }
                zmask[i] = mk;
            }
            const bool need_de = wg || (a.d_pos != nullptr);   // embedding gradient: only for g_B or the sample position
            __syncthreads();
            // d occ -> dh4 = Wo^T d occ
            const float wo = __ldg(a.P.g_Wo + lane);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                sDH[lane * LD + col0 + s] = wo * sDRaw[s * 4 + 3];
                if (wg) {                                // layer inputs / c_g are operands of the weight gradients only
                    sC[lane * LD + col0 + s] = (m0 + s < M) ? f.save[SL.cg * M + (m0 + s) * 32 + lane] : 0.f;
                    sIn[lane * LD + col0 + s] = (m0 + s < M) ? f.save[SL.gh * M + (4ll * M + m0 + s) * 32 + lane] : 0.f;
                }
            }
            // geometry embedding (recomputed for dW0 / dW3) + the accumulator of its gradient
            if (need_de) {
                for (int s = 0; s < SPW; ++s) {
                    const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]), z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int j = lane + 32 * t;
                        float v = 0.f;
                        if (wg && j < PSL_GEO_EMB)
                            v = sinf(fmaf(z, __ldg(a.P.g_B + 2 * 93 + j), fmaf(y, __ldg(a.P.g_B + 93 + j), x * __ldg(a.P.g_B + j))));
                        sEmb[j * LD + col0 + s] = v;
                        sDE[j * LD + col0 + s] = 0.f;
                    }
                }
            }
            __syncthreads();
            if (wg) {   // d Wo (1,32), d bo
                if (threadIdx.x < 32) {
                    float t = 0.f;
                    for (int c = 0; c < TS; ++c) t = fmaf(smem[B_DRAW + (c >> 3) * 32 + (c & 7) * 4 + 3], sIn[threadIdx.x * LD + c], t);
                    part[GR_gWo + threadIdx.x] += t;
                } else if (threadIdx.x == 32) {
                    float t = 0.f;
                    for (int c = 0; c < TS; ++c) t += smem[B_DRAW + (c >> 3) * 32 + (c & 7) * 4 + 3];
                    part[GR_gbo] += t;
                }
            }
            float dcg[1][8];
#pragma unroll
            for (int s = 0; s < 8; ++s) dcg[0][s] = 0.f;
            float* dh = sDH;                         // rows 0..31
            float* dhn = sDH + 32 * LD;              // rows 32..63
#pragma unroll 1
            for (int i = 4; i >= 0; --i) {
                const int K = geo_k(i);
                const float* Wi = i == 0 ? gW0 : i == 1 ? gW1 : i == 2 ? gW2 : i == 3 ? gW3 : gW4;
                // (a) dc += Fc_i^T dh
                dense8T<1>(dcg, dh + col0, 32, gWc + 1024 * i, 32, 0, 32, lane);
                __syncthreads();
                if (wg) {
                    outer_acc<2, 2>(dh, sC, 32, part + GR_gWc + 1024 * i, 32, 0);
                    rowsum_acc(dh, 32, part + GR_gbc + 32 * i);
                }
                __syncthreads();
                // (c) dz = dh * relu'(z)
                {
                    const unsigned mk = i == 4 ? zmask[4] : i == 3 ? zmask[3] : i == 2 ? zmask[2] : i == 1 ? zmask[1] : zmask[0];
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        if (!((mk >> s) & 1u)) dh[lane * LD + col0 + s] = 0.f;
                }
                // (d) layer input
                if (wg && i >= 1) {
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        sIn[lane * LD + col0 + s] = (m0 + s < M) ? f.save[SL.gh * M + ((long long)(i - 1) * M + m0 + s) * 32 + lane] : 0.f;
                }
                __syncthreads();
                if (wg) {
                    if (i == 0) outer_acc<2, 6>(dh, sEmb, 93, part + GR_gW0, 93, 0);
                    else if (i == 3) {
                        outer_acc<2, 6>(dh, sEmb, 93, part + GR_gW3, 125, 0);
                        outer_acc<2, 2>(dh, sIn, 32, part + GR_gW3, 125, 93);
                    } else outer_acc<2, 2>(dh, sIn, 32, part + GR_gW(i), 32, 0);
                    rowsum_acc(dh, 32, part + GR_gb + 32 * i);
                }
                // (f) d_in = W_i^T dz
                if (need_de && (i == 0 || i == 3)) {
                    float de[3][8];
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int s = 0; s < 8; ++s) de[j][s] = 0.f;
                    dense8T<3>(de, dh + col0, 32, Wi, K, 0, 93, lane);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (lane + 32 * j < 93)
#pragma unroll
                            for (int s = 0; s < 8; ++s) sDE[(lane + 32 * j) * LD + col0 + s] += de[j][s];
                }
                if (i >= 1) {
                    float din[1][8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) din[0][s] = 0.f;
                    dense8T<1>(din, dh + col0, 32, Wi, K, i == 3 ? 93 : 0, 32, lane);
#pragma unroll
                    for (int s = 0; s < 8; ++s) dhn[lane * LD + col0 + s] = din[0][s];
                }
                __syncthreads();
                float* t = dh; dh = dhn; dhn = t;
            }
            // geometry embedding gradient: sin only.  d arg = de * cos(arg);  d pos += 2 pi B d arg;  d B[c][j] += 2 pi p_c d arg
            if (need_de) {
                float gB[3][3];
#pragma unroll
                for (int t = 0; t < 3; ++t) { gB[t][0] = 0.f; gB[t][1] = 0.f; gB[t][2] = 0.f; }
                for (int s = 0; s < SPW; ++s) {
                    const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]), z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
                    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int j = lane + 32 * t;
                        if (j < PSL_GEO_EMB) {
                            const float b0 = __ldg(a.P.g_B + j), b1 = __ldg(a.P.g_B + 93 + j), b2 = __ldg(a.P.g_B + 2 * 93 + j);
                            const float da = sDE[j * LD + col0 + s] * cosf(fmaf(z, b2, fmaf(y, b1, x * b0)));
                            gx = fmaf(da, b0, gx); gy = fmaf(da, b1, gy); gz = fmaf(da, b2, gz);
                            gB[t][0] = fmaf(da, x, gB[t][0]); gB[t][1] = fmaf(da, y, gB[t][1]); gB[t][2] = fmaf(da, z, gB[t][2]);
                        }
                    }
                    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
                    if (lane == 0) { sDP[s * 4] += kTwoPi * gx; sDP[s * 4 + 1] += kTwoPi * gy; sDP[s * 4 + 2] += kTwoPi * gz; }
                }
                if (wg) {   // fixed-order reduction of the 8 warps through shared memory (sU is free now)
                    __syncthreads();
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int j = lane + 32 * t;
                        if (j < PSL_GEO_EMB) { sU[(warp * 3 + 0) * 96 + j] = gB[t][0]; sU[(warp * 3 + 1) * 96 + j] = gB[t][1]; sU[(warp * 3 + 2) * 96 + j] = gB[t][2]; }
                    }
                    __syncthreads();
                    for (int e = threadIdx.x; e < 279; e += blockDim.x) {
                        const int c = e / 93, j = e - 93 * c;
                        float t = 0.f;
                        for (int w = 0; w < NWARP; ++w) t += sU[(w * 3 + c) * 96 + j];
                        part[GR_gB + e] += t;
                    }
                }
            }
            __syncwarp();
            // d c_g -> output (zero where no neighbours) and d wn += dc_g . geo_feats[I_k]
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float dcl = sHas[s] ? dcg[0][s] : 0.f;
                if (a.d_cg && m0 + s < M) a.d_cg[(m0 + s) * 32 + lane] = dcl;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = sI[s * 8 + k];
                    const float fv = idx >= 0 ? __ldg(f.geo_feats + (size_t)idx * 32 + lane) : 0.f;
                    const float dot = warp_sum(dcl * fv);
                    if (lane == 0) sDWn[s * 8 + k] += dot;
                }
            }
            __syncwarp();
        }

        // ================= B6: IDW weights -> d_pos (tracker path: D recomputed from cloud_pos, decoder.py:143-148) ===
        if (a.d_pos) {
            if (f.cfg.is_tracker) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
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
                        const long long m = m0 + s;
                        const float Dv = f.D[m * 8 + k];
                        if (f.cfg.weighting == PSL_WEIGHT_EXPO) dD = dw * wr * (-10.0f / sqrtf(Dv));
                        else dD = -dw * wr * wr;
                        // D = sum (c - p)^2  ->  dD/dp = -2 (c - p)
                        const float cx = __ldg(f.cloud_pos + (size_t)idx * 3) - sP[s * 4];
                        const float cy = __ldg(f.cloud_pos + (size_t)idx * 3 + 1) - sP[s * 4 + 1];
                        const float cz = __ldg(f.cloud_pos + (size_t)idx * 3 + 2) - sP[s * 4 + 2];
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
            if (lane < 24) {
                const int s = lane / 3, c = lane - 3 * s;
                if (m0 + s < M) a.d_pos[(m0 + s) * 3 + c] = sDP[s * 4 + c];
            }
        }
        __syncthreads();
    }
}

// fixed-order reduction of the per-CTA partial buffers into the caller's gradient tensors
struct ReduceJob { float* dst; int off, n; };
struct ReduceJobs { ReduceJob j[64]; int n; };
__global__ void k_reduce_partials(const float* __restrict__ partial, int n_cta, ReduceJobs jobs) {
    const ReduceJob J = jobs.j[blockIdx.y];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < J.n; e += gridDim.x * blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // 4 interleaved accumulators, combined in a fixed order
        int c = 0;
        for (; c + 4 <= n_cta; c += 4) {
            s0 += partial[(size_t)c * GR_TOTAL + J.off + e];
            s1 += partial[(size_t)(c + 1) * GR_TOTAL + J.off + e];
            s2 += partial[(size_t)(c + 2) * GR_TOTAL + J.off + e];
            s3 += partial[(size_t)(c + 3) * GR_TOTAL + J.off + e];
        }
        for (; c < n_cta; ++c) s0 += partial[(size_t)c * GR_TOTAL + J.off + e];
        J.dst[e] = (s0 + s1) + (s2 + s3);
    }
}

}  // namespace psl

using namespace psl;

static long long bwd_grid(long long m) {
    const long long n_tiles = (m + TS - 1) / TS;
    return n_tiles < sm_count() ? n_tiles : sm_count();
}

extern "C" size_t psl_decode_bwd_ws_bytes(int64_t m) {
    if (m <= 0) return 256;
    return sizeof(float) * (size_t)GR_TOTAL * (size_t)bwd_grid(m) + 256;
}

extern "C" int psl_decode_bwd(const psl_decode_cfg* cfg, const psl_decoder_params* P, const float* packed,
                              const float* pos, int64_t m, const int32_t* I, const float* D, const int32_t* nnum,
                              const double* r2, const float* cloud_pos, const float* geo_feats, const float* col_feats,
                              const float* exposure_affine, const float* raw, const float* save, const float* d_raw,
                              float* d_pos, float* d_cg, float* wn, float* d_colpair, const psl_decoder_grads* G,
                              float* d_exposure_affine, const float* dwn_extra, const float* dpos_extra, void* ws,
                              size_t ws_bytes, psl_stream_t stream) {
    PSL_REQUIRE(cfg && P && pos && I && D && nnum && geo_feats && raw && save && d_raw && ws, "NULL argument");
    PSL_REQUIRE(m >= 0, "m < 0");
    const bool color = cfg->stage == PSL_STAGE_COLOR;
    PSL_REQUIRE(!color || col_feats, "colour stage needs col_feats");
    PSL_REQUIRE(!(color && cfg->encode_rel_pos) || cloud_pos, "rel-pos encoding needs cloud_pos");
    PSL_REQUIRE(!(cfg->is_tracker && d_pos) || cloud_pos, "tracker backward needs cloud_pos");
    PSL_REQUIRE(cfg->rgb_mode != PSL_RGB_AFFINE_SIGMOID || exposure_affine, "affine mode needs exposure_affine");
    if (m == 0) return 0;
    PSL_REQUIRE(ws_bytes >= psl_decode_bwd_ws_bytes(m), "workspace too small");
    cudaStream_t st = as_stream(stream);
    BwdArgs a{};
    a.f.cfg = *cfg; a.f.packed = packed; a.f.pos = pos; a.f.m = m; a.f.I = I; a.f.D = D; a.f.nnum = nnum; a.f.r2 = r2;
    a.f.cloud_pos = cloud_pos; a.f.geo_feats = geo_feats; a.f.col_feats = col_feats; a.f.affine = exposure_affine;
    a.f.raw = const_cast<float*>(raw); a.f.save = const_cast<float*>(save);
    a.P = *P;
    a.d_raw = d_raw; a.d_pos = d_pos; a.d_cg = d_cg; a.wn_out = wn;
    const bool rel = color && cfg->encode_rel_pos;
    a.d_colpair = rel ? d_colpair : nullptr;
    a.d_cc = (color && !rel) ? d_colpair : nullptr;      // without the neighbour MLP the caller gets d_cc (m,32) here
    a.partial = static_cast<float*>(ws);
    a.dwn_extra = dwn_extra; a.dpos_extra = dpos_extra;
    bool wg = false, wc = false;
    if (G) {
        wg = G->g_B || G->g_Wo || G->g_bo;
        wc = G->c_Brel || G->c_N1 || G->c_N2 || G->c_n1b || G->c_n2b || G->c_Wo || G->c_bo;
        for (int i = 0; i < 5; ++i) {
            wg = wg || G->g_W[i] || G->g_b[i] || G->g_Wc[i] || G->g_bc[i];
            wc = wc || G->c_W[i] || G->c_b[i] || G->c_Wc[i] || G->c_bc[i];
        }
    }
    if (d_exposure_affine) wc = true;
    a.want_geo_params = wg; a.want_col_params = wc && color;
    if (!color && (cfg->reserved & PSL_GEO_MMA_BIT)) {
        PSL_REQUIRE(!wg, "the tensor-core geometry backward returns data gradients only (clear bit 1 of cfg.reserved)");
        PSL_REQUIRE(packed, "NULL argument");
        return geo_bwd_mma(cfg, packed, pos, m, I, D, nnum, r2, cloud_pos, geo_feats, save, d_raw, d_pos, d_cg, wn, dwn_extra,
                           dpos_extra, st);
    }
    const long long n_tiles = (m + TS - 1) / TS;
    const long long grid = bwd_grid(m);
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_decode_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BWD_BYTES));
    {
        TimingScope ts(T_DECODE_BWD, st, 2);   // memset + kernel
        PSL_CHECK_CUDA(cudaMemsetAsync(a.partial, 0, sizeof(float) * (size_t)GR_TOTAL * grid, st));
        k_decode_bwd<<<(unsigned)grid, NWARP * 32, SM_BWD_BYTES, st>>>(a, n_tiles);
        PSL_CHECK_CUDA(cudaGetLastError());
    }
    if (G || d_exposure_affine) {
        ReduceJobs J;
        J.n = 0;
        auto add = [&](float* dst, int off, int n) { if (dst) { J.j[J.n].dst = dst; J.j[J.n].off = off; J.j[J.n].n = n; ++J.n; } };
        if (G) {
            add(G->g_B, GR_gB, 279);
            for (int i = 0; i < 5; ++i) {
                add(G->g_W[i], GR_gW(i), 32 * geo_k(i));
                add(G->g_b[i], GR_gb + 32 * i, 32);
                add(G->g_Wc[i], GR_gWc + 1024 * i, 1024);
                add(G->g_bc[i], GR_gbc + 32 * i, 32);
            }
            add(G->g_Wo, GR_gWo, 32);
            add(G->g_bo, GR_gbo, 1);
            if (color) {
                add(G->c_Brel, GR_cBrel, 30);
                add(G->c_N1, GR_cN1, 128 * 52); add(G->c_n1b, GR_cn1b, 128);
                add(G->c_N2, GR_cN2, 4096); add(G->c_n2b, GR_cn2b, 32);
                for (int i = 0; i < 5; ++i) {
                    add(G->c_W[i], GR_cW(i), 128 * col_k(i));
                    add(G->c_b[i], GR_cb + 128 * i, 128);
                    add(G->c_Wc[i], GR_cWc + 4096 * i, 4096);
                    add(G->c_bc[i], GR_cbc + 128 * i, 128);
                }
                add(G->c_Wo, GR_cWo, 384);
                add(G->c_bo, GR_cbo, 3);
            }
        }
        if (color) add(d_exposure_affine, GR_aff, 12);
        if (J.n > 0) {
            TimingScope ts(T_REDUCE, st);
            k_reduce_partials<<<dim3(16, J.n), 256, 0, st>>>(a.partial, (int)grid, J);
            PSL_CHECK_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

// d_pos (m,3) += dpos_add + chain rule of the IDW weights applied to dwn (m,8): the part of the tracker's position gradient that comes
// from another kernel's gradient on the normalised weights (decoder.py:143-163), as a separate pass so that the two branch backwards
// can run concurrently (see psl_idw_chain in include/pointslam_b200.h)
extern "C" int psl_idw_chain(const psl_decode_cfg* cfg, const float* pos, int64_t m, const int32_t* I, const float* D, const double* r2,
                             const float* cloud_pos, const float* dwn, const float* dpos_add, float* d_pos, psl_stream_t stream) {
    PSL_REQUIRE(cfg && pos && I && D && cloud_pos && dwn && d_pos, "NULL argument");
    PSL_REQUIRE(m >= 0, "m < 0");
    PSL_REQUIRE(r2 == nullptr || cfg->r2_group >= 1, "r2_group must be >= 1");
    if (m == 0) return 0;
    return geo_idw_chain(cfg, pos, m, I, D, r2, cloud_pos, dwn, dpos_add, d_pos, as_stream(stream));
}

