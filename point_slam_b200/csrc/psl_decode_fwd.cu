// K2+K3 forward: IDW interpolation, per-neighbour colour MLP, geometry + colour trunk, fused per 64-sample tile.
// Activations never leave shared memory (except the optional copy kept for the backward); weights are staged
// per layer from the packed blob (L2 resident).  fp32 FFMA path: exact to the reference's fp32 arithmetic.
//
// Follows src/conv_onet/models/decoder.py:130-222 (geometry), :341-449 (colour), :476-518 (POINT.forward).
#include "psl_decode.cuh"

namespace psl {

// ------------------------------------------------------------------------------------------------
// parameter packing
// ------------------------------------------------------------------------------------------------
struct PackJob {
    const float* src;
    int dst_off, dst_ld, na, nb, src_ld, src_col0, transpose;
};
constexpr int MAX_JOBS = 64;
struct PackJobs {
    PackJob j[MAX_JOBS];
    int n;
};

__global__ void k_pack(PackJobs jobs, float* __restrict__ dst) {
    const PackJob& J = jobs.j[blockIdx.y];
    const int total = J.na * J.nb;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int a = e / J.nb, b = e - a * J.nb;
        const float v = J.transpose ? J.src[(size_t)b * J.src_ld + J.src_col0 + a]
                                    : J.src[(size_t)a * J.src_ld + J.src_col0 + b];
        dst[J.dst_off + a * J.dst_ld + b] = v;
    }
}

static void add_job(PackJobs& P, const float* src, int dst_off, int dst_ld, int na, int nb, int src_ld, int src_col0,
                    int transpose) {
    PackJob& j = P.j[P.n++];
    j.src = src; j.dst_off = dst_off; j.dst_ld = dst_ld; j.na = na; j.nb = nb; j.src_ld = src_ld;
    j.src_col0 = src_col0; j.transpose = transpose;
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
constexpr int SM_W = 0;
constexpr int SM_A = SM_W + SW_FLOATS;
constexpr int SM_B = SM_A + 128 * LD;
constexpr int SM_E = SM_B + 128 * LD;
constexpr int SM_CC = SM_E + 40 * LD;
constexpr int SM_CG = SM_CC + 32 * LD;
constexpr int SM_WN = SM_CG + 32 * LD;                 // [NWARP][8][8]
constexpr int SM_I = SM_WN + NWARP * 64;               // [NWARP][8][8] int
constexpr int SM_P = SM_I + NWARP * 64;                // [NWARP][8][4]
constexpr int SM_OUT = SM_P + NWARP * 32;              // [NWARP][8][4]
constexpr int SM_HAS = SM_OUT + NWARP * 32;            // [NWARP][8] int
constexpr int SM_FWD_FLOATS = SM_HAS + NWARP * 8;
constexpr size_t SM_FWD_BYTES = sizeof(float) * SM_FWD_FLOATS;

template <bool SAVE>
__global__ void __launch_bounds__(NWARP * 32, 1) k_decode_fwd(DecodeArgs a, long long n_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem + SM_W;
    float* sA = smem + SM_A;
    float* sB = smem + SM_B;
    float* sE = smem + SM_E;
    float* sCc = smem + SM_CC;
    float* sCg = smem + SM_CG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sWn = smem + SM_WN + warp * 64;
    int* sI = reinterpret_cast<int*>(smem + SM_I) + warp * 64;
    float* sP = smem + SM_P + warp * 32;
    float* sOut = smem + SM_OUT + warp * 32;
    int* sHas = reinterpret_cast<int*>(smem + SM_HAS) + warp * 8;
    const int col0 = warp * SPW;
    const long long M = a.m;
    const bool color = a.cfg.stage == PSL_STAGE_COLOR;
    const bool rel = color && a.cfg.encode_rel_pos;
    const SaveLayout SL = save_layout(color, a.cfg.encode_rel_pos);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long m0 = tile * TS + col0;
        // ---------------- P0: sample meta + normalised IDW weights ---------------------------------------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
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
            const float wn = __fdiv_rn(w, fmaxf(sum, 1e-12f));
            sWn[s * 8 + k] = wn;
            sI[s * 8 + k] = (w != 0.f) ? idx : -1;
            if (k == 0) {
                const int has = (m < M) && (a.nnum[m] >= a.cfg.min_nn);
                sHas[s] = has;
                if (m < M) a.has_nb[m] = (unsigned char)has;
            }
        }
        if (lane < 24) {
            const int s = lane / 3, c = lane - 3 * s;
            const long long m = m0 + s;
            sP[s * 4 + c] = (m < M) ? a.pos[m * 3 + c] : 0.f;
        }
        __syncwarp();

        // ---------------- P1: geometry feature interpolation (decoder.py:164-171) -------------------------
        for (int s = 0; s < SPW; ++s) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = sI[s * 8 + k];
                if (idx >= 0) acc = fmaf(sWn[s * 8 + k], __ldg(a.geo_feats + (size_t)idx * 32 + lane), acc);
            }
            if (!sHas[s]) acc = a.rand_geo[lane];
            sCg[lane * LD + col0 + s] = acc;
            if (SAVE && m0 + s < M) a.save[SL.cg * M + (m0 + s) * 32 + lane] = acc;
        }

        // ---------------- P2: geometry trunk ----------------------------------------------------------------
        __syncthreads();
        if (color || tile == (long long)blockIdx.x)      // geometry-only launches keep the geometry stage resident
            stage_weights(sW, a.packed + OFF_GEO, G_SIZE);
        __syncthreads();
        {
            // Fourier embedding sin(2 pi p B), 93 channels (+3 zero rows) into sB
            for (int s = 0; s < SPW; ++s) {
                const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]),
                            z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int j = lane + 32 * t;
                    float v = 0.f;
                    if (j < PSL_GEO_EMB) {
                        const float arg = fmaf(z, sW[G_B + 2 * 96 + j], fmaf(y, sW[G_B + 96 + j], x * sW[G_B + j]));
                        v = sinf(arg);
                    }
                    sB[j * LD + col0 + s] = v;
                }
            }
            __syncwarp();
            float h[1][8];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                float acc[1][8], fc[1][8];
#pragma unroll
                for (int s = 0; s < 8; ++s) { acc[0][s] = sW[G_BIAS + 32 * i + lane]; fc[0][s] = sW[G_BIASC + 32 * i + lane]; }
                const float* hin = sA + ((i & 1) ? 0 : 32) * LD + col0;      // layer i reads what layer i-1 wrote
                if (i == 0) dense8<1>(acc, sB + col0, 96, sW + G_L0, lane);
                else if (i == 1) dense8<1>(acc, hin, 32, sW + G_L1, lane);
                else if (i == 2) dense8<1>(acc, hin, 32, sW + G_L2, lane);
                else if (i == 3) { dense8<1>(acc, sB + col0, 96, sW + G_L3, lane); dense8<1>(acc, hin, 32, sW + G_L3 + 96 * 32, lane); }
                else dense8<1>(acc, hin, 32, sW + G_L4, lane);
                dense8<1>(fc, sCg + col0, 32, sW + G_FC + 1024 * i, lane);
                float* hout = sA + ((i & 1) ? 32 : 0) * LD + col0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    h[0][s] = __fadd_rn(fmaxf(acc[0][s], 0.f), fc[0][s]);
                    hout[lane * LD + s] = h[0][s];
                    if (SAVE && m0 + s < M) {
                        a.save[SL.gz * M + ((long long)i * M + m0 + s) * 32 + lane] = acc[0][s];
                        a.save[SL.gh * M + ((long long)i * M + m0 + s) * 32 + lane] = h[0][s];
                    }
                }
                __syncwarp();
            }
            const float wo = sW[G_WO + lane], bo = sW[G_BO];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float o = warp_sum(wo * h[0][s]) + bo;
                if (lane == s) sOut[s * 4 + 3] = o;
            }
        }

        if (color) {
            // ---------------- P3: colour feature interpolation (+ per-neighbour MLP, decoder.py:372-388) ----
            if (rel) {
                __syncthreads();
                stage_weights(sW, a.packed + OFF_NBR, N_SIZE);
                __syncthreads();
                for (int s = 0; s < SPW; ++s) {
                    const long long m = m0 + s;
                    {   // x = [sin, cos](2 pi (x_i - p) Brel) (20) ++ col_feats[I] (32), one column per neighbour
                        const int r = lane & 7;
                        const int idx = sI[s * 8 + r];
                        float rx = 0.f, ry = 0.f, rz = 0.f;
                        if (idx >= 0) {
                            rx = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)idx * 3), sP[s * 4]));
                            ry = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)idx * 3 + 1), sP[s * 4 + 1]));
                            rz = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)idx * 3 + 2), sP[s * 4 + 2]));
                        }
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const int jj = (lane >> 3) + 4 * t;
                            if (jj < PSL_REL_EMB) {
                                float sn = 0.f, cs = 0.f;
                                if (idx >= 0) {
                                    const float arg = fmaf(rz, sW[N_BREL + 24 + jj],
                                                           fmaf(ry, sW[N_BREL + 12 + jj], rx * sW[N_BREL + jj]));
                                    sincosf(arg, &sn, &cs);
                                }
                                sA[jj * LD + col0 + r] = sn;
                                sA[(10 + jj) * LD + col0 + r] = cs;
                            }
                        }
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int id2 = sI[s * 8 + rr];
                            sA[(20 + lane) * LD + col0 + rr] = id2 >= 0 ? __ldg(a.col_feats + (size_t)id2 * 32 + lane) : 0.f;
                        }
                    }
                    __syncwarp();
                    float acc[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int r = 0; r < 8; ++r) acc[j][r] = sW[N_B1 + lane + 32 * j];
                    dense8<4>(acc, sA + col0, 52, sW + N_W1, lane);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            if (SAVE && m < M) a.save[SL.nz1 * M + (m * 8 + r) * 128 + lane + 32 * j] = acc[j][r];
                            sB[(lane + 32 * j) * LD + col0 + r] = softplus100(acc[j][r]);
                        }
                    __syncwarp();
                    float f[1][8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) f[0][r] = sW[N_B2 + lane];
                    dense8<1>(f, sB + col0, 128, sW + N_W2, lane);
                    float cc = 0.f;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (SAVE && m < M) a.save[SL.nf * M + (m * 8 + r) * 32 + lane] = f[0][r];
                        cc = fmaf(sWn[s * 8 + r], f[0][r], cc);
                    }
                    if (!sHas[s]) cc = a.rand_col[lane];
                    sCc[lane * LD + col0 + s] = cc;
                    if (SAVE && m < M) a.save[SL.cc * M + m * 32 + lane] = cc;
                    __syncwarp();
                }
            } else {
                for (int s = 0; s < SPW; ++s) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int idx = sI[s * 8 + k];
                        if (idx >= 0) acc = fmaf(sWn[s * 8 + k], __ldg(a.col_feats + (size_t)idx * 32 + lane), acc);
                    }
                    if (!sHas[s]) acc = a.rand_col[lane];
                    sCc[lane * LD + col0 + s] = acc;
                    if (SAVE && m0 + s < M) a.save[SL.cc * M + (m0 + s) * 32 + lane] = acc;
                }
            }

            // ---------------- P4: colour trunk (decoder.py:411-431) -----------------------------------------
            float h[4][8];
#pragma unroll 1
            for (int i = 0; i < 5; ++i) {
                __syncthreads();
                stage_weights(sW, a.packed + OFF_COL(i), CL_SIZE(i));
                __syncthreads();
                if (i == 0 && lane < PSL_COL_EMB) {
                    for (int s = 0; s < SPW; ++s) {
                        const float x = __fmul_rn(kTwoPi, sP[s * 4]), y = __fmul_rn(kTwoPi, sP[s * 4 + 1]),
                                    z = __fmul_rn(kTwoPi, sP[s * 4 + 2]);
                        const float arg = fmaf(z, sW[CL_X(0) + 40 + lane], fmaf(y, sW[CL_X(0) + 20 + lane], x * sW[CL_X(0) + lane]));
                        float sn, cs;
                        sincosf(arg, &sn, &cs);
                        sE[lane * LD + col0 + s] = sn;
                        sE[(20 + lane) * LD + col0 + s] = cs;
                    }
                }
                __syncwarp();
                float acc[4][8], fc[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        acc[j][s] = sW[CL_B(i) + lane + 32 * j];
                        fc[j][s] = sW[CL_BC(i) + lane + 32 * j];
                    }
                const float* hin = ((i & 1) ? sA : sB) + col0;          // layer i-1 wrote here
                float* hout = ((i & 1) ? sB : sA) + col0;
                if (i == 0) dense8<4>(acc, sE + col0, 40, sW, lane);
                else if (i == 3) { dense8<4>(acc, sE + col0, 40, sW, lane); dense8<4>(acc, hin, 128, sW + 40 * 128, lane); }
                else dense8<4>(acc, hin, 128, sW, lane);
                dense8<4>(fc, sCc + col0, 32, sW + CL_FC(i), lane);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        h[j][s] = __fadd_rn(softplus100(acc[j][s]), fc[j][s]);
                        hout[(lane + 32 * j) * LD + s] = h[j][s];
                        if (SAVE && m0 + s < M) {
                            a.save[SL.cz * M + ((long long)i * M + m0 + s) * 128 + lane + 32 * j] = acc[j][s];
                        }
                    }
                __syncwarp();
            }
            // output layer (128 -> 3); sW still holds the layer-4 stage with Wo^T [128][4] and bo
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) part = fmaf(sW[CL_X(4) + (lane + 32 * j) * 4 + c], h[j][s], part);
                    const float o = warp_sum(part) + sW[CL_X(4) + 512 + c];
                    if (lane == s) sOut[s * 4 + c] = o;
                }
            }
        }
        __syncwarp();
        if (lane < SPW && m0 + lane < M) {
            float r = 0.f, g = 0.f, b = 0.f;
            const float occ = sOut[lane * 4 + 3];
            if (color) {
                r = sOut[lane * 4]; g = sOut[lane * 4 + 1]; b = sOut[lane * 4 + 2];
                if (a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                    const float* A = a.affine;        // rot (3x3 row-major), trans (3):  out @ rot + trans
                    const float r2 = fmaf(b, A[6], fmaf(g, A[3], r * A[0])) + A[9];
                    const float g2 = fmaf(b, A[7], fmaf(g, A[4], r * A[1])) + A[10];
                    const float b2 = fmaf(b, A[8], fmaf(g, A[5], r * A[2])) + A[11];
                    r = r2; g = g2; b = b2;
                }
                if (a.cfg.rgb_mode != PSL_RGB_RAW) { r = sigmoidf_(r); g = sigmoidf_(g); b = sigmoidf_(b); }
            }
            if (a.cfg.reserved & 1) a.raw[(m0 + lane) * 4 + 3] = occ;        // occupancy only: rgb belongs to a concurrent colour kernel
            else reinterpret_cast<float4*>(a.raw)[m0 + lane] = make_float4(r, g, b, occ);
        }
        __syncwarp();
    }
}

}  // namespace psl

using namespace psl;

extern "C" size_t psl_packed_params_floats(void) { return (size_t)PACKED_FLOATS + geo_mma_floats(); }

extern "C" size_t psl_decode_save_floats_per_sample(const psl_decode_cfg* cfg) {
    if (!cfg) return 0;
    if (cfg->stage == PSL_STAGE_GEOMETRY && (cfg->reserved & PSL_GEO_MMA_BIT)) return (size_t)GEO_MMA_SAVE_WORDS;
    return (size_t)save_layout(cfg->stage == PSL_STAGE_COLOR, cfg->encode_rel_pos).total;
}

extern "C" int psl_pack_params(const psl_decoder_params* P, float* packed, psl_stream_t stream) {
    PSL_REQUIRE(P && packed, "NULL argument");
    cudaStream_t st = as_stream(stream);
    PSL_CHECK_CUDA(cudaMemsetAsync(packed, 0, sizeof(float) * PACKED_FLOATS, st));
    PackJobs J;
    J.n = 0;
    // geometry
    add_job(J, P->g_B, OFF_GEO + G_B, 96, 3, 93, 93, 0, 0);
    add_job(J, P->g_W[0], OFF_GEO + G_L0, 32, 93, 32, 93, 0, 1);
    add_job(J, P->g_W[1], OFF_GEO + G_L1, 32, 32, 32, 32, 0, 1);
    add_job(J, P->g_W[2], OFF_GEO + G_L2, 32, 32, 32, 32, 0, 1);
    add_job(J, P->g_W[3], OFF_GEO + G_L3, 32, 93, 32, 125, 0, 1);
    add_job(J, P->g_W[3], OFF_GEO + G_L3 + 96 * 32, 32, 32, 32, 125, 93, 1);
    add_job(J, P->g_W[4], OFF_GEO + G_L4, 32, 32, 32, 32, 0, 1);
    for (int i = 0; i < 5; ++i) {
        add_job(J, P->g_Wc[i], OFF_GEO + G_FC + 1024 * i, 32, 32, 32, 32, 0, 1);
        add_job(J, P->g_b[i], OFF_GEO + G_BIAS + 32 * i, 32, 1, 32, 32, 0, 0);
        add_job(J, P->g_bc[i], OFF_GEO + G_BIASC + 32 * i, 32, 1, 32, 32, 0, 0);
    }
    add_job(J, P->g_Wo, OFF_GEO + G_WO, 32, 1, 32, 32, 0, 0);
    add_job(J, P->g_bo, OFF_GEO + G_BO, 1, 1, 1, 1, 0, 0);
    // neighbour MLP
    add_job(J, P->c_Brel, OFF_NBR + N_BREL, 12, 3, 10, 10, 0, 0);
    add_job(J, P->c_N1, OFF_NBR + N_W1, 128, 52, 128, 52, 0, 1);
    add_job(J, P->c_N2, OFF_NBR + N_W2, 32, 128, 32, 128, 0, 1);
    add_job(J, P->c_n1b, OFF_NBR + N_B1, 128, 1, 128, 128, 0, 0);
    add_job(J, P->c_n2b, OFF_NBR + N_B2, 32, 1, 32, 32, 0, 0);
    // colour trunk
    for (int i = 0; i < 5; ++i) {
        const int K = col_k(i), off = OFF_COL(i);
        add_job(J, P->c_W[i], off, 128, K, 128, K, 0, 1);
        add_job(J, P->c_Wc[i], off + CL_FC(i), 128, 32, 128, 32, 0, 1);
        add_job(J, P->c_b[i], off + CL_B(i), 128, 1, 128, 128, 0, 0);
        add_job(J, P->c_bc[i], off + CL_BC(i), 128, 1, 128, 128, 0, 0);
    }
    add_job(J, P->c_B, OFF_COL(0) + CL_X(0), 20, 3, 20, 20, 0, 0);
    add_job(J, P->c_Wo, OFF_COL(4) + CL_X(4), 4, 128, 3, 128, 0, 1);
    add_job(J, P->c_bo, OFF_COL(4) + CL_X(4) + 512, 3, 1, 3, 3, 0, 0);
    for (int i = 0; i < J.n; ++i) PSL_REQUIRE(J.j[i].src != nullptr, "NULL parameter pointer");
    TimingScope ts(T_PACK, st, 2);
    k_pack<<<dim3(8, J.n), 256, 0, st>>>(J, packed);
    PSL_CHECK_CUDA(cudaGetLastError());
    return geo_mma_pack(P, packed, st);            // pre-split B fragments of the geometry matrices (psl_geo_mma.cu)
}

extern "C" int psl_decode_fwd(const psl_decode_cfg* cfg, const float* packed, const float* pos, int64_t m,
                              const int32_t* I, const float* D, const int32_t* nnum, const double* r2,
                              const float* cloud_pos, const float* geo_feats, const float* col_feats,
                              const float* rand_geo, const float* rand_col, const float* exposure_affine, float* raw,
                              uint8_t* has_nb, float* save, psl_stream_t stream) {
    PSL_REQUIRE(cfg && packed && pos && I && D && nnum && geo_feats && rand_geo && raw && has_nb, "NULL argument");
    PSL_REQUIRE(m >= 0, "m < 0");
    PSL_REQUIRE(cfg->stage == PSL_STAGE_GEOMETRY || (col_feats && rand_col), "colour stage needs col_feats/rand_col");
    PSL_REQUIRE(!(cfg->stage == PSL_STAGE_COLOR && cfg->encode_rel_pos) || cloud_pos, "rel-pos encoding needs cloud_pos");
    PSL_REQUIRE(cfg->rgb_mode != PSL_RGB_AFFINE_SIGMOID || exposure_affine, "affine mode needs exposure_affine");
    PSL_REQUIRE(r2 == nullptr || cfg->r2_group >= 1, "r2_group must be >= 1");
    if (m == 0) return 0;
    if (cfg->stage == PSL_STAGE_GEOMETRY && (cfg->reserved & PSL_GEO_MMA_BIT))
        return geo_fwd_mma(cfg, packed, pos, m, I, D, nnum, r2, geo_feats, rand_geo, raw, has_nb, save, as_stream(stream));
    DecodeArgs a{};
    a.cfg = *cfg; a.packed = packed; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.geo_feats = geo_feats; a.col_feats = col_feats; a.rand_geo = rand_geo;
    a.rand_col = rand_col; a.affine = exposure_affine; a.raw = raw; a.has_nb = has_nb; a.save = save;
    const long long n_tiles = (m + TS - 1) / TS;
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_decode_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_FWD_BYTES));
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_decode_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_FWD_BYTES));
    long long blocks = n_tiles < sm_count() ? n_tiles : sm_count();
    TimingScope ts(T_DECODE_FWD, as_stream(stream));
    if (save) k_decode_fwd<true><<<(unsigned)blocks, NWARP * 32, SM_FWD_BYTES, as_stream(stream)>>>(a, n_tiles);
    else k_decode_fwd<false><<<(unsigned)blocks, NWARP * 32, SM_FWD_BYTES, as_stream(stream)>>>(a, n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
