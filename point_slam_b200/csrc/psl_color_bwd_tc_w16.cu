// 3xTF32 colour backward, data gradients (the A/B baseline of psl_color_bwd_h2.cu, selected with PSL_H2_BWD=0): the round-1
// kernel with SIXTEEN worker warps -- four threads per sample row.  Thread (row r, quarter q) owns columns
// 32q..32q+31 of every 128-wide gradient plane (two 16-column chunks), and the steps the production kernel runs on its half-0
// warps only are spread over the quarters:
//   q0  dL/d(output) operand, colour-embedding -> d_pos, rel-pos backward of every neighbour (d_pos, dBrel), d_pos store
//   q1  df_k operand columns 0-15,  feature-gradient columns 0-11 of the pair
//   q2  df_k operand columns 16-31, feature-gradient columns 12-27
//   q3  d(w_k) dot products, dccT store, feature-gradient columns 28-31, the no-neighbour-MLP outputs
// Every output element is produced by the same arithmetic in the same order as in the production kernel, so results must be
// bit-identical.  Same operand blob, TMEM regions, shared-memory map, producer and MMA issue order (psl_color_bwd_tc.cuh).
//
// Verified bit-identical to the round-1 8-worker-warp kernel on hardware (profiles/r02_s1_*) before that kernel was removed.
#include "psl_color_bwd_tc.cuh"

namespace psl {
namespace cbt16 {

using namespace cbt;
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) { tc::mbar_wait_p(bar, parity); }

constexpr int NWORK16 = 512, NTHR16 = 576;      // warps 0-15 workers, 16 bulk-copy producer, 17 TMEM allocator + MMA issuer

__device__ __forceinline__ void worker_signal16(uint64_t* a_ready) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncwarp();                                    // one arrival per warp: 512 per-thread arrivals on one word serialise
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(a_ready);
}

__global__ void __launch_bounds__(NTHR16, 1) k_color_bwd_tc_w16(Args a, long long n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sVec = reinterpret_cast<float*>(smem + SB_VEC);
    float* sAff = reinterpret_cast<float*>(smem + SB_AFF);
    float* sRed = reinterpret_cast<float*>(smem + SB_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SB_BAR);
    uint64_t* full = bars; uint64_t* empty = bars + 2; uint64_t* nbrw_full = bars + 4;
    uint64_t* a_ready = bars + 5; uint64_t* d_ready = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool rel = a.cfg.encode_rel_pos != 0;
    const TSave TL = tsave_layout(a.m, a.cfg.encode_rel_pos);
    const TBwd BL = tbwd_layout(a.m, a.cfg.encode_rel_pos);

    if (threadIdx.x == 0) {
        tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::mbar_init(&empty[0], 1); tc::mbar_init(&empty[1], 1);
        tc::mbar_init(nbrw_full, 1); tc::mbar_init(a_ready, NWORK16 / 32); tc::mbar_init(d_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < BV_SIZE; i += NTHR16) sVec[i] = a.blob[BB_VEC + i];
    if (threadIdx.x < 12) sAff[threadIdx.x] = a.affine ? a.affine[threadIdx.x] : 0.f;
    if (threadIdx.x < 128) sRed[threadIdx.x] = 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        if (lane == 0) {
            if (rel) {
                tc::mbar_expect_tx(nbrw_full, 98304);
                for (int i = 0; i < 3; ++i) tc::bulk_g2s(smem + SB_NBRW + i * 32768, a.blob + BB_N2T + i * 8192, 32768, nbrw_full);
            }
            uint32_t cnt = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int q = 0; q < NMAT; ++q) {
                    const int N = mat_n(q), K = mat_k(q), kc = K < 32 ? K : 32;
                    for (int c = 0; c * kc < K; ++c, ++cnt) {
                        const int st = cnt & 1;
                        const uint32_t bytes = (uint32_t)(2 * N * kc * 4);
                        mbar_wait_wd(&empty[st], ((cnt >> 1) & 1) ^ 1);
                        tc::mbar_expect_tx(&full[st], bytes);
                        tc::bulk_g2s(smem + SB_RING + st * 32768, a.blob + BB_TRUNK + mat_off(q) + c * 2 * N * kc, bytes, &full[st]);
                    }
                }
            }
        }
    } else if (warp == 17) {
        if (lane == 0) {
            uint32_t pa = 0, cnt = 0;
            const uint32_t n2t = tc::smem_u32(smem + SB_NBRW), n1t = n2t + 2 * 128 * 32 * 4;
            // one streamed matrix: A = (a_hi, a_lo) columns, K reduction, D columns d, N rows of B
            auto run_mat = [&](int q, uint32_t a_hi, uint32_t a_lo, uint32_t d, uint32_t first_acc) {
                const int N = mat_n(q), K = mat_k(q), kc = K < 32 ? K : 32;
                const uint32_t idesc = tc::make_idesc_tf32(128, N), lbo = (uint32_t)N * 16u;
                for (int c = 0; c * kc < K; ++c, ++cnt) {
                    const int st = cnt & 1;
                    mbar_wait_wd(&full[st], (cnt >> 1) & 1);
                    const uint32_t rb = tc::smem_u32(smem + SB_RING + st * 32768);
                    for (int j = 0; j < kc / 8; ++j) {
                        const uint64_t bh = tc::make_smem_desc(rb + j * 2 * lbo, lbo, 128);
                        const uint64_t bl = tc::make_smem_desc(rb + (uint32_t)N * kc * 4 + j * 2 * lbo, lbo, 128);
                        const uint32_t ac = (c == 0 && j == 0) ? first_acc : 1u;
                        const uint32_t ko = (uint32_t)(c * kc + 8 * j);
                        tc::mma_tf32_ts(tmem + d, tmem + a_hi + ko, bh, idesc, ac);
                        tc::mma_tf32_ts(tmem + d, tmem + a_lo + ko, bh, idesc, 1);
                        tc::mma_tf32_ts(tmem + d, tmem + a_hi + ko, bl, idesc, 1);
                    }
                    tc::mma_commit(&empty[st]);
                }
            };
            if (rel) mbar_wait_wd(nbrw_full, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                // output layer: A = dout (P[0:16] / R[0:16])
                mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(0, TP, TR, TQ, 0);               // dh_4 -> Q
                run_mat(1, TP, TR, TDC, 0);              // dc    = dout G_out
                tc::mma_commit(d_ready);
                int q = 2;
                for (int l = 4; l >= 1; --l) {
                    mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                    const uint32_t cur = (l & 1) ? TP : TQ, oth = (l & 1) ? TQ : TP;     // dz_l lives in cur, dh_{l-1} goes to oth
                    run_mat(q++, cur, TR, oth, 0);
                    run_mat(q++, cur, TR, TDC, 1);
                    if (l == 3) run_mat(q++, cur, TR, TDE, 0);
                    tc::mma_commit(d_ready);
                }
                mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(11, TQ, TR, TDE, 1);             // de += dz_0 L_0
                tc::mma_commit(d_ready);
                if (rel) {
                    const uint32_t id128 = tc::make_idesc_tf32(128, 128), id64 = tc::make_idesc_tf32(128, 64);
                    for (int k = 0; k < 8; ++k) {
                        mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 4; ++j) {            // dh1 = df N2   (A: P[0:32]/R[0:32], K = 32) -> Q
                            const uint64_t bh = tc::make_smem_desc(n2t + j * 2 * 2048, 2048, 128);
                            const uint64_t bl = tc::make_smem_desc(n2t + 128 * 32 * 4 + j * 2 * 2048, 2048, 128);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TP + 8 * j, bh, id128, j > 0);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TR + 8 * j, bh, id128, 1);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TP + 8 * j, bl, id128, 1);
                        }
                        tc::mma_commit(d_ready);
                        mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 16; ++j) {           // dx = dz1 N1   (A: Q / R, K = 128) -> DX (64 cols)
                            const uint64_t bh = tc::make_smem_desc(n1t + j * 2 * 1024, 1024, 128);
                            const uint64_t bl = tc::make_smem_desc(n1t + 64 * 128 * 4 + j * 2 * 1024, 1024, 128);
                            tc::mma_tf32_ts(tmem + TDX, tmem + TQ + 8 * j, bh, id64, j > 0);
                            tc::mma_tf32_ts(tmem + TDX, tmem + TR + 8 * j, bh, id64, 1);
                            tc::mma_tf32_ts(tmem + TDX, tmem + TQ + 8 * j, bl, id64, 1);
                        }
                        tc::mma_commit(d_ready);
                    }
                }
            }
        }
    } else {
        // =============================== workers: 4 threads per sample row =============================================
        const int r = 32 * (warp & 3) + lane, q = warp >> 2;
        const uint32_t lb = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t pd = 0;
        const float* Br = sVec + BV_BREL; const float* Bc = sVec + BV_BC;
        float brel_acc = 0.f;                                    // lane e < 30 of the quarter-0 warps: d Brel[e/10][e%10]
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long m = tile * TM + r;
            const bool inb = m < a.m;
            float px = 0.f, py = 0.f, pz = 0.f, wn[8];
            int idx[8];
            bool has = false;
            {
                float sum = 0.f, w[8], tle = -1.f;
                if (inb) {
                    px = a.pos[m * 3]; py = a.pos[m * 3 + 1]; pz = a.pos[m * 3 + 2];
                    tle = thr_le_of(a.r2 ? a.r2[m / a.cfg.r2_group] : a.cfg.r2_scalar);
                    has = a.nnum[m] >= a.cfg.min_nn;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    idx[k] = inb ? a.I[m * 8 + k] : -1;
                    w[k] = inb ? idw_raw(a.D[m * 8 + k], idx[k], tle, a.cfg.weighting) : 0.f;
                    sum += fabsf(w[k]);
                }
                const float den = fmaxf(sum, 1e-12f);
#pragma unroll
                for (int k = 0; k < 8; ++k) { wn[k] = __fdiv_rn(w[k], den); if (w[k] == 0.f) idx[k] = -1; }
                if (q == 0 && inb && a.wn_out) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) a.wn_out[m * 8 + k] = (has && idx[k] >= 0) ? wn[k] : 0.f;
                }
            }
            float dpx = 0.f, dpy = 0.f, dpz = 0.f;
            // ---- dL/d(colour output) -> A operand (16 columns): quarter 0 ---------------------------------------------
            if (q == 0) {
                float g[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) g[j] = 0.f;
                if (inb) {
                    const float4 dr = reinterpret_cast<const float4*>(a.d_raw)[m];
                    g[0] = dr.x; g[1] = dr.y; g[2] = dr.z;
                    if (a.cfg.rgb_mode != PSL_RGB_RAW) {
                        const float4 rv = reinterpret_cast<const float4*>(a.raw)[m];
                        g[0] *= rv.x * (1.0f - rv.x); g[1] *= rv.y * (1.0f - rv.y); g[2] *= rv.z * (1.0f - rv.z);
                    }
                    if (a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                        const float4 op = *reinterpret_cast<const float4*>(a.tsave + TL.outpre + (tile * 128 + r) * 4);
                        float* af = a.tbwd + BL.aff + (tile * 128 + r) * 12;           // d rot[a][b] = out_a g_b ; d trans = g
                        af[0] = op.x * g[0]; af[1] = op.x * g[1]; af[2] = op.x * g[2];
                        af[3] = op.y * g[0]; af[4] = op.y * g[1]; af[5] = op.y * g[2];
                        af[6] = op.z * g[0]; af[7] = op.z * g[1]; af[8] = op.z * g[2];
                        af[9] = g[0]; af[10] = g[1]; af[11] = g[2];
                        const float o0 = sAff[0] * g[0] + sAff[1] * g[1] + sAff[2] * g[2];
                        const float o1 = sAff[3] * g[0] + sAff[4] * g[1] + sAff[5] * g[2];
                        const float o2 = sAff[6] * g[0] + sAff[7] * g[1] + sAff[8] * g[2];
                        g[0] = o0; g[1] = o1; g[2] = o2;
                    }
                } else if (a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                    float* af = a.tbwd + BL.aff + (tile * 128 + r) * 12;
#pragma unroll
                    for (int j = 0; j < 12; ++j) af[j] = 0.f;
                }
                if (a.want_wgrad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) a.tbwd[BL.doutT + (tile * 16 + j) * 128 + r] = g[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) tc::split_tf32(g[j], g[j], lo[j]);
                tc::tmem_st16(lb + TP, g);
                tc::tmem_st16(lb + TR, lo);
            }
            worker_signal16(a_ready);
            // ---- trunk: dh_l -> (store) -> dz_l = dh_l * softplus'(z_l) -> A operand: columns 32q .. 32q+31, two chunks ------
#pragma unroll 1
            for (int l = 4; l >= 0; --l) {
                // saved pre-activations of the first 16 columns are requested BEFORE waiting for the layer's MMAs, those of the
                // second 16 while the first chunk is being processed
                const long long off0 = (((long long)l * n_tiles + tile) * 128 + 32 * q) * 128 + r;
                float zc[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) zc[j] = a.tsave[TL.zT + off0 + j * 128];
                mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                const uint32_t reg = (l & 1) ? TP : TQ;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int c0 = 32 * q + 16 * c;
                    float v[16], zn[16];
                    tc::tmem_ld16(lb + reg + c0, v);
                    const long long off = off0 + (long long)(16 * c) * 128;
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) zn[j] = a.tsave[TL.zT + off + (16 + j) * 128];
                    }
                    if (a.want_wgrad) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) a.tbwd[BL.dhT + off + j * 128] = v[j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) tc::split_tf32(v[j] * sp_grad_fast(zc[j]), v[j], zc[j]);   // zc <- lo
                    tc::tmem_st16(lb + reg + c0, v);
                    tc::tmem_st16(lb + TR + c0, zc);
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) zc[j] = zn[j];
                    }
                }
                worker_signal16(a_ready);
            }
            // ---- dc, de -----------------------------------------------------------------------------------------------------
            mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
            // dc stays in TMEM (region DC is not written again before the next tile's output layer): the quarters that need it
            // re-load their columns where they use them, so nothing of it lives across the neighbour loop
            if (q == 3) {
                float dcv[32];
                tc::tmem_ld32(lb + TDC, dcv);
#pragma unroll
                for (int j = 0; j < 32; ++j) dcv[j] = has ? dcv[j] : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) a.tbwd[BL.dccT + (tile * 32 + j) * 128 + r] = dcv[j];
                if (!rel && inb) {
                    if (a.d_colpair) {
                        float4* dst = reinterpret_cast<float4*>(a.d_colpair + m * 32);
#pragma unroll
                        for (int g = 0; g < 8; ++g) dst[g] = make_float4(dcv[4 * g], dcv[4 * g + 1], dcv[4 * g + 2], dcv[4 * g + 3]);
                    }
                    if (a.dwn_col) {                       // IDW gradient of the plain (no neighbour MLP) interpolation
#pragma unroll 1
                        for (int k = 0; k < 8; ++k) {
                            float dot = 0.f;
                            if (idx[k] >= 0) {
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)idx[k] * 32) + g);
                                    dot = fmaf(dcv[4 * g], f4.x, dot); dot = fmaf(dcv[4 * g + 1], f4.y, dot);
                                    dot = fmaf(dcv[4 * g + 2], f4.z, dot); dot = fmaf(dcv[4 * g + 3], f4.w, dot);
                                }
                            }
                            a.dwn_col[m * 8 + k] = dot;
                        }
                    }
                }
            } else if (q == 0 && a.dpos_col) {             // quarter 0: colour Fourier embedding, d arg_j = dsin_j cos - dcos_j sin
                float e0[32], e1[16];
                tc::tmem_ld32(lb + TDE, e0);
                tc::tmem_ld16(lb + TDE + 32, e1);
                const float x = __fmul_rn(kTwoPi, px), y = __fmul_rn(kTwoPi, py), z = __fmul_rn(kTwoPi, pz);
                float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                for (int j = 0; j < 20; ++j) {
                    float sn, cs;
                    sincos_embed(fmaf(z, Bc[40 + j], fmaf(y, Bc[20 + j], x * Bc[j])), &sn, &cs);
                    const float dcos = j < 12 ? e0[20 + j] : e1[j - 12];
                    const float da = e0[j] * cs - dcos * sn;
                    gx = fmaf(da, Bc[j], gx); gy = fmaf(da, Bc[20 + j], gy); gz = fmaf(da, Bc[40 + j], gz);
                }
                dpx += kTwoPi * gx; dpy += kTwoPi * gy; dpz += kTwoPi * gz;
            }
            // ---- neighbour MLP backward ------------------------------------------------------------------------------------
            if (rel) {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    const int id = idx[k];
                    // df_k = wn_k dc -> A (32 columns: quarters 1 and 2, 16 each); d wn_k = dc . f_k (quarter 3)
                    if (q == 1 || q == 2) {
                        float v[16], lo[16];
                        tc::tmem_ld16(lb + TDC + 16 * (q - 1), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) tc::split_tf32(wn[k] * (has ? v[j] : 0.f), v[j], lo[j]);
                        tc::tmem_st16(lb + TP + 16 * (q - 1), v);
                        tc::tmem_st16(lb + TR + 16 * (q - 1), lo);
                    } else if (q == 3 && a.dwn_col) {
                        float dcv[32];
                        tc::tmem_ld32(lb + TDC, dcv);
                        float dot = 0.f;
                        const float4* fr = reinterpret_cast<const float4*>(a.tsave + TL.f + ((tile * 128 + r) * 8 + k) * 32);
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 f4 = fr[g];
                            dot = fmaf(has ? dcv[4 * g] : 0.f, f4.x, dot); dot = fmaf(has ? dcv[4 * g + 1] : 0.f, f4.y, dot);
                            dot = fmaf(has ? dcv[4 * g + 2] : 0.f, f4.z, dot); dot = fmaf(has ? dcv[4 * g + 3] : 0.f, f4.w, dot);
                        }
                        if (inb) a.dwn_col[m * 8 + k] = dot;
                    }
                    worker_signal16(a_ready);
                    // z1 of this neighbour: first 16 columns requested before waiting for the MMAs, the rest during the first chunk
                    const long long off0 = ((tile * 8 + k) * 128 + 32 * q) * 128 + r;
                    float zc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) zc[j] = a.tsave[TL.z1T + off0 + j * 128];
                    // dh1 -> dz1 = dh1 * softplus'(z1) (store for dN1) -> A
                    mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        const int c0 = 32 * q + 16 * c;
                        float v[16], zn[16];
                        tc::tmem_ld16(lb + TQ + c0, v);
                        const long long off = off0 + (long long)(16 * c) * 128;
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) zn[j] = a.tsave[TL.z1T + off + (16 + j) * 128];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] *= sp_grad_fast(zc[j]);
                        if (a.want_wgrad) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) a.tbwd[BL.dz1T + off + j * 128] = v[j];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) tc::split_tf32(v[j], v[j], zc[j]);                      // zc <- lo
                        tc::tmem_st16(lb + TQ + c0, v);
                        tc::tmem_st16(lb + TR + c0, zc);
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) zc[j] = zn[j];
                        }
                    }
                    worker_signal16(a_ready);
                    // dx: columns [0,20) rel-pos embedding, [20,52) feature gradient of the pair
                    mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    const bool live = has && id >= 0 && inb;
                    if (q == 0) {                          // rel-pos embedding: d arg, d rel (-> -d pos), d Brel
                        float dx[24];
                        {
                            float t16[16], t8[8];
                            tc::tmem_ld16(lb + TDX, t16);
                            tc::tmem_ld8(lb + TDX + 16, t8);
#pragma unroll
                            for (int j = 0; j < 16; ++j) dx[j] = t16[j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) dx[16 + j] = t8[j];
                        }
                        float rx = 0.f, ry = 0.f, rz = 0.f;
                        if (id >= 0) {
                            rx = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3), px));
                            ry = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 1), py));
                            rz = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 2), pz));
                        }
                        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                        for (int jj = 0; jj < 10; ++jj) {
                            float da = 0.f;
                            if (live) {
                                float sn, cs;
                                sincos_embed(fmaf(rz, Br[24 + jj], fmaf(ry, Br[12 + jj], rx * Br[jj])), &sn, &cs);
                                da = dx[jj] * cs - dx[10 + jj] * sn;
                            }
                            gx = fmaf(da, Br[jj], gx); gy = fmaf(da, Br[12 + jj], gy); gz = fmaf(da, Br[24 + jj], gz);
                            if (a.want_wgrad) {            // d Brel[c][jj] += (2 pi rel_c) d arg: lane (c*10 + jj) keeps the sum
                                const float t0 = warp_sum(da * rx), t1 = warp_sum(da * ry), t2 = warp_sum(da * rz);
                                if (lane == jj) brel_acc += t0;
                                if (lane == 10 + jj) brel_acc += t1;
                                if (lane == 20 + jj) brel_acc += t2;
                            }
                        }
                        dpx -= kTwoPi * gx; dpy -= kTwoPi * gy; dpz -= kTwoPi * gz;
                    } else if (q == 1) {                   // feature gradient columns 0-11  = dx[20..31]
                        float t16[16];
                        tc::tmem_ld16(lb + TDX + 16, t16);
                        if (a.d_colpair && inb) {
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32);
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                dst[g] = live ? make_float4(t16[4 + 4 * g], t16[5 + 4 * g], t16[6 + 4 * g], t16[7 + 4 * g]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    } else if (q == 2) {                   // feature gradient columns 12-27 = dx[32..47]
                        float t16[16];
                        tc::tmem_ld16(lb + TDX + 32, t16);
                        if (a.d_colpair && inb) {
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32) + 3;
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                dst[g] = live ? make_float4(t16[4 * g], t16[4 * g + 1], t16[4 * g + 2], t16[4 * g + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    } else {                               // feature gradient columns 28-31 = dx[48..51]
                        float t8[8];
                        tc::tmem_ld8(lb + TDX + 48, t8);
                        if (a.d_colpair && inb) {
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32) + 7;
                            dst[0] = live ? make_float4(t8[0], t8[1], t8[2], t8[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
            if (q == 0 && inb && a.dpos_col) { a.dpos_col[m * 3] = dpx; a.dpos_col[m * 3 + 1] = dpy; a.dpos_col[m * 3 + 2] = dpz; }
        }
        if (q == 0 && lane < 30) sRed[(warp & 3) * 32 + lane] = brel_acc;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 30 && a.part_brel)
        a.part_brel[blockIdx.x * 32 + threadIdx.x] = sRed[threadIdx.x] + sRed[32 + threadIdx.x] + sRed[64 + threadIdx.x] + sRed[96 + threadIdx.x];
    if (warp == 17) tc::tmem_dealloc(tmem, 512);
}

}  // namespace cbt16
}  // namespace psl

using namespace psl;

// same contract as psl_color_bwd_tc (include/pointslam_b200.h); experiment build of the worker side, see the file header
extern "C" int psl_color_bwd_tc(const psl_decode_cfg* cfg, const float* bwd_blob, const float* pos, int64_t m,
                                    const int32_t* I, const float* D, const int32_t* nnum, const double* r2,
                                    const float* cloud_pos, const float* col_feats, const float* exposure_affine,
                                    const float* raw, const float* d_raw, const float* tsave, float* tbwd, float* d_colpair,
                                    float* wn_out, float* dwn_col, float* dpos_col, int32_t want_wgrad, int32_t* grid_out,
                                    psl_stream_t stream) {
    PSL_REQUIRE(cfg && bwd_blob && pos && I && D && nnum && col_feats && raw && d_raw && tsave && tbwd, "NULL argument");
    PSL_REQUIRE(!cfg->encode_rel_pos || cloud_pos, "rel-pos encoding needs cloud_pos");
    if (m == 0) return 0;
    cbt::Args a{};
    a.cfg = *cfg; a.blob = bwd_blob; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.col_feats = col_feats; a.affine = exposure_affine; a.raw = raw; a.d_raw = d_raw;
    a.tsave = tsave; a.tbwd = tbwd; a.d_colpair = d_colpair; a.wn_out = wn_out; a.dwn_col = dwn_col; a.dpos_col = dpos_col;
    a.part_brel = tbwd + tbwd_layout(m, cfg->encode_rel_pos).total;
    a.want_wgrad = want_wgrad;
    const long long n_tiles = (m + cbt::TM - 1) / cbt::TM;
    PSL_CHECK_CUDA(cudaFuncSetAttribute(cbt16::k_color_bwd_tc_w16, cudaFuncAttributeMaxDynamicSharedMemorySize, cbt::SB_TOTAL));
    const long long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (grid_out) *grid_out = (int32_t)grid;
    TimingScope ts(T_COLOR_BWD_TC, as_stream(stream));
    cbt16::k_color_bwd_tc_w16<<<(unsigned)grid, cbt16::NTHR16, cbt::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
