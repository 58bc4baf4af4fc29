// Backward DATA path of the colour branch on tcgen05 (3xTF32, gradients resident in TMEM as the A operand):
//   dL/d(rgb) -> output layer -> 5 trunk layers -> dL/dc, dL/d(embedding) -> per-neighbour MLP (8 slots) ->
//   dL/d col_feats[I] (per pair), dL/d(IDW weights), dL/d(pos) (Fourier + rel-pos parts), dL/d Brel.
// With the fc_c injection folded (see psl_color_tc.cu) every layer needs ONE A operand, dz_l = dh_l * softplus'(z_l):
//   dh_{l-1} = dz_l L_l[:, act]      dc += dz_l (L_l[:, act] Fc_{l-1})      de += dz_l L_l[:, emb]
// dh_l (and dz1 of the neighbour MLP) are also written to HBM tile-transposed for the weight-gradient GEMMs
// (psl_wgrad_tc.cu).  Same warp roles / barriers / weight ring as the forward kernel.
//
// Autograd semantics: src/conv_onet/models/decoder.py:341-449.
#include "psl_decode.cuh"
#include "psl_tc.cuh"
#include "psl_tc_layout.cuh"
#include "psl_color_bwd_tc.cuh"

namespace psl {
namespace cbt {

// ---------------------------------------------------------------------------------------------------------------------
// packing (uses the folded rows written by ctc::k_tc_fold: row n of layer l = [e | act | G] with G = L_act Fc_{l-1})
// ---------------------------------------------------------------------------------------------------------------------
struct PackArgs { psl_decoder_params P; const float* fold; int fold_ld; float* blob; };

__device__ __forceinline__ void put(float* blob, int q, int n, int k, float v) {
    // matrix q, row n (< mat_n), reduction index k (< mat_k): chunk of 32 k (or 16 for q < 2)
    const int N = mat_n(q), K = mat_k(q);
    const int kc = K < 32 ? K : 32;
    const int chunk = k / kc, kk = k - chunk * kc;
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    float* base = blob + BB_TRUNK + mat_off(q) + chunk * 2 * N * kc;
    const uint32_t o = tc::canon_off_floats(n, kk, N);
    base[o] = hi;
    base[N * kc + o] = lo;
}

__global__ void k_bwd_pack(PackArgs a) {
    const int job = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const float* F = a.fold;
    const int ld = a.fold_ld;
    if (job == 0) {                  // WoT [n'=128][o=16] and GoT [c=32][o=16] from fold layer 5 (rows o < 3 live)
        if (e < 128 * 16) { const int n = e >> 4, o = e & 15; put(a.blob, 0, n, o, o < 3 ? F[((size_t)5 * 128 + o) * ld + n] : 0.f); }
        if (e < 32 * 16) { const int c = e >> 4, o = e & 15; put(a.blob, 1, c, o, o < 3 ? F[((size_t)5 * 128 + o) * ld + 128 + c] : 0.f); }
    } else if (job <= 4) {           // l = 5 - job in {4,3,2,1}: LaT_l [k_in][n], GT_l [c][n], (l == 3) LeT_3 [j][n]
        const int l = 5 - job;
        const int qa = l == 4 ? 2 : (l == 3 ? 4 : (l == 2 ? 7 : 9));
        const int ne = (l == 3) ? 40 : 0;
        if (e < 128 * 128) { const int kin = e >> 7, n = e & 127; put(a.blob, qa, kin, n, F[((size_t)l * 128 + n) * ld + ne + kin]); }
        if (e < 32 * 128) { const int c = e >> 7, n = e & 127; put(a.blob, qa + 1, c, n, F[((size_t)l * 128 + n) * ld + ne + 128 + c]); }
        if (l == 3 && e < 48 * 128) { const int j = e >> 7, n = e & 127; put(a.blob, 6, j, n, j < 40 ? F[((size_t)3 * 128 + n) * ld + j] : 0.f); }
    } else if (job == 5) {           // LeT_0
        if (e < 48 * 128) { const int j = e >> 7, n = e & 127; put(a.blob, 11, j, n, j < 40 ? F[(size_t)n * ld + j] : 0.f); }
    } else if (job == 6) {           // N2T [hid 128][c 32] = N2[c][hid] ; N1T [j 64][hid 128] = N1[hid][j]
        if (e < 128 * 32) {
            const int hid = e >> 5, c = e & 31;
            float hi, lo;
            tc::split_tf32(a.P.c_N2[c * 128 + hid], hi, lo);
            const uint32_t o = tc::canon_off_floats(hid, c, 128);
            a.blob[BB_N2T + o] = hi; a.blob[BB_N2T + 128 * 32 + o] = lo;
        }
        if (e < 64 * 128) {
            const int j = e >> 7, hid = e & 127;
            float hi, lo;
            tc::split_tf32(j < 52 ? a.P.c_N1[hid * 52 + j] : 0.f, hi, lo);
            const uint32_t o = tc::canon_off_floats(j, hid, 64);
            a.blob[BB_N1T + o] = hi; a.blob[BB_N1T + 64 * 128 + o] = lo;
        }
    } else {
        if (e < 30) a.blob[BB_VEC + BV_BREL + (e / 10) * 12 + (e % 10)] = a.P.c_Brel[e];
        if (e < 60) a.blob[BB_VEC + BV_BC + e] = a.P.c_B[e];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void worker_signal(uint64_t* a_ready) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    tc::mbar_arrive(a_ready);
}

__global__ void __launch_bounds__(NTHR, 1) k_color_bwd_tc(Args a, long long n_tiles) {
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
        tc::mbar_init(nbrw_full, 1); tc::mbar_init(a_ready, NWORK); tc::mbar_init(d_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < BV_SIZE; i += NTHR) sVec[i] = a.blob[BB_VEC + i];
    if (threadIdx.x < 12) sAff[threadIdx.x] = a.affine ? a.affine[threadIdx.x] : 0.f;
    if (threadIdx.x < 128) sRed[threadIdx.x] = 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
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
                        tc::mbar_wait(&empty[st], ((cnt >> 1) & 1) ^ 1);
                        tc::mbar_expect_tx(&full[st], bytes);
                        tc::bulk_g2s(smem + SB_RING + st * 32768, a.blob + BB_TRUNK + mat_off(q) + c * 2 * N * kc, bytes, &full[st]);
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            uint32_t pa = 0, cnt = 0;
            const uint32_t n2t = tc::smem_u32(smem + SB_NBRW), n1t = n2t + 2 * 128 * 32 * 4;
            // one streamed matrix: A = (a_hi, a_lo) columns, K reduction, D columns d, N rows of B
            auto run_mat = [&](int q, uint32_t a_hi, uint32_t a_lo, uint32_t d, uint32_t first_acc) {
                const int N = mat_n(q), K = mat_k(q), kc = K < 32 ? K : 32;
                const uint32_t idesc = tc::make_idesc_tf32(128, N), lbo = (uint32_t)N * 16u;
                for (int c = 0; c * kc < K; ++c, ++cnt) {
                    const int st = cnt & 1;
                    tc::mbar_wait(&full[st], (cnt >> 1) & 1);
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
            if (rel) tc::mbar_wait(nbrw_full, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                // output layer: A = dout (P[0:16] / R[0:16])
                tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(0, TP, TR, TQ, 0);               // dh_4 -> Q
                run_mat(1, TP, TR, TDC, 0);              // dc    = dout G_out
                tc::mma_commit(d_ready);
                int q = 2;
                for (int l = 4; l >= 1; --l) {
                    tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                    const uint32_t cur = (l & 1) ? TP : TQ, oth = (l & 1) ? TQ : TP;     // dz_l lives in cur, dh_{l-1} goes to oth
                    run_mat(q++, cur, TR, oth, 0);
                    run_mat(q++, cur, TR, TDC, 1);
                    if (l == 3) run_mat(q++, cur, TR, TDE, 0);
                    tc::mma_commit(d_ready);
                }
                tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(11, TQ, TR, TDE, 1);             // de += dz_0 L_0
                tc::mma_commit(d_ready);
                if (rel) {
                    const uint32_t id128 = tc::make_idesc_tf32(128, 128), id64 = tc::make_idesc_tf32(128, 64);
                    for (int k = 0; k < 8; ++k) {
                        tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 4; ++j) {            // dh1 = df N2   (A: P[0:32]/R[0:32], K = 32) -> Q
                            const uint64_t bh = tc::make_smem_desc(n2t + j * 2 * 2048, 2048, 128);
                            const uint64_t bl = tc::make_smem_desc(n2t + 128 * 32 * 4 + j * 2 * 2048, 2048, 128);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TP + 8 * j, bh, id128, j > 0);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TR + 8 * j, bh, id128, 1);
                            tc::mma_tf32_ts(tmem + TQ, tmem + TP + 8 * j, bl, id128, 1);
                        }
                        tc::mma_commit(d_ready);
                        tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
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
        // =============================== workers ======================================================================
        const int r = 32 * (warp & 3) + lane, h = warp >> 2;
        const uint32_t lb = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t pd = 0;
        const float* Br = sVec + BV_BREL; const float* Bc = sVec + BV_BC;
        float brel_acc = 0.f;                                    // lane e < 30 of half-0 warps: d Brel[e/10][e%10]
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
                if (h == 0 && inb && a.wn_out) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) a.wn_out[m * 8 + k] = (has && idx[k] >= 0) ? wn[k] : 0.f;
                }
            }
            float dpx = 0.f, dpy = 0.f, dpz = 0.f;
            // ---- dL/d(colour output) -> A operand (16 columns) --------------------------------------------------------
            if (h == 0) {
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
            worker_signal(a_ready);
            // ---- trunk: dh_l -> (store) -> dz_l = dh_l * softplus'(z_l) -> A operand ------------------------------------
#pragma unroll 1
            for (int l = 4; l >= 0; --l) {
                // the saved pre-activations of the first 32 columns are requested BEFORE waiting for the layer's MMAs, those of the
                // second 32 while the first half is being processed: their latency is off the critical path
                const long long off0 = (((long long)l * n_tiles + tile) * 128 + 64 * h) * 128 + r;
                float zc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) zc[j] = a.tsave[TL.zT + off0 + j * 128];
                tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                const uint32_t reg = (l & 1) ? TP : TQ;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int c0 = 64 * h + 32 * c;
                    float v[32], zn[32];
                    tc::tmem_ld32(lb + reg + c0, v);
                    const long long off = off0 + (long long)(32 * c) * 128;
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) zn[j] = a.tsave[TL.zT + off + (32 + j) * 128];
                    }
                    if (a.want_wgrad) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) a.tbwd[BL.dhT + off + j * 128] = v[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) tc::split_tf32(v[j] * sp_grad_fast(zc[j]), v[j], zc[j]);   // zc <- lo
                    tc::tmem_st32(lb + reg + c0, v);
                    tc::tmem_st32(lb + TR + c0, zc);
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) zc[j] = zn[j];
                    }
                }
                worker_signal(a_ready);
            }
            // ---- dc, de -----------------------------------------------------------------------------------------------------
            tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
            float dcv[32];
            tc::tmem_ld32(lb + TDC, dcv);
#pragma unroll
            for (int j = 0; j < 32; ++j) dcv[j] = has ? dcv[j] : 0.f;
            if (h == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a.tbwd[BL.dccT + (tile * 32 + j) * 128 + r] = dcv[j];
                if (!rel && inb && a.d_colpair) {
                    float4* dst = reinterpret_cast<float4*>(a.d_colpair + m * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = make_float4(dcv[4 * q], dcv[4 * q + 1], dcv[4 * q + 2], dcv[4 * q + 3]);
                }
                if (a.dpos_col) {                          // colour Fourier embedding: d arg_j = dsin_j cos - dcos_j sin
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
            }
            // ---- IDW gradient of the plain (no neighbour MLP) interpolation, or the neighbour MLP backward --------------------
            if (!rel) {
                if (h == 0 && inb && a.dwn_col) {
#pragma unroll 1
                    for (int k = 0; k < 8; ++k) {
                        float dot = 0.f;
                        if (idx[k] >= 0) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)idx[k] * 32) + q);
                                dot = fmaf(dcv[4 * q], f4.x, dot); dot = fmaf(dcv[4 * q + 1], f4.y, dot);
                                dot = fmaf(dcv[4 * q + 2], f4.z, dot); dot = fmaf(dcv[4 * q + 3], f4.w, dot);
                            }
                        }
                        a.dwn_col[m * 8 + k] = dot;
                    }
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    const int id = idx[k];
                    // df_k = wn_k dc -> A (32 columns); d wn_k = dc . f_k
                    if (h == 0) {
                        float v[32], lo[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) tc::split_tf32(wn[k] * dcv[j], v[j], lo[j]);
                        tc::tmem_st32(lb + TP, v);
                        tc::tmem_st32(lb + TR, lo);
                        if (a.dwn_col) {
                            float dot = 0.f;
                            const float4* fr = reinterpret_cast<const float4*>(a.tsave + TL.f + ((tile * 128 + r) * 8 + k) * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 f4 = fr[q];
                                dot = fmaf(dcv[4 * q], f4.x, dot); dot = fmaf(dcv[4 * q + 1], f4.y, dot);
                                dot = fmaf(dcv[4 * q + 2], f4.z, dot); dot = fmaf(dcv[4 * q + 3], f4.w, dot);
                            }
                            if (inb) a.dwn_col[m * 8 + k] = dot;
                        }
                    }
                    worker_signal(a_ready);
                    // z1 of this neighbour: first 32 columns requested before waiting for the MMAs, the rest during the first half
                    const long long off0 = ((tile * 8 + k) * 128 + 64 * h) * 128 + r;
                    float zc[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) zc[j] = a.tsave[TL.z1T + off0 + j * 128];
                    // dh1 -> dz1 = dh1 * softplus'(z1) (store for dN1) -> A
                    tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        const int c0 = 64 * h + 32 * c;
                        float v[32], zn[32];
                        tc::tmem_ld32(lb + TQ + c0, v);
                        const long long off = off0 + (long long)(32 * c) * 128;
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) zn[j] = a.tsave[TL.z1T + off + (32 + j) * 128];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= sp_grad_fast(zc[j]);
                        if (a.want_wgrad) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) a.tbwd[BL.dz1T + off + j * 128] = v[j];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) tc::split_tf32(v[j], v[j], zc[j]);                      // zc <- lo
                        tc::tmem_st32(lb + TQ + c0, v);
                        tc::tmem_st32(lb + TR + c0, zc);
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) zc[j] = zn[j];
                        }
                    }
                    worker_signal(a_ready);
                    // dx: columns [0,20) rel-pos embedding, [20,52) feature  (half 0: 0..31, half 1: 32..63)
                    tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    float dx[32];
                    tc::tmem_ld32(lb + TDX + 32 * h, dx);
                    const bool live = has && id >= 0 && inb;
                    if (a.d_colpair && inb) {
                        float* dst = a.d_colpair + ((size_t)m * 8 + k) * 32;
                        if (h == 0) {
#pragma unroll
                            for (int q = 0; q < 3; ++q)
                                reinterpret_cast<float4*>(dst)[q] = live ? make_float4(dx[20 + 4 * q], dx[21 + 4 * q], dx[22 + 4 * q], dx[23 + 4 * q])
                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
#pragma unroll
                            for (int q = 0; q < 5; ++q)
                                reinterpret_cast<float4*>(dst)[3 + q] = live ? make_float4(dx[4 * q], dx[4 * q + 1], dx[4 * q + 2], dx[4 * q + 3])
                                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    if (h == 0) {                          // rel-pos embedding: d arg, d rel (-> -d pos), d Brel
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
                            // d Brel[c][jj] += (2 pi rel_c) d arg : warp reduction, lane (c*10 + jj) keeps the sum.  Only the
                            // weight-gradient pass reads it (k_wgrad_finalize): skipped for the tracker (240 reductions per tile)
                            if (a.want_wgrad) {
                                const float t0 = warp_sum(da * rx), t1 = warp_sum(da * ry), t2 = warp_sum(da * rz);
                                if (lane == jj) brel_acc += t0;
                                if (lane == 10 + jj) brel_acc += t1;
                                if (lane == 20 + jj) brel_acc += t2;
                            }
                        }
                        dpx -= kTwoPi * gx; dpy -= kTwoPi * gy; dpz -= kTwoPi * gz;
                    }
                }
            }
            if (h == 0 && inb && a.dpos_col) { a.dpos_col[m * 3] = dpx; a.dpos_col[m * 3 + 1] = dpy; a.dpos_col[m * 3 + 2] = dpz; }
        }
        if (h == 0 && lane < 30) sRed[(warp & 3) * 32 + lane] = brel_acc;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 30 && a.part_brel)
        a.part_brel[blockIdx.x * 32 + threadIdx.x] = sRed[threadIdx.x] + sRed[32 + threadIdx.x] + sRed[64 + threadIdx.x] + sRed[96 + threadIdx.x];
    if (warp == 9) tc::tmem_dealloc(tmem, 512);
}

}  // namespace cbt
}  // namespace psl

using namespace psl;

extern "C" size_t psl_tc_bwd_blob_floats(void) { return (size_t)cbt::BB_TOTAL; }
extern "C" size_t psl_tc_save_floats(int64_t m, int32_t encode_rel_pos) { return (size_t)tsave_layout(m, encode_rel_pos).total; }
extern "C" size_t psl_tc_bwd_tmp_floats(int64_t m, int32_t encode_rel_pos) {
    return (size_t)tbwd_layout(m, encode_rel_pos).total + 32 * 148 + 64;
}

// lay out the transposed / folded weights of the colour branch for the backward kernel.  `tc_blob` must hold the result of
// psl_tc_pack_params for the SAME parameters (its folded fp32 rows are reused).
extern "C" int psl_tc_bwd_pack_params(const psl_decoder_params* P, const float* tc_blob, size_t tc_fold_offset_floats,
                                      float* bwd_blob, psl_stream_t stream) {
    PSL_REQUIRE(P && tc_blob && bwd_blob, "NULL argument");
    cudaStream_t st = as_stream(stream);
    cbt::PackArgs pa;
    pa.P = *P; pa.fold = tc_blob + tc_fold_offset_floats; pa.fold_ld = 200; pa.blob = bwd_blob;
    TimingScope ts(T_PACK, st, 1);
    cbt::k_bwd_pack<<<dim3(64, 8), 256, 0, st>>>(pa);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// data-gradient pass of the colour branch.  Outputs: d_colpair (m,8,32) [rel] or (m,32) [no neighbour MLP], wn_out (m,8),
// dwn_col (m,8), dpos_col (m,3) (NULL = not wanted), tbwd (psl_tc_bwd_tmp_floats; dhT / dz1T / doutT / dccT / affine terms
// for psl_wgrad_tc when want_wgrad), and the per-CTA partial of dL/dBrel at tbwd + tbwd_layout.total.
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
    static bool attr_set = false;
    if (!attr_set) {
        PSL_CHECK_CUDA(cudaFuncSetAttribute(cbt::k_color_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, cbt::SB_TOTAL));
        attr_set = true;
    }
    const long long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (grid_out) *grid_out = (int32_t)grid;
    TimingScope ts(T_COLOR_BWD_TC, as_stream(stream));
    cbt::k_color_bwd_tc<<<(unsigned)grid, cbt::NTHR, cbt::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
