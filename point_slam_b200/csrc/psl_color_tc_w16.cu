// 3xTF32 colour forward (the A/B baseline of the f16-plane kernel psl_color_h2.cu, selected with PSL_H2=0, and the producer of the
// FFMA-layout activations when the tensor-core backward is switched off): SIXTEEN worker warps -- four threads per sample row, thread (row r, quarter q) owns columns 32q..32q+31
// of every 128-wide layer and walks them as two 16-column chunks (tcgen05.ld.x16 -> bias -> save -> softplus -> hi/lo split ->
// 2 x tcgen05.st.x16), so that the epilogues fit the 112-register cap of an 18-warp CTA.  Same operand blob, TMEM regions,
// shared-memory map, bulk-copy producer and MMA issue order as the production kernel (psl_color_tc.cuh); only the worker side
// and the warp numbering differ.  Motivation and plan: DESIGN.md section 7 item 1, profiles/r01g_summary.md (8 worker warps
// keep the issue slots 22-27 % busy).
//
// Verified bit-identical to the round-1 8-worker-warp kernel on hardware (profiles/r02_s1_*) before that kernel was removed.
#include "psl_color_tc.cuh"

namespace psl {
namespace ctc16 {

using namespace ctc;

constexpr int NWORK16 = 512, NTHR16 = 576;      // warps 0-15 workers, 16 bulk-copy producer, 17 TMEM allocator + MMA issuer

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) { tc::mbar_wait_p(bar, parity); }
__device__ __forceinline__ void worker_signal16(uint64_t* a_ready) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncwarp();                                    // one arrival per warp: 512 per-thread arrivals on one word serialise
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(a_ready);
}

template <int SAVE>
__global__ void __launch_bounds__(NTHR16, 1) k_color_fwd_tc_w16(Args a, long long n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sVec = reinterpret_cast<float*>(smem + SB_VEC);
    float* sRand = reinterpret_cast<float*>(smem + SB_RAND);
    float* sEhi = reinterpret_cast<float*>(smem + SB_E);
    float* sElo = sEhi + 128 * 40;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SB_BAR);
    uint64_t* full = bars;              // [2]
    uint64_t* empty = bars + 2;         // [2]
    uint64_t* nbrw_full = bars + 4;
    uint64_t* a_ready = bars + 5;       // workers -> MMA (count 512)
    uint64_t* d_ready = bars + 6;       // MMA -> workers (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool rel = a.cfg.encode_rel_pos != 0;
    const SaveLayout SL = save_layout(1, a.cfg.encode_rel_pos);
    const long long M = a.m;
    const TSave TL = tsave_layout(a.m, a.cfg.encode_rel_pos);

    if (threadIdx.x == 0) {
        tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
        tc::mbar_init(&empty[0], 1); tc::mbar_init(&empty[1], 1);
        tc::mbar_init(nbrw_full, 1);
        tc::mbar_init(a_ready, NWORK16 / 32);
        tc::mbar_init(d_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < V_SIZE; i += NTHR16) sVec[i] = a.blob[TB_VEC + i];
    if (threadIdx.x < 32) sRand[threadIdx.x] = a.rand_col[threadIdx.x];
    if (threadIdx.x < 12) sRand[32 + threadIdx.x] = a.affine ? a.affine[threadIdx.x] : 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // =============================== bulk-copy producer (as in psl_color_tc.cu) ==================================
        if (lane == 0) {
            if (rel) {
                tc::mbar_expect_tx(nbrw_full, 98304);
                for (int i = 0; i < 3; ++i)
                    tc::bulk_g2s(smem + SB_NBRW + i * 32768, a.blob + TB_N1 + i * 8192, 32768, nbrw_full);
            }
            uint32_t cnt = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < NLAYER; ++l) {
                    const int N = l_n(l), ks = l_ks(l);
                    for (int c = 0; c * 4 < ks; ++c, ++cnt) {
                        const int st = cnt & 1;
                        const uint32_t bytes = (uint32_t)(min(4, ks - 4 * c) * 16 * N * 4);
                        mbar_wait_wd(&empty[st], ((cnt >> 1) & 1) ^ 1);
                        tc::mbar_expect_tx(&full[st], bytes);
                        tc::bulk_g2s(smem + SB_RING + st * 32768, a.blob + TB_TRUNK + l_off(l) + c * 4 * 16 * N, bytes, &full[st]);
                    }
                }
            }
        }
    } else if (warp == 17) {
        // =============================== MMA issuer (as in psl_color_tc.cu) ==========================================
        if (lane == 0) {
            uint32_t pa = 0, cnt = 0;
            const uint32_t n1 = tc::smem_u32(smem + SB_NBRW), n2 = n1 + 2 * 128 * 64 * 4;
            const uint32_t ehi = tc::smem_u32(sEhi), elo = tc::smem_u32(sElo);
            const uint32_t id128 = tc::make_idesc_tf32(128, 128), id32 = tc::make_idesc_tf32(128, 32), id16 = tc::make_idesc_tf32(128, 16);
            if (rel) mbar_wait_wd(nbrw_full, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (rel) {
                    for (int k = 0; k < 8; ++k) {
                        mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 7; ++j) {            // z1 = x N1^T        (K = 56)
                            const uint64_t bh = tc::make_smem_desc(n1 + j * 2 * 2048, 2048, 128);
                            const uint64_t bl = tc::make_smem_desc(n1 + 128 * 64 * 4 + j * 2 * 2048, 2048, 128);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 8 * j, bh, id128, j > 0);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 64 + 8 * j, bh, id128, 1);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 8 * j, bl, id128, 1);
                        }
                        tc::mma_commit(d_ready);
                        mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 16; ++j) {           // f = softplus(z1) N2^T   (K = 128, N = 32)
                            const uint64_t bh = tc::make_smem_desc(n2 + j * 2 * 512, 512, 128);
                            const uint64_t bl = tc::make_smem_desc(n2 + 32 * 128 * 4 + j * 2 * 512, 512, 128);
                            tc::mma_tf32_ts(tmem + TSP, tmem + TP + 8 * j, bh, id32, j > 0);
                            tc::mma_tf32_ts(tmem + TSP, tmem + TQ + 8 * j, bh, id32, 1);
                            tc::mma_tf32_ts(tmem + TSP, tmem + TP + 8 * j, bl, id32, 1);
                        }
                        tc::mma_commit(d_ready);
                    }
                }
                for (int l = 0; l < NLAYER; ++l) {
                    mbar_wait_wd(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                    const int N = l_n(l), ks = l_ks(l), ne = l_ne(l), na = l_na(l);
                    const uint32_t idesc = l == 5 ? id16 : id128;
                    const uint32_t dcol = (l & 1) ? TQ : TP;
                    const uint32_t acol = (l & 1) ? TP : TQ;
                    const uint32_t lbo = (uint32_t)N * 16u;
                    for (int j = 0; j < ks; ++j) {
                        const int cpos = j & 3;
                        const int st = cnt & 1;
                        if (cpos == 0) mbar_wait_wd(&full[st], (cnt >> 1) & 1);
                        const int ksc = min(4, ks - (j - cpos));
                        const uint32_t rb = tc::smem_u32(smem + SB_RING + st * 32768);
                        const uint64_t bh = tc::make_smem_desc(rb + cpos * 2 * lbo, lbo, 128);
                        const uint64_t bl = tc::make_smem_desc(rb + (uint32_t)N * 8 * ksc * 4 + cpos * 2 * lbo, lbo, 128);
                        const uint32_t acc = j > 0;
                        if (j < ne) {
                            const uint64_t ah = tc::make_smem_desc(ehi + j * 2 * 2048, 2048, 128);
                            const uint64_t al = tc::make_smem_desc(elo + j * 2 * 2048, 2048, 128);
                            tc::mma_tf32_ss(tmem + dcol, ah, bh, idesc, acc);
                            tc::mma_tf32_ss(tmem + dcol, al, bh, idesc, 1);
                            tc::mma_tf32_ss(tmem + dcol, ah, bl, idesc, 1);
                        } else {
                            uint32_t ahc, alc;
                            if (j < ne + na) { ahc = acol + 8 * (j - ne); alc = TR + 8 * (j - ne); }
                            else { ahc = TCC + 8 * (j - ne - na); alc = TCC + 32 + 8 * (j - ne - na); }
                            tc::mma_tf32_ts(tmem + dcol, tmem + ahc, bh, idesc, acc);
                            tc::mma_tf32_ts(tmem + dcol, tmem + alc, bh, idesc, 1);
                            tc::mma_tf32_ts(tmem + dcol, tmem + ahc, bl, idesc, 1);
                        }
                        if (cpos == 3 || j == ks - 1) { tc::mma_commit(&empty[st]); ++cnt; }
                    }
                    tc::mma_commit(d_ready);
                }
            }
        }
    } else {
        // =============================== workers: 4 threads per sample row =============================================
        const int r = 32 * (warp & 3) + lane, q = warp >> 2;                    // q = column quarter (0..3)
        const uint32_t lb = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t pd = 0;
        const float* b1 = sVec + V_B1; const float* b2 = sVec + V_B2; const float* Bc = sVec + V_BC; const float* Br = sVec + V_BREL;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long m = tile * TM + r;
            const bool inb = m < a.m;
            float px = 0.f, py = 0.f, pz = 0.f, wn[8];
            int idx[8];
            bool has = false;
            {
                float sum = 0.f, w[8];
                float tle = -1.f;
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
            }
            float cacc[8];                                      // channels 8q .. 8q+7 of the interpolated feature
#pragma unroll
            for (int j = 0; j < 8; ++j) cacc[j] = 0.f;
            if (rel) {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    // ---- x_k = [sin(10) | cos(10) | col_feats[I_k](32) | 0(12)], this thread's columns 16q .. 16q+15, hi/lo -> R
                    //      quarter 0: sin 0-9, cos 0-5;  1: cos 6-9, feat 0-11;  2: feat 12-27;  3: feat 28-31, zeros
                    float xv[16], lo[16];
                    const int id = idx[k];
#pragma unroll
                    for (int j = 0; j < 16; ++j) xv[j] = 0.f;
                    if (q < 2) {
                        float rx = 0.f, ry = 0.f, rz = 0.f;
                        if (id >= 0) {
                            rx = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3), px));
                            ry = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 1), py));
                            rz = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 2), pz));
                        }
                        if (q == 0) {
#pragma unroll
                            for (int jj = 0; jj < 10; ++jj) {
                                float sn = 0.f, cs = 0.f;
                                if (id >= 0) sincos_embed(fmaf(rz, Br[24 + jj], fmaf(ry, Br[12 + jj], rx * Br[jj])), &sn, &cs);
                                xv[jj] = sn;
                                if (jj < 6) xv[10 + jj] = cs;
                            }
                        } else {
#pragma unroll
                            for (int jj = 6; jj < 10; ++jj) {
                                float sn = 0.f, cs = 0.f;
                                if (id >= 0) sincos_embed(fmaf(rz, Br[24 + jj], fmaf(ry, Br[12 + jj], rx * Br[jj])), &sn, &cs);
                                xv[jj - 6] = cs;
                            }
#pragma unroll
                            for (int g = 0; g < 3; ++g) {
                                float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (id >= 0) f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + g);
                                xv[4 + 4 * g] = f4.x; xv[5 + 4 * g] = f4.y; xv[6 + 4 * g] = f4.z; xv[7 + 4 * g] = f4.w;
                            }
                        }
                    } else if (q == 2) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (id >= 0) f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + 3 + g);
                            xv[4 * g] = f4.x; xv[4 * g + 1] = f4.y; xv[4 * g + 2] = f4.z; xv[4 * g + 3] = f4.w;
                        }
                    } else {
                        float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (id >= 0) f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + 7);
                        xv[0] = f4.x; xv[1] = f4.y; xv[2] = f4.z; xv[3] = f4.w;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) tc::split_tf32(xv[j], xv[j], lo[j]);
                    tc::tmem_st16(lb + TR + 16 * q, xv);
                    tc::tmem_st16(lb + TR + 64 + 16 * q, lo);
                    worker_signal16(a_ready);
                    if (k < 7) {                           // next neighbour's feature row / position: into L1 while the MMAs run
                        const int idn = idx[k + 1];
                        if (idn >= 0) {
                            if (q > 0) tc::prefetch_l1(a.col_feats + (size_t)idn * 32);
                            if (q < 2) tc::prefetch_l1(a.cloud_pos + (size_t)idn * 3);
                        }
                    }
                    // ---- z1 + b1 -> softplus -> hi (in place, P) / lo (Q): columns 32q .. 32q+31, two chunks of 16
                    mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        const int c0 = 32 * q + 16 * c;
                        tc::tmem_ld16(lb + TP + c0, xv);
#pragma unroll
                        for (int j = 0; j < 16; ++j) xv[j] += b1[c0 + j];
                        if (SAVE == 1 && inb) {
                            float4* dst = reinterpret_cast<float4*>(a.save + SL.nz1 * M + (m * 8 + k) * 128 + c0);
#pragma unroll
                            for (int g = 0; g < 4; ++g) dst[g] = make_float4(xv[4 * g], xv[4 * g + 1], xv[4 * g + 2], xv[4 * g + 3]);
                        }
                        if (SAVE == 2) {
                            float* dst = a.tsave + TL.z1T + ((tile * 8 + k) * 128 + c0) * 128 + r;
#pragma unroll
                            for (int j = 0; j < 16; ++j) dst[j * 128] = xv[j];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) tc::split_tf32(softplus100_fast(xv[j]), xv[j], lo[j]);
                        tc::tmem_st16(lb + TP + c0, xv);
                        tc::tmem_st16(lb + TQ + c0, lo);
                    }
                    worker_signal16(a_ready);
                    // ---- f = D2 + b2 ; c += wn_k f      (this thread: channels 8q .. 8q+7)
                    mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    float f[8];
                    tc::tmem_ld8(lb + TSP + 8 * q, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += b2[8 * q + j];
                    if ((SAVE == 1 && inb) || SAVE == 2) {
                        float4* dst = SAVE == 1 ? reinterpret_cast<float4*>(a.save + SL.nf * M + (m * 8 + k) * 32 + 8 * q)
                                                : reinterpret_cast<float4*>(a.tsave + TL.f + ((tile * 128 + r) * 8 + k) * 32 + 8 * q);
                        dst[0] = make_float4(f[0], f[1], f[2], f[3]);
                        dst[1] = make_float4(f[4], f[5], f[6], f[7]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) cacc[j] = fmaf(wn[k], f[j], cacc[j]);
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    if (idx[k] < 0) continue;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)idx[k] * 32 + 8 * q) + g);
                        cacc[4 * g] = fmaf(wn[k], f4.x, cacc[4 * g]); cacc[4 * g + 1] = fmaf(wn[k], f4.y, cacc[4 * g + 1]);
                        cacc[4 * g + 2] = fmaf(wn[k], f4.z, cacc[4 * g + 2]); cacc[4 * g + 3] = fmaf(wn[k], f4.w, cacc[4 * g + 3]);
                    }
                }
            }
            {   // ---- c (hi/lo) -> TMEM region C ; colour embedding (hi/lo) -> shared memory A operand
                float chi[8], clo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) cacc[j] = has ? cacc[j] : sRand[8 * q + j];
                if (SAVE == 1 && inb) {
                    float4* dst = reinterpret_cast<float4*>(a.save + SL.cc * M + m * 32 + 8 * q);
                    dst[0] = make_float4(cacc[0], cacc[1], cacc[2], cacc[3]);
                    dst[1] = make_float4(cacc[4], cacc[5], cacc[6], cacc[7]);
                }
                if (SAVE == 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a.tsave[TL.cT + (tile * 32 + 8 * q + j) * 128 + r] = cacc[j];
                    if (q == 0) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) a.tsave[TL.wnT + (tile * 8 + k) * 128 + r] = wn[k];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) tc::split_tf32(cacc[j], chi[j], clo[j]);
                tc::tmem_st8(lb + TCC + 8 * q, chi);
                tc::tmem_st8(lb + TCC + 32 + 8 * q, clo);
                // embedding columns: [sin 0-19 | cos 0-19]; quarter 0: sin 0-9, 1: sin 10-19, 2: cos 0-9, 3: cos 10-19
                const float x = __fmul_rn(kTwoPi, px), y = __fmul_rn(kTwoPi, py), z = __fmul_rn(kTwoPi, pz);
                const int j0 = 10 * (q & 1), col0 = 20 * (q >> 1) + j0;
#pragma unroll 2
                for (int j = 0; j < 10; ++j) {
                    const float arg = fmaf(z, Bc[40 + j0 + j], fmaf(y, Bc[20 + j0 + j], x * Bc[j0 + j]));
                    float ehi, elo;
                    tc::split_tf32(q < 2 ? sin_embed(arg) : cos_embed(arg), ehi, elo);
                    const uint32_t o = tc::canon_off_floats(r, col0 + j, 128);
                    sEhi[o] = ehi; sElo[o] = elo;
                }
                tc::fence_proxy_async();
                worker_signal16(a_ready);
            }
            // ---- trunk epilogues: z + b' -> softplus -> hi in place / lo in R: columns 32q .. 32q+31, two chunks of 16
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                const uint32_t dcol = (l & 1) ? TQ : TP;
                const float* bias = sVec + V_BIAS + 128 * l;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int c0 = 32 * q + 16 * c;
                    float v[16], lo[16];
                    tc::tmem_ld16(lb + dcol + c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += bias[c0 + j];
                    if (SAVE == 1 && inb) {
                        float4* dst = reinterpret_cast<float4*>(a.save + SL.cz * M + ((long long)l * M + m) * 128 + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) dst[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    }
                    if (SAVE == 2) {
                        float* dst = a.tsave + TL.zT + (((long long)l * n_tiles + tile) * 128 + c0) * 128 + r;
#pragma unroll
                        for (int j = 0; j < 16; ++j) dst[j * 128] = v[j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) tc::split_tf32(softplus100_fast(v[j]), v[j], lo[j]);
                    tc::tmem_st16(lb + dcol + c0, v);
                    tc::tmem_st16(lb + TR + c0, lo);
                }
                worker_signal16(a_ready);
            }
            // ---- output layer (quarter 0 writes the pixel)
            mbar_wait_wd(d_ready, pd); pd ^= 1; tc::fence_after_sync();
            if (q == 0) {
                float o[8];
                tc::tmem_ld8(lb + TQ, o);
                if (inb) {
                    float cr = o[0] + sVec[V_BOUT], cg = o[1] + sVec[V_BOUT + 1], cb = o[2] + sVec[V_BOUT + 2];
                    if (SAVE == 2) *reinterpret_cast<float4*>(a.tsave + TL.outpre + (tile * 128 + r) * 4) = make_float4(cr, cg, cb, 0.f);
                    if (a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                        const float* A = sRand + 32;
                        const float r2 = fmaf(cb, A[6], fmaf(cg, A[3], cr * A[0])) + A[9];
                        const float g2 = fmaf(cb, A[7], fmaf(cg, A[4], cr * A[1])) + A[10];
                        const float b2_ = fmaf(cb, A[8], fmaf(cg, A[5], cr * A[2])) + A[11];
                        cr = r2; cg = g2; cb = b2_;
                    }
                    if (a.cfg.rgb_mode != PSL_RGB_RAW) { cr = sigmoidf_(cr); cg = sigmoidf_(cg); cb = sigmoidf_(cb); }
                    a.raw[m * 4] = cr; a.raw[m * 4 + 1] = cg; a.raw[m * 4 + 2] = cb;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 17) tc::tmem_dealloc(tmem, 512);
}

}  // namespace ctc16
}  // namespace psl

using namespace psl;

// same contract as psl_color_fwd_tc (include/pointslam_b200.h); experiment build of the worker side, see the file header
extern "C" int psl_color_fwd_tc(const psl_decode_cfg* cfg, const float* tc_blob, const float* pos, int64_t m,
                                    const int32_t* I, const float* D, const int32_t* nnum, const double* r2,
                                    const float* cloud_pos, const float* col_feats, const float* rand_col,
                                    const float* exposure_affine, float* raw, float* save, float* tsave, psl_stream_t stream) {
    PSL_REQUIRE(cfg && tc_blob && pos && I && D && nnum && col_feats && rand_col && raw, "NULL argument");
    PSL_REQUIRE(!cfg->encode_rel_pos || cloud_pos, "rel-pos encoding needs cloud_pos");
    PSL_REQUIRE(cfg->rgb_mode != PSL_RGB_AFFINE_SIGMOID || exposure_affine, "affine mode needs exposure_affine");
    PSL_REQUIRE(!(save && tsave), "pass at most one of save / tsave");
    if (m == 0) return 0;
    ctc::Args a{};
    a.cfg = *cfg; a.blob = tc_blob; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.col_feats = col_feats; a.rand_col = rand_col; a.affine = exposure_affine; a.raw = raw; a.save = save; a.tsave = tsave;
    const long long n_tiles = (m + ctc::TM - 1) / ctc::TM;
    PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc16::k_color_fwd_tc_w16<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
    PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc16::k_color_fwd_tc_w16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
    PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc16::k_color_fwd_tc_w16<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
    const long long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    TimingScope ts(T_COLOR_FWD_TC, as_stream(stream));
    if (tsave) ctc16::k_color_fwd_tc_w16<2><<<(unsigned)grid, ctc16::NTHR16, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    else if (save) ctc16::k_color_fwd_tc_w16<1><<<(unsigned)grid, ctc16::NTHR16, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    else ctc16::k_color_fwd_tc_w16<0><<<(unsigned)grid, ctc16::NTHR16, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
