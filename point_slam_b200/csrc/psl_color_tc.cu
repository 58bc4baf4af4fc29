// Colour branch of the decoder on the 5th-generation tensor cores (inference forward, round 1):
//   per-neighbour MLP (52 -> 128 -> 32, 8 neighbour slots) + colour trunk (5 x 128, skip-cat) + output layer,
//   tcgen05.mma kind::tf32 with 3xTF32 error compensation (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo), fp32 accumulators in
//   TMEM, ACTIVATIONS RESIDENT IN TMEM as the A operand (128 samples = 128 lanes), weights in shared memory
//   (neighbour MLP resident, trunk streamed by 1-D bulk async copies through a 2-stage mbarrier ring).
// The per-layer feature injection h = act(z) + Fc c + bc (decoder.py:422-430) is folded into the next layer's GEMM:
//   z_{i+1} = L_{i+1} act(z_i) + (L_{i+1} Fc_i) c + (L_{i+1} bc_i + b_{i+1}),  so one epilogue per layer.
// Warp roles: 0-7 workers (2 threads per sample row: gathers, embeddings, bias + softplus + hi/lo split epilogues),
//             8 bulk-copy producer, 9 TMEM allocator + single-thread MMA issuer.
//
// Replaces (forward only) MLP_color.get_feature_at_pos / forward, src/conv_onet/models/decoder.py:341-449.
#include "psl_decode.cuh"
#include "psl_tc.cuh"
#include "psl_tc_layout.cuh"
#include "psl_color_tc.cuh"

namespace psl {
namespace ctc {

// ---------------------------------------------------------------------------------------------------------------------
// weight folding + packing
// ---------------------------------------------------------------------------------------------------------------------
struct FoldArgs { psl_decoder_params P; float* fold; float* blob; int job0; };

// grid (6, 128), block 64: row n of layer l:  [e part | act part | (L_act Fc_{l-1}) ] and the folded bias
__global__ void k_tc_fold(FoldArgs a) {
    const int l = blockIdx.x, n = blockIdx.y, t = threadIdx.x;
    float* row = a.fold + ((size_t)l * 128 + n) * FOLD_LD;
    const int N = l_n(l);
    const bool live = l < 5 ? true : (n < 3);
    const float* L = l < 5 ? a.P.c_W[l] : a.P.c_Wo;             // (128, K_l) or (3,128)
    const int ldl = l < 5 ? col_k(l) : 128;
    const int hoff = l == 3 ? 40 : 0;                            // where the act part starts inside L's columns
    if (n >= N) return;
    if (l == 0) {
        for (int k = t; k < 40; k += blockDim.x) row[k] = L[n * ldl + k];
        if (t == 0) a.blob[TB_VEC + V_BIAS + n] = a.P.c_b[0][n];
        return;
    }
    const int ne = l_ne(l) * 8;
    for (int k = t; k < ne; k += blockDim.x) row[k] = live ? L[n * ldl + k] : 0.f;
    for (int k = t; k < 128; k += blockDim.x) row[ne + k] = live ? L[n * ldl + hoff + k] : 0.f;
    const float* Fc = a.P.c_Wc[l - 1];                           // (128, 32)
    const float* bc = a.P.c_bc[l - 1];
    if (t < 32) {
        float s = 0.f;
        if (live)
            for (int k = 0; k < 128; ++k) s = fmaf(L[n * ldl + hoff + k], Fc[k * 32 + t], s);
        row[ne + 128 + t] = s;
    } else if (t == 32) {
        float s = 0.f;
        if (live) {
            s = l < 5 ? a.P.c_b[l][n] : a.P.c_bo[n];
            for (int k = 0; k < 128; ++k) s = fmaf(L[n * ldl + hoff + k], bc[k], s);
        }
        a.blob[TB_VEC + (l < 5 ? V_BIAS + 128 * l : V_BOUT) + n] = s;
    }
}

// split into tf32 hi/lo planes and lay out the canonical chunk images
__global__ void k_tc_pack(FoldArgs a) {
    const int job = blockIdx.y + a.job0;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (job < NLAYER) {
        const int l = job, N = l_n(l), Kf = l_ks(l) * 8;
        if (e >= N * Kf) return;
        const int n = e / Kf, kf = e - n * Kf;
        float hi, lo;
        tc::split_tf32(a.fold[((size_t)l * 128 + n) * FOLD_LD + kf], hi, lo);
        const int chunk = kf >> 5, kk = kf & 31;
        const int ksc = min(4, l_ks(l) - 4 * chunk);
        float* base = a.blob + TB_TRUNK + l_off(l) + chunk * 4 * 16 * N;
        const uint32_t o = tc::canon_off_floats(n, kk, N);
        base[o] = hi;
        base[N * 8 * ksc + o] = lo;
    } else if (job == NLAYER) {             // N1 (128,52) -> canonical 128 x 64
        if (e >= 128 * 64) return;
        const int n = e >> 6, k = e & 63;
        float hi, lo;
        tc::split_tf32(k < 52 ? a.P.c_N1[n * 52 + k] : 0.f, hi, lo);
        const uint32_t o = tc::canon_off_floats(n, k, 128);
        a.blob[TB_N1 + o] = hi;
        a.blob[TB_N1 + 128 * 64 + o] = lo;
    } else if (job == NLAYER + 1) {         // N2 (32,128) -> canonical 32 x 128
        if (e >= 32 * 128) return;
        const int n = e >> 7, k = e & 127;
        float hi, lo;
        tc::split_tf32(a.P.c_N2[n * 128 + k], hi, lo);
        const uint32_t o = tc::canon_off_floats(n, k, 32);
        a.blob[TB_N2 + o] = hi;
        a.blob[TB_N2 + 32 * 128 + o] = lo;
    } else {                                // small vectors
        if (e < 128) a.blob[TB_VEC + V_B1 + e] = a.P.c_n1b[e];
        if (e < 32) a.blob[TB_VEC + V_B2 + e] = a.P.c_n2b[e];
        if (e < 60) a.blob[TB_VEC + V_BC + e] = a.P.c_B[e];
        if (e < 30) a.blob[TB_VEC + V_BREL + (e / 10) * 12 + (e % 10)] = a.P.c_Brel[e];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void worker_signal(uint64_t* a_ready) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    tc::mbar_arrive(a_ready);
}

template <int SAVE>
__global__ void __launch_bounds__(NTHR, 1) k_color_fwd_tc(Args a, long long n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sVec = reinterpret_cast<float*>(smem + SB_VEC);
    float* sRand = reinterpret_cast<float*>(smem + SB_RAND);
    float* sEhi = reinterpret_cast<float*>(smem + SB_E);
    float* sElo = sEhi + 128 * 40;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SB_BAR);
    uint64_t* full = bars;              // [2]
    uint64_t* empty = bars + 2;         // [2]
    uint64_t* nbrw_full = bars + 4;
    uint64_t* a_ready = bars + 5;       // workers -> MMA (count 256)
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
        tc::mbar_init(a_ready, NWORK);
        tc::mbar_init(d_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < V_SIZE; i += NTHR) sVec[i] = a.blob[TB_VEC + i];
    if (threadIdx.x < 32) sRand[threadIdx.x] = a.rand_col[threadIdx.x];
    if (threadIdx.x < 12) sRand[32 + threadIdx.x] = a.affine ? a.affine[threadIdx.x] : 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // =============================== bulk-copy producer ==========================================================
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
                        tc::mbar_wait(&empty[st], ((cnt >> 1) & 1) ^ 1);
                        tc::mbar_expect_tx(&full[st], bytes);
                        tc::bulk_g2s(smem + SB_RING + st * 32768, a.blob + TB_TRUNK + l_off(l) + c * 4 * 16 * N, bytes, &full[st]);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // =============================== MMA issuer (one thread) ======================================================
        if (lane == 0) {
            uint32_t pa = 0, cnt = 0;
            const uint32_t n1 = tc::smem_u32(smem + SB_NBRW), n2 = n1 + 2 * 128 * 64 * 4;
            const uint32_t ehi = tc::smem_u32(sEhi), elo = tc::smem_u32(sElo);
            const uint32_t id128 = tc::make_idesc_tf32(128, 128), id32 = tc::make_idesc_tf32(128, 32), id16 = tc::make_idesc_tf32(128, 16);
            if (rel) tc::mbar_wait(nbrw_full, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (rel) {
                    for (int k = 0; k < 8; ++k) {
                        tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                        for (int j = 0; j < 7; ++j) {            // z1 = x N1^T        (K = 56)
                            const uint64_t bh = tc::make_smem_desc(n1 + j * 2 * 2048, 2048, 128);
                            const uint64_t bl = tc::make_smem_desc(n1 + 128 * 64 * 4 + j * 2 * 2048, 2048, 128);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 8 * j, bh, id128, j > 0);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 64 + 8 * j, bh, id128, 1);
                            tc::mma_tf32_ts(tmem + TP, tmem + TR + 8 * j, bl, id128, 1);
                        }
                        tc::mma_commit(d_ready);
                        tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
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
                    tc::mbar_wait(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                    const int N = l_n(l), ks = l_ks(l), ne = l_ne(l), na = l_na(l);
                    const uint32_t idesc = l == 5 ? id16 : id128;
                    const uint32_t dcol = (l & 1) ? TQ : TP;          // layer l writes D here ...
                    const uint32_t acol = (l & 1) ? TP : TQ;          // ... and reads act(z_{l-1}) hi from the other one
                    const uint32_t lbo = (uint32_t)N * 16u;
                    for (int j = 0; j < ks; ++j) {
                        const int cpos = j & 3;
                        const int st = cnt & 1;
                        if (cpos == 0) tc::mbar_wait(&full[st], (cnt >> 1) & 1);
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
        // =============================== workers: 2 threads per sample row ==============================================
        const int r = 32 * (warp & 3) + lane, h = warp >> 2;
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
            float cacc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) cacc[j] = 0.f;
            if (rel) {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    // ---- x_k = [sin, cos](2 pi (x_i - p) Brel) (20) | col_feats[I_k] (32) | 0 (12), hi/lo -> TMEM region R
                    float xv[32], lo[32];
                    const int id = idx[k];
                    if (h == 0) {
                        float rx = 0.f, ry = 0.f, rz = 0.f;
                        if (id >= 0) {
                            rx = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3), px));
                            ry = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 1), py));
                            rz = __fmul_rn(kTwoPi, __fsub_rn(__ldg(a.cloud_pos + (size_t)id * 3 + 2), pz));
                        }
#pragma unroll
                        for (int jj = 0; jj < 10; ++jj) {
                            float sn = 0.f, cs = 0.f;
                            if (id >= 0) sincos_embed(fmaf(rz, Br[24 + jj], fmaf(ry, Br[12 + jj], rx * Br[jj])), &sn, &cs);
                            xv[jj] = sn; xv[10 + jj] = cs;
                        }
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (id >= 0) f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + q);
                            xv[20 + 4 * q] = f4.x; xv[21 + 4 * q] = f4.y; xv[22 + 4 * q] = f4.z; xv[23 + 4 * q] = f4.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (id >= 0) f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + 3 + q);
                            xv[4 * q] = f4.x; xv[4 * q + 1] = f4.y; xv[4 * q + 2] = f4.z; xv[4 * q + 3] = f4.w;
                        }
#pragma unroll
                        for (int j = 20; j < 32; ++j) xv[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) tc::split_tf32(xv[j], xv[j], lo[j]);
                    tc::tmem_st32(lb + TR + 32 * h, xv);
                    tc::tmem_st32(lb + TR + 64 + 32 * h, lo);
                    worker_signal(a_ready);
                    if (k < 7) {                           // next neighbour's feature row / position: into L1 while the MMAs run
                        const int idn = idx[k + 1];
                        if (idn >= 0) {
                            tc::prefetch_l1(a.col_feats + (size_t)idn * 32);
                            if (h == 0) tc::prefetch_l1(a.cloud_pos + (size_t)idn * 3);
                        }
                    }
                    // ---- z1 + b1 -> softplus -> hi (in place, P) / lo (Q)
                    tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        const int c0 = 64 * h + 32 * c;
                        tc::tmem_ld32(lb + TP + c0, xv);
#pragma unroll
                        for (int j = 0; j < 32; ++j) xv[j] += b1[c0 + j];
                        if (SAVE == 1 && inb) {
                            float4* dst = reinterpret_cast<float4*>(a.save + SL.nz1 * M + (m * 8 + k) * 128 + c0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) dst[q] = make_float4(xv[4 * q], xv[4 * q + 1], xv[4 * q + 2], xv[4 * q + 3]);
                        }
                        if (SAVE == 2) {
                            float* dst = a.tsave + TL.z1T + ((tile * 8 + k) * 128 + c0) * 128 + r;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dst[j * 128] = xv[j];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) tc::split_tf32(softplus100_fast(xv[j]), xv[j], lo[j]);
                        tc::tmem_st32(lb + TP + c0, xv);
                        tc::tmem_st32(lb + TQ + c0, lo);
                    }
                    worker_signal(a_ready);
                    // ---- f = D2 + b2 ; c += wn_k f
                    tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    float f[16];
                    tc::tmem_ld16(lb + TSP + 16 * h, f);
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] += b2[16 * h + j];
                    if ((SAVE == 1 && inb) || SAVE == 2) {
                        float4* dst = SAVE == 1 ? reinterpret_cast<float4*>(a.save + SL.nf * M + (m * 8 + k) * 32 + 16 * h)
                                                : reinterpret_cast<float4*>(a.tsave + TL.f + ((tile * 128 + r) * 8 + k) * 32 + 16 * h);
#pragma unroll
                        for (int q = 0; q < 4; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) cacc[j] = fmaf(wn[k], f[j], cacc[j]);
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    if (idx[k] < 0) continue;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)idx[k] * 32 + 16 * h) + q);
                        cacc[4 * q] = fmaf(wn[k], f4.x, cacc[4 * q]); cacc[4 * q + 1] = fmaf(wn[k], f4.y, cacc[4 * q + 1]);
                        cacc[4 * q + 2] = fmaf(wn[k], f4.z, cacc[4 * q + 2]); cacc[4 * q + 3] = fmaf(wn[k], f4.w, cacc[4 * q + 3]);
                    }
                }
            }
            {   // ---- c (hi/lo) -> TMEM region C ; colour embedding (hi/lo) -> shared memory A operand
                float chi[16], clo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) cacc[j] = has ? cacc[j] : sRand[16 * h + j];
                if (SAVE == 1 && inb) {
                    float4* dst = reinterpret_cast<float4*>(a.save + SL.cc * M + m * 32 + 16 * h);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(cacc[4 * q], cacc[4 * q + 1], cacc[4 * q + 2], cacc[4 * q + 3]);
                }
                if (SAVE == 2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) a.tsave[TL.cT + (tile * 32 + 16 * h + j) * 128 + r] = cacc[j];
                    if (h == 0) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) a.tsave[TL.wnT + (tile * 8 + k) * 128 + r] = wn[k];
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) tc::split_tf32(cacc[j], chi[j], clo[j]);
                tc::tmem_st16(lb + TCC + 16 * h, chi);
                tc::tmem_st16(lb + TCC + 32 + 16 * h, clo);
                const float x = __fmul_rn(kTwoPi, px), y = __fmul_rn(kTwoPi, py), z = __fmul_rn(kTwoPi, pz);
#pragma unroll 4
                for (int j = 0; j < 20; ++j) {
                    const float arg = fmaf(z, Bc[40 + j], fmaf(y, Bc[20 + j], x * Bc[j]));
                    float ehi, elo;
                    tc::split_tf32(h == 0 ? sin_embed(arg) : cos_embed(arg), ehi, elo);
                    const uint32_t o = tc::canon_off_floats(r, 20 * h + j, 128);
                    sEhi[o] = ehi; sElo[o] = elo;
                }
                tc::fence_proxy_async();
                worker_signal(a_ready);
            }
            // ---- trunk epilogues: z + b' -> softplus -> hi in place / lo in R
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                const uint32_t dcol = (l & 1) ? TQ : TP;
                const float* bias = sVec + V_BIAS + 128 * l;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int c0 = 64 * h + 32 * c;
                    float v[32], lo[32];
                    tc::tmem_ld32(lb + dcol + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bias[c0 + j];
                    if (SAVE == 1 && inb) {
                        float4* dst = reinterpret_cast<float4*>(a.save + SL.cz * M + ((long long)l * M + m) * 128 + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                    if (SAVE == 2) {
                        float* dst = a.tsave + TL.zT + (((long long)l * n_tiles + tile) * 128 + c0) * 128 + r;
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[j * 128] = v[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) tc::split_tf32(softplus100_fast(v[j]), v[j], lo[j]);
                    tc::tmem_st32(lb + dcol + c0, v);
                    tc::tmem_st32(lb + TR + c0, lo);
                }
                worker_signal(a_ready);
            }
            // ---- output layer
            tc::mbar_wait(d_ready, pd); pd ^= 1; tc::fence_after_sync();
            {
                float o[16];
                tc::tmem_ld16(lb + TQ, o);
                if (h == 0 && inb) {
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
    if (warp == 9) tc::tmem_dealloc(tmem, 512);
}

}  // namespace ctc
}  // namespace psl

using namespace psl;

extern "C" size_t psl_tc_fold_offset_floats(void) { return (size_t)ctc::TB_TOTAL; }
extern "C" size_t psl_tc_blob_floats(void) { return (size_t)ctc::TB_TOTAL + (size_t)ctc::FOLD_FLOATS; }

// fold + split + lay out the colour-branch weights for the tensor-core kernel (blob: psl_tc_blob_floats() floats)
extern "C" int psl_tc_pack_params(const psl_decoder_params* P, float* blob, psl_stream_t stream) {
    PSL_REQUIRE(P && blob, "NULL argument");
    cudaStream_t st = as_stream(stream);
    PSL_CHECK_CUDA(cudaMemsetAsync(blob, 0, sizeof(float) * ctc::TB_TOTAL, st));
    ctc::FoldArgs fa;
    fa.P = *P; fa.blob = blob; fa.fold = blob + ctc::TB_TOTAL; fa.job0 = 0;
    TimingScope ts(T_PACK, st, 3);
    ctc::k_tc_fold<<<dim3(ctc::NLAYER, 128), 64, 0, st>>>(fa);
    ctc::k_tc_pack<<<dim3((128 * 200 + 255) / 256, ctc::NLAYER + 3), 256, 0, st>>>(fa);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// the part of psl_tc_pack_params the f16-plane kernels need: folded fp32 rows (behind the blob) and the small-vector section
// (biases, Fourier bases); the tf32 chunk images are NOT written
extern "C" int psl_tc_fold_params(const psl_decoder_params* P, float* blob, psl_stream_t stream) {
    PSL_REQUIRE(P && blob, "NULL argument");
    cudaStream_t st = as_stream(stream);
    ctc::FoldArgs fa;
    fa.P = *P; fa.blob = blob; fa.fold = blob + ctc::TB_TOTAL; fa.job0 = ctc::NLAYER + 2;
    TimingScope ts(T_PACK, st, 2);
    ctc::k_tc_fold<<<dim3(ctc::NLAYER, 128), 64, 0, st>>>(fa);
    ctc::k_tc_pack<<<dim3(1, 1), 256, 0, st>>>(fa);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// colour branch on tensor cores: writes raw[:, 0:3]; raw[:, 3] / has_nb come from psl_decode_fwd(stage = geometry)
extern "C" int psl_color_fwd_tc(const psl_decode_cfg* cfg, const float* tc_blob, const float* pos, int64_t m,
                                const int32_t* I, const float* D, const int32_t* nnum, const double* r2,
                                const float* cloud_pos, const float* col_feats, const float* rand_col,
                                const float* exposure_affine, float* raw, float* save, float* tsave, psl_stream_t stream) {
    PSL_REQUIRE(cfg && tc_blob && pos && I && D && nnum && col_feats && rand_col && raw, "NULL argument");
    PSL_REQUIRE(!cfg->encode_rel_pos || cloud_pos, "rel-pos encoding needs cloud_pos");
    PSL_REQUIRE(cfg->rgb_mode != PSL_RGB_AFFINE_SIGMOID || exposure_affine, "affine mode needs exposure_affine");
    if (m == 0) return 0;
    ctc::Args a{};
    a.cfg = *cfg; a.blob = tc_blob; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.col_feats = col_feats; a.rand_col = rand_col; a.affine = exposure_affine; a.raw = raw; a.save = save; a.tsave = tsave;
    PSL_REQUIRE(!(save && tsave), "pass at most one of save / tsave");
    const long long n_tiles = (m + ctc::TM - 1) / ctc::TM;
    static bool attr_set = false;
    if (!attr_set) {
        PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc::k_color_fwd_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
        PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc::k_color_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
        PSL_CHECK_CUDA(cudaFuncSetAttribute(ctc::k_color_fwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctc::SB_TOTAL));
        attr_set = true;
    }
    const long long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    TimingScope ts(T_COLOR_FWD_TC, as_stream(stream));
    if (tsave) ctc::k_color_fwd_tc<2><<<(unsigned)grid, ctc::NTHR, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    else if (save) ctc::k_color_fwd_tc<1><<<(unsigned)grid, ctc::NTHR, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    else ctc::k_color_fwd_tc<0><<<(unsigned)grid, ctc::NTHR, ctc::SB_TOTAL, as_stream(stream)>>>(a, n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
