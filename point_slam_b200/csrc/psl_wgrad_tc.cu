// Weight gradients of the colour branch on tcgen05 (3xTF32): GEMMs whose reduction index is the SAMPLE.
//   For trunk layer l (z_l = L_l h_{l-1} + b_l, h = softplus(z) + Fc c + bc):
//     dL_l  = dz_l^T h_{l-1} = dz_l^T a_{l-1}  +  (dz_l^T c) Fc_{l-1}^T  +  (sum dz_l) bc_{l-1}^T     (a = softplus(z))
//     dFc_l = dh_l^T c ,  db_l = sum dz_l ,  dbc_l = sum dh_l ,  embedding columns: dz_l^T e
//   so the kernel accumulates  S1 = dz^T a, S2 = dz^T c, S4 = dz^T e (one MMA group, B rows [a | c | e]) and S3 = dh^T c,
//   the output layer's  a_4^T dout, c^T dout,  and the neighbour MLP's  dz1^T x, softplus(z1)^T df;
//   a small finalize kernel combines them into the reference-layout gradients.
// Operands come from the tile-transposed buffers written by the forward / backward-data kernels (psl_tc_layout.cuh), so
// staging is a 16-byte load, an elementwise function, the tf32 hi/lo split and two 16-byte shared-memory stores; all
// operands are K-major (K = sample).  Work unit = one 16-sample chunk of a tile; every CTA owns a contiguous range of units,
// stages them through a 3-deep ring of operand images (the workers fill stage i+1, i+2 while the MMAs of stage i run; the
// global operands of the next item are already in registers) and keeps its accumulators in TMEM over the whole range;
// one partial buffer per CTA, fixed-order reduction -> deterministic.
#include "psl_decode.cuh"
#include "psl_tc.cuh"
#include "psl_tc_layout.cuh"

namespace psl {
namespace wgt {

constexpr int NWORK = 256, NTHR = 288;          // warps 0-7 staging workers, warp 8 MMA issuer + TMEM allocator
constexpr int KC = 16;                          // samples per staged chunk (2 k-steps)
constexpr int NS = 3;                           // staging ring depth: workers stage chunk i+1, i+2 while the MMAs of chunk i run
constexpr int UPT = 128 / KC;                   // work units (chunks) per 128-sample tile

// ---- per-CTA partial layout (floats) ------------------------------------------------------------------------------------
__host__ __device__ constexpr int W_S1(int l) { return l * 32768; }                       // [128][128]  (l >= 1)
__host__ __device__ constexpr int W_S2(int l) { return l * 32768 + 16384; }               // [128][32]   (l >= 1)
__host__ __device__ constexpr int W_S3(int l) { return l * 32768 + 20480; }               // [128][32]
__host__ __device__ constexpr int W_S4(int l) { return l * 32768 + 24576; }               // [128][48]   (l = 0, 3)
__host__ __device__ constexpr int W_DB(int l) { return l * 32768 + 30720; }               // [128]
__host__ __device__ constexpr int W_DBC(int l) { return l * 32768 + 30848; }              // [128]
constexpr int W_U1 = 5 * 32768;                                       // [128][16]  a_4^T dout
constexpr int W_U2 = W_U1 + 2048;                                     // [128][16]  rows 0..31: c^T dout
constexpr int W_DBO = W_U2 + 2048;                                    // [16]
constexpr int W_N1 = W_DBO + 16;                                      // [128][64]  dz1^T x
constexpr int W_N2T = W_N1 + 8192;                                    // [128][32]  softplus(z1)^T df
constexpr int W_DB1 = W_N2T + 4096;                                   // [128]
constexpr int W_DB2 = W_DB1 + 128;                                    // [32]
constexpr int W_TOTAL = W_DB2 + 32;

// ---- shared memory (bytes): NS stages of {A0, A1, B, B2} ----------------------------------------------------------------
constexpr int ST_A0 = 0;                        // A operand 0: [128 rows x 16 k] hi | lo   (16 KB)
constexpr int ST_A1 = 16384;                    // A operand 1                              (16 KB)
constexpr int ST_B = 32768;                     // B operand: up to 208 rows x 16 k, hi | lo (26 KB)
constexpr int ST_B2 = ST_B + 26624;             // second small B (neighbour df / dout): 32 rows x 16 k hi | lo (4 KB)
constexpr int ST_BYTES = ST_B2 + 4096;          // 63488
constexpr int SB_POS = NS * ST_BYTES;           // [128][4] positions of the current tile
constexpr int SB_VEC = SB_POS + 2048;           // Bc [3][20] (64) | Brel [3][12] (48)
constexpr int SB_BAR = SB_VEC + 512;
constexpr int SB_IDX = SB_BAR + 128;            // [8][128] neighbour indices of the current tile (-1: no contribution)
constexpr int SB_TOTAL = SB_IDX + 4096;

struct Args {
    psl_decode_cfg cfg;
    psl_decoder_params P;
    const float* pos; long long m;
    const int* I;
    const float* cloud_pos; const float* col_feats;
    const float* tsave; const float* tbwd;
    float* partial;                 // [grid][W_TOTAL]
};

__device__ __forceinline__ float sp_fast(float z) { return softplus100_fast(z); }
__device__ __forceinline__ float spg_fast(float z) { return __fdividef(1.0f, 1.0f + __expf(-fminf(100.0f * z, 30.0f))); }

// store 4 consecutive-k values of row `row` (k0 multiple of 4) into a canonical (R rows x KC) hi|lo operand image
__device__ __forceinline__ void put4(float* base, int R, int row, int k0, float4 v) {
    float4 hi, lo;
    tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y); tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
    const uint32_t o = (uint32_t)(((k0 >> 2) * (R >> 3) + (row >> 3)) * 32 + (row & 7) * 4);
    *reinterpret_cast<float4*>(base + o) = hi;
    *reinterpret_cast<float4*>(base + R * KC + o) = lo;
}
__device__ __forceinline__ float sum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

// 128 rows x KC samples of one [channel][128 samples] plane: thread (row = tid / 2, 8 samples = 2 float4).
// MODE 0: v = x      2: v = softplus(z)
template <int MODE>
__device__ __forceinline__ float rows128(float* dstA, const float* __restrict__ X, int m0c, int tid) {
    const int row = tid >> 1, k0 = (tid & 1) * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(X + row * 128 + m0c + k0);
    const float4 x1 = *reinterpret_cast<const float4*>(X + row * 128 + m0c + k0 + 4);
    float4 v0 = x0, v1 = x1;
    if (MODE == 2) {
        v0 = make_float4(sp_fast(x0.x), sp_fast(x0.y), sp_fast(x0.z), sp_fast(x0.w));
        v1 = make_float4(sp_fast(x1.x), sp_fast(x1.y), sp_fast(x1.z), sp_fast(x1.w));
    }
    put4(dstA, 128, row, k0, v0);
    put4(dstA, 128, row, k0 + 4, v1);
    float rs = sum4(v0) + sum4(v1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    return rs;                                      // row sum over the chunk (both threads of a row hold it)
}
struct TrunkRegs { float4 h[2], z[2], zp[2], c; };
struct NbrRegs { float4 dz[2], z1[2], dcc, w4; };
struct NbrGather { float4 f4; float cx, cy, cz; int id; };      // the neighbour's feature row / position: fetched TWO items ahead

__global__ void __launch_bounds__(NTHR, 1) k_wgrad_tc(Args a, long long n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sPos = reinterpret_cast<float*>(smem + SB_POS);
    float* sVec = reinterpret_cast<float*>(smem + SB_VEC);
    int* sIdx = reinterpret_cast<int*>(smem + SB_IDX);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SB_BAR);
    uint64_t* staged = bars;            // [NS] workers -> MMA (count 256)
    uint64_t* consumed = bars + NS;     // [NS] MMA -> workers (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const bool rel = a.cfg.encode_rel_pos != 0;
    const TSave TL = tsave_layout(a.m, a.cfg.encode_rel_pos);
    const TBwd BL = tbwd_layout(a.m, a.cfg.encode_rel_pos);
    float* part = a.partial + (size_t)blockIdx.x * W_TOTAL;
    // this CTA's contiguous range of work units (unit = one KC-sample chunk of one tile)
    const long long n_units = n_tiles * UPT;
    const long long u0 = (long long)blockIdx.x * n_units / gridDim.x, u1 = (long long)(blockIdx.x + 1) * n_units / gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { tc::mbar_init(staged + s, NWORK); tc::mbar_init(consumed + s, 1); }
        tc::mbar_fence_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
    if (tid < 60) sVec[tid] = a.P.c_B[tid];
    if (tid < 30) sVec[64 + (tid / 10) * 12 + (tid % 10)] = a.P.c_Brel[tid];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t smem0 = tc::smem_u32(smem);
    unsigned it = 0;                    // staged items so far (workers and the MMA thread walk the same sequence)
    long long cur_tile = -1;

    // one MMA group: D[dcol : dcol+N] (+)= A(abase) x B(rows from brow, R rows total)
    auto mma_group = [&](uint32_t abase, uint32_t bbase, int R, int brow, int N, uint32_t dcol, uint32_t first) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N), lboA = 128u * 16u, lboB = (uint32_t)R * 16u;
#pragma unroll
        for (int j = 0; j < KC / 8; ++j) {
            const uint64_t ah = tc::make_smem_desc(abase + j * 2 * lboA, lboA, 128);
            const uint64_t al = tc::make_smem_desc(abase + 128 * KC * 4 + j * 2 * lboA, lboA, 128);
            const uint32_t bo = bbase + (uint32_t)(brow >> 3) * 128u + j * 2 * lboB;
            const uint64_t bh = tc::make_smem_desc(bo, lboB, 128);
            const uint64_t bl = tc::make_smem_desc(bo + (uint32_t)R * KC * 4, lboB, 128);
            tc::mma_tf32_ss(tmem + dcol, ah, bh, idesc, (j == 0) ? first : 1u);
            tc::mma_tf32_ss(tmem + dcol, al, bh, idesc, 1);
            tc::mma_tf32_ss(tmem + dcol, ah, bl, idesc, 1);
        }
    };
    // workers: claim the next ring stage (wait until the MMAs of its previous use are done) / hand it to the MMA thread
    auto acquire = [&]() -> unsigned char* {
        const unsigned s = it % NS, n = it / NS;
        if (n > 0) tc::mbar_wait(consumed + s, (n - 1) & 1);
        return smem + s * ST_BYTES;
    };
    auto hand_over = [&]() {
        tc::fence_proxy_async();
        tc::mbar_arrive(staged + it % NS);
        ++it;
    };
    // MMA thread: wait for the next staged item; returns the stage's shared-memory address
    auto mma_wait = [&]() -> uint32_t {
        const unsigned s = it % NS, n = it / NS;
        tc::mbar_wait(staged + s, n & 1);
        tc::fence_after_sync();
        return smem0 + s * ST_BYTES;
    };
    auto mma_done = [&]() { tc::mma_commit(consumed + it % NS); ++it; };
    // workers: every MMA issued so far has completed (commits complete in order)
    auto wait_all = [&]() {
        if (it > 0) { const unsigned l = it - 1; tc::mbar_wait(consumed + l % NS, (l / NS) & 1); }
        tc::fence_after_sync();
    };
    auto load_pos = [&](long long tile) {             // workers: positions of `tile` -> sPos (used by the embeddings)
        if (tile == cur_tile) return;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = tid; i < 128 * 3; i += NWORK) {
            const long long m = tile * 128 + i / 3;
            sPos[(i / 3) * 4 + i % 3] = m < a.m ? a.pos[m * 3 + i % 3] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_tile = tile;
    };
    // workers: resolved neighbour indices of `tile` -> sIdx, so that the per-item prefetch starts its gathers from a shared-memory
    // read instead of a weight -> index -> row chain of three dependent global loads
    long long ids_tile = -1;
    auto load_ids = [&](long long tile) {
        if (tile == ids_tile) return;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = tid; i < 8 * 128; i += NWORK) {
            const int k = i >> 7, sm = i & 127;
            const long long m = tile * 128 + sm;
            int id = -1;
            if (m < a.m && a.tsave[TL.wnT + (tile * 8 + k) * 128 + sm] != 0.f) id = a.I[m * 8 + k];
            sIdx[i] = id;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        ids_tile = tile;
    };
    // drain TMEM columns [col, col+ncols) of this thread's lane into partial rows (row = channel), both halves of the warp group
    auto drain = [&](uint32_t col, int ncols, float* dst, int ld) {
        const int r = 32 * (warp & 3) + lane, h = warp >> 2;
        const uint32_t lb = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        for (int c0 = 16 * h; c0 < ncols; c0 += 32) {
            float v[16];
            tc::tmem_ld16(lb + col + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(dst + r * ld + c0 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    };

    // =========================================== trunk layers, two groups ==================================================
    for (int grp = 0; grp < 2; ++grp) {
        const int l_hi = grp == 0 ? 4 : 2, l_lo = grp == 0 ? 3 : 0;
        float db_acc[3] = {0.f, 0.f, 0.f}, dbc_acc[3] = {0.f, 0.f, 0.f};
        const int n_l = l_hi - l_lo + 1;
        const long long n_items = (u1 - u0) * n_l;
        if (warp < 8) {
            // global operands of item j (unit u0 + j / n_l, layer l_hi - j % n_l), fetched one item ahead of their use
            TrunkRegs cur, nxt;
            auto fetch = [&](long long u, int li_, TrunkRegs& r) {
                const int l = l_hi - li_;
                const long long tile = u / UPT;
                const int m0c = (int)(u % UPT) * KC;
                const int row = tid >> 1, k0 = (tid & 1) * 8;
                const float* dhT = a.tbwd + BL.dhT + ((long long)l * n_tiles + tile) * 16384 + row * 128 + m0c + k0;
                const float* zT = a.tsave + TL.zT + ((long long)l * n_tiles + tile) * 16384 + row * 128 + m0c + k0;
                r.h[0] = *reinterpret_cast<const float4*>(dhT); r.h[1] = *reinterpret_cast<const float4*>(dhT + 4);
                r.z[0] = *reinterpret_cast<const float4*>(zT); r.z[1] = *reinterpret_cast<const float4*>(zT + 4);
                if (l >= 1) {
                    const float* zpT = a.tsave + TL.zT + ((long long)(l - 1) * n_tiles + tile) * 16384 + row * 128 + m0c + k0;
                    r.zp[0] = *reinterpret_cast<const float4*>(zpT); r.zp[1] = *reinterpret_cast<const float4*>(zpT + 4);
                }
                if (tid < 128) r.c = *reinterpret_cast<const float4*>(a.tsave + TL.cT + tile * 4096 + (tid >> 2) * 128 + m0c + 4 * (tid & 3));
            };
            if (n_items > 0) fetch(u0, 0, cur);
            long long u = u0;
            int li = 0;                                      // item j = (unit u, layer l_hi - li), walked without 64-bit divisions
            for (long long j = 0; j < n_items; ++j) {
                const int l = l_hi - li;
                const long long tile = u / UPT;
                const int m0c = (int)(u % UPT) * KC;
                if (li == 0) load_pos(tile);                 // layers 3 and 0 embed the sample position
                const bool wrap = li + 1 == n_l;
                if (j + 1 < n_items) fetch(u + (wrap ? 1 : 0), wrap ? 0 : li + 1, nxt);
                const int R = l == 0 ? 80 : 208;             // B rows: l >= 1: [a 128 | c 32 | e 48], l == 0: [e 48 | c 32]
                const int crow = l == 0 ? 48 : 128;
                unsigned char* st = acquire();
                float* sA0 = reinterpret_cast<float*>(st + ST_A0);
                float* sA1 = reinterpret_cast<float*>(st + ST_A1);
                float* sB = reinterpret_cast<float*>(st + ST_B);
                {
                    const int row = tid >> 1, k0 = (tid & 1) * 8;
                    float sdz = 0.f, sdh = 0.f;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {                                          // A0 = dz_l^T, A1 = dh_l^T
                        const float4 h = cur.h[q], z = cur.z[q];
                        const float4 d = make_float4(h.x * spg_fast(z.x), h.y * spg_fast(z.y), h.z * spg_fast(z.z), h.w * spg_fast(z.w));
                        put4(sA0, 128, row, k0 + 4 * q, d);
                        put4(sA1, 128, row, k0 + 4 * q, h);
                        sdz += sum4(d); sdh += sum4(h);
                        if (l >= 1) {                                                      // B rows 0..127 = a_{l-1}^T
                            const float4 zp = cur.zp[q];
                            put4(sB, R, row, k0 + 4 * q, make_float4(sp_fast(zp.x), sp_fast(zp.y), sp_fast(zp.z), sp_fast(zp.w)));
                        }
                    }
                    sdz += __shfl_xor_sync(0xffffffffu, sdz, 1);
                    sdh += __shfl_xor_sync(0xffffffffu, sdh, 1);
                    if (li == 0) { db_acc[0] += sdz; dbc_acc[0] += sdh; }
                    else if (li == 1) { db_acc[1] += sdz; dbc_acc[1] += sdh; }
                    else { db_acc[2] += sdz; dbc_acc[2] += sdh; }
                }
                if (tid < 128) put4(sB, R, crow + (tid >> 2), 4 * (tid & 3), cur.c);       // c^T: 32 rows x 4 float4
                if (l == 0 || l == 3) {
                    // e^T: rows j (sin) and 20+j (cos), j < 20; rows 40..47 zero.  item = (j, sample): 320 items
                    const int erow = l == 0 ? 0 : 160;
                    for (int i = tid; i < 20 * KC; i += NWORK) {
                        const int jj = i / KC, k = i - jj * KC;
                        const float* p = sPos + (m0c + k) * 4;
                        const float x = __fmul_rn(kTwoPi, p[0]), y = __fmul_rn(kTwoPi, p[1]), z = __fmul_rn(kTwoPi, p[2]);
                        float sn, cs;
                        sincos_embed(fmaf(z, sVec[40 + jj], fmaf(y, sVec[20 + jj], x * sVec[jj])), &sn, &cs);
                        float hi, lo;
                        const uint32_t o1 = tc::canon_off_floats(erow + jj, k, R), o2 = tc::canon_off_floats(erow + 20 + jj, k, R);
                        tc::split_tf32(sn, hi, lo); sB[o1] = hi; sB[R * KC + o1] = lo;
                        tc::split_tf32(cs, hi, lo); sB[o2] = hi; sB[R * KC + o2] = lo;
                    }
                    if (tid < 8 * KC) {
                        const uint32_t o = tc::canon_off_floats(erow + 40 + tid / KC, tid % KC, R);
                        sB[o] = 0.f; sB[R * KC + o] = 0.f;
                    }
                }
                hand_over();
                cur = nxt;
                if (wrap) { li = 0; ++u; } else ++li;
            }
        } else if (lane == 0) {
            int li = 0;
            for (long long j = 0; j < n_items; ++j, li = (li + 1 == n_l) ? 0 : li + 1) {
                const int l = l_hi - li;
                const uint32_t base = (uint32_t)(li == 0 ? 0 : (li == 1 ? 192 : 384));
                const int ncol1 = l == 0 ? 48 : (l == 3 ? 208 : 160);
                const int R = l == 0 ? 80 : 208, crow = l == 0 ? 48 : 128;
                const uint32_t sb = mma_wait();
                const uint32_t first = (j < n_l) ? 0u : 1u;
                mma_group(sb + ST_A0, sb + ST_B, R, 0, ncol1, base, first);                // [S1 | S2 | S4]  (l == 0: S4)
                mma_group(sb + ST_A1, sb + ST_B, R, crow, 32, base + ncol1, first);        // S3 = dh^T c
                mma_done();
            }
        }
        // ---- drain this group's accumulators into the partial buffer ------------------------------------------------------
        if (warp < 8) {
            wait_all();
            const bool any = u1 > u0;
            for (int l = l_hi, li = 0; l >= l_lo; --l, ++li) {
                const uint32_t base = (uint32_t)(li == 0 ? 0 : (li == 1 ? 192 : 384));
                if (any) {
                    if (l >= 1) {
                        drain(base, 128, part + W_S1(l), 128);
                        drain(base + 128, 32, part + W_S2(l), 32);
                        if (l == 3) drain(base + 160, 48, part + W_S4(l), 48);
                        drain(base + (l == 3 ? 208 : 160), 32, part + W_S3(l), 32);
                    } else {
                        drain(base, 48, part + W_S4(0), 48);
                        drain(base + 48, 32, part + W_S3(0), 32);
                    }
                }
                if ((tid & 1) == 0) { part[W_DB(l) + (tid >> 1)] = db_acc[li]; part[W_DBC(l) + (tid >> 1)] = dbc_acc[li]; }
            }
            tc::fence_before_sync();
        }
        __syncthreads();                               // TMEM columns are reused by the next group
        tc::fence_after_sync();
    }

    // =========================================== output layer + neighbour MLP ===============================================
    {
        float dbo_acc = 0.f, db1_acc = 0.f, db2_acc = 0.f;
        if (warp < 8) {                                 // A1 rows 32..127 stay zero for the c^T dout product (every stage)
            for (int s = 0; s < NS; ++s) {
                float* sA1 = reinterpret_cast<float*>(smem + s * ST_BYTES + ST_A1);
                for (int i = tid; i < 2 * 128 * KC; i += NWORK) sA1[i] = 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        for (long long u = u0; u < u1; ++u) {
            const long long tile = u / UPT;
            const int m0c = (int)(u % UPT) * KC;
            if (warp < 8) {
                const float* z4T = a.tsave + TL.zT + (4ll * n_tiles + tile) * 16384;
                const float* cT = a.tsave + TL.cT + tile * 4096;
                const float* doutT = a.tbwd + BL.doutT + tile * 2048;
                unsigned char* st = acquire();
                float* sA0 = reinterpret_cast<float*>(st + ST_A0);
                float* sA1 = reinterpret_cast<float*>(st + ST_A1);
                float* sB2 = reinterpret_cast<float*>(st + ST_B2);
                rows128<2>(sA0, z4T, m0c, tid);                                        // A0 = a_4^T
                if (tid < 128) {                                                       // A1 rows 0..31 = c^T
                    const int row = tid >> 2, q = tid & 3;
                    put4(sA1, 128, row, 4 * q, *reinterpret_cast<const float4*>(cT + row * 128 + m0c + 4 * q));
                }
                if (tid < 64) {                                                        // B2 = dout^T (16 rows x 4 float4)
                    const int row = tid >> 2, q = tid & 3;
                    const float4 v = *reinterpret_cast<const float4*>(doutT + row * 128 + m0c + 4 * q);
                    put4(sB2, 16, row, 4 * q, v);
                    float s = sum4(v);
                    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
                    dbo_acc += s;                                                      // threads with q == 0 keep row sums
                }
                hand_over();
            } else if (lane == 0) {
                const uint32_t sb = mma_wait();
                const uint32_t first = (u == u0) ? 0u : 1u;
                mma_group(sb + ST_A0, sb + ST_B2, 16, 0, 16, 0, first);                // U1 = a_4^T dout
                mma_group(sb + ST_A1, sb + ST_B2, 16, 0, 16, 16, first);               // U2 = [c^T; 0] dout
                mma_done();
            }
        }
        if (rel) {
            const long long n_items = (u1 - u0) * 8;
            if (warp < 8) {
                const int s = tid & 15, p = tid >> 4;
                NbrRegs cur, nxt;
                NbrGather gcur, gnxt, gnx2;
                // global operands of item j (unit u0 + j / 8, neighbour j % 8).  The dense planes are fetched one item ahead; the
                // neighbour's feature row and position -- an index read followed by a scattered, DRAM-latency gather -- two items ahead
                auto fetch = [&](long long j, NbrRegs& r) {
                    const long long u = u0 + (j >> 3);
                    const int k = (int)(j & 7);
                    const long long tile = u / UPT;
                    const int m0c = (int)(u % UPT) * KC;
                    const int row = tid >> 1, k0 = (tid & 1) * 8;
                    const float* dz1T = a.tbwd + BL.dz1T + (tile * 8 + k) * 16384 + row * 128 + m0c + k0;
                    const float* z1T = a.tsave + TL.z1T + (tile * 8 + k) * 16384 + row * 128 + m0c + k0;
                    const float* wnT = a.tsave + TL.wnT + (tile * 8 + k) * 128;
                    r.dz[0] = *reinterpret_cast<const float4*>(dz1T); r.dz[1] = *reinterpret_cast<const float4*>(dz1T + 4);
                    r.z1[0] = *reinterpret_cast<const float4*>(z1T); r.z1[1] = *reinterpret_cast<const float4*>(z1T + 4);
                    if (tid < 128) {
                        r.dcc = *reinterpret_cast<const float4*>(a.tbwd + BL.dccT + tile * 4096 + (tid >> 2) * 128 + m0c + 4 * (tid & 3));
                        r.w4 = *reinterpret_cast<const float4*>(wnT + m0c + 4 * (tid & 3));
                    }
                };
                auto gather = [&](long long j, NbrGather& r) {
                    const long long u = u0 + (j >> 3);
                    const int k = (int)(j & 7);
                    const long long tile = u / UPT;
                    const int m0c = (int)(u % UPT) * KC;
                    const long long m = tile * 128 + m0c + s;
                    int id = -1;
                    if (tile == ids_tile) id = sIdx[k * 128 + m0c + s];
                    else if (m < a.m && a.tsave[TL.wnT + (tile * 8 + k) * 128 + m0c + s] != 0.f) id = a.I[m * 8 + k];   // items of the next tile
                    r.id = id;
                    r.f4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    r.cx = r.cy = r.cz = 0.f;
                    if (id >= 0) {
                        if (p < 8) r.f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)id * 32) + p);
                        if (p < 10) { r.cx = __ldg(a.cloud_pos + (size_t)id * 3); r.cy = __ldg(a.cloud_pos + (size_t)id * 3 + 1); r.cz = __ldg(a.cloud_pos + (size_t)id * 3 + 2); }
                    }
                };
                if (n_items > 0) { fetch(0, cur); gather(0, gcur); }
                if (n_items > 1) gather(1, gnxt);
                for (long long j = 0; j < n_items; ++j) {
                    const long long u = u0 + (j >> 3);
                    const int k = (int)(j & 7);
                    const long long tile = u / UPT;
                    const int m0c = (int)(u % UPT) * KC;
                    if (k == 0) { load_pos(tile); load_ids(tile); }
                    if (j + 1 < n_items) fetch(j + 1, nxt);
                    if (j + 2 < n_items) gather(j + 2, gnx2);
                    unsigned char* st = acquire();
                    float* sA0 = reinterpret_cast<float*>(st + ST_A0);
                    float* sA1 = reinterpret_cast<float*>(st + ST_A1);
                    float* sB = reinterpret_cast<float*>(st + ST_B);
                    float* sB2 = reinterpret_cast<float*>(st + ST_B2);
                    {
                        const int row = tid >> 1, k0 = (tid & 1) * 8;
                        float rs = 0.f;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            put4(sA0, 128, row, k0 + 4 * q, cur.dz[q]);                                    // A0 = dz1^T
                            rs += sum4(cur.dz[q]);
                            const float4 z = cur.z1[q];
                            put4(sA1, 128, row, k0 + 4 * q, make_float4(sp_fast(z.x), sp_fast(z.y), sp_fast(z.z), sp_fast(z.w)));   // A1 = softplus(z1)^T
                        }
                        rs += __shfl_xor_sync(0xffffffffu, rs, 1);
                        db1_acc += rs;
                    }
                    {   // B (64 rows x 16 samples) = x_k^T : thread (sample s = tid % 16, part p = tid / 16):
                        // p < 8: feature float4 p (rows 20 + 4p ..), p < 10: sin/cos j = p (rows p, 10 + p), p >= 10: zero rows 52..63
                        if (p < 8) {
                            const float fv[4] = {gcur.f4.x, gcur.f4.y, gcur.f4.z, gcur.f4.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float hi, lo;
                                tc::split_tf32(fv[c], hi, lo);
                                const uint32_t o = tc::canon_off_floats(20 + 4 * p + c, s, 64);
                                sB[o] = hi; sB[64 * KC + o] = lo;
                            }
                        }
                        if (p < 10) {
                            float sn = 0.f, cs = 0.f;
                            if (gcur.id >= 0) {
                                const float* pp = sPos + (m0c + s) * 4;
                                const float rx = __fmul_rn(kTwoPi, __fsub_rn(gcur.cx, pp[0]));
                                const float ry = __fmul_rn(kTwoPi, __fsub_rn(gcur.cy, pp[1]));
                                const float rz = __fmul_rn(kTwoPi, __fsub_rn(gcur.cz, pp[2]));
                                sincos_embed(fmaf(rz, sVec[64 + 24 + p], fmaf(ry, sVec[64 + 12 + p], rx * sVec[64 + p])), &sn, &cs);
                            }
                            float hi, lo;
                            const uint32_t o1 = tc::canon_off_floats(p, s, 64), o2 = tc::canon_off_floats(10 + p, s, 64);
                            tc::split_tf32(sn, hi, lo); sB[o1] = hi; sB[64 * KC + o1] = lo;
                            tc::split_tf32(cs, hi, lo); sB[o2] = hi; sB[64 * KC + o2] = lo;
                        } else {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const uint32_t o = tc::canon_off_floats(52 + 2 * (p - 10) + c, s, 64);
                                sB[o] = 0.f; sB[64 * KC + o] = 0.f;
                            }
                        }
                    }
                    if (tid < 128) {   // B2 (32 rows x 4 float4) = df_k^T = wn_k * dcc^T
                        const float4 d = cur.dcc, w = cur.w4;
                        const float4 v = make_float4(d.x * w.x, d.y * w.y, d.z * w.z, d.w * w.w);
                        put4(sB2, 32, tid >> 2, 4 * (tid & 3), v);
                        float sm = sum4(v);
                        sm += __shfl_xor_sync(0xffffffffu, sm, 1); sm += __shfl_xor_sync(0xffffffffu, sm, 2);
                        db2_acc += sm;
                    }
                    hand_over();
                    cur = nxt;
                    gcur = gnxt; gnxt = gnx2;
                }
            } else if (lane == 0) {
                for (long long j = 0; j < n_items; ++j) {
                    const uint32_t sb = mma_wait();
                    const uint32_t first = (j == 0) ? 0u : 1u;
                    mma_group(sb + ST_A0, sb + ST_B, 64, 0, 64, 32, first);                // dN1   = dz1^T x
                    mma_group(sb + ST_A1, sb + ST_B2, 32, 0, 32, 96, first);               // dN2^T = softplus(z1)^T df
                    mma_done();
                }
            }
        }
        if (warp < 8) {
            wait_all();
            if (u1 > u0) {
                drain(0, 16, part + W_U1, 16);
                drain(16, 16, part + W_U2, 16);
            }
            if (tid < 64 && (tid & 3) == 0) part[W_DBO + (tid >> 2)] = dbo_acc;
            if (rel) {
                if (u1 > u0) {
                    drain(32, 64, part + W_N1, 64);
                    drain(96, 32, part + W_N2T, 32);
                }
                if ((tid & 1) == 0) part[W_DB1 + (tid >> 1)] = db1_acc;
                if (tid < 128 && (tid & 3) == 0) part[W_DB2 + (tid >> 2)] = db2_acc;
            }
            tc::fence_before_sync();
        }
    }
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// fixed-order reduction of the per-CTA partials + combination into the reference-layout gradients
// ---------------------------------------------------------------------------------------------------------------------
// one thread per element; four interleaved accumulators (CTA c goes to accumulator c % 4) keep 4 loads in flight and are
// combined in a fixed order, so the result does not depend on the launch configuration
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int n_cta, float* __restrict__ sums) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= W_TOTAL) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= n_cta; c += 4) {
        s0 += partial[(size_t)c * W_TOTAL + e];
        s1 += partial[(size_t)(c + 1) * W_TOTAL + e];
        s2 += partial[(size_t)(c + 2) * W_TOTAL + e];
        s3 += partial[(size_t)(c + 3) * W_TOTAL + e];
    }
    for (; c < n_cta; ++c) s0 += partial[(size_t)c * W_TOTAL + e];
    sums[e] = (s0 + s1) + (s2 + s3);
}

struct FinArgs {
    psl_decoder_params P;
    psl_decoder_grads G;
    const float* sums;
    const float* part_brel; int n_cta_a;          // per-CTA dBrel partials of the backward-data kernel
    const float* aff; long long aff_rows;         // per-sample affine terms (rows x 12) or NULL
    float* d_affine;
    int rel;
};

__global__ void k_wgrad_finalize(FinArgs a) {
    const int job = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const float* S = a.sums;
    if (job < 5) {
        const int l = job, K = col_k(l);
        if (a.G.c_W[l] && e < 128 * K) {
            const int n = e / K, k = e - n * K;
            float v;
            const int ne = (l == 0 || l == 3) ? 40 : 0;
            if (k < ne) v = S[W_S4(l) + n * 48 + k];
            else {
                const int kk = k - ne;                                         // column of h_{l-1}
                v = S[W_S1(l) + n * 128 + kk];
                const float* Fc = a.P.c_Wc[l - 1];
                for (int j = 0; j < 32; ++j) v = fmaf(S[W_S2(l) + n * 32 + j], Fc[kk * 32 + j], v);
                v = fmaf(S[W_DB(l) + n], a.P.c_bc[l - 1][kk], v);
            }
            a.G.c_W[l][e] = v;
        }
        if (a.G.c_Wc[l] && e < 128 * 32) a.G.c_Wc[l][e] = S[W_S3(l) + e];
        if (a.G.c_b[l] && e < 128) a.G.c_b[l][e] = S[W_DB(l) + e];
        if (a.G.c_bc[l] && e < 128) a.G.c_bc[l][e] = S[W_DBC(l) + e];
    } else if (job == 5) {
        if (a.G.c_Wo && e < 3 * 128) {                                         // dWo[o][n] = U1[n][o] + Fc_4[n][:] . U2[:][o] + bc_4[n] dbo[o]
            const int o = e / 128, n = e - o * 128;
            float v = S[W_U1 + n * 16 + o];
            for (int j = 0; j < 32; ++j) v = fmaf(a.P.c_Wc[4][n * 32 + j], S[W_U2 + j * 16 + o], v);
            v = fmaf(a.P.c_bc[4][n], S[W_DBO + o], v);
            a.G.c_Wo[e] = v;
        }
        if (a.G.c_bo && e < 3) a.G.c_bo[e] = S[W_DBO + e];
    } else if (job == 6) {
        if (a.rel) {
            if (a.G.c_N1 && e < 128 * 52) a.G.c_N1[e] = S[W_N1 + (e / 52) * 64 + e % 52];
            if (a.G.c_N2 && e < 32 * 128) a.G.c_N2[e] = S[W_N2T + (e % 128) * 32 + e / 128];
            if (a.G.c_n1b && e < 128) a.G.c_n1b[e] = S[W_DB1 + e];
            if (a.G.c_n2b && e < 32) a.G.c_n2b[e] = S[W_DB2 + e];
            if (a.G.c_Brel && e < 30) {
                float s = 0.f;
                for (int c = 0; c < a.n_cta_a; ++c) s += a.part_brel[c * 32 + e];
                a.G.c_Brel[e] = s;
            }
        } else {
            if (a.G.c_N1 && e < 128 * 52) a.G.c_N1[e] = 0.f;
            if (a.G.c_N2 && e < 32 * 128) a.G.c_N2[e] = 0.f;
            if (a.G.c_n1b && e < 128) a.G.c_n1b[e] = 0.f;
            if (a.G.c_n2b && e < 32) a.G.c_n2b[e] = 0.f;
            if (a.G.c_Brel && e < 30) a.G.c_Brel[e] = 0.f;
        }
    } else if (job == 7) {
        if (a.d_affine && a.aff && e < 12) {
            float s = 0.f;
            for (long long r = 0; r < a.aff_rows; ++r) s += a.aff[r * 12 + e];
            a.d_affine[e] = s;
        }
    }
}

}  // namespace wgt
}  // namespace psl

using namespace psl;

static long long wgrad_grid(long long m) {
    const long long n_units = (m + 127) / 128 * wgt::UPT;
    return n_units < sm_count() ? n_units : sm_count();
}

extern "C" size_t psl_wgrad_tc_ws_floats(int64_t m) { return (size_t)wgt::W_TOTAL * (size_t)(wgrad_grid(m) + 1) + 64; }

// weight gradients of the colour branch from the buffers left by psl_color_fwd_tc(tsave) and psl_color_bwd_tc(want_wgrad = 1).
// grads: only the c_* entries are written (NULL = skip).  n_cta_bwd: *grid_out of psl_color_bwd_tc.
extern "C" int psl_wgrad_tc(const psl_decode_cfg* cfg, const psl_decoder_params* P, const float* pos, int64_t m, const int32_t* I,
                            const float* cloud_pos, const float* col_feats, const float* tsave, const float* tbwd,
                            int32_t n_cta_bwd, const psl_decoder_grads* G, float* d_exposure_affine, float* ws,
                            size_t ws_floats, psl_stream_t stream) {
    PSL_REQUIRE(cfg && P && pos && I && col_feats && tsave && tbwd && G && ws, "NULL argument");
    PSL_REQUIRE(!cfg->encode_rel_pos || cloud_pos, "rel-pos encoding needs cloud_pos");
    if (m == 0) return 0;
    PSL_REQUIRE(ws_floats >= psl_wgrad_tc_ws_floats(m), "workspace too small");
    cudaStream_t st = as_stream(stream);
    wgt::Args a{};
    a.cfg = *cfg; a.P = *P; a.pos = pos; a.m = m; a.I = I; a.cloud_pos = cloud_pos; a.col_feats = col_feats;
    a.tsave = tsave; a.tbwd = tbwd; a.partial = ws;
    const long long n_tiles = (m + 127) / 128, grid = wgrad_grid(m);
    PSL_CHECK_CUDA(cudaFuncSetAttribute(wgt::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, wgt::SB_TOTAL));
    float* sums = ws + (size_t)wgt::W_TOTAL * grid;
    {
        TimingScope ts(T_WGRAD_TC, st);
        wgt::k_wgrad_tc<<<(unsigned)grid, wgt::NTHR, wgt::SB_TOTAL, st>>>(a, n_tiles);
        PSL_CHECK_CUDA(cudaGetLastError());
    }
    TimingScope ts(T_REDUCE, st, 2);
    wgt::k_wgrad_reduce<<<(wgt::W_TOTAL + 255) / 256, 256, 0, st>>>(ws, (int)grid, sums);
    wgt::FinArgs f{};
    f.P = *P; f.G = *G; f.sums = sums; f.rel = cfg->encode_rel_pos;
    const TBwd BL = tbwd_layout(m, cfg->encode_rel_pos);
    f.part_brel = tbwd + BL.total; f.n_cta_a = n_cta_bwd;
    f.aff = (cfg->rgb_mode == PSL_RGB_AFFINE_SIGMOID) ? tbwd + BL.aff : nullptr;
    f.aff_rows = n_tiles * 128; f.d_affine = d_exposure_affine;
    wgt::k_wgrad_finalize<<<dim3((128 * 168 + 255) / 256, 8), 256, 0, st>>>(f);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
