// Backward DATA path of the colour branch, round 2: the gradient planes as f16 hi/lo pairs (kind::f16 MMAs, see psl_color_h2.cu).
//   dL/d(rgb) -> output layer -> 5 trunk layers -> dL/dc, dL/d(embedding) -> per-neighbour MLP (8 slots) ->
//   dL/d col_feats[I] (per pair), dL/d(IDW weights), dL/d(pos) (Fourier + rel-pos parts), dL/d Brel.
// Same mathematics, buffers and outputs as psl_color_bwd_tc(_w16).cu; what changes:
//   * operands: weights x 64 as f16 hi/lo planes (half the bytes streamed per tile, K = 16 per MMA at twice the tf32 rate);
//     gradients carry a PER-ROW power-of-two scale s_r, chosen from the row's dL/d(rgb) so that its largest entry lands in
//     [256, 512): the backward is linear in the gradient, so the scale rides through every layer (dz = dh * softplus' is
//     element-wise) and is divided out where a true gradient leaves the kernel.  f16 then covers 2^7 of growth and 2^-22 of
//     decay per row; rows whose gradient is exactly zero use s_r = 1.
//   * weight ring: 8 stages of 16 KB (two K = 16 steps of a 128-row matrix) instead of 2 x 32 KB, so the producer runs a whole
//     layer ahead of the single tile in flight;
//   * one mbarrier arrival per worker warp, bare try_wait loops, both 16-column TMEM loads of an epilogue issued before the first
//     value is used.
// TMEM (one tile): D_H [0,128) | A hi [128,192) | A lo [192,256) | D_C [256,288) | D_E / D_X [320,384).
//
// Autograd semantics: src/conv_onet/models/decoder.py:341-449.
#include "psl_color_bwd_tc.cuh"
#include "psl_color_tc.cuh"

namespace psl {
namespace cbh {

using namespace cbt;             // Args, mat_n, mat_k, BV_*, sp_grad_fast

constexpr int NWORKER = 512, NTHREADS = 576;
constexpr float W_SCALE = 64.0f, INV_W = 1.0f / 64.0f;
constexpr uint32_t T_DH = 0, T_AH = 128, T_AL = 192, T_DC = 256, T_DE = 320;

// ---- operand blob (bytes): [vectors | N2T units | N1T units | 12 streamed matrices as K = 16 units] -----------------------------
__host__ __device__ constexpr int m_units(int q) { return mat_k(q) / 16; }
__host__ __device__ constexpr int m_unit(int q) { return mat_n(q) * 64; }                 // [hi N*32 B | lo N*32 B]
__host__ __device__ constexpr int m_off(int q) {
    int o = 0;
    for (int i = 0; i < q; ++i) o += m_units(i) * m_unit(i);
    return o;
}
constexpr int HB_VEC = 0;
constexpr int HB_N2T = 512;                            // BV_SIZE floats = 448 B, padded
constexpr int HB_N1T = HB_N2T + 2 * 128 * 64;          // N2T: 128 rows (hid) x K = 32 (c): 2 units
constexpr int HB_MAT = HB_N1T + 8 * 64 * 64;           // N1T: 64 rows (j) x K = 128 (hid): 8 units
constexpr int HB_TOTAL = HB_MAT + m_off(NMAT);
static_assert(BV_SIZE * 4 <= 512 && HB_MAT % 16 == 0, "blob layout");

// ---- shared memory -----------------------------------------------------------------------------------------------------------
constexpr int NSTAGE = 8, STAGE_BYTES = 16384;
constexpr int S_NBRW = 0;                              // N2T 16 KB | N1T 32 KB
constexpr int S_RING = 49152;
constexpr int S_IDX = S_RING + NSTAGE * STAGE_BYTES;   // idx[8][128] int: neighbour ids of the tile's rows (-1 = no weight)
constexpr int S_WN = S_IDX + 4096;                     // wn[8][128] float: normalised IDW weights
constexpr int S_VEC = S_WN + 4096;
constexpr int S_AFF = S_VEC + 512;
constexpr int S_RED = S_AFF + 64;
constexpr int S_BAR = S_RED + 4 * 32 * 4;
constexpr int S_TOTAL = S_BAR + 32 * 8;
static_assert(S_TOTAL <= 227 * 1024, "shared memory over budget");

struct PackArgs { psl_decoder_params P; const float* fold; unsigned char* hb; };

__device__ __forceinline__ void put_pair(unsigned char* base, int N, int n, int k, float v0, float v1) {   // k even
    uint32_t hi, lo;
    tc::split_h2_f16(v0 * W_SCALE, v1 * W_SCALE, hi, lo);
    unsigned char* unit = base + (k >> 4) * N * 64;
    const uint32_t o = tc::canon_off_h(n, k & 15, N) * 2;
    *reinterpret_cast<uint32_t*>(unit + o) = hi;
    *reinterpret_cast<uint32_t*>(unit + N * 32 + o) = lo;
}

// same matrices as cbt::k_bwd_pack (rows of the fp32 folded layers, transposed), two k per thread
__global__ void k_bwd_pack_h(PackArgs a) {
    const int job = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const float* F = a.fold;
    const int ld = ctc::FOLD_LD;
    unsigned char* M = a.hb + HB_MAT;
    if (job == 0) {                  // WoT [n = 128][o = 16], GoT [c = 32][o = 16] from folded layer 5 (rows o < 3 live)
        if (e < 128 * 8) { const int n = e >> 3, o = 2 * (e & 7); put_pair(M + m_off(0), 128, n, o, o < 3 ? F[((size_t)5 * 128 + o) * ld + n] : 0.f, o + 1 < 3 ? F[((size_t)5 * 128 + o + 1) * ld + n] : 0.f); }
        if (e < 32 * 8) { const int c = e >> 3, o = 2 * (e & 7); put_pair(M + m_off(1), 32, c, o, o < 3 ? F[((size_t)5 * 128 + o) * ld + 128 + c] : 0.f, o + 1 < 3 ? F[((size_t)5 * 128 + o + 1) * ld + 128 + c] : 0.f); }
    } else if (job <= 4) {           // l = 5 - job: LaT_l [k_in][n], GT_l [c][n], (l == 3) LeT_3 [j][n]
        const int l = 5 - job;
        const int qa = l == 4 ? 2 : (l == 3 ? 4 : (l == 2 ? 7 : 9));
        const int ne = (l == 3) ? 40 : 0;
        const float* R = F + (size_t)l * 128 * ld;
        if (e < 128 * 64) { const int kin = e >> 6, n = 2 * (e & 63); put_pair(M + m_off(qa), 128, kin, n, R[(size_t)n * ld + ne + kin], R[(size_t)(n + 1) * ld + ne + kin]); }
        if (e < 32 * 64) { const int c = e >> 6, n = 2 * (e & 63); put_pair(M + m_off(qa + 1), 32, c, n, R[(size_t)n * ld + ne + 128 + c], R[(size_t)(n + 1) * ld + ne + 128 + c]); }
        if (l == 3 && e < 48 * 64) { const int j = e >> 6, n = 2 * (e & 63); put_pair(M + m_off(6), 48, j, n, j < 40 ? R[(size_t)n * ld + j] : 0.f, j < 40 ? R[(size_t)(n + 1) * ld + j] : 0.f); }
    } else if (job == 5) {           // LeT_0
        if (e < 48 * 64) { const int j = e >> 6, n = 2 * (e & 63); put_pair(M + m_off(11), 48, j, n, j < 40 ? F[(size_t)n * ld + j] : 0.f, j < 40 ? F[(size_t)(n + 1) * ld + j] : 0.f); }
    } else if (job == 6) {           // N2T [hid 128][c 32] = N2[c][hid] ; N1T [j 64][hid 128] = N1[hid][j]
        if (e < 128 * 16) { const int hid = e >> 4, c = 2 * (e & 15); put_pair(a.hb + HB_N2T, 128, hid, c, a.P.c_N2[c * 128 + hid], a.P.c_N2[(c + 1) * 128 + hid]); }
        if (e < 64 * 64) { const int j = e >> 6, hid = 2 * (e & 63); put_pair(a.hb + HB_N1T, 64, j, hid, j < 52 ? a.P.c_N1[hid * 52 + j] : 0.f, j < 52 ? a.P.c_N1[(hid + 1) * 52 + j] : 0.f); }
    } else {
        float* v = reinterpret_cast<float*>(a.hb + HB_VEC);
        if (e < 30) v[BV_BREL + (e / 10) * 12 + (e % 10)] = a.P.c_Brel[e];
        if (e < 60) v[BV_BC + e] = a.P.c_B[e];
    }
}

__device__ __forceinline__ void st8u(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void signal(uint64_t* a_ready) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(a_ready);
}
// 32 consecutive TMEM columns, both 16-column loads in flight before the first value is used
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r0[16], r1[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
                   "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                   "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                 : "r"(taddr + 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(r0[j]); v[16 + j] = __uint_as_float(r1[j]); }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_color_bwd_h2(Args a, const unsigned char* __restrict__ hb, long long n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sVec = reinterpret_cast<float*>(smem + S_VEC);
    float* sAff = reinterpret_cast<float*>(smem + S_AFF);
    float* sRed = reinterpret_cast<float*>(smem + S_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_BAR);
    uint64_t* full = bars;                  // [8]
    uint64_t* empty = bars + 8;             // [8]
    uint64_t* nbrw_full = bars + 16;
    uint64_t* a_ready = bars + 17;          // workers -> MMA (16 arrivals: one per worker warp)
    uint64_t* d_ready = bars + 18;          // MMA -> workers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool rel = a.cfg.encode_rel_pos != 0;
    const TSave TL = tsave_layout(a.m, a.cfg.encode_rel_pos);
    const TBwd BL = tbwd_layout(a.m, a.cfg.encode_rel_pos);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(nbrw_full, 1); tc::mbar_init(a_ready, NWORKER / 32); tc::mbar_init(d_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < BV_SIZE; i += NTHREADS) sVec[i] = reinterpret_cast<const float*>(hb + HB_VEC)[i];
    if (threadIdx.x < 12) sAff[threadIdx.x] = a.affine ? a.affine[threadIdx.x] : 0.f;
    if (threadIdx.x < 128) sRed[threadIdx.x] = 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // =============================== bulk-copy producer: 12 matrices per tile, two K = 16 units per stage =================
        if (lane == 0) {
            if (rel) {
                tc::mbar_expect_tx(nbrw_full, 49152);
                for (int i = 0; i < 3; ++i) tc::bulk_g2s(smem + S_NBRW + i * 16384, hb + HB_N2T + i * 16384, 16384, nbrw_full);
            }
            uint32_t cnt = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int q = 0; q < NMAT; ++q) {
                    const int nu = m_units(q), ub = m_unit(q);
                    for (int j = 0; j < nu; j += 2, ++cnt) {
                        const int st = cnt & (NSTAGE - 1);
                        const uint32_t bytes = (uint32_t)(min(2, nu - j) * ub);
                        tc::mbar_wait_p(&empty[st], ((cnt / NSTAGE) & 1) ^ 1);
                        tc::mbar_expect_tx(&full[st], bytes);
                        tc::bulk_g2s(smem + S_RING + st * STAGE_BYTES, hb + HB_MAT + m_off(q) + j * ub, bytes, &full[st]);
                    }
                }
        }
    } else if (warp == 17) {
        // =============================== MMA issuer ==========================================================================
        if (lane == 0) {
            uint32_t pa = 0, cnt = 0;
            const uint32_t hi128 = tc::desc_hi(128);
            const uint32_t ring0 = tc::smem_u32(smem + S_RING);
            // one streamed matrix: A planes (K = mat_k(q) gradient columns in TMEM), B from the ring, D columns d
            auto run_mat = [&](int q, uint32_t d, uint32_t first_acc) {
                const int N = mat_n(q), nu = m_units(q);
                const uint32_t idesc = tc::make_idesc_f16(128, N, tc::FMT_F16, tc::FMT_F16), lbo = (uint32_t)N * 16u, ub16 = (uint32_t)m_unit(q) >> 4;
                for (int j = 0; j < nu; ++j) {
                    const int u = j & 1, st = cnt & (NSTAGE - 1);
                    if (u == 0) { tc::mbar_wait_p(&full[st], (cnt / NSTAGE) & 1); tc::fence_after_sync(); }
                    const uint32_t b_ = tc::desc_lo(ring0 + st * STAGE_BYTES, lbo) + (uint32_t)u * ub16;
                    const uint64_t bh = tc::desc_of(b_, hi128), bl = tc::desc_of(b_ + (ub16 >> 1), hi128);
                    const uint32_t ac = j == 0 ? first_acc : 1u;
                    tc::mma_f16_ts(tmem + d, tmem + T_AH + 8 * j, bh, idesc, ac);
                    tc::mma_f16_ts(tmem + d, tmem + T_AL + 8 * j, bh, idesc, 1);
                    tc::mma_f16_ts(tmem + d, tmem + T_AH + 8 * j, bl, idesc, 1);
                    if (u == 1 || j == nu - 1) { tc::mma_commit(&empty[st]); ++cnt; }
                }
            };
            if (rel) tc::mbar_wait_p(nbrw_full, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                tc::mbar_wait_p(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(0, T_DH, 0);                       // dh_4 = dout Wo
                run_mat(1, T_DC, 0);                       // dc   = dout G_out
                tc::mma_commit(d_ready);
                int q = 2;
                for (int l = 4; l >= 1; --l) {
                    tc::mbar_wait_p(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                    run_mat(q++, T_DH, 0);                 // dh_{l-1} = dz_l La_l
                    run_mat(q++, T_DC, 1);                 // dc      += dz_l G_l
                    if (l == 3) run_mat(q++, T_DE, 0);     // de       = dz_3 Le_3
                    tc::mma_commit(d_ready);
                }
                tc::mbar_wait_p(a_ready, pa); pa ^= 1; tc::fence_after_sync();
                run_mat(11, T_DE, 1);                      // de += dz_0 Le_0
                tc::mma_commit(d_ready);
                if (rel) {
                    const uint32_t id128 = tc::make_idesc_f16(128, 128, tc::FMT_F16, tc::FMT_F16), id64 = tc::make_idesc_f16(128, 64, tc::FMT_F16, tc::FMT_F16);
                    const uint32_t n2 = tc::desc_lo(tc::smem_u32(smem + S_NBRW), 2048), n1 = tc::desc_lo(tc::smem_u32(smem + S_NBRW + 16384), 1024);
                    for (int k = 0; k < 8; ++k) {
                        tc::mbar_wait_p(a_ready, pa); pa ^= 1; tc::fence_after_sync();
#pragma unroll
                        for (int j = 0; j < 2; ++j) {            // dh1 = df N2   (K = 32) -> D_H
                            const uint64_t bh = tc::desc_of(n2 + j * 512u, hi128), bl = tc::desc_of(n2 + j * 512u + 256u, hi128);
                            tc::mma_f16_ts(tmem + T_DH, tmem + T_AH + 8 * j, bh, id128, j > 0);
                            tc::mma_f16_ts(tmem + T_DH, tmem + T_AL + 8 * j, bh, id128, 1);
                            tc::mma_f16_ts(tmem + T_DH, tmem + T_AH + 8 * j, bl, id128, 1);
                        }
                        tc::mma_commit(d_ready);
                        tc::mbar_wait_p(a_ready, pa); pa ^= 1; tc::fence_after_sync();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {            // dx = dz1 N1   (K = 128) -> D_X (64 columns)
                            const uint64_t bh = tc::desc_of(n1 + j * 256u, hi128), bl = tc::desc_of(n1 + j * 256u + 128u, hi128);
                            tc::mma_f16_ts(tmem + T_DE, tmem + T_AH + 8 * j, bh, id64, j > 0);
                            tc::mma_f16_ts(tmem + T_DE, tmem + T_AL + 8 * j, bh, id64, 1);
                            tc::mma_f16_ts(tmem + T_DE, tmem + T_AH + 8 * j, bl, id64, 1);
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
        int* sIdx = reinterpret_cast<int*>(smem + S_IDX);
        float* sWn = reinterpret_cast<float*>(smem + S_WN);
        float brel_acc = 0.f;                                    // lane e < 30 of the quarter-0 warps: d Brel[e/10][e%10]
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long m = tile * TM + r;
            const bool inb = m < a.m;
            float px = 0.f, py = 0.f, pz = 0.f;
            bool has = false;
            if (inb) {
                px = a.pos[m * 3]; py = a.pos[m * 3 + 1]; pz = a.pos[m * 3 + 2];
                has = a.nnum[m] >= a.cfg.min_nn;
            }
            worker_bar();                                    // the previous tile's readers of sIdx / sWn are done
            if (q == 0) {                                    // IDW weights once per row -> shared memory
                float sum = 0.f, w[8], tle = -1.f;
                int idx[8];
                if (inb) tle = thr_le_of(a.r2 ? a.r2[m / a.cfg.r2_group] : a.cfg.r2_scalar);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    idx[k] = inb ? a.I[m * 8 + k] : -1;
                    w[k] = inb ? idw_raw(a.D[m * 8 + k], idx[k], tle, a.cfg.weighting) : 0.f;
                    sum += fabsf(w[k]);
                }
                const float den = fmaxf(sum, 1e-12f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wnk = __fdiv_rn(w[k], den);
                    const int idk = w[k] == 0.f ? -1 : idx[k];
                    sWn[k * 128 + r] = wnk;
                    sIdx[k * 128 + r] = idk;
                    if (inb && a.wn_out) a.wn_out[m * 8 + k] = (has && idk >= 0) ? wnk : 0.f;
                }
            }
            worker_bar();
            float dpx = 0.f, dpy = 0.f, dpz = 0.f;
            // ---- dL/d(colour output): every quarter derives the row's scale from it, quarter 0 builds the A operand ----------
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (inb) {
                const float4 dr = reinterpret_cast<const float4*>(a.d_raw)[m];
                g0 = dr.x; g1 = dr.y; g2 = dr.z;
                if (a.cfg.rgb_mode != PSL_RGB_RAW) {
                    const float4 rv = reinterpret_cast<const float4*>(a.raw)[m];
                    g0 *= rv.x * (1.0f - rv.x); g1 *= rv.y * (1.0f - rv.y); g2 *= rv.z * (1.0f - rv.z);
                }
            }
            if (q == 0 && a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                float* af = a.tbwd + BL.aff + (tile * 128 + r) * 12;           // d rot[a][b] = out_a g_b ; d trans = g
                if (inb) {
                    const float4 op = *reinterpret_cast<const float4*>(a.tsave + TL.outpre + (tile * 128 + r) * 4);
                    af[0] = op.x * g0; af[1] = op.x * g1; af[2] = op.x * g2;
                    af[3] = op.y * g0; af[4] = op.y * g1; af[5] = op.y * g2;
                    af[6] = op.z * g0; af[7] = op.z * g1; af[8] = op.z * g2;
                    af[9] = g0; af[10] = g1; af[11] = g2;
                } else {
#pragma unroll
                    for (int j = 0; j < 12; ++j) af[j] = 0.f;
                }
            }
            if (a.cfg.rgb_mode == PSL_RGB_AFFINE_SIGMOID) {
                const float o0 = sAff[0] * g0 + sAff[1] * g1 + sAff[2] * g2;
                const float o1 = sAff[3] * g0 + sAff[4] * g1 + sAff[5] * g2;
                const float o2 = sAff[6] * g0 + sAff[7] * g1 + sAff[8] * g2;
                g0 = o0; g1 = o1; g2 = o2;
            }
            // per-row power-of-two scale: largest |g| -> [256, 512); the same value in all four quarter threads of the row
            float s_r = 1.0f;
            {
                const float gm = fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2)));
                if (gm > 0.f && gm < 1e30f) {
                    int e = 8 - (((int)(__float_as_uint(gm) >> 23) & 0xff) - 127);
                    e = e < -100 ? -100 : (e > 100 ? 100 : e);
                    s_r = __uint_as_float((uint32_t)(e + 127) << 23);
                }
            }
            const float inv_true = INV_W / s_r;                  // accumulator -> true gradient
            if (q == 0) {
                if (a.want_wgrad) {
                    a.tbwd[BL.doutT + (tile * 16 + 0) * 128 + r] = g0; a.tbwd[BL.doutT + (tile * 16 + 1) * 128 + r] = g1;
                    a.tbwd[BL.doutT + (tile * 16 + 2) * 128 + r] = g2;
#pragma unroll
                    for (int j = 3; j < 16; ++j) a.tbwd[BL.doutT + (tile * 16 + j) * 128 + r] = 0.f;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { hi[j] = 0u; lo[j] = 0u; }
                tc::split_h2_f16(g0 * s_r, g1 * s_r, hi[0], lo[0]);
                tc::split_h2_f16(g2 * s_r, 0.f, hi[1], lo[1]);
                st8u(lb + T_AH, hi);
                st8u(lb + T_AL, lo);
            }
            signal(a_ready);
            // ---- trunk: dh_l -> (store) -> dz_l = dh_l * softplus'(z_l) -> A planes: this thread's columns 32q .. 32q+31 ------------
#pragma unroll 1
            for (int l = 4; l >= 0; --l) {
                // the saved pre-activations are requested BEFORE waiting for the layer's MMAs
                const long long off0 = (((long long)l * n_tiles + tile) * 128 + 32 * q) * 128 + r;
                float zc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) zc[j] = a.tsave[TL.zT + off0 + j * 128];
                tc::mbar_wait_p(d_ready, pd); pd ^= 1; tc::fence_after_sync();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float v[16];
                    tc::tmem_ld16(lb + T_DH + 32 * q + 16 * c, v);
                    if (a.want_wgrad) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) a.tbwd[BL.dhT + off0 + (16 * c + j) * 128] = v[j] * inv_true;
                    }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e0 = 16 * c + 2 * j;
                        tc::split_h2_f16(v[2 * j] * INV_W * sp_grad_fast(zc[e0]), v[2 * j + 1] * INV_W * sp_grad_fast(zc[e0 + 1]), hi[j], lo[j]);
                    }
                    st8u(lb + T_AH + 16 * q + 8 * c, hi);
                    st8u(lb + T_AL + 16 * q + 8 * c, lo);
                }
                signal(a_ready);
            }
            // ---- dc, de -----------------------------------------------------------------------------------------------------
            tc::mbar_wait_p(d_ready, pd); pd ^= 1; tc::fence_after_sync();
            // dc stays in TMEM (D_C is not written again before the next tile's output layer): the quarters that need it re-load
            // their columns where they use them
            if (q == 3) {
                float dcv[32];
                ld32(lb + T_DC, dcv);
#pragma unroll
                for (int j = 0; j < 32; ++j) dcv[j] = has ? dcv[j] * inv_true : 0.f;
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
                            const int idk = sIdx[k * 128 + r];
                            if (idk >= 0) {
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    const float4 f4 = __ldg(reinterpret_cast<const float4*>(a.col_feats + (size_t)idk * 32) + g);
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
                ld32(lb + T_DE, e0);
                tc::tmem_ld16(lb + T_DE + 32, e1);
                const float x = __fmul_rn(kTwoPi, px), y = __fmul_rn(kTwoPi, py), z = __fmul_rn(kTwoPi, pz);
                float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                for (int j = 0; j < 20; ++j) {
                    float sn, cs;
                    sincos_embed(fmaf(z, Bc[40 + j], fmaf(y, Bc[20 + j], x * Bc[j])), &sn, &cs);
                    const float dcos = j < 12 ? e0[20 + j] : e1[j - 12];
                    const float da = (e0[j] * cs - dcos * sn) * inv_true;
                    gx = fmaf(da, Bc[j], gx); gy = fmaf(da, Bc[20 + j], gy); gz = fmaf(da, Bc[40 + j], gz);
                }
                dpx += kTwoPi * gx; dpy += kTwoPi * gy; dpz += kTwoPi * gz;
            }
            // ---- neighbour MLP backward ------------------------------------------------------------------------------------
            if (rel) {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    const int id = sIdx[k * 128 + r];
                    // df_k = wn_k dc -> A planes (32 columns: quarters 1 and 2, 16 each); d wn_k = dc . f_k (quarter 3)
                    if (q == 1 || q == 2) {
                        float v[16];
                        tc::tmem_ld16(lb + T_DC + 16 * (q - 1), v);
                        uint32_t hi[8], lo[8];
                        const float sc = has ? sWn[k * 128 + r] * INV_W : 0.f;   // dc accumulator / 64 = s_r dc ; times the IDW weight
#pragma unroll
                        for (int j = 0; j < 8; ++j) tc::split_h2_f16(sc * v[2 * j], sc * v[2 * j + 1], hi[j], lo[j]);
                        st8u(lb + T_AH + 8 * (q - 1), hi);
                        st8u(lb + T_AL + 8 * (q - 1), lo);
                    } else if (q == 3 && a.dwn_col) {
                        float dcv[32];
                        ld32(lb + T_DC, dcv);
                        float dot = 0.f;
                        const float4* fr = reinterpret_cast<const float4*>(a.tsave + TL.f + ((tile * 128 + r) * 8 + k) * 32);
                        const float sc = has ? inv_true : 0.f;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 f4 = fr[g];
                            dot = fmaf(sc * dcv[4 * g], f4.x, dot); dot = fmaf(sc * dcv[4 * g + 1], f4.y, dot);
                            dot = fmaf(sc * dcv[4 * g + 2], f4.z, dot); dot = fmaf(sc * dcv[4 * g + 3], f4.w, dot);
                        }
                        if (inb) a.dwn_col[m * 8 + k] = dot;
                    }
                    signal(a_ready);
                    // z1 of this neighbour: requested before waiting for the MMAs
                    const long long off0 = ((tile * 8 + k) * 128 + 32 * q) * 128 + r;
                    float zc[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) zc[j] = a.tsave[TL.z1T + off0 + j * 128];
                    // dh1 -> dz1 = dh1 * softplus'(z1) (store for dN1) -> A planes
                    tc::mbar_wait_p(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    {
                        const float is = 1.0f / s_r;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            float v[16];
                            tc::tmem_ld16(lb + T_DH + 32 * q + 16 * c, v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = v[j] * INV_W * sp_grad_fast(zc[16 * c + j]);      // s_r dz1
                            if (a.want_wgrad) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) a.tbwd[BL.dz1T + off0 + (16 * c + j) * 128] = v[j] * is;
                            }
                            uint32_t hi[8], lo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) tc::split_h2_f16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                            st8u(lb + T_AH + 16 * q + 8 * c, hi);
                            st8u(lb + T_AL + 16 * q + 8 * c, lo);
                        }
                    }
                    signal(a_ready);
                    // dx: columns [0,20) rel-pos embedding, [20,52) feature gradient of the pair
                    tc::mbar_wait_p(d_ready, pd); pd ^= 1; tc::fence_after_sync();
                    const bool live = has && id >= 0 && inb;
                    if (q == 0) {                          // rel-pos embedding: d arg, d rel (-> -d pos), d Brel
                        float dx[24];
                        {
                            float t16[16], t8[8];
                            tc::tmem_ld16(lb + T_DE, t16);
                            tc::tmem_ld8(lb + T_DE + 16, t8);
#pragma unroll
                            for (int j = 0; j < 16; ++j) dx[j] = t16[j] * inv_true;
#pragma unroll
                            for (int j = 0; j < 8; ++j) dx[16 + j] = t8[j] * inv_true;
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
                        tc::tmem_ld16(lb + T_DE + 16, t16);
                        if (a.d_colpair && inb) {
                            const float sc = live ? inv_true : 0.f;
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32);
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                dst[g] = make_float4(sc * t16[4 + 4 * g], sc * t16[5 + 4 * g], sc * t16[6 + 4 * g], sc * t16[7 + 4 * g]);
                        }
                    } else if (q == 2) {                   // feature gradient columns 12-27 = dx[32..47]
                        float t16[16];
                        tc::tmem_ld16(lb + T_DE + 32, t16);
                        if (a.d_colpair && inb) {
                            const float sc = live ? inv_true : 0.f;
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32) + 3;
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                dst[g] = make_float4(sc * t16[4 * g], sc * t16[4 * g + 1], sc * t16[4 * g + 2], sc * t16[4 * g + 3]);
                        }
                    } else {                               // feature gradient columns 28-31 = dx[48..51]
                        float t8[8];
                        tc::tmem_ld8(lb + T_DE + 48, t8);
                        if (a.d_colpair && inb) {
                            const float sc = live ? inv_true : 0.f;
                            float4* dst = reinterpret_cast<float4*>(a.d_colpair + ((size_t)m * 8 + k) * 32) + 7;
                            dst[0] = make_float4(sc * t8[0], sc * t8[1], sc * t8[2], sc * t8[3]);
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

}  // namespace cbh
}  // namespace psl

using namespace psl;

extern "C" size_t psl_h2_bwd_blob_bytes(void) { return (size_t)cbh::HB_TOTAL; }

// f16 hi/lo images of the transposed / folded colour-branch weights for psl_color_bwd_h2; `tc_blob` must hold the result of
// psl_tc_pack_params for the SAME parameters (its folded fp32 rows are reused).
extern "C" int psl_h2_bwd_pack_params(const psl_decoder_params* P, const float* tc_blob, void* h2_bwd_blob, psl_stream_t stream) {
    PSL_REQUIRE(P && tc_blob && h2_bwd_blob, "NULL argument");
    cudaStream_t st = as_stream(stream);
    cbh::PackArgs pa{*P, tc_blob + ctc::TB_TOTAL, static_cast<unsigned char*>(h2_bwd_blob)};
    TimingScope ts(T_PACK, st, 1);
    cbh::k_bwd_pack_h<<<dim3(32, 8), 256, 0, st>>>(pa);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// data-gradient pass of the colour branch on f16 planes: contract of psl_color_bwd_tc (include/pointslam_b200.h)
extern "C" int psl_color_bwd_h2(const psl_decode_cfg* cfg, const void* h2_bwd_blob, const float* pos, int64_t m, const int32_t* I,
                                const float* D, const int32_t* nnum, const double* r2, const float* cloud_pos, const float* col_feats,
                                const float* exposure_affine, const float* raw, const float* d_raw, const float* tsave, float* tbwd,
                                float* d_colpair, float* wn_out, float* dwn_col, float* dpos_col, int32_t want_wgrad, int32_t* grid_out,
                                psl_stream_t stream) {
    PSL_REQUIRE(cfg && h2_bwd_blob && pos && I && D && nnum && col_feats && raw && d_raw && tsave && tbwd, "NULL argument");
    PSL_REQUIRE(!cfg->encode_rel_pos || cloud_pos, "rel-pos encoding needs cloud_pos");
    if (m == 0) return 0;
    cbt::Args a{};
    a.cfg = *cfg; a.pos = pos; a.m = m; a.I = I; a.D = D; a.nnum = nnum; a.r2 = r2;
    a.cloud_pos = cloud_pos; a.col_feats = col_feats; a.affine = exposure_affine; a.raw = raw; a.d_raw = d_raw;
    a.tsave = tsave; a.tbwd = tbwd; a.d_colpair = d_colpair; a.wn_out = wn_out; a.dwn_col = dwn_col; a.dpos_col = dpos_col;
    a.part_brel = tbwd + tbwd_layout(m, cfg->encode_rel_pos).total;
    a.want_wgrad = want_wgrad;
    const long long n_tiles = (m + cbt::TM - 1) / cbt::TM;
    // per launch: the attribute belongs to the device the launch goes to
    PSL_CHECK_CUDA(cudaFuncSetAttribute(cbh::k_color_bwd_h2, cudaFuncAttributeMaxDynamicSharedMemorySize, cbh::S_TOTAL));
    const long long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (grid_out) *grid_out = (int32_t)grid;
    TimingScope ts(T_COLOR_BWD_TC, as_stream(stream));
    cbh::k_color_bwd_h2<<<(unsigned)grid, cbh::NTHREADS, cbh::S_TOTAL, as_stream(stream)>>>(a, static_cast<const unsigned char*>(h2_bwd_blob), n_tiles);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
