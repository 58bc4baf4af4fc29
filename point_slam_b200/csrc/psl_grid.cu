// K0: spatial hash build.   K1: fused ray-march + exact radius-kNN (warp per ray, shared-memory
// staging of candidate cells, register top-8 lists merged with warp shuffles).
//
// Replaces faiss GpuIndexIVFFlat train/add/search as used by src/neural_point.py:37-41,161-164,189-197
// and the sample placement of src/utils/Renderer.py:133-174.
#include <cub/device/device_merge.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <float.h>

#include "psl_common.cuh"
#include "psl_grid.cuh"

namespace psl {

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
__global__ void k_cell_keys(const float* __restrict__ pos, int n, float inv_cell, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    keys[i] = cell_key(cell_coord(x, inv_cell), cell_coord(y, inv_cell), cell_coord(z, inv_cell));
    vals[i] = (uint32_t)i;
}

__global__ void k_gather_sorted(const float* __restrict__ pos, const uint32_t* __restrict__ order,
                                const uint64_t* __restrict__ keys, int n, float4* __restrict__ out,
                                unsigned long long* __restrict__ n_cells) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int head = 0;
    if (j < n) {
        const uint32_t i = order[j];
        out[j] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __uint_as_float(i));
        head = (j == 0) || (keys[j] != keys[j - 1]);
    }
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_cells, (unsigned long long)__popc(b));
}

// ---- incremental rebuild (append): only the NEW points are keyed and sorted, then merged into the sorted copy ----------------
__global__ void k_cell_keys_new(const float* __restrict__ pos, int n_old, int k_new, float inv_cell, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k_new) return;
    const float* p = pos + 3 * (size_t)(n_old + i);
    keys[i] = cell_key(cell_coord(p[0], inv_cell), cell_coord(p[1], inv_cell), cell_coord(p[2], inv_cell));
    vals[i] = (uint32_t)(n_old + i);
}
__global__ void k_gather_new(const float* __restrict__ pos, const uint32_t* __restrict__ order, int k_new, float4* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k_new) return;
    const uint32_t i = order[j];
    out[j] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __uint_as_float(i));
}
__global__ void k_count_cells(const uint64_t* __restrict__ keys, int n, unsigned long long* __restrict__ n_cells) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int head = j < n && (j == 0 || keys[j] != keys[j - 1]);
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_cells, (unsigned long long)__popc(b));
}

__device__ __forceinline__ uint32_t table_find_or_insert(uint64_t* tkeys, uint32_t mask, uint64_t key) {
    uint32_t slot = (uint32_t)mix64(key) & mask;
    for (;;) {
        const unsigned long long prev =
            atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + slot), (unsigned long long)kEmptyKey,
                      (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) return slot;
        slot = (slot + 1) & mask;
    }
}

__global__ void k_hash_heads(const uint64_t* __restrict__ keys, int n, uint64_t* tkeys, uint32_t* tvals,
                             uint32_t mask) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == 0 || keys[j] != keys[j - 1]) {
        const uint32_t slot = table_find_or_insert(tkeys, mask, keys[j]);
        tvals[2 * slot] = (uint32_t)j;
    }
}

__global__ void k_hash_tails(const uint64_t* __restrict__ keys, int n, const uint64_t* __restrict__ tkeys,
                             uint32_t* tvals, uint32_t mask) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == n - 1 || keys[j] != keys[j + 1]) {
        const uint64_t key = keys[j];
        uint32_t slot = (uint32_t)mix64(key) & mask;
        while (tkeys[slot] != key) slot = (slot + 1) & mask;
        tvals[2 * slot + 1] = (uint32_t)(j + 1) - tvals[2 * slot];
    }
}

// ------------------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------------------
constexpr int KNN_WARPS = 8;
#ifndef PSL_KNN_CAP
#define PSL_KNN_CAP 512        // candidates staged per warp before a merge; -DPSL_KNN_CAP=256 halves the shared memory per CTA
#endif                         // (82 -> 49 KB: 3 CTAs per SM instead of 2) at the price of more merges -- A/B with PSL_LIB
constexpr int KNN_CAP = PSL_KNN_CAP;
static_assert(KNN_CAP >= 256 && KNN_CAP % 128 == 0, "KNN_CAP: multiple of 128, at least 256");
constexpr unsigned long long KEY_INF = ~0ull;

__device__ __forceinline__ void list_insert(unsigned long long (&k)[8], unsigned long long x) {
    // k ascending; x < k[7] guaranteed by caller
    k[7] = x;
#pragma unroll
    for (int i = 7; i > 0; --i) {
        const unsigned long long a = k[i - 1], b = k[i];
        const bool sw = b < a;
        k[i - 1] = sw ? b : a;
        k[i] = sw ? a : b;
    }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

struct KnnArgs {
    // RAYS mode
    const float* rays_o; const float* rays_d; const float* gt_depth; const float* t_vals; const float* z_override;
    float near_s, far_s;
    float* z_vals; float* pos_out;
    // POS mode
    const float* pos_in;
    // common
    long long m;            // total queries
    int seg;                // queries per segment (S for rays)
    const double* r2; double r2_scalar; int r2_group;
    int* I; float* D; int* nnum;
    unsigned long long* stats;   // optional counters: [0] candidates staged, [1] warp passes, [2] queries, [3] cells probed
};

template <bool RAYS>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn(GridDev g, KnnArgs a, long long n_work, int n_chunks) {
    grid_resolve(g);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_cand_all = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* s_best_all =
        reinterpret_cast<unsigned long long*>(smem_raw + sizeof(float4) * KNN_WARPS * KNN_CAP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* s_cand = s_cand_all + warp * KNN_CAP;
    unsigned long long* s_best = s_best_all + warp * 32 * 8;

    for (long long w = (long long)blockIdx.x * KNN_WARPS + warp; w < n_work; w += (long long)gridDim.x * KNN_WARPS) {
        const long long segi = w / n_chunks;
        const int chunk = (int)(w - segi * n_chunks);
        const int s0 = chunk * 32;
        int cnt = a.seg - s0;
        cnt = cnt > 32 ? 32 : cnt;
        const long long mbase = segi * a.seg + s0;
        if (mbase + cnt > a.m) cnt = (int)(a.m - mbase);
        const bool valid = lane < cnt;
        const long long mq = mbase + lane;

        // ---- the lane's query point and radius ---------------------------------------------------
        float qx = 0.f, qy = 0.f, qz = 0.f, tle = -1.f, tlt = -1.f, rf = 0.f;
        if (valid) {
            if (RAYS) {
                const float ox = a.rays_o[3 * segi], oy = a.rays_o[3 * segi + 1], oz = a.rays_o[3 * segi + 2];
                const float dx = a.rays_d[3 * segi], dy = a.rays_d[3 * segi + 1], dz = a.rays_d[3 * segi + 2];
                const float dep = a.gt_depth[segi];
                float z;
                if (dep > 0.f) {
                    const float t = a.t_vals[s0 + lane];
                    z = __fadd_rn(__fmul_rn(__fmul_rn(a.near_s, dep), __fsub_rn(1.0f, t)),
                                  __fmul_rn(__fmul_rn(a.far_s, dep), t));
                } else {
                    z = a.z_override ? a.z_override[mq] : 0.f;
                }
                qx = __fadd_rn(ox, __fmul_rn(dx, z));
                qy = __fadd_rn(oy, __fmul_rn(dy, z));
                qz = __fadd_rn(oz, __fmul_rn(dz, z));
                a.z_vals[mq] = z;
                a.pos_out[3 * mq] = qx; a.pos_out[3 * mq + 1] = qy; a.pos_out[3 * mq + 2] = qz;
            } else {
                qx = a.pos_in[3 * mq]; qy = a.pos_in[3 * mq + 1]; qz = a.pos_in[3 * mq + 2];
            }
            const double r2 = a.r2 ? a.r2[mq / a.r2_group] : a.r2_scalar;
            tle = thr_le_of(r2);
            tlt = thr_lt_of(r2);
            rf = sqrtf(fmaxf(tlt, 0.f)) * 1.001f + 1e-5f;
        }
        // Two-pass search.  Pass 0 looks only inside r_small (a few times the typical 8th-neighbour distance): if a
        // sample finds 8 points with D <= r_small^2 these ARE its 8 nearest (every point inside that ball was seen), so
        // the result is exact; samples that do not fill their list force pass 1 with the full query radius.
        float tins = tle, rcur = rf;               // insertion threshold / search radius of the current pass
        bool small_pass = false;
        if (g.r_small > 0.f) {
            const float ts = g.r_small * g.r_small;
            if (valid && ts < tle) { tins = ts; rcur = g.r_small * 1.001f + 1e-5f; }
            small_pass = __any_sync(0xffffffffu, valid && ts < tle);
        }
      for (int pass = small_pass ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1) { tins = tle; rcur = rf; }
        // ---- warp AABB (already dilated by the radii) and its cell range --------------------------
        float bx0 = valid ? qx - rcur : FLT_MAX, by0 = valid ? qy - rcur : FLT_MAX, bz0 = valid ? qz - rcur : FLT_MAX;
        float bx1 = valid ? qx + rcur : -FLT_MAX, by1 = valid ? qy + rcur : -FLT_MAX, bz1 = valid ? qz + rcur : -FLT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
            by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
            bz0 = fminf(bz0, __shfl_xor_sync(0xffffffffu, bz0, o));
            bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
            bz1 = fmaxf(bz1, __shfl_xor_sync(0xffffffffu, bz1, o));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s_best[lane * 8 + k] = KEY_INF;
        __syncwarp();

        int nstaged = 0;
        auto flush = [&]() {
            __syncwarp();
            if (a.stats && lane == 0) atomicAdd(a.stats, (unsigned long long)nstaged);
            for (int s = 0; s < cnt; ++s) {
                const float sx = __shfl_sync(0xffffffffu, qx, s), sy = __shfl_sync(0xffffffffu, qy, s),
                            sz = __shfl_sync(0xffffffffu, qz, s);
                const float sle = __shfl_sync(0xffffffffu, tins, s);
                unsigned long long keys[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) keys[k] = KEY_INF;
                const unsigned long long worst = s_best[s * 8 + 7];
                float bound = sle;
                if (worst != KEY_INF) bound = fminf(bound, __uint_as_float((unsigned)(worst >> 32)));
                if (lane < 8) {
                    const unsigned long long prev = s_best[s * 8 + lane];
                    if (prev != KEY_INF) list_insert(keys, prev);
                }
                for (int j = lane; j < nstaged; j += 32) {
                    const float4 c = s_cand[j];
                    const float d = sqdist_canonical(c.x, c.y, c.z, sx, sy, sz);
                    if (d <= bound) {
                        const unsigned long long key =
                            ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)__float_as_uint(c.w);
                        if (key < keys[7]) list_insert(keys, key);
                    }
                }
                unsigned long long mine = KEY_INF;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const unsigned long long mn = warp_min_u64(keys[0]);
                    if (mn != KEY_INF && keys[0] == mn) {
#pragma unroll
                        for (int k = 0; k < 7; ++k) keys[k] = keys[k + 1];
                        keys[7] = KEY_INF;
                    }
                    if (lane == r) mine = mn;
                }
                __syncwarp();
                if (lane < 8) s_best[s * 8 + lane] = mine;
                __syncwarp();
            }
            nstaged = 0;
        };

        if (cnt > 0 && g.n > 0) {
            const int cx0 = cell_coord(bx0, g.inv_cell), cy0 = cell_coord(by0, g.inv_cell), cz0 = cell_coord(bz0, g.inv_cell);
            const int cx1 = cell_coord(bx1, g.inv_cell), cy1 = cell_coord(by1, g.inv_cell), cz1 = cell_coord(bz1, g.inv_cell);
            const int nx = cx1 - cx0 + 1, ny = cy1 - cy0 + 1, nz = cz1 - cz0 + 1;
            long long ncell = (long long)nx * ny * nz;
            if (nx <= 0 || ny <= 0 || nz <= 0 || ncell > (1ll << 22)) ncell = 0;   // NaN / absurd radius: no neighbours
            if (a.stats && lane == 0) {
                atomicAdd(a.stats + 1, 1ull);
                atomicAdd(a.stats + 3, (unsigned long long)ncell);
                if (pass == (small_pass ? 0 : 1)) atomicAdd(a.stats + 2, (unsigned long long)cnt);
            }
            for (long long base = 0; base < ncell; base += 32) {
                const long long ci = base + lane;
                uint2 rng = make_uint2(0u, 0u);
                if (ci < ncell) {
                    const int ix = (int)(ci % nx), iy = (int)((ci / nx) % ny), iz = (int)(ci / ((long long)nx * ny));
                    rng = grid_lookup(g, cell_key(cx0 + ix, cy0 + iy, cz0 + iz));
                }
                int incl = (int)rng.y;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - (int)rng.y;
                // candidates of these 32 cells, 4 x 32 at a time: the four point loads are issued before any of them is
                // consumed (one L2 round trip per 128 candidates instead of one per 32)
                for (int t0 = 0; t0 < total; t0 += 128) {
                    if (nstaged > KNN_CAP - 128) flush();
                    float4 pt[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        pt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t0 + 32 * u < total) {                      // warp-uniform
                            const int t = t0 + 32 * u + lane;
                            int c = 0;
#pragma unroll
                            for (int step = 16; step > 0; step >>= 1) {
                                const int v = __shfl_sync(0xffffffffu, incl, c + step - 1);
                                if (v <= t) c += step;
                            }
                            c = c > 31 ? 31 : c;
                            const int cstart = __shfl_sync(0xffffffffu, (int)rng.x, c);
                            const int cexcl = __shfl_sync(0xffffffffu, excl, c);
                            if (t < total) pt[u] = __ldg(g.pts + cstart + (t - cexcl));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (t0 + 32 * u < total) {
                            const bool keep = (t0 + 32 * u + lane < total) && pt[u].x >= bx0 && pt[u].x <= bx1 && pt[u].y >= by0 &&
                                              pt[u].y <= by1 && pt[u].z >= bz0 && pt[u].z <= bz1;
                            const unsigned bal = __ballot_sync(0xffffffffu, keep);
                            if (keep) s_cand[nstaged + __popc(bal & ((1u << lane) - 1u))] = pt[u];
                            nstaged += __popc(bal);
                        }
                    }
                }
            }
            if (nstaged > 0) flush();
        }
        __syncwarp();
        if (pass == 0) {
            const bool done = !valid || tins >= tle || s_best[lane * 8 + 7] != KEY_INF;
            if (__all_sync(0xffffffffu, done)) break;
        }
      }
        __syncwarp();
        // ---- write results: 4 queries x 8 slots per pass (128 B coalesced) -------------------------
        for (int sb = 0; sb < cnt; sb += 4) {
            const int s = sb + (lane >> 3), k = lane & 7;
            const bool ok = s < cnt;
            const unsigned long long key = ok ? s_best[s * 8 + k] : KEY_INF;
            const float d = key == KEY_INF ? FLT_MAX : __uint_as_float((unsigned)(key >> 32));
            const int idx = key == KEY_INF ? -1 : (int)(unsigned)(key & 0xffffffffu);
            const float slt = __shfl_sync(0xffffffffu, tlt, ok ? s : 0);
            const unsigned bal = __ballot_sync(0xffffffffu, ok && key != KEY_INF && d < slt);
            if (ok) {
                const long long mo = mbase + s;
                a.I[mo * 8 + k] = idx;
                a.D[mo * 8 + k] = d;
                if (k == 0) a.nnum[mo] = __popc((bal >> (lane & 24)) & 0xffu);
            }
        }
        __syncwarp();
    }
}

template <bool RAYS>
static int launch_knn(const GridDev& g, const KnnArgs& a, cudaStream_t st) {
    if (a.m == 0) return 0;
    const int n_chunks = (a.seg + 31) / 32;
    const long long n_seg = (a.m + a.seg - 1) / a.seg;
    const long long n_work = n_seg * n_chunks;
    const size_t smem = sizeof(float4) * KNN_WARPS * KNN_CAP + sizeof(unsigned long long) * KNN_WARPS * 32 * 8;
    // per launch: the attribute belongs to the device / context the launch goes to (a process may drive several GPUs)
    PSL_CHECK_CUDA(cudaFuncSetAttribute(k_knn<RAYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (n_work + KNN_WARPS - 1) / KNN_WARPS;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    TimingScope ts(T_KNN, st);
    k_knn<RAYS><<<(unsigned)blocks, KNN_WARPS * 32, smem, st>>>(g, a, n_work, n_chunks);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace psl

using namespace psl;

extern "C" size_t psl_grid_sort_ws_bytes(int64_t n) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 63);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al(sizeof(uint64_t) * n) + 2 * al(sizeof(uint32_t) * n) + al(cub_bytes) + 256;
}

extern "C" int psl_grid_sort(const float* cloud_pos, int64_t n, float cell, float* sorted_pts, uint64_t* sorted_keys,
                             void* ws, size_t ws_bytes, int64_t* n_cells_host, psl_stream_t stream) {
    PSL_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
    PSL_REQUIRE(cell > 0.f, "cell size must be positive");
    PSL_REQUIRE(n_cells_host != nullptr, "n_cells_host is NULL");
    *n_cells_host = 0;
    if (n == 0) return 0;
    PSL_REQUIRE(ws_bytes >= psl_grid_sort_ws_bytes(n), "workspace too small");
    cudaStream_t st = as_stream(stream);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    unsigned char* w = static_cast<unsigned char*>(ws);
    uint64_t* keys_in = reinterpret_cast<uint64_t*>(w); w += al(sizeof(uint64_t) * n);
    uint32_t* vals_in = reinterpret_cast<uint32_t*>(w); w += al(sizeof(uint32_t) * n);
    uint32_t* vals_out = reinterpret_cast<uint32_t*>(w); w += al(sizeof(uint32_t) * n);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(w); w += 256;
    void* cub_tmp = w;
    size_t cub_bytes = ws_bytes - (size_t)(w - static_cast<unsigned char*>(ws));
    const int tb = 256, nb = (int)((n + tb - 1) / tb);
    k_cell_keys<<<nb, tb, 0, st>>>(cloud_pos, (int)n, 1.0f / cell, keys_in, vals_in);
    PSL_CHECK_CUDA(cudaGetLastError());
    PSL_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, sorted_keys, vals_in, vals_out, (int)n,
                                                   0, 63, st));
    PSL_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    k_gather_sorted<<<nb, tb, 0, st>>>(cloud_pos, vals_out, sorted_keys, (int)n, reinterpret_cast<float4*>(sorted_pts),
                                       counter);
    PSL_CHECK_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    PSL_CHECK_CUDA(cudaMemcpyAsync(&h, counter, sizeof(h), cudaMemcpyDeviceToHost, st));
    PSL_CHECK_CUDA(cudaStreamSynchronize(st));
    *n_cells_host = (int64_t)h;
    return 0;
}

static size_t append_cub_bytes(int64_t n_total, int64_t k_new, size_t* sort_b, size_t* merge_b) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int)k_new, 0, 63);
    cub::DeviceMerge::MergePairs(nullptr, b, (const uint64_t*)nullptr, (const float4*)nullptr, (int)(n_total - k_new), (const uint64_t*)nullptr,
                                 (const float4*)nullptr, (int)k_new, (uint64_t*)nullptr, (float4*)nullptr);
    if (sort_b) *sort_b = a;
    if (merge_b) *merge_b = b;
    return a > b ? a : b;
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t psl_grid_append_ws_bytes(int64_t n_total, int64_t k_new) {
    return 2 * al256(sizeof(uint64_t) * k_new) + 2 * al256(sizeof(uint32_t) * k_new) + al256(sizeof(float4) * k_new) +
           al256(sizeof(uint64_t) * n_total) + al256(sizeof(float4) * n_total) + al256(append_cub_bytes(n_total, k_new, nullptr, nullptr)) + 512;
}

// Rebuild after an append WITHOUT re-sorting the cloud: the first n_old points are already in (sorted_pts, sorted_keys); the k_new
// appended points (cloud_pos[n_old : n_old + k_new]) are keyed, sorted among themselves and merged in (stable: equal keys keep the old
// points first, and new indices are larger than old ones -- the result is bit-identical to psl_grid_sort on the whole cloud).
// sorted_pts / sorted_keys are updated IN PLACE (they must have room for n_old + k_new entries); *n_cells_host as psl_grid_sort.
extern "C" int psl_grid_append(const float* cloud_pos, int64_t n_old, int64_t k_new, float cell, float* sorted_pts, uint64_t* sorted_keys,
                               void* ws, size_t ws_bytes, int64_t* n_cells_host, psl_stream_t stream) {
    PSL_REQUIRE(cloud_pos && sorted_pts && sorted_keys && ws && n_cells_host, "NULL argument");
    PSL_REQUIRE(n_old > 0 && k_new > 0 && n_old + k_new < (1ll << 31), "append needs an existing sorted cloud and at least one new point");
    PSL_REQUIRE(cell > 0.f, "cell size must be positive");
    const int64_t n = n_old + k_new;
    PSL_REQUIRE(ws_bytes >= psl_grid_append_ws_bytes(n, k_new), "workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned char* w = static_cast<unsigned char*>(ws);
    uint64_t* k_in = reinterpret_cast<uint64_t*>(w); w += al256(sizeof(uint64_t) * k_new);
    uint64_t* k_srt = reinterpret_cast<uint64_t*>(w); w += al256(sizeof(uint64_t) * k_new);
    uint32_t* v_in = reinterpret_cast<uint32_t*>(w); w += al256(sizeof(uint32_t) * k_new);
    uint32_t* v_srt = reinterpret_cast<uint32_t*>(w); w += al256(sizeof(uint32_t) * k_new);
    float4* p_new = reinterpret_cast<float4*>(w); w += al256(sizeof(float4) * k_new);
    uint64_t* k_out = reinterpret_cast<uint64_t*>(w); w += al256(sizeof(uint64_t) * n);
    float4* p_out = reinterpret_cast<float4*>(w); w += al256(sizeof(float4) * n);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(w); w += 256;
    void* cub_tmp = w;
    size_t sort_b = 0, merge_b = 0;
    append_cub_bytes(n, k_new, &sort_b, &merge_b);
    const int tb = 256;
    TimingScope ts(T_MAP, st, 5);
    k_cell_keys_new<<<(int)((k_new + tb - 1) / tb), tb, 0, st>>>(cloud_pos, (int)n_old, (int)k_new, 1.0f / cell, k_in, v_in);
    PSL_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_b, k_in, k_srt, v_in, v_srt, (int)k_new, 0, 63, st));
    k_gather_new<<<(int)((k_new + tb - 1) / tb), tb, 0, st>>>(cloud_pos, v_srt, (int)k_new, p_new);
    PSL_CHECK_CUDA(cub::DeviceMerge::MergePairs(cub_tmp, merge_b, (const uint64_t*)sorted_keys, reinterpret_cast<const float4*>(sorted_pts),
                                                (int)n_old, (const uint64_t*)k_srt, (const float4*)p_new, (int)k_new, k_out, p_out, {}, st));
    PSL_CHECK_CUDA(cudaMemcpyAsync(sorted_keys, k_out, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, st));
    PSL_CHECK_CUDA(cudaMemcpyAsync(sorted_pts, p_out, sizeof(float4) * n, cudaMemcpyDeviceToDevice, st));
    PSL_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    k_count_cells<<<(int)((n + tb - 1) / tb), tb, 0, st>>>(sorted_keys, (int)n, counter);
    PSL_CHECK_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    PSL_CHECK_CUDA(cudaMemcpyAsync(&h, counter, sizeof(h), cudaMemcpyDeviceToHost, st));
    PSL_CHECK_CUDA(cudaStreamSynchronize(st));
    *n_cells_host = (int64_t)h;
    return 0;
}

extern "C" int psl_grid_hash(const uint64_t* sorted_keys, int64_t n, uint64_t* table_keys, uint32_t* table_vals,
                             uint32_t capacity, psl_stream_t stream) {
    PSL_REQUIRE(capacity && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
    cudaStream_t st = as_stream(stream);
    PSL_CHECK_CUDA(cudaMemsetAsync(table_keys, 0xff, sizeof(uint64_t) * capacity, st));
    PSL_CHECK_CUDA(cudaMemsetAsync(table_vals, 0, sizeof(uint32_t) * 2 * capacity, st));
    if (n == 0) return 0;
    const int tb = 256, nb = (int)((n + tb - 1) / tb);
    k_hash_heads<<<nb, tb, 0, st>>>(sorted_keys, (int)n, table_keys, table_vals, capacity - 1);
    k_hash_tails<<<nb, tb, 0, st>>>(sorted_keys, (int)n, table_keys, table_vals, capacity - 1);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int psl_knn_query(const psl_grid* grid_host, const float* pos, int64_t m, const double* r2, double r2_scalar,
                             int32_t r2_group, int32_t* I, float* D, int32_t* nnum, psl_stream_t stream) {
    GridDev g;
    if (int e = make_grid_dev(grid_host, &g)) return e;
    PSL_REQUIRE(m >= 0, "m < 0");
    PSL_REQUIRE(r2_group >= 1, "r2_group must be >= 1");
    KnnArgs a{};
    a.pos_in = pos; a.m = m; a.seg = r2_group; a.r2 = r2; a.r2_scalar = r2_scalar; a.r2_group = r2_group;
    a.I = I; a.D = D; a.nnum = nnum;
    return launch_knn<false>(g, a, as_stream(stream));
}

extern "C" int psl_raymarch_knn(const psl_grid* grid_host, const float* rays_o, const float* rays_d,
                                const float* gt_depth, int64_t n_rays, int32_t n_samples, const float* t_vals,
                                float near_surface, float far_surface, const float* z_override, const double* r2_ray,
                                double r2_scalar, float* z_vals, float* pos, int32_t* I, float* D, int32_t* nnum,
                                psl_stream_t stream) {
    GridDev g;
    if (int e = make_grid_dev(grid_host, &g)) return e;
    PSL_REQUIRE(n_rays >= 0 && n_samples >= 1, "bad ray/sample count");
    KnnArgs a{};
    a.rays_o = rays_o; a.rays_d = rays_d; a.gt_depth = gt_depth; a.t_vals = t_vals; a.z_override = z_override;
    a.near_s = near_surface; a.far_s = far_surface; a.z_vals = z_vals; a.pos_out = pos;
    a.m = n_rays * n_samples; a.seg = n_samples; a.r2 = r2_ray; a.r2_scalar = r2_scalar; a.r2_group = n_samples;
    a.I = I; a.D = D; a.nnum = nnum;
    return launch_knn<true>(g, a, as_stream(stream));
}

extern "C" int psl_raymarch_knn_stats(const psl_grid* grid_host, const float* rays_o, const float* rays_d,
                                      const float* gt_depth, int64_t n_rays, int32_t n_samples, const float* t_vals,
                                      float near_surface, float far_surface, const float* z_override, const double* r2_ray,
                                      double r2_scalar, float* z_vals, float* pos, int32_t* I, float* D, int32_t* nnum,
                                      uint64_t* stats, psl_stream_t stream) {
    GridDev g;
    if (int e = make_grid_dev(grid_host, &g)) return e;
    PSL_REQUIRE(n_rays >= 0 && n_samples >= 1 && stats, "bad ray/sample count or NULL stats");
    KnnArgs a{};
    a.rays_o = rays_o; a.rays_d = rays_d; a.gt_depth = gt_depth; a.t_vals = t_vals; a.z_override = z_override;
    a.near_s = near_surface; a.far_s = far_surface; a.z_vals = z_vals; a.pos_out = pos;
    a.m = n_rays * n_samples; a.seg = n_samples; a.r2 = r2_ray; a.r2_scalar = r2_scalar; a.r2_group = n_samples;
    a.I = I; a.D = D; a.nnum = nnum; a.stats = reinterpret_cast<unsigned long long*>(stats);
    return launch_knn<true>(g, a, as_stream(stream));
}
