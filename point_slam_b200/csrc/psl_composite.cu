// Alpha composite (src/common.py:298-336) with the "-100 where no neighbours" masking of
// src/utils/Renderer.py:189-190, its backward, the ray backward of p = o + d z, the ray validity mask
// (decoder.py:200-201) and the deterministic feature-gradient scatter.
#include <cub/device/device_radix_sort.cuh>

#include "psl_composite.cuh"

namespace psl {

__global__ void k_composite_fwd(const float4* __restrict__ raw, const unsigned char* __restrict__ has_nb,
                                const float* __restrict__ z_vals, long long R, int S, float coef,
                                float* __restrict__ depth, float* __restrict__ var, float* __restrict__ rgb,
                                float* __restrict__ weights) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const RayOut o = composite_fwd_ray(raw, has_nb, z_vals, r, S, coef, weights);
    depth[r] = o.depth;
    var[r] = o.var;
    rgb[r * 3] = o.r; rgb[r * 3 + 1] = o.g; rgb[r * 3 + 2] = o.b;
}

__global__ void k_composite_bwd(const float4* __restrict__ raw, const unsigned char* __restrict__ has_nb,
                                const float* __restrict__ z_vals, long long R, int S, float coef,
                                const float* __restrict__ d_depth, const float* __restrict__ d_var,
                                const float* __restrict__ d_rgb, float4* __restrict__ d_raw) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    composite_bwd_ray(raw, has_nb, z_vals, r, S, coef, d_depth ? d_depth[r] : 0.f, d_var ? d_var[r] : 0.f, d_rgb ? d_rgb[r * 3] : 0.f,
                      d_rgb ? d_rgb[r * 3 + 1] : 0.f, d_rgb ? d_rgb[r * 3 + 2] : 0.f, d_raw);
}

__global__ void k_rays_bwd(const float* __restrict__ d_pos, const float* __restrict__ z_vals, long long R, int S,
                           float* __restrict__ d_o, float* __restrict__ d_d) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    for (int s = 0; s < S; ++s) {
        const float z = z_vals[r * S + s];
        const float gx = d_pos[(r * S + s) * 3], gy = d_pos[(r * S + s) * 3 + 1], gz = d_pos[(r * S + s) * 3 + 2];
        ox += gx; oy += gy; oz += gz;
        dx += z * gx; dy += z * gy; dz += z * gz;
    }
    if (d_o) { d_o[r * 3] = ox; d_o[r * 3 + 1] = oy; d_o[r * 3 + 2] = oz; }
    if (d_d) { d_d[r * 3] = dx; d_d[r * 3 + 1] = dy; d_d[r * 3 + 2] = dz; }
}

__global__ void k_ray_mask(const unsigned char* __restrict__ has_nb, long long R, int S, int min_count,
                           unsigned char* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int c = 0;
    for (int s = 0; s < S; ++s) c += has_nb[r * S + s] ? 1 : 0;
    out[r] = c >= min_count;
}

// ---- deterministic feature-gradient scatter ----------------------------------------------------------------
// key = output row of the pair: the point index, or row_map[point] when only a subset of rows is optimised (pairs of
// unmapped points, row_map < 0, are dropped together with the zero-weight ones)
__global__ void k_pair_keys(const int* __restrict__ I, const float* __restrict__ wn, long long n_pairs, unsigned n_points,
                            const int* __restrict__ row_map, unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int idx = I[p];
    if (idx >= 0 && row_map) idx = row_map[idx];
    keys[p] = (idx >= 0 && wn[p] != 0.f) ? (unsigned)idx : n_points;
    vals[p] = (unsigned)p;
}

// One warp per window of 32 consecutive sorted pairs: segment heads are found with one coalesced key load + ballot, then
// each head's segment is summed cooperatively (lane = channel) in PAIR ORDER -- the order the previous one-warp-per-pair
// version used, so results are bit-identical -- with the (key, pair) entries of up to 32 pairs fetched by one coalesced load
// and the gathers of four pairs in flight before their (ordered) accumulation.
__global__ void k_scatter_segments(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals,
                                   long long n_pairs, unsigned n_points, const float* __restrict__ wn,
                                   const float* __restrict__ d_cg, const float* __restrict__ d_colpair,
                                   const float* __restrict__ d_cc, float* __restrict__ d_geo, float* __restrict__ d_col) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_win = (n_pairs + 31) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_win; w += nwarps) {
        const long long p0 = w << 5, pj = p0 + lane;
        const unsigned k = pj < n_pairs ? keys[pj] : 0xffffffffu;
        unsigned kprev = __shfl_up_sync(0xffffffffu, k, 1);
        if (lane == 0) kprev = p0 > 0 ? keys[p0 - 1] : 0xffffffffu;
        unsigned heads = __ballot_sync(0xffffffffu, pj < n_pairs && k < n_points && (pj == 0 || kprev != k));
        while (heads) {
            const int h = __ffs(heads) - 1;
            heads &= heads - 1;
            const unsigned key = __shfl_sync(0xffffffffu, k, h);
            float ag = 0.f, ac = 0.f;
            for (long long j0 = p0 + h;; j0 += 32) {
                const long long j = j0 + lane;
                const bool in = j < n_pairs;
                const unsigned kj = in ? keys[j] : 0xffffffffu;
                const unsigned pid_l = in ? vals[j] : 0u;
                const unsigned same = __ballot_sync(0xffffffffu, kj == key);
                const int cnt = same == 0xffffffffu ? 32 : __ffs(~same) - 1;      // leading pairs of this segment
                const float w_l = lane < cnt ? wn[pid_l] : 0.f;
                for (int t0 = 0; t0 < cnt; t0 += 4) {
                    float wv[4], g[4], c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u < cnt ? t0 + u : t0;                   // (clamped lanes are not accumulated)
                        const unsigned pid = __shfl_sync(0xffffffffu, pid_l, t);
                        wv[u] = __shfl_sync(0xffffffffu, w_l, t);
                        const long long m = pid >> 3;
                        g[u] = d_geo ? d_cg[m * 32 + lane] : 0.f;
                        c[u] = d_col ? (d_colpair ? d_colpair[(long long)pid * 32 + lane] : d_cc[m * 32 + lane]) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (t0 + u < cnt) {
                            if (d_geo) ag = fmaf(wv[u], g[u], ag);
                            if (d_col) ac = d_colpair ? __fadd_rn(ac, c[u]) : fmaf(wv[u], c[u], ac);
                        }
                    }
                }
                if (cnt < 32) break;
            }
            if (d_geo) d_geo[(long long)key * 32 + lane] = ag;
            if (d_col) d_col[(long long)key * 32 + lane] = ac;
        }
    }
}

}  // namespace psl

using namespace psl;

static inline unsigned nblk(long long n, int tb) { return (unsigned)((n + tb - 1) / tb); }

extern "C" int psl_composite_fwd(const float* raw, const uint8_t* has_nb, const float* z_vals, int64_t n_rays,
                                 int32_t n_samples, float coef, float* depth, float* var, float* rgb, float* weights,
                                 psl_stream_t stream) {
    PSL_REQUIRE(raw && has_nb && z_vals && depth && var && rgb, "NULL argument");
    if (n_rays == 0) return 0;
    TimingScope ts(T_COMPOSITE, as_stream(stream));
    k_composite_fwd<<<nblk(n_rays, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(raw), has_nb, z_vals,
                                                                     n_rays, n_samples, coef, depth, var, rgb, weights);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int psl_composite_bwd(const float* raw, const uint8_t* has_nb, const float* z_vals, int64_t n_rays,
                                 int32_t n_samples, float coef, const float* d_depth, const float* d_var,
                                 const float* d_rgb, float* d_raw, psl_stream_t stream) {
    PSL_REQUIRE(raw && has_nb && z_vals && d_raw, "NULL argument");
    PSL_REQUIRE(n_samples <= MAX_S, "n_samples > 64 not supported by the composite backward");
    if (n_rays == 0) return 0;
    TimingScope ts(T_COMPOSITE, as_stream(stream));
    k_composite_bwd<<<nblk(n_rays, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(raw), has_nb, z_vals,
                                                                     n_rays, n_samples, coef, d_depth, d_var, d_rgb,
                                                                     reinterpret_cast<float4*>(d_raw));
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int psl_rays_bwd(const float* d_pos, const float* z_vals, int64_t n_rays, int32_t n_samples, float* d_rays_o,
                            float* d_rays_d, psl_stream_t stream) {
    PSL_REQUIRE(d_pos && z_vals, "NULL argument");
    if (n_rays == 0) return 0;
    TimingScope ts(T_COMPOSITE, as_stream(stream));
    k_rays_bwd<<<nblk(n_rays, 128), 128, 0, as_stream(stream)>>>(d_pos, z_vals, n_rays, n_samples, d_rays_o, d_rays_d);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int psl_ray_mask(const uint8_t* has_nb, int64_t n_rays, int32_t n_samples, int32_t min_count,
                            uint8_t* ray_mask, psl_stream_t stream) {
    PSL_REQUIRE(has_nb && ray_mask, "NULL argument");
    if (n_rays == 0) return 0;
    TimingScope ts(T_COMPOSITE, as_stream(stream));
    k_ray_mask<<<nblk(n_rays, 128), 128, 0, as_stream(stream)>>>(has_nb, n_rays, n_samples, min_count, ray_mask);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t psl_feat_scatter_ws_bytes(int64_t m) {
    const long long np = m * 8;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)np, 0, 32);
    return 4 * al256(sizeof(unsigned) * np) + al256(cub_bytes) + 256;
}

extern "C" int psl_feat_scatter(const int32_t* I, int64_t m, int64_t n_points, const float* wn, const float* d_cg,
                                const float* d_colpair, const float* d_cc, float* d_geo, float* d_col, void* ws,
                                size_t ws_bytes, psl_stream_t stream) {
    return psl_feat_scatter_mapped(I, m, nullptr, n_points, wn, d_cg, d_colpair, d_cc, d_geo, d_col, ws, ws_bytes, stream);
}

extern "C" int psl_feat_scatter_mapped(const int32_t* I, int64_t m, const int32_t* row_map, int64_t n_points, const float* wn,
                                       const float* d_cg, const float* d_colpair, const float* d_cc, float* d_geo, float* d_col,
                                       void* ws, size_t ws_bytes, psl_stream_t stream) {
    PSL_REQUIRE(I && wn && ws, "NULL argument");
    PSL_REQUIRE(!d_geo || d_cg, "d_geo needs d_cg");
    PSL_REQUIRE(!d_col || d_colpair || d_cc, "d_col needs d_colpair or d_cc");
    PSL_REQUIRE(m * 8 < (1ll << 31) && n_points < (1ll << 31), "too many pairs/points");
    if (m == 0 || (!d_geo && !d_col)) return 0;
    PSL_REQUIRE(ws_bytes >= psl_feat_scatter_ws_bytes(m), "workspace too small");
    cudaStream_t st = as_stream(stream);
    const long long np = m * 8;
    unsigned char* w = static_cast<unsigned char*>(ws);
    unsigned* k_in = reinterpret_cast<unsigned*>(w); w += al256(sizeof(unsigned) * np);
    unsigned* k_out = reinterpret_cast<unsigned*>(w); w += al256(sizeof(unsigned) * np);
    unsigned* v_in = reinterpret_cast<unsigned*>(w); w += al256(sizeof(unsigned) * np);
    unsigned* v_out = reinterpret_cast<unsigned*>(w); w += al256(sizeof(unsigned) * np);
    size_t cub_bytes = ws_bytes - (size_t)(w - static_cast<unsigned char*>(ws));
    TimingScope ts(T_SCATTER, st, 6);   // key kernel + 4 radix-sort passes + segment kernel
    k_pair_keys<<<nblk(np, 256), 256, 0, st>>>(I, wn, np, (unsigned)n_points, row_map, k_in, v_in);
    int bits = 1;
    while ((1ll << bits) <= n_points) ++bits;
    PSL_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w, cub_bytes, k_in, k_out, v_in, v_out, (int)np, 0, bits, st));
    long long blocks = (((np + 31) / 32) * 32 + 255) / 256;      // one warp per window of 32 sorted pairs
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_scatter_segments<<<(unsigned)blocks, 256, 0, st>>>(k_out, v_out, np, (unsigned)n_points, wn, d_cg, d_colpair, d_cc,
                                                        d_geo, d_col);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
