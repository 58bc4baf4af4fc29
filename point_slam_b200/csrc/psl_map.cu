// Map maintenance between renders: the two O(N) host-side steps the reference's mapper
// runs once per mapped frame, as device kernels with no host round trip of the cloud.
//
//   psl_add_points       NeuralPointCloud.add_neural_points, src/neural_point.py:91-167: depth mask, surface points
//                        o + d*depth, "no indexed point inside the add radius" test against the spatial hash (exact,
//                        canonical fp32 distance, strict float64 compare), ORDER-PRESERVING compaction, N_add points per
//                        kept ray straight into the tail of the (capacity-doubling) position buffer.
//   psl_frustum_select   Mapper.get_mask_from_c2w, src/Mapper.py:120-168: float64 projection of every point (the
//                        reference projects a Python list, i.e. float64 positions, with a float32 inverse pose),
//                        bilinear sensor-depth lookup with the arithmetic of cv2.remap(INTER_LINEAR) (1/32-pixel
//                        fixed-point coordinates, border taps = 0), zero depth -> largest sampled depth, the
//                        image / depth-consistency tests, and the ascending index list (np.where(mask)[0]).
// Integer / index outputs are bit-exact against the reference (tests/golden/aux.npz, tests/golden/frustum.npz).
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <limits.h>

#include "psl_common.cuh"
#include "psl_grid.cuh"

namespace psl {

static inline size_t al256m(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned nblkm(long long n, int tb) { return (unsigned)((n + tb - 1) / tb); }

// ------------------------------------------------------------------------------------------------------------------
// add_neural_points
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_add_valid(const float* __restrict__ depth, int n, int* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) valid[i] = depth[i] > 0.f ? 1 : 0;
}

// one WARP per ray: is there NO indexed point with D < r^2 around the surface point?  (neural_point.py:118-121,199-213)
// Lanes look up the cells of the query ball (<= 27 for r <= cell), the cell ranges are flattened with a warp prefix sum and
// the candidate points are tested 32 at a time (coalesced float4 loads, canonical fp32 distance); the warp leaves at the first
// hit.  (The first version walked the cells with one thread per ray: 83-94 us for 6000 rays at 4.8 % issue utilisation,
// profiles/r01g_summary.md.)
__global__ void __launch_bounds__(256) k_add_probe(GridDev g, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                   const float* __restrict__ depth, int n, const int* __restrict__ valid_rank,
                                                   const double* __restrict__ r2, double r2_scalar, int* __restrict__ keep) {
    grid_resolve(g);
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        const float dep = depth[i];
        if (!(dep > 0.f)) { if (lane == 0) keep[i] = 0; continue; }
        if (g.n == 0) { if (lane == 0) keep[i] = 1; continue; }       // untrained index: every valid ray is kept (:116)
        const float qx = __fadd_rn(rays_o[3 * i], __fmul_rn(rays_d[3 * i], dep));
        const float qy = __fadd_rn(rays_o[3 * i + 1], __fmul_rn(rays_d[3 * i + 1], dep));
        const float qz = __fadd_rn(rays_o[3 * i + 2], __fmul_rn(rays_d[3 * i + 2], dep));
        const double rr = r2 ? r2[valid_rank[i]] : r2_scalar;        // dynamic radii are given for the depth > 0 rays only
        const float tlt = thr_lt_of(rr);
        const float rf = sqrtf(fmaxf(tlt, 0.f)) * 1.001f + 1e-5f;
        const int cx0 = cell_coord(qx - rf, g.inv_cell), cy0 = cell_coord(qy - rf, g.inv_cell), cz0 = cell_coord(qz - rf, g.inv_cell);
        const int cx1 = cell_coord(qx + rf, g.inv_cell), cy1 = cell_coord(qy + rf, g.inv_cell), cz1 = cell_coord(qz + rf, g.inv_cell);
        const int nx = cx1 - cx0 + 1, ny = cy1 - cy0 + 1, nz = cz1 - cz0 + 1;
        long long ncell = (long long)nx * ny * nz;
        if (nx <= 0 || ny <= 0 || nz <= 0 || ncell > (1ll << 15)) ncell = 0;   // NaN / absurd radius: no neighbours
        bool found = false;
        for (long long base = 0; base < ncell && !found; base += 32) {
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
            for (int t0 = 0; t0 < total && !found; t0 += 32) {
                const int t = t0 + lane;
                int c = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, c + step - 1);
                    if (v <= t) c += step;
                }
                c = c > 31 ? 31 : c;
                const int cstart = __shfl_sync(0xffffffffu, (int)rng.x, c);
                const int cexcl = __shfl_sync(0xffffffffu, excl, c);
                bool hit = false;
                if (t < total) {
                    const float4 p = __ldg(g.pts + cstart + (t - cexcl));
                    hit = sqdist_canonical(p.x, p.y, p.z, qx, qy, qz) < tlt;
                }
                found = __any_sync(0xffffffffu, hit);
            }
        }
        if (lane == 0) keep[i] = found ? 0 : 1;
    }
}

struct AddEmitArgs {
    const float* rays_o; const float* rays_d; const float* depth; const float* color;
    const int* keep; const int* keep_rank; const int* valid; const int* valid_rank;
    int n, n_add, fixed_interval;
    float near_s, far_s;
    const float* steps;                 // n_add floats: linspace(0,1,N_add), or linspace(-0.04,0.04,N_add) for the fixed interval
    float* new_pos; float* input_pos; float* input_rgb; int* counts;
};

__global__ void k_add_emit(AddEmitArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (i == a.n - 1) {
        a.counts[0] = a.valid_rank[i] + a.valid[i];
        a.counts[1] = a.keep_rank[i] + a.keep[i];
    }
    if (!a.keep[i]) return;
    const long long j = a.keep_rank[i];
    const float ox = a.rays_o[3 * i], oy = a.rays_o[3 * i + 1], oz = a.rays_o[3 * i + 2];
    const float dx = a.rays_d[3 * i], dy = a.rays_d[3 * i + 1], dz = a.rays_d[3 * i + 2];
    const float dep = a.depth[i];
    if (a.input_pos) {
        a.input_pos[3 * j] = __fadd_rn(ox, __fmul_rn(dx, dep));
        a.input_pos[3 * j + 1] = __fadd_rn(oy, __fmul_rn(dy, dep));
        a.input_pos[3 * j + 2] = __fadd_rn(oz, __fmul_rn(dz, dep));
    }
    if (a.input_rgb && a.color) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a.input_rgb[3 * j + c] = __fmul_rn(a.color[3 * i + c], 255.0f);      // :108
    }
    for (int s = 0; s < a.n_add; ++s) {
        const float t = a.steps[s];
        const float z = a.fixed_interval ? __fadd_rn(dep, t)                                                // :133
                                         : __fadd_rn(__fmul_rn(__fmul_rn(a.near_s, dep), __fsub_rn(1.0f, t)),
                                                     __fmul_rn(__fmul_rn(a.far_s, dep), t));               // :135-137
        float* o = a.new_pos + (j * a.n_add + s) * 3;
        o[0] = __fadd_rn(ox, __fmul_rn(dx, z));
        o[1] = __fadd_rn(oy, __fmul_rn(dy, z));
        o[2] = __fadd_rn(oz, __fmul_rn(dz, z));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// frustum feature selection
// ------------------------------------------------------------------------------------------------------------------
struct FrustumArgs {
    const float* pos; long long n;
    double w[12];                       // rows 0..2 of the float32 inverse pose, promoted to float64
    double fx, fy, cx, cy;
    const float* depth; int H, W, edge;
    float* dsamp; uint8_t* in_img; unsigned* dmax_key; uint8_t* mask;
};

__device__ __forceinline__ int cv_round_x32(float a) {         // cvRound(a * INTER_TAB_SIZE); non-finite / overflow -> INT_MIN
    const float r = rintf(__fmul_rn(a, 32.0f));
    if (!(r >= -2147483648.0f && r < 2147483648.0f)) return INT_MIN;
    return (int)r;
}
__device__ __forceinline__ unsigned float_order_key(float f) {  // monotone float -> unsigned
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ double cam_row(const double* w, double x, double y, double z) {   // Mapper.py:138 (no FMA)
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w[0], x), __dmul_rn(w[1], y)), __dmul_rn(w[2], z)), w[3]);
}

__global__ void k_frustum_project(FrustumArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned key = 0u;
    if (i < a.n) {
        const double x = (double)a.pos[3 * i], y = (double)a.pos[3 * i + 1], z = (double)a.pos[3 * i + 2];
        const double xc = -cam_row(a.w, x, y, z), yc = cam_row(a.w + 4, x, y, z), zc = cam_row(a.w + 8, x, y, z);
        const double zz = __dadd_rn(zc, 1e-5);                                                          // :144
        const float u = __double2float_rn(__ddiv_rn(__dadd_rn(__dmul_rn(a.fx, xc), __dmul_rn(a.cx, zc)), zz));
        const float v = __double2float_rn(__ddiv_rn(__dadd_rn(__dmul_rn(a.fy, yc), __dmul_rn(a.cy, zc)), zz));
        // cv2.remap(depth, u, v, INTER_LINEAR), border constant 0 (Mapper.py:150-154)
        const int sx = cv_round_x32(u), sy = cv_round_x32(v);
        const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
        const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int xx = ix + (q & 1), yy = iy + (q >> 1);
            t[q] = (xx >= 0 && xx < a.W && yy >= 0 && yy < a.H) ? __ldg(a.depth + (long long)yy * a.W + xx) : 0.f;
        }
        const float w0 = __fmul_rn(1.0f - fy, 1.0f - fx), w1 = __fmul_rn(1.0f - fy, fx), w2 = __fmul_rn(fy, 1.0f - fx), w3 = __fmul_rn(fy, fx);
        const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0], w0), __fmul_rn(t[1], w1)), __fmul_rn(t[2], w2)), __fmul_rn(t[3], w3));
        a.dsamp[i] = d;
        const float e = (float)a.edge;
        a.in_img[i] = (u < (float)(a.W - a.edge) && u > e && v < (float)(a.H - a.edge) && v > e) ? 1 : 0;      // :156-157
        key = float_order_key(d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
    if ((threadIdx.x & 31) == 0 && key) atomicMax(a.dmax_key, key);        // integer max: order independent
}

__global__ void k_frustum_mask(FrustumArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double x = (double)a.pos[3 * i], y = (double)a.pos[3 * i + 1], z = (double)a.pos[3 * i + 2];
    const double nz = -__dadd_rn(cam_row(a.w + 8, x, y, z), 1e-5);
    float d = a.dsamp[i];
    if (d == 0.f) d = float_from_order_key(*a.dmax_key);                                                // :159-160
    a.mask[i] = (a.in_img[i] && 0.0 <= nz && nz <= (double)__fadd_rn(d, 0.5f)) ? 1 : 0;                 // :162
}

}  // namespace psl

using namespace psl;

extern "C" size_t psl_add_points_ws_bytes(int64_t n) {
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (int)n);
    return 4 * al256m(sizeof(int) * (size_t)n) + al256m(cub_bytes) + 256;
}

extern "C" int psl_add_points(const psl_grid* grid_host, const float* rays_o, const float* rays_d, const float* gt_depth,
                              const float* gt_color, int64_t n, const double* r2_valid, double r2_scalar, int32_t n_add,
                              int32_t fixed_interval, float near_surface, float far_surface, const float* steps,
                              float* new_pos, float* input_pos, float* input_rgb, int32_t* counts, void* ws, size_t ws_bytes,
                              psl_stream_t stream) {
    GridDev g;
    if (int e = make_grid_dev(grid_host, &g)) return e;
    PSL_REQUIRE(rays_o && rays_d && gt_depth && steps && new_pos && counts && ws, "NULL argument");
    PSL_REQUIRE(n >= 0 && n < (1ll << 30) && n_add >= 1, "bad ray count / N_add");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        PSL_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
        return 0;
    }
    PSL_REQUIRE(ws_bytes >= psl_add_points_ws_bytes(n), "workspace too small");
    unsigned char* w = static_cast<unsigned char*>(ws);
    int* valid = reinterpret_cast<int*>(w); w += al256m(sizeof(int) * n);
    int* valid_rank = reinterpret_cast<int*>(w); w += al256m(sizeof(int) * n);
    int* keep = reinterpret_cast<int*>(w); w += al256m(sizeof(int) * n);
    int* keep_rank = reinterpret_cast<int*>(w); w += al256m(sizeof(int) * n);
    size_t cub_bytes = ws_bytes - (size_t)(w - static_cast<unsigned char*>(ws));
    TimingScope ts(T_MAP, st, 5);
    k_add_valid<<<nblkm(n, 256), 256, 0, st>>>(gt_depth, (int)n, valid);
    PSL_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w, cub_bytes, valid, valid_rank, (int)n, st));
    {
        long long blocks = (n * 32 + 255) / 256;                   // one warp per ray
        const long long cap = (long long)sm_count() * 8;
        k_add_probe<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(g, rays_o, rays_d, gt_depth, (int)n, valid_rank, r2_valid,
                                                                            r2_scalar, keep);
    }
    PSL_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w, cub_bytes, keep, keep_rank, (int)n, st));
    AddEmitArgs a{};
    a.rays_o = rays_o; a.rays_d = rays_d; a.depth = gt_depth; a.color = gt_color;
    a.keep = keep; a.keep_rank = keep_rank; a.valid = valid; a.valid_rank = valid_rank;
    a.n = (int)n; a.n_add = n_add; a.fixed_interval = fixed_interval; a.near_s = near_surface; a.far_s = far_surface;
    a.steps = steps; a.new_pos = new_pos; a.input_pos = input_pos; a.input_rgb = input_rgb; a.counts = counts;
    k_add_emit<<<nblkm(n, 256), 256, 0, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t psl_frustum_select_ws_bytes(int64_t n) {
    size_t cub_bytes = 0;
    cub::CountingInputIterator<int64_t> it(0);
    cub::DeviceSelect::Flagged(nullptr, cub_bytes, it, (const uint8_t*)nullptr, (int64_t*)nullptr, (int32_t*)nullptr, (int)n);
    return al256m(sizeof(float) * (size_t)n) + al256m((size_t)n) + al256m(cub_bytes) + 512;
}

extern "C" int psl_frustum_select(const float* cloud_pos, int64_t n, const double* w2c_host, double fx, double fy, double cx,
                                  double cy, const float* depth, int32_t H, int32_t W, int32_t edge, uint8_t* mask,
                                  int64_t* indices, int32_t* count, void* ws, size_t ws_bytes, psl_stream_t stream) {
    PSL_REQUIRE(w2c_host && depth && mask && count && ws, "NULL argument");
    PSL_REQUIRE(n >= 0 && n < (1ll << 31) && H > 0 && W > 0 && H < 32767 && W < 32767, "bad sizes");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        PSL_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        return 0;
    }
    PSL_REQUIRE(cloud_pos != nullptr, "NULL argument");
    PSL_REQUIRE(ws_bytes >= psl_frustum_select_ws_bytes(n), "workspace too small");
    unsigned char* w = static_cast<unsigned char*>(ws);
    FrustumArgs a{};
    a.pos = cloud_pos; a.n = n;
    for (int i = 0; i < 12; ++i) a.w[i] = w2c_host[i];
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.depth = depth; a.H = H; a.W = W; a.edge = edge; a.mask = mask;
    a.dsamp = reinterpret_cast<float*>(w); w += al256m(sizeof(float) * n);
    a.in_img = w; w += al256m((size_t)n);
    a.dmax_key = reinterpret_cast<unsigned*>(w); w += 256;
    size_t cub_bytes = ws_bytes - (size_t)(w - static_cast<unsigned char*>(ws));
    TimingScope ts(T_MAP, st, 4);
    PSL_CHECK_CUDA(cudaMemsetAsync(a.dmax_key, 0, sizeof(unsigned), st));
    k_frustum_project<<<nblkm(n, 256), 256, 0, st>>>(a);
    k_frustum_mask<<<nblkm(n, 256), 256, 0, st>>>(a);
    PSL_CHECK_CUDA(cudaGetLastError());
    if (indices) {
        cub::CountingInputIterator<int64_t> it(0);
        PSL_CHECK_CUDA(cub::DeviceSelect::Flagged(w, cub_bytes, it, mask, indices, count, (int)n, st));
    }
    return 0;
}
