// Buffer layouts shared by the tensor-core training kernels (forward save -> backward data -> weight gradients).
// "Tile-transposed" tensors are stored per 128-sample tile as [tile][channel][128 samples]: a worker thread (= one
// sample) writes/reads them one channel at a time, so a warp touches 32 consecutive floats (coalesced), and the
// weight-gradient GEMM -- whose reduction index is the sample -- finds its operands K-major without a transpose.
#pragma once
namespace psl {
struct TSave { long long zT, z1T, cT, wnT, f, outpre, total; };
__host__ __device__ inline TSave tsave_layout(long long M, int rel) {
    const long long T = (M + 127) / 128;
    TSave L{};
    long long o = 0;
    L.zT = o; o += 5 * T * 16384;                 // trunk pre-activations   [5][T][128 ch][128]
    L.z1T = o; if (rel) o += T * 8 * 16384;       // neighbour-MLP pre-act.  [T][8][128 ch][128]
    L.cT = o; o += T * 4096;                      // interpolated feature c  [T][32][128]
    L.wnT = o; o += T * 1024;                     // normalised IDW weights  [T][8][128]
    L.f = o; if (rel) o += T * 128 * 256;         // neighbour-MLP outputs   (T*128, 8, 32) row-major
    L.outpre = o; o += T * 128 * 4;               // colour output before the exposure affine / sigmoid (T*128, 4)
    L.total = o;
    return L;
}
struct TBwd { long long dhT, doutT, dz1T, dccT, aff, total; };
__host__ __device__ inline TBwd tbwd_layout(long long M, int rel) {
    const long long T = (M + 127) / 128;
    TBwd L{};
    long long o = 0;
    L.dhT = o; o += 5 * T * 16384;                // dL/dh_l                  [5][T][128][128]
    L.doutT = o; o += T * 16 * 128;               // dL/d(colour output)      [T][16][128] (rows 3..15 zero)
    L.dz1T = o; if (rel) o += T * 8 * 16384;      // neighbour-MLP dL/dz1     [T][8][128][128]
    L.dccT = o; o += T * 4096;                    // dL/dc (0 where no nbrs)  [T][32][128]
    L.aff = o; o += T * 128 * 12;                 // per-sample exposure-affine gradient terms (T*128, 12)
    L.total = o;
    return L;
}
}  // namespace psl
