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

