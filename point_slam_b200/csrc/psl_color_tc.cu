// Colour branch of the decoder on the 5th-generation tensor cores: weight folding and operand packing shared by the forward
// kernels (3xTF32: psl_color_tc_w16.cu; f16 planes: psl_color_h2.cu).  The forward they serve (round 1 text follows):
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

