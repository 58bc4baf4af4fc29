// Per-ray alpha composite (src/common.py:298-336) with the "-100 where no neighbours" masking of src/utils/Renderer.py:189-190 and
// its backward, as device functions shared by the stand-alone kernels (psl_composite.cu) and the fused render tail (psl_shell.cu).
#pragma once
#include "psl_common.cuh"

namespace psl {

constexpr int MAX_S = 64;

struct RayOut { float depth, var, r, g, b; };

__device__ __forceinline__ RayOut composite_fwd_ray(const float4* __restrict__ raw, const unsigned char* __restrict__ has_nb,
                                                    const float* __restrict__ z_vals, long long r, int S, float coef,
                                                    float* __restrict__ weights) {
    float T = 1.f, wsum = 0.f, A = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    for (int s = 0; s < S; ++s) {
        const float4 v = raw[r * S + s];
        const float occ = has_nb[r * S + s] ? v.w : -100.0f;
        const float al = sigmoidf_(__fmul_rn(coef, occ));
        const float w = __fmul_rn(al, T);
        T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, al), 1e-10f));
        const float z = z_vals[r * S + s];
        wsum = __fadd_rn(wsum, w);
        A = __fadd_rn(A, __fmul_rn(w, z));
        cr = __fadd_rn(cr, __fmul_rn(w, v.x));
        cg = __fadd_rn(cg, __fmul_rn(w, v.y));
        cb = __fadd_rn(cb, __fmul_rn(w, v.z));
        if (weights) weights[r * S + s] = w;
    }
    wsum = __fadd_rn(wsum, 1e-10f);
    const float d = __fdiv_rn(A, wsum);
    float vv = 0.f, T2 = 1.f;
    for (int s = 0; s < S; ++s) {
        const float4 v = raw[r * S + s];
        const float occ = has_nb[r * S + s] ? v.w : -100.0f;
        const float al = sigmoidf_(__fmul_rn(coef, occ));
        const float w = __fmul_rn(al, T2);
        T2 = __fmul_rn(T2, __fadd_rn(__fsub_rn(1.0f, al), 1e-10f));
        const float t = __fsub_rn(z_vals[r * S + s], d);
        vv = __fadd_rn(vv, __fmul_rn(__fmul_rn(w, t), t));
    }
    RayOut o;
    o.depth = d; o.var = vv;
    o.r = __fdiv_rn(cr, wsum); o.g = __fdiv_rn(cg, wsum); o.b = __fdiv_rn(cb, wsum);
    return o;
}

// gd0 / gv / (gr, gg, gb): incoming gradients of depth / var / rgb of ray r
__device__ __forceinline__ void composite_bwd_ray(const float4* __restrict__ raw, const unsigned char* __restrict__ has_nb,
                                                  const float* __restrict__ z_vals, long long r, int S, float coef, float gd0,
                                                  float gv, float gr, float gg, float gb, float4* __restrict__ d_raw) {
    float al[MAX_S], Tr[MAX_S];
    float T = 1.f, wsum = 0.f, A = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    for (int s = 0; s < S; ++s) {
        const float4 v = raw[r * S + s];
        const float occ = has_nb[r * S + s] ? v.w : -100.0f;
        const float a_ = sigmoidf_(coef * occ);
        al[s] = a_; Tr[s] = T;
        const float w = a_ * T;
        T *= (1.0f - a_) + 1e-10f;
        wsum += w; A += w * z_vals[r * S + s];
        cr += w * v.x; cg += w * v.y; cb += w * v.z;
    }
    wsum += 1e-10f;
    const float inv = 1.0f / wsum;
    const float dep = A * inv, Rr = cr * inv, Rg = cg * inv, Rb = cb * inv;
    float sw = 0.f;                                   // sum_s w_s (z_s - depth)
    for (int s = 0; s < S; ++s) sw += al[s] * Tr[s] * (z_vals[r * S + s] - dep);
    const float gd = gd0 - 2.0f * gv * sw;
    const float dA = gd * inv, dBr = gr * inv, dBg = gg * inv, dBb = gb * inv;
    const float dWs = -(gd * dep + gr * Rr + gg * Rg + gb * Rb) * inv;
    float G = 0.f;                                    // sum_{s>j} dT_s * T_s
    for (int s = S - 1; s >= 0; --s) {
        const float4 v = raw[r * S + s];
        const float z = z_vals[r * S + s];
        const float w = al[s] * Tr[s];
        const float t = z - dep;
        const float dw = gv * t * t + dA * z + dBr * v.x + dBg * v.y + dBb * v.z + dWs;
        const float u = (1.0f - al[s]) + 1e-10f;
        const float da = dw * Tr[s] - G / u;
        G += dw * al[s] * Tr[s];
        const float docc = da * al[s] * (1.0f - al[s]) * coef;
        d_raw[r * S + s] = make_float4(w * dBr, w * dBg, w * dBb, docc);
    }
}

}  // namespace psl
