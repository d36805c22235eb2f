// VJP of normalize_spline_params + rational_quadratic_spline_forward (utils.py:37-141) for one (event, dim) row:
// shared by the stage-style kernel of zf_train.cu (theta from HBM) and the fused conditioner-recompute kernel of
// zf_chain.cu (theta from tensor memory).  Derivation: SURVEY.md Appendix A.
#pragma once
#include "zf_math.cuh"

namespace zf {

__device__ __forceinline__ float squareplus_grad(float a) {  // d/da 0.5*(a + sqrt(a^2+4))
    return 0.5f * (1.0f + a * rsqrtf(a * a + 4.0f));
}

// row: raw theta (3K-1) in shared memory, overwritten by its cotangent.  Returns d/dx.
// KT > 0: K known at compile time (loops unrolled); KT == 0: runtime K.
template <int KT>
__device__ __forceinline__ float rqs_row_backward(float* row, int K_rt, float x, float gy, float gld, const KnotNorm& kn) {
    const int K = KT > 0 ? KT : K_rt;
    RqsBin b;
    rqs_locate<KT>(row, K, true, x, kn, b);
    const int P = 3 * K - 1;
    const bool oob = (x < 0.f) || (x >= 1.f);
    const int idx = b.idx;
    if (oob || idx >= K || !(x == x)) {  // identity branch (or the reference's NaN corner): no parameter gradient
        for (int p = 0; p < P; ++p) row[p] = 0.f;
        return oob ? gy : 0.f;
    }
    const float xk = b.ks, w = b.bs, h = b.bo, d0 = b.dk, d1 = b.dkp1;
    const float s = h / w;
    const float xi_raw = (x - xk) / w;
    const bool clipped = !(xi_raw > kEps && xi_raw < kOneMinusEps);
    const float xi = fminf(fmaxf(xi_raw, kEps), kOneMinusEps);
    const float az = 1.0f - xi;
    const float beta = d1 + d0 - 2.0f * s;
    const float u = s * xi + d0 * az;
    const float num = h * xi * u;
    const float den = s + beta * xi * az;
    const float Dn = den + kEps;
    const float v = d1 * xi + 2.0f * s * az;
    const float num2 = xi * v + d0 * az * az;

    // adjoints (Appendix A)
    const float g_yk = gy;
    const float g_num = gy / Dn;
    const float g_den = -gy * num / (Dn * Dn) - 2.0f * gld / Dn;
    const float g_num2 = gld / (num2 + kEps);
    float g_s = gld * 2.0f / (s + kEps);
    float g_h = g_num * xi * u;
    float g_xi = g_num * h * u;
    const float g_u = g_num * h * xi;
    g_s += g_u * xi;
    g_xi += g_u * s;
    float g_d0 = g_u * az;
    float g_az = g_u * d0;
    g_s += g_den;
    const float g_beta = g_den * xi * az;
    g_xi += g_den * beta * az;
    g_az += g_den * beta * xi;
    float g_d1 = g_beta;
    g_d0 += g_beta;
    g_s -= 2.0f * g_beta;
    g_xi += g_num2 * v;
    const float g_v = g_num2 * xi;
    g_d0 += g_num2 * az * az;
    g_az += g_num2 * d0 * 2.0f * az;
    g_d1 += g_v * xi;
    g_xi += g_v * d1;
    g_s += g_v * 2.0f * az;
    g_az += g_v * 2.0f * s;
    g_xi -= g_az;
    const float g_xr = clipped ? 0.f : g_xi;
    const float g_x = g_xr / w;
    const float g_xk = -g_xr / w;
    float g_w = -g_xr * xi_raw / w;
    g_h += g_s / w;
    g_w -= g_s * s / w;

    // slopes first (their raw values are needed before the row is overwritten)
    const float c_lo = (idx >= 1) ? row[2 * K + idx - 1] : 0.f;
    const float c_hi = (idx + 1 <= K - 1) ? row[2 * K + idx] : 0.f;

    // widths / heights: W_j = kappa*(s_j/S + c); cotangent of W_j is g_lt (j<idx), g_at (j==idx), 0 otherwise
    const float kappa = kn.rden;
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
        float* pr = row + blk * K;
        const float g_lt = blk == 0 ? g_xk : g_yk;
        const float g_at = blk == 0 ? g_w : g_h;
        float S = 0.f, Slt = 0.f, s_at = 0.f;
        constexpr int KS = KT > 0 ? KT : 1;
        if (KT > 0) {
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const float aj = pr[j];
                const float sj = 0.5f * (aj + sqrtf(fmaf(aj, aj, 4.0f)));   // fp32-tolerance path: plain sqrt
                S += sj;
                Slt += (j < idx) ? sj : 0.f;
                s_at = (j == idx) ? sj : s_at;
            }
        } else {
            for (int j = 0; j < K; ++j) {
                const float sj = squareplus_rn(pr[j]);
                S += sj;
                if (j < idx) Slt += sj;
                if (j == idx) s_at = sj;
            }
        }
        const float A = (g_lt * Slt + g_at * s_at) / S;
        const float ks = kappa / S;
        if (KT > 0) {
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const float a = pr[j];
                const float gW = (j < idx) ? g_lt : ((j == idx) ? g_at : 0.f);
                // d squareplus/da = 0.5*(1 + a/sqrt(a^2+4)) = s/(2s - a) ... use the rsqrt form
                pr[j] = ks * (gW - A) * squareplus_grad(a);
            }
        } else {
            for (int j = 0; j < K; ++j) {
                const float a = pr[j];
                const float gW = (j < idx) ? g_lt : ((j == idx) ? g_at : 0.f);
                pr[j] = ks * (gW - A) * squareplus_grad(a);
            }
        }
    }
    for (int j = 0; j < K - 1; ++j) row[2 * K + j] = 0.f;
    if (idx >= 1) row[2 * K + idx - 1] = g_d0 * squareplus_grad(c_lo);
    if (idx + 1 <= K - 1) row[2 * K + idx] = g_d1 * squareplus_grad(c_hi);
    return g_x;
}

}  // namespace zf
