/* CPU oracle, plain C restatement of the spline stage.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, scalar and in the reference's operation order, what zenflow computes for one
 * (event, transformed dim) from the raw conditioner output theta[3K-1]:
 *   utils.py:18-34   squareplus, softmax_with_threshold
 *   utils.py:235-250 _knots (sequential cumsum), _index
 *   utils.py:205-232 _compute_rqs_input (gathers, boundary derivative 1, fill-mode NaN at idx == K)
 *   utils.py:121-139 rational_quadratic_spline_forward, utils.py:191-201 inverse
 * It is a second, independent implementation next to oracle/zenflow_oracle.py: tests assert that the two
 * agree on every bin index bit for bit (tests/test_oracle_c.py), which pins the operation order the CUDA
 * kernels must reproduce.  Compile WITHOUT fp contraction / fast-math (see Makefile).
 * Parity unpinned against a running JAX (not installable here), see zenflow_oracle.py.
 */
#include <math.h>
#include <stdint.h>

#define EPSF 1e-5f

static float squareplusf(float x) { return 0.5f * (x + sqrtf(x * x + 4.0f)); }

typedef struct { int idx; float ks, bs, ko, bo, dk, dkp1; } bin_t;

/* search == 0: bins searched on the cumsum of the first K-block (forward); 1: second block (inverse) */
static void locate(const float* th, int K, int search_second, float v, bin_t* o) {
    const float c = (float)(1e-5 / (1.0 - (double)K * 1e-5));
    const float den = (float)(1.0 + (1e-5 / (1.0 - (double)K * 1e-5)) * (double)K);
    const float* ps = th + (search_second ? K : 0);
    const float* po = th + (search_second ? 0 : K);
    float sum = 0.f;
    for (int j = 0; j < K; ++j) { float t = squareplusf(ps[j]); sum = j == 0 ? t : sum + t; }
    /* idx = clip(sum_j [knot_j <= v] - 1, 0, K) over the K+1 knots (utils.py:246-249) */
    float acc = 0.f;
    int count = (0.0f <= v) ? 1 : 0;
    float knots[65], bins[65];
    knots[0] = 0.f;
    for (int j = 0; j < K; ++j) {
        float w = (squareplusf(ps[j]) / sum + c) / den;
        bins[j] = w;
        acc = acc + w;
        knots[j + 1] = acc;
        if (acc <= v) ++count;
    }
    int idx = count - 1;
    if (idx < 0) idx = 0;
    if (idx > K) idx = K;
    o->idx = idx;
    o->ks = knots[idx];
    o->bs = idx < K ? bins[idx] : NAN;
    sum = 0.f;
    for (int j = 0; j < K; ++j) { float t = squareplusf(po[j]); sum = j == 0 ? t : sum + t; }
    acc = 0.f;
    float ko = 0.f, bo = NAN;
    for (int j = 0; j < K; ++j) {
        float h = (squareplusf(po[j]) / sum + c) / den;
        if (j == idx) { ko = acc; bo = h; }
        acc = acc + h;
    }
    if (idx == K) ko = acc;
    o->ko = ko;
    o->bo = bo;
    const float* sl = th + 2 * K;
    o->dk = (idx == 0 || idx == K) ? 1.0f : squareplusf(sl[idx - 1]);
    o->dkp1 = (idx + 1 == K) ? 1.0f : (idx + 1 > K ? NAN : squareplusf(sl[idx]));
}

void zo_rqs_forward(const float* theta, const float* x, int64_t M, int d, int K, float* y, float* log_det, int32_t* idx) {
    const int P = 3 * K - 1;
    for (int64_t m = 0; m < M; ++m) {
        float ldsum = 0.f;
        for (int j = 0; j < d; ++j) {
            const float v = x[m * d + j];
            bin_t b;
            locate(theta + (m * d + j) * P, K, 0, v, &b);
            const float xk = b.ks, dxk = b.bs, yk = b.ko, dyk = b.bo, dk = b.dk, dkp1 = b.dkp1;
            const float sk = dyk / dxk;
            const int oob = (v < 0.f) || (v >= 1.f);
            float z = (v - xk) / dxk;
            if (z == z) { if (z < EPSF) z = EPSF; if (z > (float)(1.0 - 1e-5)) z = (float)(1.0 - 1e-5); }
            const float az = 1.0f - z;
            const float num = dyk * z * (sk * z + dk * az);
            const float den = sk + (dkp1 + dk - 2.0f * sk) * z * az;
            const float yy = yk + num / (den + EPSF);
            const float num2 = z * (dkp1 * z + 2.0f * sk * az) + dk * (az * az);
            const float l = 2.0f * logf(sk + EPSF) + logf(num2 + EPSF) - 2.0f * logf(den + EPSF);
            y[m * d + j] = oob ? v : yy;
            ldsum += oob ? 0.f : l;
            if (idx) idx[m * d + j] = b.idx;
        }
        log_det[m] = ldsum;
    }
}

void zo_rqs_inverse(const float* theta, const float* y, int64_t M, int d, int K, float* x, int32_t* idx) {
    const int P = 3 * K - 1;
    for (int64_t m = 0; m < M; ++m)
        for (int j = 0; j < d; ++j) {
            const float v = y[m * d + j];
            bin_t b;
            locate(theta + (m * d + j) * P, K, 1, v, &b);
            const float yk = b.ks, dyk = b.bs, xk = b.ko, dxk = b.bo, dk = b.dk, dkp1 = b.dkp1;
            const float sk = dyk / dxk;
            const int oob = (v < 0.f) || (v >= 1.f);
            const float beta = dkp1 + dk - 2.0f * sk;
            const float a = dyk * (sk - dk) + (v - yk) * beta;
            const float bq = dyk * dk - (v - yk) * beta;
            const float c = -sk * (v - yk);
            const float z = 2.0f * c / (-bq - sqrtf(bq * bq - 4.0f * a * c));
            x[m * d + j] = oob ? v : z * dxk + xk;
            if (idx) idx[m * d + j] = b.idx;
        }
}
