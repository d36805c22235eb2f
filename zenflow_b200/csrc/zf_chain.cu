// Whole-chain eval passes: one kernel applies every bijector of a Chain to a tile of samples
// that stays in shared memory from the first ShiftBounds to the latent log-pdf.
//
//   zf_chain_forward  <- Chain.__call__(train=False)        bijectors.py:104-111
//   zf_chain_inverse  <- Chain.inverse                      bijectors.py:113-116
//   zf_flow_log_prob  <- Flow.__call__(train=False)         flow.py:22-48
// per step:
//   ShiftBounds.__call__/inverse                            bijectors.py:164-240
//   Roll                                                    bijectors.py:288-297 (column renaming only)
//   NeuralSplineCoupling._spline_params/__call__/inverse    bijectors.py:329-371
//     eval BatchNorm -> Dense+swish ... -> Dense            (fp32 FFMA register-tiled GEMM)
//     -> normalize_spline_params + rqs forward/inverse      (zf_math.cuh), theta never leaves smem
//
// Data layout: the tile is feature-major in shared memory (xs[col][m]) in *physical* column
// order; a Roll only changes the logical->physical rotation carried by the following steps.
// Parameters are re-packed per call by a tiny kernel into zero-padded, 16-byte aligned blocks
// in the caller's workspace (the FLAX leaves themselves are never modified).
#include "zf_common.cuh"
#include "zf_math.cuh"

#include <float.h>
#include <math.h>
#include <vector>

namespace zf {

void count_launch();

constexpr int TM = 64;        // samples per tile
constexpr int KC = 16;        // k rows per staged weight chunk
constexpr int NCOL = 128;     // output columns per GEMM pass
constexpr int kChainThreads = 256;
constexpr int kStepKindShiftBounds = 0, kStepKindCoupling = 2;

__host__ __device__ inline int ru(int x, int m) { return (x + m - 1) / m * m; }

// Device-side description of one non-Roll step; lives at the head of the workspace.
struct StepDesc {
    int kind;
    int rot;       // logical column j is physical column (j - rot) mod D while this step runs
    int K, n_hidden, F, d;
    int hidden[ZF_MAX_LAYERS];
    int off_bn;                     // float offsets into the workspace
    int off_W[ZF_MAX_LAYERS + 1];
    int off_b[ZF_MAX_LAYERS + 1];
    int off_sb;
    int pad[2];
};
static_assert(sizeof(StepDesc) % 16 == 0, "StepDesc must keep the packed blocks 16-byte aligned");

struct PackJob {
    StepDesc desc;
    int step_index;
    int D;
    // coupling
    const float *bn_scale, *bn_bias, *bn_mean, *bn_var;
    const float* kernel[ZF_MAX_LAYERS + 1];
    const float* bias[ZF_MAX_LAYERS + 1];
    int Kin[ZF_MAX_LAYERS + 1];
    int N[ZF_MAX_LAYERS + 1];
    // shift bounds
    int sb_kind[ZF_MAX_DIM];
    double lo[ZF_MAX_DIM], hi[ZF_MAX_DIM];
    const float *xmin, *xmax;
};

constexpr int kSbStride = 8;  // floats per column: kind, a, b, xmin, xmax, mul, log(mul), -

__global__ void __launch_bounds__(256) pack_step_kernel(const __grid_constant__ PackJob job, float* ws) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsz = gridDim.x * blockDim.x;
    const StepDesc& s = job.desc;
    if (gtid == 0) reinterpret_cast<StepDesc*>(ws)[job.step_index] = s;

    if (s.kind == kStepKindShiftBounds) {
        for (int i = gtid; i < job.D; i += gsz) {
            float* t = ws + s.off_sb + i * kSbStride;
            const int kind = job.sb_kind[i];
            float a = (float)job.lo[i], b = (float)job.hi[i];
            float xmin = 0.f, xmax = 0.f, mul;
            if (kind == ZF_BOUND_BOTH) {
                mul = (float)(1.0 / (job.hi[i] - job.lo[i]));      // bijectors.py:189
            } else {
                xmin = job.xmin[i];
                xmax = job.xmax[i];
                mul = __fdiv_rn(1.0f, __fsub_rn(xmax, xmin));      // bijectors.py:265
            }
            t[0] = (float)kind; t[1] = a; t[2] = b; t[3] = xmin; t[4] = xmax;
            t[5] = mul; t[6] = logf(mul); t[7] = 0.f;
        }
        return;
    }

    // BatchNorm (eval): y = (x - mean) * (rsqrt(var + eps) * scale) + bias
    const int F = s.F, F_p = ru(F, KC);
    for (int f = gtid; f < F_p; f += gsz) {
        float mul = 0.f, mean = 0.f, bias = 0.f;
        if (f < F) {
            mul = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(job.bn_var[f], 1e-5f))) * job.bn_scale[f];
            mean = job.bn_mean[f];
            bias = job.bn_bias[f];
        }
        ws[s.off_bn + f] = mul;
        ws[s.off_bn + F_p + f] = mean;
        ws[s.off_bn + 2 * F_p + f] = bias;
    }
    // hidden layers: W (Kin, N) -> [Kin_p][N_p] zero padded
    for (int l = 0; l < s.n_hidden; ++l) {
        const int Kin = job.Kin[l], N = job.N[l];
        const int Kin_p = ru(Kin, KC), N_p = ru(N, KC);
        const float* W = job.kernel[l];
        for (int e = gtid; e < Kin_p * N_p; e += gsz) {
            int k = e / N_p, n = e - k * N_p;
            ws[s.off_W[l] + e] = (k < Kin && n < N) ? W[(size_t)k * N + n] : 0.f;
        }
        for (int n = gtid; n < N_p; n += gsz) ws[s.off_b[l] + n] = n < N ? job.bias[l][n] : 0.f;
    }
    // last layer: (Kin, d*P) -> [d][Kin_p][Pp]
    {
        const int L = s.n_hidden;
        const int Kin = job.Kin[L], Kin_p = ru(Kin, KC);
        const int P = 3 * s.K - 1, Pp = ru(P, 4), d = s.d;
        const float* W = job.kernel[L];
        const int per = Kin_p * Pp;
        for (int e = gtid; e < d * per; e += gsz) {
            int jj = e / per, r = e - jj * per;
            int k = r / Pp, p = r - k * Pp;
            ws[s.off_W[L] + e] = (k < Kin && p < P) ? W[(size_t)k * (d * P) + jj * P + p] : 0.f;
        }
        for (int e = gtid; e < d * Pp; e += gsz) {
            int jj = e / Pp, p = e - jj * Pp;
            ws[s.off_b[L] + e] = p < P ? job.bias[L][jj * P + p] : 0.f;
        }
    }
}

enum ChainMode : int { kModeForward = 0, kModeLogProb = 1, kModeInverse = 2 };

struct ChainArgs {
    const float* x;     // (M, D) input (x for forward/log_prob, z for inverse)
    const float* c;     // (M, C) or null
    float* y;           // (M, D) output or null
    float* log_det;     // (M,) or null (forward)
    float* lp;          // (M,) (log_prob)
    const float* ws;    // workspace: StepDesc[n_steps] then packed parameters
    long long M;
    int D, C;
    int n_steps;
    int rot_total;      // rotation after the last step
    int act_rows;
    int mode;
    int acc_log_det;    // log_det[m] += instead of =
    LatentConst lc;
};

__device__ __forceinline__ int pmod(int a, int D) {
    int r = a % D;
    return r < 0 ? r + D : r;
}

// acc[4][8] += act_in[Kin_p][TM]^T (samples tx*4..+3) x W[Kin_p][ncols] (cols ty*4..+3, 64+ty*4..+3)
// W rows are streamed from global (L2) through a cp.async double buffer.
__device__ __forceinline__ void gemm_pass(const float* __restrict__ act_in, int Kin_p,
                                          const float* __restrict__ Wg, int ldw, int ncols,
                                          float* __restrict__ wst, float (&acc)[4][8], int tid) {
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int nchunks = Kin_p / KC;
    const int c4 = ncols >> 2;
    auto load_chunk = [&](int c, int buf) {
        const float* src = Wg + (size_t)c * KC * ldw;
        float* dst = wst + buf * (KC * NCOL);
        for (int q = tid; q < KC * c4; q += kChainThreads) {
            int row = q / c4, col = (q - row * c4) << 2;
            cp_async_16(dst + row * NCOL + col, src + (size_t)row * ldw + col);
        }
    };
    load_chunk(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            load_chunk(c + 1, (c + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* wb = wst + (c & 1) * (KC * NCOL);
        const float* ab = act_in + c * KC * TM;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(ab + kk * TM + tx * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(wb + kk * NCOL + ty * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(wb + kk * NCOL + 64 + ty * 4);
            const float a_[4] = {av.x, av.y, av.z, av.w};
            const float b_[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        __syncthreads();
    }
}

template <bool INVERSE>
__device__ __forceinline__ void spline_rows(float* th, int Pst, int K, const KnotNorm& kn, float* xcol,
                                            float& ldc, int tid) {
    // one thread per sample of the tile (tid < TM): theta row -> bin -> transform
    float* row = th + tid * Pst;
    const float v = xcol[tid];
    RqsBin b;
    switch (K) {
        case 16: rqs_locate<16>(row, K, !INVERSE, v, kn, b); break;
        case 32: rqs_locate<32>(row, K, !INVERSE, v, kn, b); break;
        default: rqs_locate<0>(row, K, !INVERSE, v, kn, b); break;
    }
    if (!INVERSE) {
        float y, ld;
        rqs_eval_forward(v, b, y, ld);
        xcol[tid] = y;
        ldc += ld;
    } else {
        xcol[tid] = rqs_eval_inverse(v, b);
    }
}

template <bool INVERSE>
__device__ __forceinline__ void run_coupling(const StepDesc& s, const float* __restrict__ wsf, int D, int C,
                                             float* xs, const float* cs, float* act0, float* act1,
                                             float* wst, float& ld_acc, int tid) {
    const int d = s.d, F = s.F, F_p = ru(F, KC), rot = s.rot;
    // ---- conditioner input: hstack(xc, c) then eval BatchNorm (bijectors.py:341-342)
    {
        const float* bn = wsf + s.off_bn;
        for (int e = tid; e < F_p * TM; e += kChainThreads) {
            const int f = e / TM, m = e - f * TM;
            float h = 0.f;
            if (f < F) {
                const float v = (f < D - d) ? xs[pmod(d + f - rot, D) * TM + m] : cs[(f - (D - d)) * TM + m];
                h = (v - bn[F_p + f]) * bn[f] + bn[2 * F_p + f];
            }
            act0[e] = h;
        }
    }
    __syncthreads();

    float* cur = act0;
    float* nxt = act1;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][8];
    int Kin_p = F_p;
    // ---- hidden layers: Dense + swish (bijectors.py:343-345)
    for (int l = 0; l < s.n_hidden; ++l) {
        const int N_p = ru(s.hidden[l], KC);
        const float* W = wsf + s.off_W[l];
        const float* bias = wsf + s.off_b[l];
        for (int n0 = 0; n0 < N_p; n0 += NCOL) {
            const int ncols = min(NCOL, N_p - n0);
            gemm_pass(cur, Kin_p, W + n0, N_p, ncols, wst, acc, tid);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = (j < 4) ? (ty * 4 + j) : (64 + ty * 4 + (j - 4));
                if (n < ncols) {
                    const float bj = bias[n0 + n];
                    float4 o;
                    o.x = swishf(acc[0][j] + bj);
                    o.y = swishf(acc[1][j] + bj);
                    o.z = swishf(acc[2][j] + bj);
                    o.w = swishf(acc[3][j] + bj);
                    *reinterpret_cast<float4*>(nxt + (n0 + n) * TM + tx * 4) = o;
                }
            }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
        Kin_p = N_p;
    }
    // ---- last Dense, one transformed dim at a time; theta rows go to the dead buffer
    const int K = s.K, P = 3 * K - 1, Pp = ru(P, 4), Pst = P | 1;
    const KnotNorm kn = make_knot_norm(K);
    const int L = s.n_hidden;
    float ldc = 0.f;
    for (int jj = 0; jj < d; ++jj) {
        const float* W = wsf + s.off_W[L] + (size_t)jj * Kin_p * Pp;
        const float* bias = wsf + s.off_b[L] + jj * Pp;
        gemm_pass(cur, Kin_p, W, Pp, Pp, wst, acc, tid);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = (j < 4) ? (ty * 4 + j) : (64 + ty * 4 + (j - 4));
            if (p < P) {
                const float bj = bias[p];
#pragma unroll
                for (int i = 0; i < 4; ++i) nxt[(tx * 4 + i) * Pst + p] = acc[i][j] + bj;
            }
        }
        __syncthreads();
        if (tid < TM) spline_rows<INVERSE>(nxt, Pst, K, kn, xs + pmod(jj - rot, D) * TM, ldc, tid);
        __syncthreads();
    }
    ld_acc += ldc;  // Chain: log_det += ld   (bijectors.py:110)
}

template <bool INVERSE>
__device__ __forceinline__ void run_shift_bounds(const StepDesc& s, const float* __restrict__ wsf, int D,
                                                 float* xs, float& ld_acc, int tid) {
    if (tid < TM) {
        float ldc = 0.f;
        for (int i = 0; i < D; ++i) {
            const float* t = wsf + s.off_sb + i * kSbStride;
            const int kind = (int)t[0];
            const float a = t[1], b = t[2], xmin = t[3], xmax = t[4], mul = t[5], logmul = t[6];
            float* px = xs + pmod(i - s.rot, D) * TM + tid;
            const float v = *px;
            if (!INVERSE) {  // bijectors.py:183-207
                float z, ld;
                if (kind == ZF_BOUND_BOTH) {
                    z = __fmul_rn(__fsub_rn(v, a), mul);
                    ld = logmul;
                } else {
                    float u = v;
                    if (kind == ZF_BOUND_LOWER) u = logf(__fadd_rn(__fsub_rn(v, a), FLT_MIN));
                    if (kind == ZF_BOUND_UPPER) u = logf(__fadd_rn(__fsub_rn(b, v), FLT_MIN));
                    z = clip_nanprop(__fmul_rn(__fsub_rn(u, xmin), mul), 0.f, 1.f);
                    ld = (kind == ZF_BOUND_NONE) ? logmul : (logmul - u);
                }
                *px = z;
                ldc += ld;
            } else {  // bijectors.py:214-238
                float x;
                if (kind == ZF_BOUND_BOTH) {
                    x = __fadd_rn(__fmul_rn(v, b), __fmul_rn(__fsub_rn(1.f, v), a));
                } else {
                    float u = __fadd_rn(__fmul_rn(v, xmax), __fmul_rn(__fsub_rn(1.f, v), xmin));
                    if (kind == ZF_BOUND_LOWER) x = expf(u) + a;
                    else if (kind == ZF_BOUND_UPPER) x = b - expf(u);
                    else x = u;
                }
                *px = x;
            }
        }
        ld_acc += ldc;
    }
    __syncthreads();
}

template <bool INVERSE>
__global__ void __launch_bounds__(kChainThreads, 2) chain_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int D = a.D, C = a.C;
    float* xs = smem;
    float* cs = xs + D * TM;
    float* act0 = cs + C * TM;
    float* act1 = act0 + a.act_rows * TM;
    float* wst = act1 + a.act_rows * TM;
    const StepDesc* steps = reinterpret_cast<const StepDesc*>(a.ws);
    const float* wsf = a.ws;

    const long long n_tiles = (a.M + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long m0 = tile * TM;
        const int nm = (int)min((long long)TM, a.M - m0);
        // ---- load the tile, feature-major; padding samples sit at 0.5 and are never stored
        const int rot_in = INVERSE ? a.rot_total : 0;
        for (int e = tid; e < TM * D; e += kChainThreads) {
            const int m = e / D, j = e - m * D;
            const float v = (m < nm) ? a.x[m0 * D + e] : 0.5f;
            xs[pmod(j - rot_in, D) * TM + m] = v;
        }
        for (int e = tid; e < TM * C; e += kChainThreads) {
            const int m = e / C, j = e - m * C;
            cs[j * TM + m] = (m < nm) ? a.c[m0 * C + e] : 0.f;
        }
        __syncthreads();

        float ld_acc = 0.f;  // Chain: log_det = zeros(M)   (bijectors.py:107)
        for (int si = 0; si < a.n_steps; ++si) {
            const StepDesc& s = steps[INVERSE ? (a.n_steps - 1 - si) : si];
            if (s.kind == kStepKindShiftBounds) run_shift_bounds<INVERSE>(s, wsf, D, xs, ld_acc, tid);
            else run_coupling<INVERSE>(s, wsf, D, C, xs, cs, act0, act1, wst, ld_acc, tid);
        }

        // ---- store
        if (a.mode == kModeLogProb) {
            if (tid < nm) {  // flow.py:46-47
                float lat = 0.f;
                for (int j = 0; j < D; ++j) lat += latent_logpdf(xs[pmod(j - a.rot_total, D) * TM + tid], a.lc);
                a.lp[m0 + tid] = nan_to_num_lp(lat + ld_acc);
            }
        } else {
            const int rot_out = INVERSE ? 0 : a.rot_total;
            if (a.y) {
                for (int e = tid; e < nm * D; e += kChainThreads) {
                    const int m = e / D, j = e - m * D;
                    a.y[m0 * D + e] = xs[pmod(j - rot_out, D) * TM + m];
                }
            }
            if (!INVERSE && a.log_det && tid < nm) a.log_det[m0 + tid] = a.acc_log_det ? a.log_det[m0 + tid] + ld_acc : ld_acc;
        }
        __syncthreads();
    }
}

// ---- host side ------------------------------------------------------------------------

struct Plan {
    std::vector<PackJob> jobs;
    int rot_total = 0;
    int act_rows = KC;
    size_t ws_floats = 0;
};

static int build_plan(const zf_chain* chain, Plan& plan) {
    ZF_REQUIRE(chain != nullptr, "chain is NULL");
    const int D = chain->dim, C = chain->cdim;
    ZF_REQUIRE(D >= 1 && D <= ZF_MAX_DIM, "dim must be in [1, %d] (got %d)", ZF_MAX_DIM, D);
    ZF_REQUIRE(C >= 0 && C <= 1024, "cdim must be in [0, 1024] (got %d)", C);
    ZF_REQUIRE(chain->n_ops >= 0 && (chain->n_ops == 0 || chain->ops), "ops is NULL");
    int rot = 0;
    for (int i = 0; i < chain->n_ops; ++i) {
        const zf_op& op = chain->ops[i];
        if (op.kind == ZF_OP_ROLL) {
            rot = ((rot + op.shift) % D + D) % D;
            continue;
        }
        PackJob job{};
        job.D = D;
        job.desc.rot = rot;
        if (op.kind == ZF_OP_SHIFT_BOUNDS) {
            ZF_REQUIRE(op.shift_bounds, "op %d: shift_bounds is NULL", i);
            const zf_shift_bounds& sb = *op.shift_bounds;
            job.desc.kind = kStepKindShiftBounds;
            bool need_stats = false;
            for (int j = 0; j < D; ++j) {
                ZF_REQUIRE(sb.kind[j] >= 0 && sb.kind[j] <= 3, "op %d: bad bound kind for column %d", i, j);
                job.sb_kind[j] = sb.kind[j];
                job.lo[j] = sb.lo[j];
                job.hi[j] = sb.hi[j];
                if (sb.kind[j] != ZF_BOUND_BOTH) need_stats = true;
            }
            ZF_REQUIRE(!need_stats || (sb.xmin && sb.xmax), "op %d: xmin/xmax are NULL", i);
            job.xmin = sb.xmin;
            job.xmax = sb.xmax;
        } else if (op.kind == ZF_OP_COUPLING) {
            ZF_REQUIRE(op.coupling, "op %d: coupling is NULL", i);
            const zf_coupling& cp = *op.coupling;
            const int d = D / 2;
            ZF_REQUIRE(d > 0 && d < D, "NeuralSplineCoupling needs dim >= 2 (bijectors.py:326)");
            ZF_REQUIRE(cp.knots >= 1, "op %d: knots must be >= 1", i);
            ZF_REQUIRE(cp.n_hidden >= 0 && cp.n_hidden <= ZF_MAX_LAYERS, "op %d: at most %d hidden layers", i,
                       ZF_MAX_LAYERS);
            const int P = 3 * cp.knots - 1;
            if (ru(P, 4) > NCOL)
                return fail(ZF_ERR_UNSUPPORTED, "op %d: knots=%d needs %d columns per dim; the fused kernel holds %d",
                            i, cp.knots, P, NCOL);
            job.desc.kind = kStepKindCoupling;
            job.desc.K = cp.knots;
            job.desc.n_hidden = cp.n_hidden;
            job.desc.F = D - d + C;
            job.desc.d = d;
            ZF_REQUIRE(cp.bn_scale && cp.bn_bias && cp.bn_mean && cp.bn_var, "op %d: BatchNorm leaf is NULL", i);
            job.bn_scale = cp.bn_scale; job.bn_bias = cp.bn_bias; job.bn_mean = cp.bn_mean; job.bn_var = cp.bn_var;
            int kin = job.desc.F;
            plan.act_rows = std::max(plan.act_rows, ru(kin, KC));
            for (int l = 0; l <= cp.n_hidden; ++l) {
                ZF_REQUIRE(cp.kernel[l] && cp.bias[l], "op %d: Dense_%d leaf is NULL", i, l);
                ZF_REQUIRE((reinterpret_cast<uintptr_t>(cp.kernel[l]) & 3) == 0, "op %d: Dense_%d kernel misaligned", i, l);
                const int n = (l < cp.n_hidden) ? cp.hidden[l] : d * P;
                ZF_REQUIRE(n >= 1, "op %d: layer width must be positive", i);
                if (l < cp.n_hidden) job.desc.hidden[l] = n;
                job.kernel[l] = cp.kernel[l];
                job.bias[l] = cp.bias[l];
                job.Kin[l] = kin;
                job.N[l] = n;
                if (l < cp.n_hidden) plan.act_rows = std::max(plan.act_rows, ru(n, KC));
                kin = n;
            }
            plan.act_rows = std::max(plan.act_rows, ru((P | 1), 4));
        } else {
            return fail(ZF_ERR_INVALID, "op %d: unknown kind %d", i, op.kind);
        }
        plan.jobs.push_back(job);
    }
    plan.rot_total = rot;

    // workspace layout: StepDesc array, then per-step blocks (all multiples of 4 floats)
    size_t off = (sizeof(StepDesc) * std::max<size_t>(plan.jobs.size(), 1) + 15) / 16 * 4;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return (int)o; };
    for (size_t si = 0; si < plan.jobs.size(); ++si) {
        PackJob& job = plan.jobs[si];
        job.step_index = (int)si;
        StepDesc& s = job.desc;
        if (s.kind == kStepKindShiftBounds) {
            s.off_sb = take((size_t)D * kSbStride);
        } else {
            const int F_p = ru(s.F, KC);
            s.off_bn = take((size_t)3 * F_p);
            for (int l = 0; l < s.n_hidden; ++l) {
                const int Kin_p = ru(job.Kin[l], KC), N_p = ru(job.N[l], KC);
                s.off_W[l] = take((size_t)Kin_p * N_p);
                s.off_b[l] = take((size_t)N_p);
            }
            const int L = s.n_hidden;
            const int Kin_p = ru(job.Kin[L], KC), Pp = ru(3 * s.K - 1, 4);
            s.off_W[L] = take((size_t)s.d * Kin_p * Pp);
            s.off_b[L] = take((size_t)s.d * Pp);
        }
        if (off > (size_t)0x7fffffff) return fail(ZF_ERR_UNSUPPORTED, "packed parameters exceed 2^31 floats");
    }
    plan.ws_floats = off;
    return ZF_OK;
}

static LatentConst make_latent(int kind, float peakness) {
    LatentConst lc{};
    lc.kind = kind;
    lc.p1 = (float)((double)peakness - 1.0);
    lc.betaln = (float)(2.0 * lgamma((double)peakness) - lgamma(2.0 * (double)peakness));
    lc.lognorm = (float)log(2.0 * M_PI * 0.1 * 0.1);
    lc.logmass = (float)log(0.5 * (erf(5.0 / sqrt(2.0)) - erf(-5.0 / sqrt(2.0))));
    return lc;
}

static int run_chain(cudaStream_t stream, const zf_chain* chain, int mode, int latent_kind, float peakness,
                     const float* x, const float* c, long long M, float* y, float* log_det, float* lp,
                     void* workspace, size_t workspace_bytes, int acc_log_det = 0) {
    Plan plan;
    if (int rc = build_plan(chain, plan)) return rc;
    ZF_REQUIRE(M >= 0, "M must be >= 0");
    if (M == 0) return ZF_OK;
    ZF_REQUIRE(x != nullptr, "input tensor is NULL");
    ZF_REQUIRE(chain->cdim == 0 || c != nullptr, "chain has cdim=%d but c is NULL", chain->cdim);
    if (mode == kModeLogProb) {
        ZF_REQUIRE(lp != nullptr, "log_prob output is NULL");
        ZF_REQUIRE(latent_kind >= 0 && latent_kind <= 3, "unknown latent kind %d", latent_kind);
        ZF_REQUIRE(latent_kind != ZF_LATENT_BETA || peakness >= 1.f, "peakness must be at least 1 (distributions.py:96-97)");
    }
    const size_t need = plan.ws_floats * sizeof(float);
    if (workspace_bytes < need || (need && !workspace))
        return fail(ZF_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    ZF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;

    float* ws = static_cast<float*>(workspace);
    for (const PackJob& job : plan.jobs) {
        pack_step_kernel<<<32, 256, 0, stream>>>(job, ws);
        count_launch();
    }
    ZF_CUDA_CHECK(cudaGetLastError());

    ChainArgs a{};
    a.x = x; a.c = c; a.y = y; a.log_det = log_det; a.lp = lp;
    a.ws = ws; a.M = M; a.D = chain->dim; a.C = chain->cdim;
    a.n_steps = (int)plan.jobs.size();
    a.rot_total = plan.rot_total;
    a.act_rows = plan.act_rows;
    a.mode = mode;
    a.acc_log_det = acc_log_det;
    a.lc = make_latent(latent_kind, peakness);

    const size_t smem = ((size_t)(a.D + a.C) * TM + 2 * (size_t)a.act_rows * TM + 2 * KC * NCOL) * sizeof(float);
    if (smem > (size_t)di.max_smem_optin)
        return fail(ZF_ERR_UNSUPPORTED, "chain tile needs %zu bytes of shared memory (limit %d): layers too wide", smem,
                    di.max_smem_optin);
    const int bps = (2 * (smem + 1024) <= (size_t)di.max_smem_optin) ? 2 : 1;
    const long long n_tiles = (M + TM - 1) / TM;
    const unsigned grid = (unsigned)std::min<long long>(n_tiles, (long long)di.sm_count * bps);

    if (mode == kModeInverse) {
        ZF_CUDA_CHECK(cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chain_kernel<true><<<grid, kChainThreads, smem, stream>>>(a);
    } else {
        ZF_CUDA_CHECK(cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chain_kernel<false><<<grid, kChainThreads, smem, stream>>>(a);
    }
    count_launch();
    ZF_CUDA_CHECK(cudaGetLastError());
    return ZF_OK;
}

}  // namespace zf

extern "C" size_t zf_chain_workspace_bytes(const zf_chain* chain, int64_t M) {
    (void)M;
    zf::Plan plan;
    if (zf::build_plan(chain, plan) != ZF_OK) return 0;
    return plan.ws_floats * sizeof(float);
}

extern "C" int zf_chain_forward(void* stream, const zf_chain* chain, const float* x, const float* c, int64_t M,
                                float* y, float* log_det, void* workspace, size_t workspace_bytes) {
    return zf::run_chain((cudaStream_t)stream, chain, zf::kModeForward, 0, 0.f, x, c, (long long)M, y, log_det,
                         nullptr, workspace, workspace_bytes);
}

extern "C" int zf_chain_forward_acc(void* stream, const zf_chain* chain, const float* x, const float* c, int64_t M,
                                    float* y, float* log_det, void* workspace, size_t workspace_bytes) {
    return zf::run_chain((cudaStream_t)stream, chain, zf::kModeForward, 0, 0.f, x, c, (long long)M, y, log_det,
                         nullptr, workspace, workspace_bytes, 1);
}

extern "C" int zf_chain_inverse(void* stream, const zf_chain* chain, const float* z, const float* c, int64_t M,
                                float* x, void* workspace, size_t workspace_bytes) {
    ZF_REQUIRE(x != nullptr || M == 0, "output tensor is NULL");
    return zf::run_chain((cudaStream_t)stream, chain, zf::kModeInverse, 0, 0.f, z, c, (long long)M, x, nullptr,
                         nullptr, workspace, workspace_bytes);
}

extern "C" int zf_flow_log_prob(void* stream, const zf_chain* chain, int32_t latent_kind, float peakness,
                                const float* x, const float* c, int64_t M, float* log_prob, void* workspace,
                                size_t workspace_bytes) {
    return zf::run_chain((cudaStream_t)stream, chain, zf::kModeLogProb, latent_kind, peakness, x, c, (long long)M,
                         nullptr, nullptr, log_prob, workspace, workspace_bytes);
}
