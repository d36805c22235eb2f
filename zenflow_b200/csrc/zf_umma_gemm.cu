// tcgen05 GEMM family for the train step's conditioner recompute and VJP (fp32 in / fp32 out,
// 3xTF32 split products, fp32-class accuracy):
//   mode 0 (NN): C[m][n]  = sum_k opA(A[m][k]) * B[k][n] + bias[n]                 (Dense forward)
//   mode 1 (NT): C[m][k]  = (sum_n A[m][n] * B[k][n]) * swish'(Z[m][k])           (grad wrt Dense input)
//   mode 2 (TN): C[k][n] += sum_m opA(A[m][k]) * B[m][n];  colsum[n] += sum_m B[m][n]   (grad wrt kernel, bias)
// opA = swish when a_swish (stored pre-activations are re-activated on load).
//
// One CTA computes a 128 x TN output tile.  8 loader warps read the fp32 operands from global memory,
// apply opA, split x = hi + lo (tf32) and write K-major core-matrix images (zf_umma.cuh) of a 32-deep
// reduction chunk into a shared-memory ring; warp 8 issues tcgen05.mma.kind::tf32 with both operands
// from shared memory; the main products accumulate in TMEM columns [0,TN), the two cross products in
// [TN,2TN) (the tensor core's accumulator truncates per step, see zf_chain.cu); the loader warps then
// turn into the epilogue (TMEM -> registers -> global, bias / swish' / atomics).
#include "zf_umma.cuh"

#include <algorithm>
#include <stdlib.h>

namespace zf {

void count_launch();

struct UGemmArgs {
    const float* A; long long lda;
    const float* B; long long ldb;
    float* C; long long ldc;
    const float* bias;
    float* colsum;
    const float* Z; long long ldz;
    int a_swish;
    long long I, J, R;   // output rows, output cols, reduction length
    long long r_slab;    // mode 2: reduction rows per CTA (gridDim.z slabs)
};

constexpr int UG_THREADS = 288;  // 8 loader/epilogue warps + 1 MMA warp
template <int TN> struct UgCfg { static constexpr int KC = (TN == 256) ? 32 : 16; static constexpr int STAGES = (TN == 256) ? 2 : 3; };

__device__ __forceinline__ float ug_swish(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ug_swish_grad(float z) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-z));
    return s * (1.0f + z * (1.0f - s));
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}

// One 16-byte unit (4 consecutive reduction indices r4*4..+3 of image row `row`) of an operand image.
//   RCONTIG: X[(row0+row)*ld + r0 + r]      (reduction index contiguous in memory)
//   else   : X[(r0+r)*ld + row0 + row]      (transposed on load)
template <bool RCONTIG>
__device__ __forceinline__ float4 ug_load_unit(const float* __restrict__ X, long long ld, long long row, long long row_max,
                                               long long r, long long r_max, bool vec_ok) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef ZF_GEMM_EXP_NOLOAD   // timing experiment only
    return make_float4((float)row, (float)r, 1.f, 2.f);
#endif
    if (row < row_max) {
        if (RCONTIG && vec_ok && r + 3 < r_max) {
            o = __ldg(reinterpret_cast<const float4*>(X + row * ld + r));
        } else {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (r + i < r_max) v[i] = RCONTIG ? __ldg(X + row * ld + r + i) : __ldg(X + (r + i) * ld + row);
            }
            o = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    return o;
}

// Which 16-byte unit (image row, k-quad r4) thread `tid` handles in round q of a chunk of ROWS x R4 units.
//   transposed operands (reduction index strided in memory): lanes vary the row, so the four scalar loads of a
//     unit are coalesced along the row index and 8 lanes fill one 128-byte core matrix of the image;
//   reduction-contiguous operands: a quarter-warp takes 8 rows of one k-quad (conflict-free image stores) and the
//     four quarter-warps take 4 consecutive k-quads, so a warp load reads 64 contiguous bytes from each of 8 rows
//     (full sectors, 8 cache lines) instead of 16 bytes from each of 32 rows.
template <bool RCONTIG, int ROWS, int R4>
__device__ __forceinline__ void ug_unit(int tid, int q, int& row, int& r4) {
    const int u = tid + q * 256;
    if (!RCONTIG || R4 < 4) {
        row = u % ROWS;
        r4 = u / ROWS;
    } else {
        const int lane = u & 31, w = u >> 5;            // w: warp-sized group index in [0, ROWS * R4 / 32)
        constexpr int GROUPS_PER_ROWBLOCK = R4 / 4;      // groups that share the same 8 rows
        row = (w / GROUPS_PER_ROWBLOCK) * 8 + (lane & 7);
        r4 = (w % GROUPS_PER_ROWBLOCK) * 4 + (lane >> 3);
    }
}

template <int MODE, int TN>
__global__ void __launch_bounds__(UG_THREADS, (TN == 256) ? 1 : 2) umma_gemm_kernel(const __grid_constant__ UGemmArgs g) {
    constexpr int UG_KC = UgCfg<TN>::KC;       // reduction depth per ring stage
    constexpr int A_FLOATS = 128 * UG_KC;      // one image (hi or lo) of the A chunk
    constexpr int B_FLOATS = TN * UG_KC;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    constexpr int STAGES = UgCfg<TN>::STAGES;
    constexpr bool A_RCONTIG = (MODE != 2);    // NN, NT: A[m][r]; TN: A[r][i]
    constexpr bool B_RCONTIG = (MODE == 1);    // NT: B[k][n] = [j][r]; NN, TN: B[r][j]

    extern __shared__ __align__(128) float smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_FLOATS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    uint64_t* full = bars;            // [STAGES] count 256
    uint64_t* empty = bars + 3;       // [STAGES] count 1 (tcgen05.commit)
    uint64_t* done = bars + 6;        // count 1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long i0 = (long long)blockIdx.x * 128, j0 = (long long)blockIdx.y * TN;
    long long rbeg = 0, rend = g.R;
    if (MODE == 2) {
        rbeg = (long long)blockIdx.z * g.r_slab;
        rend = (rbeg + g.r_slab < g.R) ? rbeg + g.r_slab : g.R;
    }
    const int n_chunks = (int)((rend - rbeg + UG_KC - 1) / UG_KC);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 256); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 8) umma::tmem_alloc(tmem_slot, 2 * TN);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------------------------ loaders
        float csum = 0.f;  // mode 2: column sum of B for this thread's fixed column
        constexpr int R4 = UG_KC / 4;                               // 16-byte units per image row per chunk
        constexpr int UA = 128 * R4 / 256, UB = TN * R4 / 256;      // units per thread per chunk
        const bool vecA = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && ((rbeg & 3) == 0);
        const bool vecB = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && ((rbeg & 3) == 0);
        // the whole next chunk is fetched into registers while the current one is converted and stored
        auto fetch = [&](int c, float4 (&ra)[UA], float4 (&rb)[UB]) {
            const long long r0 = rbeg + (long long)c * UG_KC;
#pragma unroll
            for (int q = 0; q < UA; ++q) {
                int row, r4;
                ug_unit<A_RCONTIG, 128, R4>(tid, q, row, r4);
                ra[q] = ug_load_unit<A_RCONTIG>(g.A, g.lda, i0 + row, g.I, r0 + r4 * 4, rend, vecA);
            }
#pragma unroll
            for (int q = 0; q < UB; ++q) {
                int row, r4;
                ug_unit<B_RCONTIG, TN, R4>(tid, q, row, r4);
                rb[q] = ug_load_unit<B_RCONTIG>(g.B, g.ldb, j0 + row, g.J, r0 + r4 * 4, rend, vecB);
            }
        };
        auto store = [&](float* st, const float4 (&ra)[UA], const float4 (&rb)[UB]) {
#pragma unroll
            for (int q = 0; q < UA; ++q) {
                int row, r4;
                ug_unit<A_RCONTIG, 128, R4>(tid, q, row, r4);
                float4 v = ra[q];
                if (g.a_swish) { v.x = ug_swish(v.x); v.y = ug_swish(v.y); v.z = ug_swish(v.z); v.w = ug_swish(v.w); }
                float4 hi, lo;
                umma::split_tf32(v.x, hi.x, lo.x); umma::split_tf32(v.y, hi.y, lo.y);
                umma::split_tf32(v.z, hi.z, lo.z); umma::split_tf32(v.w, hi.w, lo.w);
                const int off = (r4 * 16 + (row >> 3)) * 32 + (row & 7) * 4;
#ifdef ZF_GEMM_EXP_NOSTORE   // timing experiment only
                if (hi.x == 1234.5f && lo.y == 0.25f)
#endif
                {
                *reinterpret_cast<float4*>(st + off) = hi;
                *reinterpret_cast<float4*>(st + A_FLOATS + off) = lo;
                }
            }
#pragma unroll
            for (int q = 0; q < UB; ++q) {
                int row, r4;
                ug_unit<B_RCONTIG, TN, R4>(tid, q, row, r4);
                const float4 v = rb[q];
                if (MODE == 2) csum += (v.x + v.y) + (v.z + v.w);
                float4 hi, lo;
                umma::split_tf32(v.x, hi.x, lo.x); umma::split_tf32(v.y, hi.y, lo.y);
                umma::split_tf32(v.z, hi.z, lo.z); umma::split_tf32(v.w, hi.w, lo.w);
                const int off = (r4 * (TN / 8) + (row >> 3)) * 32 + (row & 7) * 4;
#ifdef ZF_GEMM_EXP_NOSTORE
                if (hi.x == 1234.5f && lo.y == 0.25f)
#endif
                {
                *reinterpret_cast<float4*>(st + 2 * A_FLOATS + off) = hi;
                *reinterpret_cast<float4*>(st + 2 * A_FLOATS + B_FLOATS + off) = lo;
                }
            }
        };
        float4 ra0[UA], rb0[UB], ra1[UA], rb1[UB];
        uint32_t stage = 0, phase = 0;
        if (n_chunks > 0) fetch(0, ra0, rb0);
        for (int c = 0; c < n_chunks; c += 2) {
            if (c + 1 < n_chunks) fetch(c + 1, ra1, rb1);
            mbar_wait(&empty[stage], phase ^ 1u);
            store(smem + (size_t)stage * STAGE_FLOATS, ra0, rb0);
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
            umma::mbar_arrive(&full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            if (c + 1 < n_chunks) {
                if (c + 2 < n_chunks) fetch(c + 2, ra0, rb0);
                mbar_wait(&empty[stage], phase ^ 1u);
                store(smem + (size_t)stage * STAGE_FLOATS, ra1, rb1);
                fence_proxy_async_smem();
                umma::mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        if (MODE == 2 && g.colsum && blockIdx.x == 0) {
            // in the transposed B loader a thread always serves column (tid % TN) (+ 0 or 128 for TN=128 pairs)
            const int col = (TN == 256) ? tid : (tid & 127);
            if (TN == 128) {  // two threads share a column: combine through shuffle-free atomics
                if (j0 + col < g.J) atomicAdd(&g.colsum[j0 + col], csum);
            } else if (j0 + col < g.J) {
                atomicAdd(&g.colsum[j0 + col], csum);
            }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(done, 0);
        umma::fence_after_sync();
        const int q = warp & 3, half = warp >> 2;
        const long long i = i0 + q * 32 + lane;
        constexpr int HALF_COLS = TN / 2;
#pragma unroll 1
        for (int n0 = half * HALF_COLS; n0 < (half + 1) * HALF_COLS; n0 += 16) {
            float v[16], w[16];
            umma::ld16(umma::taddr(tb, q * 32, n0), v);
            umma::ld16(umma::taddr(tb, q * 32, TN + n0), w);
            umma::wait_ld();
#ifdef ZF_GEMM_EXP_NOEPI
            if (v[0] == 1234.5f)
#endif
            if (i < g.I) {
                float x[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) x[t] = v[t] + w[t];
                const long long jb = j0 + n0;
                const bool full16 = jb + 15 < g.J;
                if (MODE != 2 && full16 && (g.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 &&
                    (MODE != 0 || !g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) &&
                    (MODE == 0 || !g.Z || ((g.ldz & 3) == 0 && (reinterpret_cast<uintptr_t>(g.Z) & 15) == 0))) {
                    float4* dst = reinterpret_cast<float4*>(g.C + i * g.ldc + jb);
#pragma unroll
                    for (int t4 = 0; t4 < 4; ++t4) {
                        float4 o = make_float4(x[t4 * 4], x[t4 * 4 + 1], x[t4 * 4 + 2], x[t4 * 4 + 3]);
                        if (MODE == 0) {
                            if (g.bias) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + jb) + t4);  // bias + jb is 16B aligned when jb % 4 == 0
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            }
                        } else if (g.Z) {
                            const float4 zv = *(reinterpret_cast<const float4*>(g.Z + i * g.ldz + jb) + t4);
                            o.x *= ug_swish_grad(zv.x); o.y *= ug_swish_grad(zv.y);
                            o.z *= ug_swish_grad(zv.z); o.w *= ug_swish_grad(zv.w);
                        }
                        dst[t4] = o;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const long long j = jb + t;
                        if (j >= g.J) continue;
                        float xv = x[t];
                        if (MODE == 0) {
                            if (g.bias) xv += g.bias[j];
                            g.C[i * g.ldc + j] = xv;
                        } else if (MODE == 1) {
                            if (g.Z) xv *= ug_swish_grad(g.Z[i * g.ldz + j]);
                            g.C[i * g.ldc + j] = xv;
                        } else {
                            atomicAdd(&g.C[i * g.ldc + j], xv);
                        }
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma::instr_desc_tf32(TN);
        const uint32_t lbo_a = 16u * 128u, lbo_b = (uint32_t)(TN / 8) * 128u;
        uint32_t stage = 0, phase = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full[stage], phase);
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint32_t base = smem_u32(smem + (size_t)stage * STAGE_FLOATS);
                const uint32_t a_hi = base, a_lo = base + A_FLOATS * 4u;
                const uint32_t b_hi = base + 2u * A_FLOATS * 4u, b_lo = b_hi + B_FLOATS * 4u;
#pragma unroll
                for (int ks = 0; ks < UG_KC / 8; ++ks) {
                    const uint64_t dah = umma::smem_desc_kmajor(a_hi + ks * 2 * lbo_a, lbo_a, 128u);
                    const uint64_t dal = umma::smem_desc_kmajor(a_lo + ks * 2 * lbo_a, lbo_a, 128u);
                    const uint64_t dbh = umma::smem_desc_kmajor(b_hi + ks * 2 * lbo_b, lbo_b, 128u);
                    const uint64_t dbl = umma::smem_desc_kmajor(b_lo + ks * 2 * lbo_b, lbo_b, 128u);
                    const bool first = (c | ks) == 0;
                    mma_tf32_ss(tb + TN, dal, dbh, idesc, !first);
                    mma_tf32_ss(tb + TN, dah, dbl, idesc, true);
                    mma_tf32_ss(tb, dah, dbh, idesc, !first);
                }
                umma::commit(&empty[stage]);
                if (c == n_chunks - 1) umma::commit(done);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (n_chunks == 0 && lane == 0) umma::mbar_arrive(done);  // nothing to accumulate (never launched that way)
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tb, 2 * TN);
}

template <int MODE, int TN>
static int launch_one(cudaStream_t st, const UGemmArgs& g) {
    constexpr int STAGES = UgCfg<TN>::STAGES, UG_KC = UgCfg<TN>::KC;
    const size_t smem = ((size_t)STAGES * (2 * 128 * UG_KC + 2 * TN * UG_KC)) * sizeof(float) + 8 * 8 + 16;
    dim3 grid((unsigned)((g.I + 127) / 128), (unsigned)((g.J + TN - 1) / TN), 1);
    if (MODE == 2) grid.z = (unsigned)((g.R + g.r_slab - 1) / g.r_slab);
    if (grid.x == 0 || grid.y == 0 || g.R <= 0) return ZF_OK;
    ZF_CUDA_CHECK(cudaFuncSetAttribute(umma_gemm_kernel<MODE, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_gemm_kernel<MODE, TN><<<grid, UG_THREADS, smem, st>>>(g);
    count_launch();
    ZF_CUDA_CHECK(cudaGetLastError());
    return ZF_OK;
}

// Same argument convention as the FFMA gemm of zf_train.cu.  Wide outputs use 256-column tiles.
int launch_umma_gemm(cudaStream_t st, int mode, const float* A, long long lda, const float* B, long long ldb, float* C,
                     long long ldc, const float* bias, float* colsum, const float* Z, long long ldz, int a_swish,
                     long long I, long long J, long long R, long long r_slab) {
    UGemmArgs g{A, lda, B, ldb, C, ldc, bias, colsum, Z, ldz, a_swish, I, J, R, r_slab};
    const bool wide = J > 128;
    if (mode == 2) {
        // grad-weight: the output is tiny, the reduction (samples) is split into slabs so that at least
        // ~3 waves of CTAs are in flight; each CTA adds its partial tile with atomics
        const long long tiles = ((I + 127) / 128) * ((J + (wide ? 255 : 127)) / (wide ? 256 : 128));
        long long slabs = std::max<long long>(1, (3 * 148 + tiles - 1) / tiles);
        long long rs = (R + slabs - 1) / slabs;
        rs = std::max<long long>(256, (rs + 31) / 32 * 32);
        g.r_slab = rs;
    }
    if (mode == 0) return wide ? launch_one<0, 256>(st, g) : launch_one<0, 128>(st, g);
    if (mode == 1) return wide ? launch_one<1, 256>(st, g) : launch_one<1, 128>(st, g);
    return wide ? launch_one<2, 256>(st, g) : launch_one<2, 128>(st, g);
}

}  // namespace zf

extern "C" int zf_selftest_umma_gemm(void* stream, int32_t mode, const float* A, int64_t lda, const float* B, int64_t ldb,
                                     float* C, int64_t ldc, const float* bias, float* colsum, const float* Z, int64_t ldz,
                                     int32_t a_swish, int64_t I, int64_t J, int64_t R, int64_t r_slab) {
    ZF_REQUIRE(A && B && C && mode >= 0 && mode <= 2 && I >= 1 && J >= 1 && R >= 1, "selftest_umma_gemm: bad argument");
    ZF_REQUIRE(mode != 2 || (I <= 128 && r_slab >= 32 && r_slab % 32 == 0), "selftest_umma_gemm: mode 2 needs I <= 128 and r_slab % 32 == 0");
    return zf::launch_umma_gemm((cudaStream_t)stream, mode, A, lda, B, ldb, C, ldc, bias, colsum, Z, ldz, a_swish, I, J, R, r_slab);
}
