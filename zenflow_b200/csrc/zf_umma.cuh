// tcgen05 / TMEM wrappers (sm_100a) for the conditioner GEMMs: 3xTF32 split products with the
// activation operand A living in tensor memory (lane = event, column = k), the weight operand B in
// shared memory (K-major, no swizzle: 8-row x 16-byte core matrices), fp32 accumulators in tensor
// memory.  PTX forms and descriptor bit layouts follow the PTX ISA / CUTLASS's cute/arch/*sm100*.
#pragma once
#include "zf_common.cuh"

#include <cuda_fp16.h>

namespace zf {
namespace umma {

// ---- tensor memory management (warp-collective) -------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tensor-memory address: lane in bits [16,32), column in bits [0,16)
__device__ __forceinline__ uint32_t taddr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// 32 lanes x 32 bit, 8 / 16 / 32 consecutive columns per lane (warp-collective; lane i <-> TMEM lane base+i)
__device__ __forceinline__ void ld8(uint32_t a, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16(uint32_t a, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st8(uint32_t a, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

__device__ __forceinline__ void st16(uint32_t a, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// ---- 3xTF32 split -------------------------------------------------------------------------------
// hi = x rounded to tf32 (round to nearest, ties away from zero: exactly cvt.rna.tf32.f32, but as two
// full-rate integer ops instead of a trip through the conversion unit); lo = x - hi is exact in fp32
// (|lo| <= 2^-11 |x|, at most 13 significant bits) and is handed to the tensor core as is: reading it as
// tf32 drops at most its last 3 bits, i.e. <= 2^-22 |x|, the same bound as rounding it.
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = to_tf32(x);
    lo = x - hi;
}

// ---- 3xFP16 split ----------------------------------------------------------------------------------
// The same idea on kind::f16 (twice the tensor rate of kind::tf32, half the operand bytes): hi = x rounded to
// fp16 (11 significant bits, like tf32), lo' = (x - hi) * 2^11 rounded to fp16.  hi * hi products are exact in
// the fp32 accumulator; the two cross products are accumulated SEPARATELY at scale 2^11 and added as
// cross * 2^-11 in the epilogue (the scaling keeps lo' in fp16's normal range whenever hi is).  Valid for
// |x| < 65504 (fp16 range); tiny |x| < 2^-14 lose relative but not absolute accuracy (error < 2^-25 + 2^-36).
constexpr float kF16LoScale = 2048.0f, kF16LoUnscale = 1.0f / 2048.0f;
// two consecutive reduction indices (k even in the low half-word) -> one 32-bit word of hi parts, one of lo' parts
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((x0 - hf.x) * kF16LoScale, (x1 - hf.y) * kF16LoScale);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__host__ __device__ constexpr uint32_t instr_desc_f16(int N) {   // kind::f16, fp16 x fp16 -> fp32, K-major, M = 128
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem, fp16 pairs] * B[smem, fp16]; K = 16 per instruction; issued by ONE thread
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
// ---- bf16 x 2 split (train-step GEMMs on event-row images) -------------------------------------------------
// hi = x rounded to bf16, lo = (x - hi) rounded to bf16: 16 significant bits, fp32's exponent range (gradients as
// small as 1 / global_count stay normal).  Products hi*hi + hi*lo + lo*hi on kind::f16 with bf16 operands.
// idesc: a_format / b_format = 1 (bf16); a_mn / b_mn: the operand is MN-major in shared memory (bits 15 / 16).
__host__ __device__ constexpr uint32_t instr_desc_bf16(int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    // round-to-nearest-even bf16 of both, packed (x0 in the low half-word)
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
    uint32_t l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
    hi = h;
    lo = l;
}
// D[tmem] (+)= A[smem] * B[smem]; K = 16 per instruction; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
// D[tmem] = A[tmem, fp16 pairs] * B[smem] + D * 2^-SCALE (scale-input-d): folds accumulated lo' cross products (scale 2^11)
// under the main product in ONE accumulator
template <int SCALE>
__device__ __forceinline__ void mma_f16_ts_scaled(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %5, %6, %7}, p, %8;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(z), "r"(z), "r"(z), "r"(z), "n"(SCALE)
        : "memory");
}
__device__ __forceinline__ void st4u(uint32_t a, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}
// half-word offset of element (r, k) of an R-row operand in an "event-row" image: [r/8][k/8][k%8][r%8] is the MN-major
// reading of the same bytes a K-major image of the transposed operand has
__host__ __device__ inline int mn_image_index_bf16(int r, int k, int Ktile) { return ((r >> 3) * (Ktile >> 3) + (k >> 3)) * 64 + (k & 7) * 8 + (r & 7); }
// half-word offset of element (n, k) in an fp16 B image of N rows: [k/8][n/8][n%8][k%8]
__host__ __device__ inline int b_image_index_f16(int n, int k, int N) { return ((k >> 3) * (N >> 3) + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7); }
__device__ __forceinline__ void st8u(uint32_t a, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// K-major, SWIZZLE_NONE: in 16-byte units the operand is ((8,n),2):((1,SBO),LBO): 8 rows of one core
// matrix are 16 B apart, 8-row groups SBO apart, the two 16-byte K halves of one MMA (K = 8 tf32) LBO apart.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;  // descriptor version (Blackwell)
    return d;         // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N in [16, 256] step 16
__host__ __device__ constexpr uint32_t instr_desc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
// same, with an output-lane mask: bit i of (m0..m3) set = tensor-memory lane i of D is NOT written
__device__ __forceinline__ void mma_tf32_ts_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                   bool accumulate, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// float offset of element (n, k) in a B image of N rows: [k/4][n/8][n%8][k%4]
__host__ __device__ inline int b_image_index(int n, int k, int N) { return ((k >> 2) * (N >> 3) + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3); }

}  // namespace umma
}  // namespace zf
