// tc_common.cuh — tcgen05 / TMEM / mbarrier building blocks for the sm_100a tensor-core kernels.
//
// Operand tiles live in shared memory in ONE format, the "row-block" format
//     element (row r, column c) of an R-row tile  ->  byte  (c / 4) * (R * 16) + r * 16 + (c % 4) * 4
// i.e. an array [c / 4][r] of float4.  A block of 8 rows x 4 columns is exactly one tcgen05
// "core matrix" (8 x 16 bytes, rows 16 bytes apart) of the un-swizzled canonical layouts, so the
// same tile can be handed to the tensor core in either role just by swapping the two strides
// of its shared-memory descriptor:
//   * rows = the M/N index, columns = K  ("K-major"):   SBO = 128 (next 8 rows), LBO = R*16 (next 4 k)
//         one kind::tf32 instruction consumes K = 8 -> advance the start address by 2 * R*16
//   * rows = K, columns = the M/N index  ("MN-major"):   SBO = R*16 (next 4 m/n), LBO = 128 (next 8 k)
//         one instruction consumes 8 rows -> advance the start address by 128
// A thread that owns row r writes 4 consecutive columns with one 16-byte st.shared; a warp's 32
// rows are 512 contiguous bytes, so tile stores are bank-conflict free.
#ifndef NAVPPO_TC_COMMON_CUH_
#define NAVPPO_TC_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- shared-memory matrix descriptor (un-swizzled), cute::UMMA::SmemDescriptor bit layout:
// [0,14) start >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
// [46,48) version = 1 (Blackwell) | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

// ---- instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor):
// [4,6) D format 1 = F32 | [7,10) A format 2 = TF32 | [10,13) B format 2 = TF32 |
// [15] A major (0 K, 1 MN) | [16] B major | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- instruction descriptor for kind::f16 with BF16 operands, fp32 accumulate
// (A / B format 1 = BF16)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// All previously issued MMAs of this thread arrive on the mbarrier when they complete.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (thread = lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then tmem_ld_wait() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (thread = lane)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A is K-major in tensor memory (lane = row, two bf16 per 32-bit column)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// round-to-nearest fp32 -> tf32 (kept in an fp32 container); the tensor core would otherwise
// truncate the low 13 mantissa bits
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

// byte offset of element (r, c) in a row-block tile of R rows
__device__ __forceinline__ uint32_t rb_off(int R, int r, int c) {
  return (uint32_t)((c >> 2) * (R * 16) + r * 16 + (c & 3) * 4);
}

// ---- bf16 row-block tiles: [c / 8][r] of 16-byte granules holding 8 bf16 columns
__device__ __forceinline__ uint32_t rb16_off(int R, int r, int c) {
  return (uint32_t)((c >> 3) * (R * 16) + r * 16 + (c & 7) * 2);
}

// x = hi + lo with hi = bf16(x) (round to nearest even), lo = bf16(x - hi): 16 mantissa bits
__device__ __forceinline__ void split_bf16(float x, uint16_t* hi, uint16_t* lo) {
  uint16_t h, l;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
  const float hf = __uint_as_float((uint32_t)h << 16);
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
  *hi = h; *lo = l;
}
// ---- packed fp32 pairs (FADD2 / FMUL2: one issue slot for two IEEE operations — the epilogues are issue-bound)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// two floats -> packed bf16x2 (lo half = a, hi half = b) and the packed residuals
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t* hi, uint32_t* lo) {
  uint32_t h;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  float la, lb;
  unpack2(sub2(pack2(a, b), pack2(__uint_as_float(h << 16), __uint_as_float(h & 0xFFFF0000u))), la, lb);
  uint32_t l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(lb), "f"(la));
  *hi = h; *lo = l;
}
// H = lrelu(v + b) of four values (slope `leak` < 1: lrelu(x) = max(x, leak x)), packed: hi[0..1], lo[0..1]
__device__ __forceinline__ void bias_lrelu_split4(const float* v, float4 b, float leak, uint32_t* hi, uint32_t* lo) {
  const uint64_t lk = pack2(leak, leak);
  const uint64_t p0 = add2(pack2(v[0], v[1]), pack2(b.x, b.y)), p1 = add2(pack2(v[2], v[3]), pack2(b.z, b.w));
  const uint64_t q0 = mul2(p0, lk), q1 = mul2(p1, lk);
  float x0, x1, x2, x3, y0, y1, y2, y3;
  unpack2(p0, x0, x1); unpack2(p1, x2, x3); unpack2(q0, y0, y1); unpack2(q1, y2, y3);
  split_bf16x2(fmaxf(x0, y0), fmaxf(x1, y1), &hi[0], &lo[0]);
  split_bf16x2(fmaxf(x2, y2), fmaxf(x3, y3), &hi[1], &lo[1]);
}

}  // namespace tc

#endif  // NAVPPO_TC_COMMON_CUH_
