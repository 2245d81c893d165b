// navppo_tc.cu — tcgen05 tensor-core path of the PPO update (NAVPPO_TF32) + a GEMM self-test.
//
// See tc_common.cuh for the shared-memory operand format and the descriptor conventions.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/navppo.h"
#include "nav_common.h"
#include "tc_common.cuh"

namespace {

extern __shared__ __align__(128) unsigned char tc_smem[];

// ----------------------------------------------------------------------------------------
// Self-test: D[128, N] = A[128, K] * B[N, K]^T with every operand role the fused kernel uses.
// a_mode / b_mode: 0 = rows are the M/N index ("K-major"), 1 = rows are K ("MN-major").
// The descriptor strides are passed in so a test can probe the hardware's conventions.
// ----------------------------------------------------------------------------------------
struct SelfTestArgs {
  const float* A;   // [128, K] row-major
  const float* B;   // [N, K] row-major
  float* D;         // [128, N]
  int N, K, a_mode, b_mode;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_major, b_major;
};

__global__ void __launch_bounds__(128) tc_selftest_kernel(SelfTestArgs p) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = tc_smem;
  unsigned char* sB = tc_smem + 128 * p.K * 4;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 128 * p.K; i += 128) {
    const int m = i / p.K, k = i % p.K;
    const uint32_t off = p.a_mode == 0 ? tc::rb_off(128, m, k) : p.a_mode == 1 ? tc::rb_off(p.K, k, m) : 4u * i;
    *reinterpret_cast<float*>(sA + off) = tc::to_tf32(p.A[i]);
  }
  for (int i = tid; i < p.N * p.K; i += 128) {
    const int n = i / p.K, k = i % p.K;
    const uint32_t off = p.b_mode == 0 ? tc::rb_off(p.N, n, k) : p.b_mode == 1 ? tc::rb_off(p.K, k, n) : 4u * i;
    *reinterpret_cast<float*>(sB + off) = tc::to_tf32(p.B[i]);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, p.N, p.a_major, p.b_major);
    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
    for (int kk = 0; kk < p.K / 8; ++kk)
      tc::mma_tf32(tmem, tc::make_desc(a0 + kk * p.a_kstep, p.a_lbo, p.a_sbo),
                   tc::make_desc(b0 + kk * p.b_kstep, p.b_lbo, p.b_sbo), idesc, kk > 0);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) p.D[(size_t)tid * p.N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// Same self-test with BF16 operands (kind::f16): `passes` = 1 (hi * hi) or 3 (hi*hi + lo*hi +
// hi*lo on the split x = hi + lo).  a_mode / b_mode: 0 = rows are the M/N index, 1 = rows are K.
struct SelfTest16Args {
  const float* A; const float* B; float* D;
  int N, K, a_mode, b_mode, passes;
};

__global__ void __launch_bounds__(128) tc_selftest_bf16_kernel(SelfTest16Args p) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t a_bytes = 128 * p.K * 2, b_bytes = p.N * p.K * 2;
  unsigned char* sAh = tc_smem;
  unsigned char* sAl = sAh + a_bytes;
  unsigned char* sBh = sAl + a_bytes;
  unsigned char* sBl = sBh + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 128 * p.K; i += 128) {
    const int m = i / p.K, k = i % p.K;
    const uint32_t off = p.a_mode == 0 ? tc::rb16_off(128, m, k) : tc::rb16_off(p.K, k, m);
    uint16_t h, l;
    tc::split_bf16(p.A[i], &h, &l);
    *reinterpret_cast<uint16_t*>(sAh + off) = h;
    *reinterpret_cast<uint16_t*>(sAl + off) = l;
  }
  for (int i = tid; i < p.N * p.K; i += 128) {
    const int n = i / p.K, k = i % p.K;
    const uint32_t off = p.b_mode == 0 ? tc::rb16_off(p.N, n, k) : tc::rb16_off(p.K, k, n);
    uint16_t h, l;
    tc::split_bf16(p.B[i], &h, &l);
    *reinterpret_cast<uint16_t*>(sBh + off) = h;
    *reinterpret_cast<uint16_t*>(sBl + off) = l;
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, p.N, p.a_mode, p.b_mode);
    // rows = M/N index: SBO = 128, LBO = R*16, 16 k per instruction = 2 granule columns
    // rows = K:         SBO = R*16, LBO = 128, 16 k per instruction = 2 groups of 8 rows
    const uint32_t a_lbo = p.a_mode == 0 ? 128 * 16 : 128, a_sbo = p.a_mode == 0 ? 128 : p.K * 16;
    const uint32_t a_step = p.a_mode == 0 ? 2 * 128 * 16 : 256;
    const uint32_t b_lbo = p.b_mode == 0 ? p.N * 16 : 128, b_sbo = p.b_mode == 0 ? 128 : p.K * 16;
    const uint32_t b_step = p.b_mode == 0 ? 2 * p.N * 16 : 256;
    uint32_t acc = 0;
    for (int kk = 0; kk < p.K / 16; ++kk) {
      const uint64_t ah = tc::make_desc(tc::smem_u32(sAh) + kk * a_step, a_lbo, a_sbo);
      const uint64_t al = tc::make_desc(tc::smem_u32(sAl) + kk * a_step, a_lbo, a_sbo);
      const uint64_t bh = tc::make_desc(tc::smem_u32(sBh) + kk * b_step, b_lbo, b_sbo);
      const uint64_t bl = tc::make_desc(tc::smem_u32(sBl) + kk * b_step, b_lbo, b_sbo);
      if (p.passes == 3) {   // small terms first
        tc::mma_bf16(tmem, al, bh, idesc, acc); acc = 1;
        tc::mma_bf16(tmem, ah, bl, idesc, acc);
      }
      tc::mma_bf16(tmem, ah, bh, idesc, acc); acc = 1;
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) p.D[(size_t)tid * p.N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" {

int navppo_tc_selftest_bf16(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                            int32_t passes, void* stream) {
  if (!A || !B || !D) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16) return nav_fail(NAVSIM_EINVAL, "N in 16..256 step 16, K in 16..128 step 16");
  if (passes != 1 && passes != 3) return nav_fail(NAVSIM_EINVAL, "passes is 1 or 3");
  SelfTest16Args p{A, B, D, N, K, a_mode, b_mode, passes};
  const size_t smem = (size_t)(128 + N) * K * 4;
  NAV_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                       const uint32_t* strides6, void* stream) {
  if (!A || !B || !D || !strides6) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (N < 16 || N > 256 || N % 16 || K < 8 || K > 128 || K % 8) return nav_fail(NAVSIM_EINVAL, "N in 16..256 step 16, K in 8..128 step 8");
  // strides6[6], [7]: major bits of the instruction descriptor (0 = K-major, 1 = MN-major).
  // a_mode / b_mode == 2: the global buffer is copied verbatim into shared memory (raw image).
  SelfTestArgs p{A, B, D, N, K, a_mode, b_mode, strides6[0], strides6[1], strides6[2], strides6[3], strides6[4], strides6[5],
                 strides6[6], strides6[7]};
  const size_t smem = (size_t)(128 + N) * K * 4;
  NAV_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

}  // extern "C"
