// tc_mma_bench.cu — standalone tcgen05 issue-rate probe (development tool, not part of the library).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/tc_mma_bench tools/tc_mma_bench.cu
//   tools/_bin/tc_mma_bench
//
// One CTA issues `reps` x K/16 back-to-back kind::f16 (BF16) MMAs of a given shape from one thread and
// reports cycles per MMA (clock64 from the first issue to the completion of the trailing commit), for
//   * A from shared memory, K-major or MN-major (the un-swizzled row-block tiles of tc_common.cuh),
//   * A from tensor memory (the "TS" form: lanes = M, two bf16 per 32-bit column),
//   * M = 128 and M = 64 (with the accumulator at TMEM lane offset 0 or 16),
// optionally while `stress` other warps stream 16-byte shared-memory stores (the epilogue's traffic).
// Every case is also checked numerically against a host product of the bf16-rounded operands.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../navbot_ppo_b200/csrc/tc_common.cuh"

struct Args {
  const float* A;   // [M, K]
  const float* B;   // [N, K]
  float* D;         // [M, N]
  long long* cyc;   // [2]: cycles, stress stores issued
  int M, N, K, a_src, b_mode, reps, stress, d_lane, issuers, stress_kind;
};

extern __shared__ __align__(128) unsigned char smem[];

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint16_t bf16_bits(float x) {
  uint16_t h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}

__global__ void __launch_bounds__(384) bench_kernel(Args p) {
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sA = smem;                       // up to 128 x 128 bf16 = 32 KB
  unsigned char* sB = smem + 32768;               // up to 256 x 128 bf16 = 64 KB
  unsigned char* sS = smem + 32768 + 65536;       // stress region 64 KB
  if (tid == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1); stop_flag = 0; }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t TM_A = 256;                      // A operand columns when a_src == 2
  if (tid < 128) {
    if (p.a_src < 2) {
      for (int i = tid; i < p.M * p.K; i += 128) {
        const int m = i / p.K, k = i % p.K;
        const uint32_t off = p.a_src == 0 ? tc::rb16_off(p.M, m, k) : tc::rb16_off(p.K, k, m);
        *reinterpret_cast<uint16_t*>(sA + off) = bf16_bits(p.A[i]);
      }
    } else {
      // thread = row m = TMEM lane; 8 columns per 16 k
      for (int kk = 0; kk < p.K / 16; ++kk) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) {
          const float a0 = tid < p.M ? p.A[tid * p.K + kk * 16 + 2 * j] : 0.f;
          const float a1 = tid < p.M ? p.A[tid * p.K + kk * 16 + 2 * j + 1] : 0.f;
          r[j] = (uint32_t)bf16_bits(a0) | ((uint32_t)bf16_bits(a1) << 16);
        }
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + TM_A + kk * 8, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < p.N * p.K; i += 128) {
      const int n = i / p.K, k = i % p.K;
      const uint32_t off = p.b_mode == 0 ? tc::rb16_off(p.N, n, k) : tc::rb16_off(p.K, k, n);
      *reinterpret_cast<uint16_t*>(sB + off) = bf16_bits(p.B[i]);
    }
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp < p.issuers) {
    uint64_t& bar = bars[warp];
    const uint32_t idesc = tc::make_idesc_bf16(p.M, p.N, p.a_src == 1 ? 1 : 0, p.b_mode);
    const uint32_t a_lbo = p.a_src == 0 ? p.M * 16 : 128, a_sbo = p.a_src == 0 ? 128 : p.K * 16;
    const uint32_t a_step = p.a_src == 0 ? 2 * p.M * 16 : 256;
    const uint32_t b_lbo = p.b_mode == 0 ? p.N * 16 : 128, b_sbo = p.b_mode == 0 ? 128 : p.K * 16;
    const uint32_t b_step = p.b_mode == 0 ? 2 * p.N * 16 : 256;
    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
    const uint32_t d = tmem + ((uint32_t)p.d_lane << 16) + (uint32_t)warp * 64;
    const int ks = p.K / 16;
    // descriptors are built once (8 k-steps at most) so that the timed loop is nothing but the MMA stream
    uint64_t ad[8], bd[8];
    uint32_t at[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k2 = kk < ks ? kk : ks - 1;
      ad[kk] = tc::make_desc(a0 + k2 * a_step, a_lbo, a_sbo);
      bd[kk] = tc::make_desc(b0 + k2 * b_step, b_lbo, b_sbo);
      at[kk] = tmem + TM_A + k2 * 8;
    }
    // the whole warp runs the loop converged; one elected lane issues (CUTLASS-style), so the MMAs are
    // back to back in the instruction stream
    __syncwarp();
    const long long t0 = clock64();
    if (p.stress_kind >= 10) {
      // kernel-like burst pattern (numerics not checked): per rep 24 TS MMAs (N = 48 / 32) + commit, 24 SS MMAs
      // (N = 32) + 2 commits, 12 SS MMAs (N = 64) + commit; 10 = as described, 11 = no commits, 12 = wait after
      // every commit, 13 = commits by re-electing for each one like the library's commit()
      const uint32_t id48 = tc::make_idesc_bf16(128, 48, 0, 1), id32 = tc::make_idesc_bf16(128, 32, 0, 0),
                     id64 = tc::make_idesc_bf16(128, 64, 0, 0);
      const uint64_t a_k = tc::make_desc(a0, 128 * 16, 128), b_k = tc::make_desc(b0, 64 * 16, 128), b_m = tc::make_desc(b0, 128, 128 * 16);
      uint32_t ph1 = 0;
      for (int rep = 0; rep < p.reps; ++rep) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 12; ++i) mma_ts(tmem + 128, tmem + TM_A + (i & 3) * 8, b_m + (i & 3) * 16, id48, i > 0);
#pragma unroll
          for (int i = 0; i < 12; ++i) mma_ts(tmem + 192, tmem + TM_A + (i & 3) * 8, b_m + (i & 3) * 16, id32, i > 0);
          if (p.stress_kind == 10 || p.stress_kind == 12) tc::mma_commit(&bars[1]);
        }
        __syncwarp();
        if (p.stress_kind == 13) { if (elect_one()) tc::mma_commit(&bars[1]); __syncwarp(); }
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 24; ++i) tc::mma_bf16(tmem + 224, a_k + (i & 7) * 256, b_k + (i & 7) * 128, id32, i > 0);
          if (p.stress_kind == 10 || p.stress_kind == 12) { tc::mma_commit(&bars[2]); tc::mma_commit(&bars[3]); }
        }
        __syncwarp();
        if (p.stress_kind == 13) { if (elect_one()) tc::mma_commit(&bars[2]); __syncwarp(); if (elect_one()) tc::mma_commit(&bars[3]); __syncwarp(); }
        if (p.stress_kind == 12) { tc::mbar_wait(&bars[2], ph1); ph1 ^= 1; }
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 12; ++i) tc::mma_bf16(tmem, a_k + (i & 1) * 256, b_k + (i & 1) * 128, id64, (i & 1));
          if (p.stress_kind == 10 || p.stress_kind == 12) tc::mma_commit(&bars[1]);
        }
        __syncwarp();
        if (p.stress_kind == 13) { if (elect_one()) tc::mma_commit(&bars[1]); __syncwarp(); }
      }
    } else if (elect_one()) {
      if (ks == 8) {
        if (p.a_src == 2) {
          for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma_ts(d, at[kk], bd[kk], idesc, kk > 0);
          }
        } else {
          for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_bf16(d, ad[kk], bd[kk], idesc, kk > 0);
          }
        }
      } else if (ks == 2) {
        for (int rep = 0; rep < p.reps; rep += 4) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) tc::mma_bf16(d, ad[kk & 1], bd[kk & 1], idesc, kk & 1);
        }
      } else {
        for (int rep = 0; rep < p.reps; rep += 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) tc::mma_bf16(d, ad[0], bd[0], idesc, 0);
        }
      }
    }
    __syncwarp();
    const long long t1 = clock64();
    if (elect_one()) tc::mma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      if (warp == 0) { p.cyc[0] = t2 - t0; p.cyc[2] = t1 - t0; }
      else p.cyc[3] = t2 - t0;
    }
    if (lane == 0)
    if (atomicAdd((int*)&stop_flag, 1) + 1 == p.issuers) stop_flag = 1000;
  } else if (warp >= 4 && warp < 4 + p.stress) {
    // keep one SM resource busy until the MMAs are done: 0 = 16-byte shared-memory stores (one 512-byte row per
    // warp instruction), 1 = FP32 multiplies + bf16x2 conversions, 2 = tcgen05.ld of 16 columns, 3 = 16-byte
    // shared-memory loads
    long long n = 0;
    uint4 v = make_uint4(tid, tid, tid, tid);
    unsigned char* base = sS + (warp - 4) * 8192 + lane * 16;
    float f0 = 1.0f + tid * 1e-3f, f1 = 0.5f, f2 = 0.25f, f3 = 2.f;
    uint32_t sink = 0;
    while (stop_flag < 1000) {
      if (p.stress_kind == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tc::smem_u32(base + i * 512)), "r"(v.x), "r"(v.y), "r"(v.z),
                       "r"(v.w)
                       : "memory");
      } else if (p.stress_kind == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f0 = f0 * 1.0001f; f1 = f1 * 0.9999f; f2 = f2 * 1.0002f; f3 = f3 * 0.9998f;
          uint32_t a_, b_;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a_) : "f"(f0), "f"(f1));
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b_) : "f"(f2), "f"(f3));
          sink ^= a_ + b_;
        }
      } else if (p.stress_kind == 2) {
        float t[16];
        tc::tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (warp >> 2) * 16, t);
        sink ^= __float_as_uint(t[0]);
      } else if (p.stress_kind == 4) {   // tcgen05.st of 16 columns
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = tid + i;
        tc::tmem_st16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (warp >> 2) * 16, r);
        tc::tmem_st_wait();
      } else if (p.stress_kind == 5) {   // an epilogue-like mix: tcgen05.ld x2, 32 multiplies, 16 conversions, tcgen05.st x2
        float t[32];
        const uint32_t ta_ = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (warp >> 2) * 32;
        tc::tmem_ld16_nowait(ta_, t);
        tc::tmem_ld16_nowait(ta_ + 16, t + 16);
        tc::tmem_ld_wait();
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a_ = t[2 * i] * f0, b_ = t[2 * i + 1] * f1;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(a_), "f"(b_));
        }
        tc::tmem_st16(ta_, r);
        tc::tmem_st16(ta_ + 16, r);
        tc::tmem_st_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint4 r;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(tc::smem_u32(base + i * 512)));
          sink ^= r.x;
        }
      }
      n += 16;
    }
    if (sink == 0x12345678u) p.D[0] = f0 + f1 + f2 + f3;
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.cyc + 1), (unsigned long long)n);
  }
  __syncthreads();
  tc::fence_after_sync();
  if (tid < 128) {
    // TMEM lane of this thread; M = 64 uses lanes 16 * (r / 16) * 2 + r % 16 (+ d_lane)
    int r = -1;
    if (p.M == 128) r = tid;
    else {
      const int l = lane - p.d_lane;
      if (l >= 0 && l < 16) r = warp * 16 + l;
    }
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (r >= 0)
        for (int i = 0; i < 16 && c0 + i < p.N; ++i) p.D[(size_t)r * p.N + c0 + i] = v[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  u = r & 0xFFFF0000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 64;
  const int only = argc > 2 ? atoi(argv[2]) : -1;   // run one case (a wrong descriptor may fault): loop over indices from a shell
  float *dA, *dB, *dD;
  long long* dC;
  cudaMalloc(&dA, 128 * 128 * 4);
  cudaMalloc(&dB, 256 * 128 * 4);
  cudaMalloc(&dD, 128 * 256 * 4);
  cudaMalloc(&dC, 32);
  const size_t sm = 32768 + 65536 + 65536;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  struct Case { int M, N, K, a_src, b_mode, stress, d_lane, issuers = 1, stress_kind = 0; };
  std::vector<Case> cases;
  const int Ns128[] = {16, 32, 48, 64, 96, 128, 256};
  for (int a_src = 0; a_src < 3; ++a_src)
    for (int N : Ns128) cases.push_back({128, N, 128, a_src, a_src == 1 ? 1 : 0, 0, 0});
  // b MN-major with a K-major (the GX / GH products)
  for (int N : {32, 128}) cases.push_back({128, N, 128, 0, 1, 0, 0});
  for (int N : {32, 128}) cases.push_back({128, N, 128, 2, 1, 0, 0});
  const int Ns64[] = {16, 32, 48, 64, 96, 128};
  for (int a_src = 0; a_src < 2; ++a_src)
    for (int N : Ns64) cases.push_back({64, N, 128, a_src, a_src == 1 ? 1 : 0, 0, 0});
  cases.push_back({64, 48, 128, 1, 1, 0, 16});
  cases.push_back({64, 32, 128, 1, 1, 0, 16});
  // short-K products (Z, GH): K = 16 / 32
  for (int K : {16, 32}) for (int N : {64, 128, 256}) cases.push_back({128, N, K, 0, 0, 0, 0});
  // with shared-memory store traffic from 4 / 8 other warps
  for (int st : {4, 8}) {
    cases.push_back({128, 48, 128, 1, 1, st, 0});
    cases.push_back({128, 128, 128, 0, 0, st, 0});
    cases.push_back({128, 32, 128, 2, 0, st, 0});
    cases.push_back({64, 48, 128, 1, 1, st, 0});
  }
  if (only <= 0) printf("%4s %4s %4s %6s %6s %6s %6s | %10s %10s %12s %10s\n", "M", "N", "K", "a_src", "b_mn", "stress", "dlane", "cyc/mma",
         "issue/mma", "stores/cyc", "max_err");
  for (int is : {2, 4})
    for (int a_src : {0, 2})
      for (int N : {16, 32, 48, 64}) cases.push_back({128, N, 128, a_src, 0, 0, 0, is});
  for (int is : {2, 4}) cases.push_back({128, 48, 128, 1, 1, 0, 0, is});
  // what slows the MMA stream down: ALU / conversion work, TMEM loads or shared-memory loads in 8 other warps?
  for (int kind : {1, 2, 3, 4, 5})
    for (int st : {8}) {
      cases.push_back({128, 64, 32, 0, 0, st, 0, 1, kind});     // Z-like: SS, N = 64
      cases.push_back({128, 32, 128, 2, 0, st, 0, 1, kind});    // U / GX-like: TS, N = 32
      cases.push_back({128, 48, 128, 1, 1, st, 0, 1, kind});    // dW-like
    }
  for (int pat : {10, 11, 12, 13}) cases.push_back({128, 64, 128, 2, 0, 0, 0, 1, pat});
  if (only == -2) { printf("%d\n", (int)cases.size()); return 0; }
  int idx = -1;
  for (const Case& c : cases) {
    if (++idx != only && only >= 0) continue;
    std::vector<float> A(c.M * c.K), B(c.N * c.K), D(c.M * c.N);
    srand(c.M * 131 + c.N * 7 + c.K + c.a_src);
    for (auto& x : A) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
    for (auto& x : B) x = bf16_round((float)rand() / RAND_MAX - 0.5f);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 256 * 4);
    cudaMemset(dC, 0, 32);
    Args p{dA, dB, dD, dC, c.M, c.N, c.K, c.a_src, c.b_mode, reps, c.stress, c.d_lane, c.issuers, c.stress_kind};
    bench_kernel<<<1, 384, sm>>>(p);   // warm (instruction cache)
    cudaMemset(dC, 0, 32);
    bench_kernel<<<1, 384, sm>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case M=%d N=%d a_src=%d: %s\n", c.M, c.N, c.a_src, cudaGetErrorString(e)); return 1; }
    long long cyc[4];
    cudaMemcpy(cyc, dC, 32, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double s = 0;
        for (int k = 0; k < c.K; ++k) s += (double)A[m * c.K + k] * B[n * c.K + k];
        err = fmax(err, fabs(s - D[m * c.N + n]));
      }
    const double nm = (double)reps * (c.K / 16);
    printf("%4d %4d %4d %6d %6d %6d %6d | %10.1f %10.1f %12.3f %10.2e %s issuers=%d other=%.1f kind=%d\n", c.M, c.N, c.K, c.a_src, c.b_mode, c.stress,
           c.d_lane, cyc[0] / nm, cyc[2] / nm, (double)cyc[1] * 1.0 / cyc[0], err, err < 1e-3 ? "ok" : "WRONG", c.issuers, cyc[3] / nm, c.stress_kind);
    fflush(stdout);
  }
  return 0;
}
