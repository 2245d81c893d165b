// navppo_kernels.cu — PPO trainer kernels for sm_100a + their C-ABI (include/navppo.h).
//
// What the reference does with ~30 PyTorch ops per epoch (project_ppo/src/ppo.py:305-393)
// plus a Python double loop (compute_rtgs, :643-671) becomes:
//   rtg_scan_kernel      reverse (gamma, lambda) scan per agent over the [H, N] rollout
//   mlp_infer_kernel     NetActor / NetCritic forward (+ Gaussian sampling / log-prob epilogue)
//   adv_*_kernel         advantage statistics and normalisation                (ppo.py:277,284)
//   mlp_grad_kernel      forward + losses + backward of one network per CTA row (ppo.py:307-386)
//   grad_reduce_kernel   fixed-order sum of the per-CTA gradient partials -> flat gradient
//   adam_kernel          both torch.optim.Adam steps                            (ppo.py:381,392)
//
// This file holds the fp32 CUDA-core arithmetic (NAVPPO_FP32): every product is an FFMA in
// fp32 like the reference's torch.float32 CPU/GPU path, so results agree with it to
// reduction-order rounding.  The tcgen05 tensor-core GEMM path lives in navppo_tc.cu.
//
// Network (net_actor.py:39-53,138-143; net_critic.py:36-48,127-129), per sample:
//   z1 = W1a x0 + b1a (512)   h1 = lrelu(z1)   u1 = x0 + W1b h1 + b1b (16)   y1 = lrelu(u1)
//   x1 = [x0 | y1] (32)
//   z2 = W2a x1 + b2a (512)   h2 = lrelu(z2)   u2 = x1 + W2b h2 + b2b (32)   y2 = lrelu(u2)
//   actor: mu = [sigmoid(w_o1 . y2 + c1), tanh(w_o2 . y2 + c2)]     critic: V = w_o . y2 + c
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cstdlib>
#include <new>
#include <string>

#include "../../include/navppo.h"
#include "nav_common.h"
#include "navsim_math.h"
#include "ppo_common.cuh"

// tensor-core gradient pass (navppo_tc.cu): fills gpart / mpart like mlp_grad_kernel
int navppo_tc_grad_launch(const ppo::GradArgs& a, int rows, int passes, float* wprep, cudaStream_t s);
size_t navppo_tc_prep_bytes();
int navppo_tc_init();
// warp-specialised version of the same kernel (navppo_tcws.cu), the default
int navppo_tcws_grad_launch(const ppo::GradArgs& a, int rows, int passes, float* wprep, cudaStream_t s);
size_t navppo_tcws_prep_bytes();
int navppo_tcws_init();
int navppo_tcws_prep_launch(const float* params, float* wprep, cudaStream_t s);
int navppo_tcws_infer_launch(const ppo::InferArgs& a, int mode, bool both_nets, int passes, const float* wprep, bool chained,
                             bool weights_ready, cudaStream_t s);
// the whole step loop of a rollout as one launch (navppo_tcws.cu); *unsupported = 1: fall back to chained kernels
int navppo_tcws_rollout_launch(navsim_t* sim, const ppo::InferArgs& a, int H, int passes, const float* wprep, float* obs_rows,
                               float* next_obs, float* rew, uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_ret,
                               float* ep_path, int32_t* ep_len, int* unsupported, cudaStream_t s);
// navsim_step_ex as a programmatic dependent launch (navsim_kernels.cu)
int navsim_step_chained(navsim_t* h, const float* act_dev, const navsim_step_out* out, void* stream);

namespace {

using namespace ppo;

extern __shared__ __align__(16) unsigned char ppo_smem[];

// ----------------------------------------------------------------------------------------
// compute_rtgs / GAE: one thread per agent walks its column backwards.
// ----------------------------------------------------------------------------------------
__global__ void rtg_scan_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ term,
                                const float* __restrict__ values, const float* __restrict__ last_value, double gamma,
                                double lam, float* __restrict__ out, int H, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // reward-to-go form with last_value: bootstrap the trailing partial episode from V(s_H)
  double acc = (!values && last_value) ? (double)last_value[n] : 0.0;
  double next_v = (values && last_value) ? (double)last_value[n] : 0.0;
  for (int t = H - 1; t >= 0; --t) {
    const size_t i = (size_t)t * N + n;
    const bool boundary = term[i] != 0;
    if (values) {
      const double live = boundary ? 0.0 : 1.0;
      const double v = (double)values[i];
      const double delta = (double)rew[i] + gamma * next_v * live - v;
      acc = delta + gamma * lam * live * acc;
      next_v = v;
    } else {
      if (boundary) acc = 0.0;                    // ppo.py:659 discounted_reward = 0
      acc = (double)rew[i] + acc * gamma;         // ppo.py:664
    }
    out[i] = (float)acc;                          // ppo.py:669
  }
}

// One residual block for one sample: u = x + Wb lrelu(Wa x + ba) + bb, two hidden units per
// trip for instruction-level parallelism.  Wa rows and WbT rows are IN contiguous floats.
template <int IN>
__device__ __forceinline__ void resblock_fwd(const float* __restrict__ sWa, const float* __restrict__ sBa,
                                             const float* __restrict__ sWbT, const float* __restrict__ sBb,
                                             const float (&x)[IN], float (&u)[IN], int j0, int j1) {
#pragma unroll 1
  for (int j = j0; j < j1; j += 2) {
    float z0 = sBa[j], z1 = sBa[j + 1];
    const float4* wa0 = reinterpret_cast<const float4*>(sWa + (size_t)j * IN);
    const float4* wa1 = reinterpret_cast<const float4*>(sWa + (size_t)(j + 1) * IN);
#pragma unroll
    for (int q = 0; q < IN / 4; ++q) {
      const float4 a = wa0[q], b = wa1[q];
      z0 = fmaf(a.x, x[4 * q], z0); z1 = fmaf(b.x, x[4 * q], z1);
      z0 = fmaf(a.y, x[4 * q + 1], z0); z1 = fmaf(b.y, x[4 * q + 1], z1);
      z0 = fmaf(a.z, x[4 * q + 2], z0); z1 = fmaf(b.z, x[4 * q + 2], z1);
      z0 = fmaf(a.w, x[4 * q + 3], z0); z1 = fmaf(b.w, x[4 * q + 3], z1);
    }
    const float h0 = lrelu(z0), h1 = lrelu(z1);
    const float4* wb0 = reinterpret_cast<const float4*>(sWbT + (size_t)j * IN);
    const float4* wb1 = reinterpret_cast<const float4*>(sWbT + (size_t)(j + 1) * IN);
#pragma unroll
    for (int q = 0; q < IN / 4; ++q) {
      const float4 a = wb0[q], b = wb1[q];
      u[4 * q] = fmaf(a.x, h0, u[4 * q]); u[4 * q + 1] = fmaf(a.y, h0, u[4 * q + 1]);
      u[4 * q + 2] = fmaf(a.z, h0, u[4 * q + 2]); u[4 * q + 3] = fmaf(a.w, h0, u[4 * q + 3]);
      u[4 * q] = fmaf(b.x, h1, u[4 * q]); u[4 * q + 1] = fmaf(b.y, h1, u[4 * q + 1]);
      u[4 * q + 2] = fmaf(b.z, h1, u[4 * q + 2]); u[4 * q + 3] = fmaf(b.w, h1, u[4 * q + 3]);
    }
  }
}

// ----------------------------------------------------------------------------------------
// Inference (NetActor / NetCritic forward + the get_action / evaluate epilogues).
//
// One CTA = 64 samples x 8 warps.  Lane l of every warp carries samples l and l + 32 of the
// tile; warp w owns hidden units 32 c + 4 w + {0..3} of each 32-unit chunk c of a residual
// block.  The weights of a chunk - 32 fc1 rows, their biases and the 32 matching fc2 columns
// (taken straight from the canonical [out][hidden] layout) - are streamed into shared memory
// with cp.async, double buffered, and every weight read is a warp-wide broadcast (one
// wavefront) used for 8 FMAs.  Per-warp partial residual sums meet in shared memory; the
// activated block output goes back to all warps through shared memory as well.
// At rollout sizes (8192 samples = 128 CTAs) this is bound by the FP32 pipe, not by weight
// latency.
// ----------------------------------------------------------------------------------------
constexpr int CF_THREADS = 256;                       // 8 warps
constexpr int CF_WARPS = CF_THREADS / 32;
constexpr int CF_SPT = 2;                             // samples per thread
constexpr int CF_TILE = 32 * CF_SPT;                  // samples per CTA
constexpr int CF_CHUNK = 4 * CF_WARPS;                // hidden units per staged chunk (32)
constexpr int CF_NCHUNK = HID / CF_CHUNK;             // 16
// shared memory (floats): [2 chunk buffers][partials: warps x X1 x tile][block output: X1 x tile]
constexpr int CF_CHUNK_FLOATS = CF_CHUNK * X1 + CF_CHUNK + X1 * CF_CHUNK;   // sized for the 32-input block
constexpr int CF_PART = 2 * CF_CHUNK_FLOATS;
constexpr int CF_YBUF = CF_PART + CF_WARPS * X1 * CF_TILE;
constexpr int CF_SMEM_FLOATS = CF_YBUF + X1 * CF_TILE;
constexpr size_t CF_SMEM = (size_t)CF_SMEM_FLOATS * sizeof(float);

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// chunk image: [CF_CHUNK rows of IN][CF_CHUNK biases][IN rows of CF_CHUNK]
template <int IN>
__device__ __forceinline__ void stage_chunk(float* dst, const float* __restrict__ Wa, const float* __restrict__ ba,
                                            const float* __restrict__ Wb, int c) {
  const int tid = threadIdx.x;
  for (int v = tid; v < CF_CHUNK * IN / 4; v += CF_THREADS) cp_async16(dst + 4 * v, Wa + (size_t)(CF_CHUNK * c) * IN + 4 * v);
  if (tid < CF_CHUNK / 4) cp_async16(dst + CF_CHUNK * IN + 4 * tid, ba + CF_CHUNK * c + 4 * tid);
  float* dB = dst + CF_CHUNK * IN + CF_CHUNK;
  for (int v = tid; v < IN * (CF_CHUNK / 4); v += CF_THREADS) {
    const int k = v / (CF_CHUNK / 4), q = v % (CF_CHUNK / 4);
    cp_async16(dB + k * CF_CHUNK + 4 * q, Wb + (size_t)k * HID + CF_CHUNK * c + 4 * q);
  }
}

// u[s][k] <- this warp's share of  Wb lrelu(Wa x + ba)  for its lanes' samples
template <int IN>
__device__ __forceinline__ void coop_resblock(float* smem, const float* __restrict__ Wa, const float* __restrict__ ba,
                                              const float* __restrict__ Wb, int warp, const float (&x)[CF_SPT][IN],
                                              float (&u)[CF_SPT][IN]) {
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
    for (int k = 0; k < IN; ++k) u[s][k] = 0.f;
  stage_chunk<IN>(smem, Wa, ba, Wb, 0);
  cp_async_commit();
#pragma unroll 1
  for (int c = 0; c < CF_NCHUNK; ++c) {
    if (c + 1 < CF_NCHUNK) {
      stage_chunk<IN>(smem + ((c + 1) & 1) * CF_CHUNK_FLOATS, Wa, ba, Wb, c + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* buf = smem + (c & 1) * CF_CHUNK_FLOATS;
    const float* A = buf + (4 * warp) * IN;
    const float* Bm = buf + CF_CHUNK * IN + CF_CHUNK + 4 * warp;
    float h[CF_SPT][4];
    {
      const float4 b4 = *reinterpret_cast<const float4*>(buf + CF_CHUNK * IN + 4 * warp);
#pragma unroll
      for (int s = 0; s < CF_SPT; ++s) { h[s][0] = b4.x; h[s][1] = b4.y; h[s][2] = b4.z; h[s][3] = b4.w; }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int q = 0; q < IN / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(A + r * IN + 4 * q);
#pragma unroll
        for (int s = 0; s < CF_SPT; ++s) {
          h[s][r] = fmaf(w.x, x[s][4 * q], h[s][r]);
          h[s][r] = fmaf(w.y, x[s][4 * q + 1], h[s][r]);
          h[s][r] = fmaf(w.z, x[s][4 * q + 2], h[s][r]);
          h[s][r] = fmaf(w.w, x[s][4 * q + 3], h[s][r]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
      for (int r = 0; r < 4; ++r) h[s][r] = lrelu(h[s][r]);
#pragma unroll
    for (int k = 0; k < IN; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(Bm + k * CF_CHUNK);
#pragma unroll
      for (int s = 0; s < CF_SPT; ++s)
        u[s][k] = fmaf(w.x, h[s][0], fmaf(w.y, h[s][1], fmaf(w.z, h[s][2], fmaf(w.w, h[s][3], u[s][k]))));
    }
    __syncthreads();   // the buffer is refilled two iterations later
  }
}

// Sum the 8 warps' partials, add skip + bias, activate: every thread finishes IN / 8 outputs
// of its two samples and publishes them in ybuf[k][sample]; returns after a CTA barrier.
template <int IN>
__device__ __forceinline__ void coop_block_output(float* smem, const float* __restrict__ bb, int warp, int lane,
                                                  const float (&x)[CF_SPT][IN], const float (&u)[CF_SPT][IN],
                                                  float (&ymine)[CF_SPT][IN / CF_WARPS]) {
  float* part = smem + CF_PART;
  float* ybuf = smem + CF_YBUF;
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
    for (int k = 0; k < IN; ++k) part[(warp * IN + k) * CF_TILE + s * 32 + lane] = u[s][k];
  __syncthreads();
  constexpr int PER = IN / CF_WARPS;
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
    for (int kk = 0; kk < PER; ++kk) {
      const int k = warp * PER + kk;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < CF_WARPS; ++w) acc += part[(w * IN + k) * CF_TILE + s * 32 + lane];
      // select x[s][k] without dynamic register indexing
      float xs = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < IN; ++k2) xs = (k2 == k) ? x[s][k2] : xs;
      const float y = lrelu(acc + (xs + __ldg(bb + k)));
      ymine[s][kk] = y;
      ybuf[k * CF_TILE + s * 32 + lane] = y;
    }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(CF_THREADS) mlp_infer_kernel(InferArgs a) {
  const int net = blockIdx.y;  // 0 actor, 1 critic
  if (MODE == INFER_FORWARD && ((net == 0 && !a.mu) || (net == 1 && !a.v))) return;
  float* smem = reinterpret_cast<float*>(ppo_smem);
  const float* __restrict__ p = a.params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = blockIdx.x * CF_TILE;
  float x1[CF_SPT][X1];
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s) {
    const int i0 = base + s * 32 + lane;
    const int ii = i0 < a.T ? i0 : a.T - 1;              // surplus lanes shadow the last sample
    const float4* o = reinterpret_cast<const float4*>(a.obs + (size_t)ii * OBS);
#pragma unroll
    for (int q = 0; q < OBS / 4; ++q) {
      const float4 v = __ldg(o + q);
      x1[s][4 * q] = v.x; x1[s][4 * q + 1] = v.y; x1[s][4 * q + 2] = v.z; x1[s][4 * q + 3] = v.w;
    }
  }
  {
    float x0[CF_SPT][OBS], u1[CF_SPT][OBS], y1[CF_SPT][OBS / CF_WARPS];
#pragma unroll
    for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
      for (int k = 0; k < OBS; ++k) x0[s][k] = x1[s][k];
    coop_resblock<OBS>(smem, p + O_W1A, p + O_B1A, p + O_W1B, warp, x0, u1);
    coop_block_output<OBS>(smem, p + O_B1B, warp, lane, x0, u1, y1);
    const float* ybuf = smem + CF_YBUF;
#pragma unroll
    for (int s = 0; s < CF_SPT; ++s)
#pragma unroll
      for (int k = 0; k < OBS; ++k) x1[s][OBS + k] = ybuf[k * CF_TILE + s * 32 + lane];
  }
  float y2[CF_SPT][X1 / CF_WARPS];
  {
    float u2[CF_SPT][X1];
    coop_resblock<X1>(smem, p + O_W2A, p + O_B2A, p + O_W2B, warp, x1, u2);
    coop_block_output<X1>(smem, p + O_B2B, warp, lane, x1, u2, y2);
  }
  // heads: each warp contributes the 4 outputs it finished; warp 0 adds the 8 partial sums
  float* part = smem + CF_PART;
  constexpr int PER = X1 / CF_WARPS;
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s) {
    float o1 = 0.f, o2 = 0.f;
#pragma unroll
    for (int kk = 0; kk < PER; ++kk) {
      const int k = warp * PER + kk;
      o1 = fmaf(__ldg(p + O_HEAD + k), y2[s][kk], o1);
      if (net == 0) o2 = fmaf(__ldg(p + O_HEAD + X1 + 1 + k), y2[s][kk], o2);
    }
    part[(warp * 2 + 0) * CF_TILE + s * 32 + lane] = o1;
    part[(warp * 2 + 1) * CF_TILE + s * 32 + lane] = o2;
  }
  __syncthreads();
  if (warp != 0) return;
#pragma unroll
  for (int s = 0; s < CF_SPT; ++s) {
    const int i = base + s * 32 + lane;
    if (i >= a.T) continue;
    float o1 = __ldg(p + O_HEAD + X1), o2 = (net == 0) ? __ldg(p + O_HEAD + 2 * X1 + 1) : 0.f;
#pragma unroll
    for (int w = 0; w < CF_WARPS; ++w) {
      o1 += part[(w * 2 + 0) * CF_TILE + s * 32 + lane];
      o2 += part[(w * 2 + 1) * CF_TILE + s * 32 + lane];
    }
    if (net == 1) {                      // critic: V = out(X), net_critic.py:129
      a.v[i] = o1;
      continue;
    }
    const float m0 = sigmoidf_(o1), m1 = tanhf(o2);  // net_actor.py:141-142
    const float var_ = (MODE == INFER_ACT && a.dyn) ? __uint_as_float(a.dyn[0]) : a.var;
    const uint32_t draw_ = (MODE == INFER_ACT && a.dyn) ? a.draw + a.dyn[1] : a.draw;
    if (MODE == INFER_FORWARD) {
      reinterpret_cast<float2*>(a.mu)[i] = make_float2(m0, m1);
    } else if (MODE == INFER_ACT) {
      float e0, e1;
      if (a.noise_in) {
        const float2 e = reinterpret_cast<const float2*>(a.noise_in)[i];
        e0 = e.x; e1 = e.y;
      } else {  // Box-Muller on two Philox words
        uint32_t r[4];
        const uint64_t agent = (uint64_t)(a.agent_off + i);
        nv_philox4x32_10(draw_, 2u, (uint32_t)agent, (uint32_t)(agent >> 32), (uint32_t)a.seed,
                         (uint32_t)(a.seed >> 32), r);
        const float uu = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
        const float vv = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(uu));
        float sn, cs;
        sincosf(6.283185307179586f * vv, &sn, &cs);
        e0 = rad * cs; e1 = rad * sn;
      }
      const float sd = sqrtf(var_);
      float a0 = fmaf(sd, e0, m0), a1 = fmaf(sd, e1, m1);       // dist.sample(), ppo.py:698-699
      a0 = fminf(fmaxf(a0, 0.f), 1.f);                          // ppo.py:701
      a1 = fminf(fmaxf(a1, -1.f), 1.f);                         // ppo.py:702
      reinterpret_cast<float2*>(a.act)[i] = make_float2(a0, a1);
      a.logp[i] = gauss_logp(a0, a1, m0, m1, var_);             // ppo.py:704 (at the clamped action)
      if (a.mu) reinterpret_cast<float2*>(a.mu)[i] = make_float2(m0, m1);
    } else {
      const float2 av = reinterpret_cast<const float2*>(a.act_in)[i];
      a.logp[i] = gauss_logp(av.x, av.y, m0, m1, a.var);        // ppo.py:734-735
    }
  }
}

// ----------------------------------------------------------------------------------------
// Advantage (ppo.py:277,284)
// ----------------------------------------------------------------------------------------
__global__ void adv_stats_kernel(const float* __restrict__ rtg, const float* __restrict__ v, int T, double* stats) {
  double s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const double a = (double)rtg[i] - (double)v[i];
    s += a; q += a * a;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
  }
  __shared__ double ws[32], wq[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { ws[w] = s; wq[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ts += ws[k]; tq += wq[k]; }
    atomicAdd(&stats[0], ts);
    atomicAdd(&stats[1], tq);
    if (blockIdx.x == 0) atomicAdd(&stats[2], (double)T);
  }
}

__global__ void adv_normalize_kernel(const float* __restrict__ rtg, const float* __restrict__ v, int T,
                                     const double* __restrict__ stats, float* __restrict__ adv) {
  const double n = stats[2];
  const double mean = stats[0] / n;
  double var = (stats[1] - n * mean * mean) / (n - 1.0);   // torch.std: unbiased
  if (!(var > 0.0)) var = 0.0;
  const float m = (float)mean, inv = (float)(1.0 / (sqrt(var) + 1e-10));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x)
    adv[i] = ((rtg[i] - v[i]) - m) * inv;
}

// ----------------------------------------------------------------------------------------
// Gradient kernel: blockIdx.y = network, each CTA walks tiles of GM samples (thread = sample
// for the per-sample chains, thread = weight patch for the weight-gradient products) and
// accumulates into its own row of the partial-gradient workspace (kernel layout).
// ----------------------------------------------------------------------------------------
constexpr int GM = 128;          // samples per tile == threads per CTA
constexpr int JC = 32;           // hidden units per chunk
constexpr int SROW = GM + 1;     // padded row of the [JC][GM] chunk buffers (bank spread)

// shared-memory carve-up (floats)
constexpr int G_SX = 0;                       // [GM][32]  x1 = [x0 | y1]
constexpr int G_SGU = G_SX + GM * X1;         // [GM][32]  g_u2, later g_u1 in [..][0:16]
constexpr int G_SH = G_SGU + GM * X1;         // [JC][SROW] h of the chunk
constexpr int G_SG = G_SH + JC * SROW;        // [JC][SROW] g_z of the chunk
constexpr int G_SWA = G_SG + JC * SROW;       // [JC][32]  fc1 rows of the chunk
constexpr int G_SWBT = G_SWA + JC * X1;       // [JC][32]  fc2 columns of the chunk
constexpr int G_SBA = G_SWBT + JC * X1;       // [JC]
constexpr int G_SRED = G_SBA + JC;            // [4 warps][72] head-gradient partials
constexpr int G_TOTAL = G_SRED + 4 * 72;
constexpr size_t GRAD_SMEM = (size_t)G_TOTAL * sizeof(float);

template <int IN>
__device__ __forceinline__ void stage_chunk(float* s, const float* __restrict__ p, int o_wa, int o_ba, int o_wb, int c) {
  // fc1 rows c*JC .. c*JC+JC-1 (contiguous in the canonical layout) and the matching fc2 columns
  for (int i = threadIdx.x; i < JC * IN; i += GM) {
    s[G_SWA + i] = p[o_wa + c * JC * IN + i];
    const int jl = i % JC, k = i / JC;        // consecutive threads -> consecutive hidden units
    s[G_SWBT + jl * IN + k] = p[o_wb + k * HID + c * JC + jl];
  }
  if (threadIdx.x < JC) s[G_SBA + threadIdx.x] = p[o_ba + c * JC + threadIdx.x];
}

template <int IN>
__device__ __forceinline__ void block_forward(float* s, const float* __restrict__ p, int o_wa, int o_ba, int o_wb,
                                              const float (&x)[IN], float (&u)[IN]) {
  for (int c = 0; c < HID / JC; ++c) {
    stage_chunk<IN>(s, p, o_wa, o_ba, o_wb, c);
    __syncthreads();
    resblock_fwd<IN>(s + G_SWA, s + G_SBA, s + G_SWBT, nullptr, x, u, 0, JC);
    __syncthreads();
  }
}

// Backward of one residual block over all hidden chunks.  gu = dL/du of this block (per
// sample, in registers AND already stored to sGU rows), x = block input (registers AND sX rows).
// gx accumulates W_a^T g_z (only when WANT_GX).  Weight gradients go to this CTA's gpart row.
template <int IN, bool WANT_GX>
__device__ __forceinline__ void block_backward(float* s, const float* __restrict__ p, float* __restrict__ grow, int o_wa,
                                               int o_ba, int o_wb, const float (&x)[IN], const float (&gu)[IN],
                                               float (&gx)[IN]) {
  const int m = threadIdx.x;
  constexpr int KQ = IN / 4;                  // outputs per thread along k (8 for IN=32, 4 for IN=16)
  const int pj = m >> 2, pk = (m & 3) * KQ;   // weight patch of this thread in phase B
  for (int c = 0; c < HID / JC; ++c) {
    stage_chunk<IN>(s, p, o_wa, o_ba, o_wb, c);
    __syncthreads();
    // ---- phase A: per-sample chain through the chunk's hidden units
#pragma unroll 1
    for (int jl = 0; jl < JC; ++jl) {
      float z = s[G_SBA + jl], gh = 0.f;
      const float4* wa = reinterpret_cast<const float4*>(s + G_SWA + jl * IN);
      const float4* wb = reinterpret_cast<const float4*>(s + G_SWBT + jl * IN);
#pragma unroll
      for (int q = 0; q < IN / 4; ++q) {
        const float4 a = wa[q], b = wb[q];
        z = fmaf(a.x, x[4 * q], z); z = fmaf(a.y, x[4 * q + 1], z);
        z = fmaf(a.z, x[4 * q + 2], z); z = fmaf(a.w, x[4 * q + 3], z);
        gh = fmaf(b.x, gu[4 * q], gh); gh = fmaf(b.y, gu[4 * q + 1], gh);
        gh = fmaf(b.z, gu[4 * q + 2], gh); gh = fmaf(b.w, gu[4 * q + 3], gh);
      }
      const float gz = gh * dlrelu(z);
      s[G_SH + jl * SROW + m] = lrelu(z);
      s[G_SG + jl * SROW + m] = gz;
      if (WANT_GX) {
#pragma unroll
        for (int q = 0; q < IN / 4; ++q) {
          const float4 a = wa[q];
          gx[4 * q] = fmaf(a.x, gz, gx[4 * q]); gx[4 * q + 1] = fmaf(a.y, gz, gx[4 * q + 1]);
          gx[4 * q + 2] = fmaf(a.z, gz, gx[4 * q + 2]); gx[4 * q + 3] = fmaf(a.w, gz, gx[4 * q + 3]);
        }
      }
    }
    __syncthreads();
    // ---- phase B: dWa[j][k] = sum_m g_z[j][m] x[m][k];  dWbT[j][k] = sum_m h[j][m] g_u[m][k]
    {
      float da[KQ], db[KQ], dbias = 0.f;
#pragma unroll
      for (int i = 0; i < KQ; ++i) { da[i] = 0.f; db[i] = 0.f; }
      const float* hrow = s + G_SH + pj * SROW;
      const float* grow_s = s + G_SG + pj * SROW;
#pragma unroll 4
      for (int mm = 0; mm < GM; ++mm) {
        const float g = grow_s[mm], h = hrow[mm];
        dbias += g;
        const float4* xv = reinterpret_cast<const float4*>(s + G_SX + mm * X1 + pk);
        const float4* gv = reinterpret_cast<const float4*>(s + G_SGU + mm * X1 + pk);
#pragma unroll
        for (int q = 0; q < KQ / 4; ++q) {
          const float4 xx = xv[q], gg = gv[q];
          da[4 * q] = fmaf(g, xx.x, da[4 * q]); da[4 * q + 1] = fmaf(g, xx.y, da[4 * q + 1]);
          da[4 * q + 2] = fmaf(g, xx.z, da[4 * q + 2]); da[4 * q + 3] = fmaf(g, xx.w, da[4 * q + 3]);
          db[4 * q] = fmaf(h, gg.x, db[4 * q]); db[4 * q + 1] = fmaf(h, gg.y, db[4 * q + 1]);
          db[4 * q + 2] = fmaf(h, gg.z, db[4 * q + 2]); db[4 * q + 3] = fmaf(h, gg.w, db[4 * q + 3]);
        }
      }
      const int j = c * JC + pj;
      float* ga = grow + o_wa + j * IN + pk;
      float* gb = grow + o_wb + j * IN + pk;   // kernel layout: fc2 transposed
#pragma unroll
      for (int i = 0; i < KQ; ++i) { ga[i] += da[i]; gb[i] += db[i]; }
      if ((m & 3) == 0) grow[o_ba + j] += dbias;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(GM) mlp_grad_kernel(GradArgs a) {
  float* s = reinterpret_cast<float*>(ppo_smem);
  const int net = blockIdx.y, m = threadIdx.x, warp = m >> 5, lane = m & 31;
  const float* __restrict__ p = a.params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  float* __restrict__ grow = a.gpart + ((size_t)net * gridDim.x + blockIdx.x) * NET_ROW;
  for (int i = m; i < NET_ROW; i += GM) grow[i] = 0.f;
  double macc[4] = {0.0, 0.0, 0.0, 0.0};
  const int ntiles = (a.T + GM - 1) / GM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int i = tile * GM + m;
    const bool valid = i < a.T;
    float x1[X1], u1[OBS], u2[X1];
    if (valid) {
      const float4* o = reinterpret_cast<const float4*>(a.obs + (size_t)i * OBS);
#pragma unroll
      for (int q = 0; q < OBS / 4; ++q) {
        const float4 t = o[q];
        x1[4 * q] = t.x; x1[4 * q + 1] = t.y; x1[4 * q + 2] = t.z; x1[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < OBS; ++k) x1[k] = 0.f;
    }
    // ---------------- forward
    {
      float x0[OBS];
#pragma unroll
      for (int k = 0; k < OBS; ++k) { x0[k] = x1[k]; u1[k] = x0[k] + p[O_B1B + k]; }
      block_forward<OBS>(s, p, O_W1A, O_B1A, O_W1B, x0, u1);
    }
#pragma unroll
    for (int k = 0; k < OBS; ++k) x1[OBS + k] = lrelu(u1[k]);
#pragma unroll
    for (int k = 0; k < X1; ++k) u2[k] = x1[k] + p[O_B2B + k];
    block_forward<X1>(s, p, O_W2A, O_B2A, O_W2B, x1, u2);
    // ---------------- heads, losses, dL/d(head pre-activations)
    float y2[X1];
#pragma unroll
    for (int k = 0; k < X1; ++k) y2[k] = lrelu(u2[k]);
    float go1 = 0.f, go2 = 0.f;     // dL/d o1, dL/d o2 (critic: go1 = dL/dV)
    if (net == 0) {
      float o1 = p[O_HEAD + X1], o2 = p[O_HEAD + 2 * X1 + 1];
#pragma unroll
      for (int k = 0; k < X1; ++k) { o1 = fmaf(p[O_HEAD + k], y2[k], o1); o2 = fmaf(p[O_HEAD + X1 + 1 + k], y2[k], o2); }
      const float m0 = sigmoidf_(o1), m1 = tanhf(o2);
      if (valid) {
        const float2 av = reinterpret_cast<const float2*>(a.act)[i];
        const float lp = gauss_logp(av.x, av.y, m0, m1, a.var);
        const float lr = lp - a.logp_old[i];
        const float ratio = expf(lr);                                          // ppo.py:316
        const float A = a.adv[i];
        const float s1 = ratio * A;                                            // ppo.py:319
        const float s2 = fminf(fmaxf(ratio, 1.f - a.clip), 1.f + a.clip) * A;  // ppo.py:320
        macc[0] += (double)(-fminf(s1, s2));                                   // ppo.py:342
        macc[2] += (double)((ratio - 1.f) - lr);                               // ppo.py:326
        macc[3] += (fabsf(ratio - 1.f) > a.clip) ? 1.0 : 0.0;                  // ppo.py:335
        // torch.min splits ties and clamp passes gradient on its closed interval:
        // d(-min(s1, s2))/d ratio = -A where s1 <= s2, 0 where the clipped branch wins
        const float g_lp = (s1 <= s2 ? -A : 0.f) * a.inv_n * ratio;
        const float gm0 = g_lp * (av.x - m0) / a.var, gm1 = g_lp * (av.y - m1) / a.var;
        go1 = gm0 * m0 * (1.f - m0);                                           // sigmoid'
        go2 = gm1 * (1.f - m1 * m1);                                           // tanh'
      }
    } else {
      float v = p[O_HEAD + X1];
#pragma unroll
      for (int k = 0; k < X1; ++k) v = fmaf(p[O_HEAD + k], y2[k], v);
      if (valid) {
        const float d = v - a.rtg[i];
        macc[1] += (double)(d * d);                                            // ppo.py:343 MSELoss
        go1 = 2.f * d * a.inv_n;
      }
    }
    // head-weight gradients: warp-reduce g_o * y2[k], fixed-order sum over the 4 warps
    {
      const int nh = (net == 0) ? 2 : 1;
      for (int hd = 0; hd < nh; ++hd) {
        const float g = hd ? go2 : go1;
        float bsum = g;
        for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
#pragma unroll
        for (int k = 0; k < X1; ++k) {
          float t = g * y2[k];
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) s[G_SRED + warp * 72 + hd * (X1 + 1) + k] = t;
        }
        if (lane == 0) s[G_SRED + warp * 72 + hd * (X1 + 1) + X1] = bsum;
      }
    }
    // dL/dy2 -> dL/du2; publish x1 and g_u2 rows for the weight-gradient products
    float gu2[X1], gx1[X1];
#pragma unroll
    for (int k = 0; k < X1; ++k) {
      const float gy = (net == 0) ? fmaf(p[O_HEAD + k], go1, p[O_HEAD + X1 + 1 + k] * go2) : p[O_HEAD + k] * go1;
      gu2[k] = gy * dlrelu(u2[k]);
      gx1[k] = gu2[k];                         // skip connection of the residual block
    }
    {
      float4* sx = reinterpret_cast<float4*>(s + G_SX + m * X1);
      float4* sg = reinterpret_cast<float4*>(s + G_SGU + m * X1);
#pragma unroll
      for (int q = 0; q < X1 / 4; ++q) {
        sx[q] = make_float4(x1[4 * q], x1[4 * q + 1], x1[4 * q + 2], x1[4 * q + 3]);
        sg[q] = make_float4(gu2[4 * q], gu2[4 * q + 1], gu2[4 * q + 2], gu2[4 * q + 3]);
      }
    }
    __syncthreads();
    {
      const int nhead = (net == 0) ? ACTOR_HEAD : CRITIC_HEAD;
      if (m < nhead)
        grow[O_HEAD + m] += ((s[G_SRED + m] + s[G_SRED + 72 + m]) + s[G_SRED + 144 + m]) + s[G_SRED + 216 + m];
      if (m < X1) {                            // fc2 bias of block 2: sum over samples of g_u2
        float t = 0.f;
        for (int mm = 0; mm < GM; ++mm) t += s[G_SGU + mm * X1 + m];
        grow[O_B2B + m] += t;
      }
    }
    // ---------------- backward, block 2
    block_backward<X1, true>(s, p, grow, O_W2A, O_B2A, O_W2B, x1, gu2, gx1);
    // ---------------- backward, block 1 (its input x0 = sX[..][0:16]; no gradient w.r.t. obs needed)
    float gu1[OBS], x0[OBS], dummy[OBS];
#pragma unroll
    for (int k = 0; k < OBS; ++k) { gu1[k] = gx1[OBS + k] * dlrelu(u1[k]); x0[k] = x1[k]; dummy[k] = 0.f; }
    {
      float4* sg = reinterpret_cast<float4*>(s + G_SGU + m * X1);
#pragma unroll
      for (int q = 0; q < OBS / 4; ++q) sg[q] = make_float4(gu1[4 * q], gu1[4 * q + 1], gu1[4 * q + 2], gu1[4 * q + 3]);
    }
    __syncthreads();
    if (m < OBS) {
      float t = 0.f;
      for (int mm = 0; mm < GM; ++mm) t += s[G_SGU + mm * X1 + m];
      grow[O_B1B + m] += t;
    }
    block_backward<OBS, false>(s, p, grow, O_W1A, O_B1A, O_W1B, x0, gu1, dummy);
  }
  // per-CTA metric partials: fixed-order sum over the CTA's threads
  __syncthreads();
  double* red = reinterpret_cast<double*>(s);  // reuse: [4][GM] doubles = 4 KB
#pragma unroll
  for (int q = 0; q < 4; ++q) red[q * GM + m] = macc[q];
  __syncthreads();
  if (m < 4) {
    double t = 0.0;
    for (int mm = 0; mm < GM; ++mm) t += red[m * GM + mm];
    a.mpart[((size_t)net * gridDim.x + blockIdx.x) * 4 + m] = t;
  }
}

// flat gradient (canonical layout) = fixed-order sum of the CTA rows; block 0 also folds the
// metric partials.
__global__ void grad_reduce_kernel(const float* __restrict__ gpart, const double* __restrict__ mpart, int rows,
                                   float inv_n, float* __restrict__ grad, double* __restrict__ metrics) {
  // thread = one element of the KERNEL layout (the rows are read as coalesced lines; the two transposed fc2 regions
  // are un-transposed by the single scattered store), rows added in order, eight loads in flight
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NAVPPO_FLAT) {
    const int net = i >= NAVPPO_CRITIC_OFFSET;
    const int k = i - (net ? NAVPPO_CRITIC_OFFSET : 0);
    const int n = net ? NAVPPO_CRITIC_PARAMS : NAVPPO_ACTOR_PARAMS;
    float t = 0.f;
    if (k < n) {
      const float* col = gpart + (size_t)net * rows * NET_ROW + k;
      int r = 0;
      for (; r + 8 <= rows; r += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = col[(size_t)(r + q) * NET_ROW];
#pragma unroll
        for (int q = 0; q < 8; ++q) t += v[q];
      }
      for (; r < rows; ++r) t += col[(size_t)r * NET_ROW];
    }
    grad[(net ? NAVPPO_CRITIC_OFFSET : 0) + (k < n ? klayout_inv(k) : k)] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) {
    const int q = threadIdx.x;
    const int net = (q == 1) ? 1 : 0;           // critic loss comes from the critic rows
    double t = 0.0;
    for (int r = 0; r < rows; ++r) t += mpart[((size_t)net * rows + r) * 4 + q];
    metrics[q] = t * (double)inv_n;
  }
}

// Where the Adam kernel takes the gradient from.
//   GRAD_LOCAL  g[i]                                             (one GPU, or after an NCCL all-reduce)
//   GRAD_PEERS  sum over the ranks, in rank order, of peer[r][i]  — every rank's flat gradient sits in a buffer
//               all ranks have mapped (NVLink peer memory); each rank reads all of them and adds them in the same
//               order, so the sum is bit-identical on every rank and the all-reduce is fused into the optimiser step
//   GRAD_NVLS   multimem.ld_reduce.add on the buffers' multicast address: the NVSwitch adds the ranks' values on
//               the way (one load per element instead of one per rank; the switch's order is its own)
enum { GRAD_LOCAL = 0, GRAD_PEERS = 1, GRAD_NVLS = 2 };
constexpr int MAX_PEERS = 16;
struct PeerPtrs {
  const float* grad[MAX_PEERS];   // this epoch's buffer of every rank (device pointers valid on THIS rank)
  const float* mc;                // multicast address of the same buffer, or null
  int world;
};

__device__ __forceinline__ float ld_peer(const float* p) {   // bypass L1: a peer's line may have changed since last epoch
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_nvls_sum(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}

// torch.optim.Adam (defaults) on the flat vector; per-block squared-gradient partials.
template <int SRC>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, PeerPtrs peers,
                                                   float* __restrict__ m1, float* __restrict__ m2, float lr_over_bc1,
                                                   float inv_sqrt_bc2, float b1, float b2, float eps,
                                                   double* __restrict__ sq_part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double sa = 0.0, sc = 0.0;
  if (i < NAVPPO_FLAT) {
    float gi;
    if (SRC == GRAD_LOCAL) {
      gi = g[i];
    } else if (SRC == GRAD_PEERS) {
      float v[MAX_PEERS];
#pragma unroll
      for (int r = 0; r < MAX_PEERS; ++r) v[r] = r < peers.world ? ld_peer(peers.grad[r] + i) : 0.f;   // all loads in flight
      gi = v[0];
#pragma unroll
      for (int r = 1; r < MAX_PEERS; ++r)
        if (r < peers.world) gi += v[r];
    } else {
      gi = ld_nvls_sum(peers.mc + i);
    }
    const float m = fmaf(1.f - b1, gi - m1[i], m1[i]);          // exp_avg.lerp_(grad, 1 - beta1)
    const float v = fmaf(b2, m2[i], (1.f - b2) * gi * gi);      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m1[i] = m; m2[i] = v;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - lr_over_bc1 * (m / denom);
    if (i < NAVPPO_CRITIC_OFFSET) sa = (double)gi * gi; else sc = (double)gi * gi;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_down_sync(0xffffffffu, sa, o);
    sc += __shfl_down_sync(0xffffffffu, sc, o);
  }
  __shared__ double wa[8], wc[8];
  if ((threadIdx.x & 31) == 0) { wa[threadIdx.x >> 5] = sa; wc[threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tc = 0.0;
    for (int k = 0; k < 8; ++k) { ta += wa[k]; tc += wc[k]; }
    sq_part[2 * blockIdx.x] = ta;
    sq_part[2 * blockIdx.x + 1] = tc;
  }
}

// Cross-GPU barrier over peer memory: rank r writes `token` into slot r of every rank's flag array and waits until
// its own array holds the token from every rank.  Ordered after the kernel that wrote this rank's gradient buffer
// (same stream) and before the Adam kernel that reads everyone's: release / acquire at system scope.
struct PeerFlags {
  uint32_t* flags[MAX_PEERS];     // flag array (MAX_PEERS words) of every rank, mapped on this rank
  int rank, world;
};
__global__ void peer_barrier_kernel(PeerFlags f, uint32_t token) {
  const int t = threadIdx.x;
  if (t < f.world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.flags[t] + f.rank), "r"(token) : "memory");
    uint32_t seen;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f.flags[f.rank] + t) : "memory");
    } while ((int32_t)(seen - token) < 0);
  }
}

__global__ void gradnorm_finalize_kernel(const double* __restrict__ sq_part, int nblocks, double* __restrict__ metrics) {
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += sq_part[2 * b + threadIdx.x];
    metrics[NAVPPO_M_ACTOR_GRAD_SQ + threadIdx.x] = t;
  }
}

}  // namespace

// ========================================================================================
// C-ABI
// ========================================================================================
struct navppo {
  navppo_cfg cfg;
  int grad_rows = 0;          // CTAs per network of mlp_grad_kernel
  float* gpart = nullptr;     // [2][grad_rows][NET_ROW]
  double* mpart = nullptr;    // [2][grad_rows][4]
  double* sq_part = nullptr;  // [adam blocks][2]
  double* adv_stats = nullptr;  // [3]
  float* grad_ws = nullptr;   // [NAVPPO_FLAT] used by navppo_update
  float* wprep = nullptr;     // tensor-core path: pre-split, pre-tiled weights
  bool tc_single_role = false;  // NAVPPO_TC_KERNEL=single: the first (single-role) tcgen05 kernel, kept as a cross-check
  bool fused_rollout = false;   // rollout: ONE persistent launch, each CTA keeps 128 robots for all H steps (NAVPPO_ROLLOUT_FUSED=0: off)
  bool chain = false;           // rollout: policy and step kernels as programmatic dependent launches (NAVPPO_ROLLOUT_CHAIN=0: plain)
  bool tc_infer = false;        // tensor-core precision: the rollout's and the update's forward passes run on the tensor cores
                                // too (NAVPPO_TC_INFER=0 keeps them on the fp32 CUDA-core kernel)
  // peer-memory gradient exchange (navppo_peer_setup): two gradient buffers + a flag array per rank
  int peer_rank = 0, peer_world = 0;
  const float* peer_grad[2][16] = {};
  const float* peer_mc[2] = {nullptr, nullptr};
  uint32_t* peer_flags[16] = {};
  int sm_count = 148;
  int64_t launches = 0;
};

namespace {

constexpr int ADAM_BLOCK = 256;
constexpr int ADAM_GRID = (NAVPPO_FLAT + ADAM_BLOCK - 1) / ADAM_BLOCK;

int check_handle(const navppo* h) {
  if (!h) return nav_fail(NAVSIM_EINVAL, "null navppo handle");
  return NAVSIM_OK;
}

template <int MODE>
int launch_infer(navppo* h, const InferArgs& a, bool both_nets, cudaStream_t s) {
  if (a.T <= 0) return nav_fail(NAVSIM_EINVAL, "T must be positive");
  if (h->tc_infer) {   // a tensor-core handle runs every network product on the tensor cores: re-tile the weights, then infer
    if (int rc = navppo_tcws_prep_launch(a.params, h->wprep, s)) return rc;
    h->launches += 2;
    return navppo_tcws_infer_launch(a, MODE, both_nets, h->cfg.precision == NAVPPO_BF16X3 ? 3 : 1, h->wprep, false, false, s);
  }
  const int grid = (a.T + CF_TILE - 1) / CF_TILE;
  mlp_infer_kernel<MODE><<<dim3(grid, both_nets ? 2 : 1), CF_THREADS, CF_SMEM, s>>>(a);
  h->launches++;
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

}  // namespace

extern "C" {

int navppo_default_cfg(navppo_cfg* cfg) {
  if (!cfg) return nav_fail(NAVSIM_EINVAL, "cfg is null");
  cfg->device = 0;
  cfg->max_samples = 1 << 20;
  cfg->precision = NAVPPO_FP32;
  cfg->reserved0 = 0;
  cfg->lr = 3e-4;         // main.py:472
  cfg->beta1 = 0.9;       // torch.optim.Adam defaults (ppo.py:116-117)
  cfg->beta2 = 0.999;
  cfg->adam_eps = 1e-8;
  cfg->clip = 0.2;        // main.py:473
  return NAVSIM_OK;
}

int navppo_create(navppo_t** out, const navppo_cfg* cfg) {
  if (!out || !cfg) return nav_fail(NAVSIM_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->max_samples <= 0) return nav_fail(NAVSIM_EINVAL, "max_samples must be positive");
  if (cfg->precision != NAVPPO_FP32 && cfg->precision != NAVPPO_BF16X3 && cfg->precision != NAVPPO_BF16)
    return nav_fail(NAVSIM_EINVAL, "unknown precision mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return nav_fail(NAVSIM_ENODEV, "no CUDA device: the PPO kernels have no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return nav_fail(NAVSIM_EINVAL, "device ordinal out of range");
  NAV_CUDA_TRY(cudaSetDevice(cfg->device));
  navppo* h = new (std::nothrow) navppo();
  if (!h) return nav_fail(NAVSIM_ENOMEM, "host allocation failed");
  h->cfg = *cfg;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  const int tiles = (cfg->max_samples + GM - 1) / GM;
  // 3 CTAs of the gradient kernel fit one SM (shared memory); the two networks share the grid
  int rows = (h->sm_count * 3) / 2;
  if (rows > tiles) rows = tiles;
  if (rows < 1) rows = 1;
  h->grad_rows = rows;
  cudaError_t e = cudaMalloc(&h->gpart, (size_t)2 * rows * NET_ROW * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->mpart, (size_t)2 * rows * 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->sq_part, (size_t)ADAM_GRID * 2 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->adv_stats, 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->grad_ws, (size_t)NAVPPO_FLAT * sizeof(float));
  if (e == cudaSuccess && cfg->precision != NAVPPO_FP32) {
    const char* which = getenv("NAVPPO_TC_KERNEL");
    h->tc_single_role = which && std::string(which) == "single";
    const char* ti = std::getenv("NAVPPO_TC_INFER");
    h->tc_infer = !(ti && std::string(ti) == "0");
    const char* fr = std::getenv("NAVPPO_ROLLOUT_FUSED");
    h->fused_rollout = h->tc_infer && !(fr && std::string(fr) == "0");
    const char* ch = std::getenv("NAVPPO_ROLLOUT_CHAIN");
    h->chain = !(ch && std::string(ch) == "0");
    const size_t wb = navppo_tc_prep_bytes() > navppo_tcws_prep_bytes() ? navppo_tc_prep_bytes() : navppo_tcws_prep_bytes();
    e = cudaMalloc(&h->wprep, wb);
    if (e == cudaSuccess && (navppo_tc_init() != NAVSIM_OK || navppo_tcws_init() != NAVSIM_OK)) {
      navppo_destroy(h);
      return NAVSIM_ECUDA;   // nav_last_error already set
    }
  }
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_infer_kernel<INFER_FORWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_infer_kernel<INFER_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_infer_kernel<INFER_EVALUATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAD_SMEM);
  if (e != cudaSuccess) {
    navppo_destroy(h);
    return nav_fail(e == cudaErrorMemoryAllocation ? NAVSIM_ENOMEM : NAVSIM_ECUDA,
                    std::string("navppo_create: ") + cudaGetErrorString(e));
  }
  *out = h;
  return NAVSIM_OK;
}

int navppo_destroy(navppo_t* h) {
  if (!h) return NAVSIM_OK;
  cudaSetDevice(h->cfg.device);
  if (h->gpart) cudaFree(h->gpart);
  if (h->mpart) cudaFree(h->mpart);
  if (h->sq_part) cudaFree(h->sq_part);
  if (h->adv_stats) cudaFree(h->adv_stats);
  if (h->grad_ws) cudaFree(h->grad_ws);
  if (h->wprep) cudaFree(h->wprep);
  delete h;
  return NAVSIM_OK;
}

int64_t navppo_launch_count(const navppo_t* h) { return h ? h->launches : 0; }

int navppo_rtg_scan(const float* rew, const uint8_t* term, const float* values, const float* last_value, double gamma,
                    double lam, float* out, int32_t H, int32_t N, void* stream) {
  if (!rew || !term || !out) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (H <= 0 || N <= 0) return nav_fail(NAVSIM_EINVAL, "H and N must be positive");
  rtg_scan_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rew, term, values, last_value, gamma, lam, out, H, N);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_forward(navppo_t* h, const float* params, const float* obs, int32_t T, float* mu, float* v, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !obs || (!mu && !v)) return nav_fail(NAVSIM_EINVAL, "null buffer");
  InferArgs a{};
  a.params = params; a.obs = obs; a.T = T; a.mu = mu; a.v = v; a.var = 1.f;
  return launch_infer<INFER_FORWARD>(h, a, v != nullptr, (cudaStream_t)stream);
}

int navppo_act(navppo_t* h, const float* params, const float* obs, int32_t N, double var, uint64_t seed,
               int64_t agent_id_offset, uint32_t draw, const float* noise_in, float* act, float* logp, float* mu_out,
               void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !obs || !act || !logp) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (!(var > 0.0)) return nav_fail(NAVSIM_EINVAL, "var must be positive");
  InferArgs a{};
  a.params = params; a.obs = obs; a.T = N; a.var = (float)var; a.seed = seed; a.agent_off = agent_id_offset;
  a.draw = draw; a.noise_in = noise_in; a.act = act; a.logp = logp; a.mu = mu_out;
  return launch_infer<INFER_ACT>(h, a, false, (cudaStream_t)stream);
}

int navppo_evaluate(navppo_t* h, const float* params, const float* obs, const float* act, int32_t T, double var,
                    float* v, float* logp, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !obs || !act || !v || !logp) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (!(var > 0.0)) return nav_fail(NAVSIM_EINVAL, "var must be positive");
  InferArgs a{};
  a.params = params; a.obs = obs; a.T = T; a.var = (float)var; a.act_in = act; a.v = v; a.logp = logp;
  return launch_infer<INFER_EVALUATE>(h, a, true, (cudaStream_t)stream);
}

int navppo_adv_stats(const float* rtg, const float* v, int32_t T, double* stats, void* stream) {
  if (!rtg || !v || !stats) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (T <= 0) return nav_fail(NAVSIM_EINVAL, "T must be positive");
  int grid = (T + 255) / 256;
  if (grid > 592) grid = 592;
  adv_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rtg, v, T, stats);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_adv_normalize(const float* rtg, const float* v, int32_t T, const double* stats, float* adv, void* stream) {
  if (!rtg || !v || !stats || !adv) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (T <= 0) return nav_fail(NAVSIM_EINVAL, "T must be positive");
  int grid = (T + 255) / 256;
  if (grid > 1184) grid = 1184;
  adv_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rtg, v, T, stats, adv);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_grad(navppo_t* h, const float* params, const float* obs, const float* act, const float* logp_old,
                const float* adv, const float* rtg, int32_t T, int64_t n_global, double var, float* grad,
                double* metrics, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !obs || !act || !logp_old || !adv || !rtg || !grad || !metrics) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (T <= 0 || T > h->cfg.max_samples) return nav_fail(NAVSIM_EINVAL, "T outside (0, max_samples]");
  if (n_global < T) return nav_fail(NAVSIM_EINVAL, "n_global must be >= T");
  if (!(var > 0.0)) return nav_fail(NAVSIM_EINVAL, "var must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles = (T + GM - 1) / GM;
  int rows = tiles < h->grad_rows ? tiles : h->grad_rows;
  GradArgs a{};
  a.params = params; a.obs = obs; a.act = act; a.logp_old = logp_old; a.adv = adv; a.rtg = rtg; a.T = T;
  a.inv_n = (float)(1.0 / (double)n_global); a.var = (float)var; a.clip = (float)h->cfg.clip;
  a.gpart = h->gpart; a.mpart = h->mpart;
  if (h->cfg.precision == NAVPPO_FP32) {
    mlp_grad_kernel<<<dim3(rows, 2), GM, GRAD_SMEM, s>>>(a);
  } else {
    // tensor-core kernel: one 256-thread CTA per SM (all of its shared memory and TMEM), the
    // two networks split the SMs
    const int per_net = h->sm_count / 2 > 0 ? h->sm_count / 2 : 1;
    rows = tiles < per_net ? tiles : per_net;
    const int passes = h->cfg.precision == NAVPPO_BF16X3 ? 3 : 1;
    if (int rc = h->tc_single_role ? navppo_tc_grad_launch(a, rows, passes, h->wprep, s)
                                   : navppo_tcws_grad_launch(a, rows, passes, h->wprep, s))
      return rc;
    h->launches++;
  }
  grad_reduce_kernel<<<(NAVPPO_FLAT + 255) / 256, 256, 0, s>>>(h->gpart, h->mpart, rows, a.inv_n, grad, metrics);
  h->launches += 2;
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_adam(navppo_t* h, float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t step,
                double* metrics, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !grad || !exp_avg || !exp_avg_sq) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (step < 1) return nav_fail(NAVSIM_EINVAL, "step is 1-based");
  cudaStream_t s = (cudaStream_t)stream;
  const double b1 = h->cfg.beta1, b2 = h->cfg.beta2;
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  adam_kernel<GRAD_LOCAL><<<ADAM_GRID, ADAM_BLOCK, 0, s>>>(params, grad, PeerPtrs{}, exp_avg, exp_avg_sq,
                                                          (float)(h->cfg.lr / bc1), (float)(1.0 / sqrt(bc2)), (float)b1,
                                                          (float)b2, (float)h->cfg.adam_eps, h->sq_part);
  h->launches++;
  if (metrics) {
    gradnorm_finalize_kernel<<<1, 32, 0, s>>>(h->sq_part, ADAM_GRID, metrics);
    h->launches++;
  }
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_peer_setup(navppo_t* h, int32_t rank, int32_t world, const uint64_t* grad_ptrs, const uint64_t* flag_ptrs,
                      const uint64_t* multicast_ptrs) {
  if (int rc = check_handle(h)) return rc;
  if (!grad_ptrs || !flag_ptrs) return nav_fail(NAVSIM_EINVAL, "null pointer table");
  if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return nav_fail(NAVSIM_EINVAL, "rank / world out of range");
  for (int b = 0; b < 2; ++b) {
    for (int r = 0; r < world; ++r) {
      if (!grad_ptrs[b * world + r]) return nav_fail(NAVSIM_EINVAL, "null peer buffer");
      h->peer_grad[b][r] = reinterpret_cast<const float*>(grad_ptrs[b * world + r]);
    }
    h->peer_mc[b] = multicast_ptrs ? reinterpret_cast<const float*>(multicast_ptrs[b]) : nullptr;
  }
  for (int r = 0; r < world; ++r) {
    if (!flag_ptrs[r]) return nav_fail(NAVSIM_EINVAL, "null peer flag array");
    h->peer_flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
  }
  h->peer_rank = rank; h->peer_world = world;
  return NAVSIM_OK;
}

int navppo_adam_peer(navppo_t* h, float* params, float* exp_avg, float* exp_avg_sq, int32_t step, int32_t buffer,
                     uint32_t token, int32_t use_multicast, double* metrics, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!params || !exp_avg || !exp_avg_sq) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (h->peer_world < 1) return nav_fail(NAVSIM_EINVAL, "navppo_peer_setup has not been called");
  if (step < 1) return nav_fail(NAVSIM_EINVAL, "step is 1-based");
  if (buffer != 0 && buffer != 1) return nav_fail(NAVSIM_EINVAL, "buffer is 0 or 1");
  if (use_multicast && !h->peer_mc[buffer]) return nav_fail(NAVSIM_EINVAL, "no multicast address was given to navppo_peer_setup");
  cudaStream_t s = (cudaStream_t)stream;
  PeerFlags f{};
  PeerPtrs pp{};
  for (int r = 0; r < h->peer_world; ++r) { f.flags[r] = h->peer_flags[r]; pp.grad[r] = h->peer_grad[buffer][r]; }
  f.rank = h->peer_rank; f.world = h->peer_world;
  pp.world = h->peer_world; pp.mc = h->peer_mc[buffer];
  peer_barrier_kernel<<<1, 32, 0, s>>>(f, token);
  const double b1 = h->cfg.beta1, b2 = h->cfg.beta2;
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  if (use_multicast)
    adam_kernel<GRAD_NVLS><<<ADAM_GRID, ADAM_BLOCK, 0, s>>>(params, nullptr, pp, exp_avg, exp_avg_sq, (float)(h->cfg.lr / bc1),
                                                           (float)(1.0 / sqrt(bc2)), (float)b1, (float)b2,
                                                           (float)h->cfg.adam_eps, h->sq_part);
  else
    adam_kernel<GRAD_PEERS><<<ADAM_GRID, ADAM_BLOCK, 0, s>>>(params, nullptr, pp, exp_avg, exp_avg_sq, (float)(h->cfg.lr / bc1),
                                                            (float)(1.0 / sqrt(bc2)), (float)b1, (float)b2,
                                                            (float)h->cfg.adam_eps, h->sq_part);
  h->launches += 2;
  if (metrics) {
    gradnorm_finalize_kernel<<<1, 32, 0, s>>>(h->sq_part, ADAM_GRID, metrics);
    h->launches++;
  }
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_rollout_ex(navppo_t* h, navsim_t* sim, const float* params, int32_t H, double var, uint64_t seed,
                      int64_t agent_id_offset, uint32_t draw0, float* obs, float* next_obs, float* act, float* logp, float* rew,
                      uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_return, float* ep_path, int32_t* ep_len,
                      const uint32_t* dyn_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!sim || !params || !obs || !next_obs || !act || !logp || !rew || !done || !arrive || !trunc)
    return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (H < 1) return nav_fail(NAVSIM_EINVAL, "H must be positive");
  if (!(var > 0.0)) return nav_fail(NAVSIM_EINVAL, "var must be positive");
  const size_t N = (size_t)navsim_num_agents(sim);
  const int tc_passes = h->cfg.precision == NAVPPO_BF16X3 ? 3 : 1;
  if (h->tc_infer) {   // the policy does not change inside a rollout: re-tile its weights once
    if (int rc = navppo_tcws_prep_launch(params, h->wprep, (cudaStream_t)stream)) return rc;
    h->launches++;
  }
  if (h->fused_rollout) {
    InferArgs a{};
    a.params = params; a.obs = obs; a.T = (int)N; a.var = (float)var; a.seed = seed; a.agent_off = agent_id_offset;
    a.draw = draw0; a.act = act; a.logp = logp; a.dyn = dyn_dev;
    int unsupported = 0;
    const int rc = navppo_tcws_rollout_launch(sim, a, H, tc_passes, h->wprep, obs, next_obs, rew, done, arrive, trunc, ep_return,
                                              ep_path, ep_len, &unsupported, (cudaStream_t)stream);
    if (rc == NAVSIM_OK) { h->launches++; return NAVSIM_OK; }
    if (!unsupported) return rc;
  }
  for (int t = 0; t < H; ++t) {
    float* o_t = obs + (size_t)t * N * OBS;
    float* o_next = (t + 1 < H) ? obs + (size_t)(t + 1) * N * OBS : next_obs;
    InferArgs a{};
    a.params = params; a.obs = o_t; a.T = (int)N; a.var = (float)var; a.seed = seed; a.agent_off = agent_id_offset;
    a.draw = draw0 + (uint32_t)t; a.act = act + (size_t)t * N * 2; a.logp = logp + (size_t)t * N; a.dyn = dyn_dev;
    if (h->tc_infer) {
      if (int rc = navppo_tcws_infer_launch(a, INFER_ACT, false, tc_passes, h->wprep, h->chain, t > 0, (cudaStream_t)stream)) return rc;
      h->launches++;
    } else if (int rc = launch_infer<INFER_ACT>(h, a, false, (cudaStream_t)stream)) return rc;
    navsim_step_out out;
    out.obs = o_next; out.rew = rew + (size_t)t * N; out.done = done + (size_t)t * N; out.arrive = arrive + (size_t)t * N;
    out.trunc = trunc + (size_t)t * N;
    out.ep_return = ep_return ? ep_return + (size_t)t * N : nullptr;
    out.ep_path = ep_path ? ep_path + (size_t)t * N : nullptr;
    out.ep_len = ep_len ? ep_len + (size_t)t * N : nullptr;
    if (h->tc_infer && h->chain) {
      if (int rc = navsim_step_chained(sim, act + (size_t)t * N * 2, &out, stream)) return rc;
    } else if (int rc = navsim_step_ex(sim, act + (size_t)t * N * 2, &out, stream)) return rc;
  }
  return NAVSIM_OK;
}

int navppo_rollout(navppo_t* h, navsim_t* sim, const float* params, int32_t H, double var, uint64_t seed,
                   int64_t agent_id_offset, uint32_t draw0, float* obs, float* next_obs, float* act, float* logp, float* rew,
                   uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_return, float* ep_path, void* stream) {
  return navppo_rollout_ex(h, sim, params, H, var, seed, agent_id_offset, draw0, obs, next_obs, act, logp, rew, done, arrive,
                           trunc, ep_return, ep_path, nullptr, nullptr, stream);
}

int navppo_update(navppo_t* h, float* params, float* exp_avg, float* exp_avg_sq, int32_t step0, const float* obs,
                  const float* act, const float* logp_old, const float* rtg, int32_t T, double var, int32_t epochs,
                  float* adv_ws, float* v_ws, double* metrics, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!adv_ws || !v_ws || !metrics) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (epochs < 0 || step0 < 0) return nav_fail(NAVSIM_EINVAL, "negative epochs / step0");
  cudaStream_t s = (cudaStream_t)stream;
  // ppo.py:275-284: V = evaluate(...); A = rtg - V; normalise.  adv_ws doubles as the
  // throw-away log-prob output of this first evaluate.
  if (int rc = navppo_evaluate(h, params, obs, act, T, var, v_ws, adv_ws, stream)) return rc;
  NAV_CUDA_TRY(cudaMemsetAsync(h->adv_stats, 0, 3 * sizeof(double), s));
  if (int rc = navppo_adv_stats(rtg, v_ws, T, h->adv_stats, stream)) return rc;
  if (int rc = navppo_adv_normalize(rtg, v_ws, T, h->adv_stats, adv_ws, stream)) return rc;
  h->launches += 2;
  for (int e = 0; e < epochs; ++e) {  // ppo.py:305
    double* mrow = metrics + (size_t)e * NAVPPO_NUM_METRICS;
    if (int rc = navppo_grad(h, params, obs, act, logp_old, adv_ws, rtg, T, (int64_t)T, var, h->grad_ws, mrow, stream)) return rc;
    if (int rc = navppo_adam(h, params, h->grad_ws, exp_avg, exp_avg_sq, step0 + e + 1, mrow, stream)) return rc;
  }
  return NAVSIM_OK;
}

}  // extern "C"
