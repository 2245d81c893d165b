// ppo_common.cuh — constants and per-sample device helpers shared by the CUDA-core
// (navppo_kernels.cu) and tensor-core (navppo_tc.cu) translation units.
#ifndef NAVPPO_COMMON_CUH_
#define NAVPPO_COMMON_CUH_

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/navppo.h"

namespace ppo {

constexpr int OBS = NAVSIM_OBS_DIM;  // 16
constexpr int X1 = 2 * OBS;          // 32
constexpr int HID = NAVPPO_HIDDEN;   // 512
constexpr float LEAK = 0.2f;         // nn.LeakyReLU(negative_slope=0.2), net_actor.py:37

// canonical offsets inside one network's flat vector (navbot_ppo_b200/layout.py)
constexpr int O_W1A = 0;                      // [512][16]
constexpr int O_B1A = O_W1A + HID * OBS;      // 8192
constexpr int O_W1B = O_B1A + HID;            // 8704   [16][512]
constexpr int O_B1B = O_W1B + OBS * HID;      // 16896
constexpr int O_W2A = O_B1B + OBS;            // 16912  [512][32]
constexpr int O_B2A = O_W2A + HID * X1;       // 33296
constexpr int O_W2B = O_B2A + HID;            // 33808  [32][512]
constexpr int O_B2B = O_W2B + X1 * HID;       // 50192
constexpr int O_HEAD = O_B2B + X1;            // 50224
constexpr int ACTOR_HEAD = 2 * (X1 + 1);      // out1.weight, out1.bias, out2.weight, out2.bias
constexpr int CRITIC_HEAD = X1 + 1;
static_assert(O_HEAD + ACTOR_HEAD == NAVPPO_ACTOR_PARAMS, "actor layout");
static_assert(O_HEAD + CRITIC_HEAD == NAVPPO_CRITIC_PARAMS, "critic layout");

// "kernel layout" of one network: same regions, but the two fc2 matrices are stored
// transposed ([hidden][out]) so that everything a hidden unit touches is contiguous.
// canonical index -> kernel-layout index
__host__ __device__ inline int klayout(int i) {
  if (i >= O_W1B && i < O_B1B) { const int r = i - O_W1B; return O_W1B + (r % HID) * OBS + r / HID; }
  if (i >= O_W2B && i < O_B2B) { const int r = i - O_W2B; return O_W2B + (r % HID) * X1 + r / HID; }
  return i;
}
// kernel-layout index -> canonical index (the inverse of klayout)
__host__ __device__ inline int klayout_inv(int k) {
  if (k >= O_W1B && k < O_B1B) { const int r = k - O_W1B; return O_W1B + (r % OBS) * HID + r / OBS; }
  if (k >= O_W2B && k < O_B2B) { const int r = k - O_W2B; return O_W2B + (r % X1) * HID + r / X1; }
  return k;
}
constexpr int NET_ROW = NAVPPO_CRITIC_OFFSET;  // 50304: padded length of one network's vector

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : LEAK * x; }
__device__ __forceinline__ float dlrelu(float x) { return x > 0.f ? 1.f : LEAK; }


__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// log N(a; mu, var I), k = 2: -1/2 |a - mu|^2 / var - ln(2 pi) - ln(var)   (ppo.py:704,735)
__device__ __forceinline__ float gauss_logp(float a0, float a1, float m0, float m1, float var) {
  const float d0 = a0 - m0, d1 = a1 - m1;
  return -0.5f * (d0 * d0 + d1 * d1) / var - 1.8378770664093453f - logf(var);
}

// arguments of one gradient pass (one epoch body over this rank's samples)
struct GradArgs {
  const float* params;
  const float* obs; const float* act; const float* logp_old; const float* adv; const float* rtg;
  int T;
  float inv_n;        // 1 / n_global
  float var, clip;
  float* gpart;       // [2][rows][NET_ROW] per-CTA gradient partials, kernel layout
  double* mpart;      // [2][rows][4] per-CTA metric sums
};

enum { INFER_FORWARD = 0, INFER_ACT = 1, INFER_EVALUATE = 2 };

// arguments of one inference pass over T samples (NetActor / NetCritic forward + the get_action / evaluate epilogues)
struct InferArgs {
  const float* params;   // flat [actor | critic]
  const float* obs;      // [T,16]
  int T;
  float var;
  // forward
  float* mu;             // [T,2] or null
  float* v;              // [T]   or null
  // act
  uint64_t seed;
  int64_t agent_off;
  uint32_t draw;
  const float* noise_in; // [T,2] or null
  float* act;            // [T,2]
  float* logp;           // [T]
  // evaluate
  const float* act_in;   // [T,2]
  // act, optional: device words {float bits of var, draw increment} read at run time, so that a captured
  // CUDA graph of the rollout can be replayed with a new variance / noise counter
  const uint32_t* dyn;
};

}  // namespace ppo

#endif  // NAVPPO_COMMON_CUH_
