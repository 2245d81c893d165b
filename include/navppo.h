/* navppo.h — C-ABI of the B200 PPO trainer kernels (policy / value residual MLPs, sampling,
 * reward-to-go, clipped-surrogate + value loss, backward, Adam).
 *
 * Drop-in boundary for the tensor work of the reference's trainer,
 *     project_ppo/src/ppo.py      class PPO                 (:37-946)
 *     project_ppo/src/net_actor.py / net_critic.py          (NetActor / NetCritic)
 * which the reference issues as ~30 separate PyTorch ops per epoch.  Each entry point names
 * the reference lines it replaces.  navbot_ppo_b200.PPO binds these through ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions (same as navsim.h): plain C, returns 0 or a negative NAVSIM_E* code with
 * nav_last_error() set; every array is DEVICE memory owned by the caller (torch tensors),
 * fp32 contiguous unless stated; work is enqueued on `stream` with no host synchronisation;
 * calls are CUDA-graph capturable.
 *
 * Parameters: both networks live in ONE flat fp32 vector of NAVPPO_FLAT elements,
 *     [0, 50290)               actor  (order: navbot_ppo_b200/layout.py ACTOR_SPEC)
 *     [50290, 50304)           zero padding (keeps the critic 128-byte aligned)
 *     [50304, 50304 + 50257)   critic (CRITIC_SPEC)
 *     [.., NAVPPO_FLAT)        zero padding
 * Gradients and the two Adam moment vectors use the same layout, so the data-parallel
 * gradient exchange of a multi-GPU run is one all-reduce over NAVPPO_FLAT floats.
 */
#ifndef NAVPPO_H_
#define NAVPPO_H_

#include <stdint.h>

#include "navsim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NAVPPO_HIDDEN 512          /* ResBlock n_neurons, net_actor.py:21 */
#define NAVPPO_ACTOR_PARAMS 50290
#define NAVPPO_CRITIC_PARAMS 50257
#define NAVPPO_CRITIC_OFFSET 50304
#define NAVPPO_FLAT 100608
#define NAVPPO_NUM_METRICS 8

/* slots of one metrics row (doubles) */
enum navppo_metric {
  NAVPPO_M_ACTOR_LOSS = 0,   /* ppo.py:342 */
  NAVPPO_M_CRITIC_LOSS = 1,  /* ppo.py:343 */
  NAVPPO_M_APPROX_KL = 2,    /* ppo.py:326 */
  NAVPPO_M_CLIP_FRAC = 3,    /* ppo.py:335 */
  NAVPPO_M_ACTOR_GRAD_SQ = 4,  /* squared L2 norm of the actor gradient  (ppo.py:352) */
  NAVPPO_M_CRITIC_GRAD_SQ = 5  /* squared L2 norm of the critic gradient (ppo.py:389) */
};

/* arithmetic of the MLP GEMMs */
enum navppo_precision {
  NAVPPO_FP32 = 0,   /* CUDA-core FFMA, the reference's fp32 arithmetic */
  NAVPPO_BF16X3 = 1, /* tcgen05 tensor cores, split-BF16 operands (hi + lo), 3 MMAs per step: ~16 mantissa bits */
  NAVPPO_BF16 = 2    /* tcgen05 tensor cores, plain BF16 operands, fp32 accumulate */
};

typedef struct navppo navppo_t;

typedef struct navppo_cfg {
  int32_t device;
  int32_t max_samples;  /* largest T of any call (sizes the gradient workspace) */
  int32_t precision;    /* enum navppo_precision: arithmetic of navppo_grad's matrix products */
  int32_t reserved0;
  double lr;            /* ppo.py:769  3e-4 on the main.py path */
  double beta1, beta2, adam_eps; /* torch.optim.Adam defaults, ppo.py:116-117 */
  double clip;          /* ppo.py:771  0.2 */
} navppo_cfg;

int navppo_default_cfg(navppo_cfg* cfg);
int navppo_create(navppo_t** out, const navppo_cfg* cfg);
int navppo_destroy(navppo_t* h);
/* kernels this handle has launched so far */
int64_t navppo_launch_count(const navppo_t* h);

/* PPO.compute_rtgs (ppo.py:643-671) on the time-major [H, N] rollout layout: one reverse
 * scan per agent, R_t = r_t + gamma R_{t+1}, restarted where term[t, n] != 0 (last step of an
 * episode: done | arrive | timeout, ppo.py:552-553) and at t = H-1 (no bootstrap, ppo.py:601).
 * With values != NULL it computes GAE(gamma, lam) advantages instead (delta_t = r_t + gamma
 * V_{t+1} - V_t, last_value[N] = V_H or NULL for 0); lam = 1 and zero values give rtg - V.
 * With values == NULL and last_value != NULL the reward-to-go of the trailing partial episode is
 * bootstrapped: R_{H-1} = r_{H-1} + gamma last_value (not in the reference; PPO's
 * `bootstrap_value` option for rollouts that continue episodes across iterations).
 * Accumulates in fp64 like the reference's Python floats, stores fp32 (ppo.py:669). */
int navppo_rtg_scan(const float* rew, const uint8_t* term, const float* values, const float* last_value, double gamma,
                    double lam, float* out, int32_t H, int32_t N, void* stream);

/* NetActor.forward / NetCritic.forward (net_actor.py:94-144, net_critic.py:83-130) on T rows.
 * mu[T,2] and / or v[T] may be NULL. */
int navppo_forward(navppo_t* h, const float* params, const float* obs, int32_t T, float* mu, float* v, void* stream);

/* PPO.get_action (ppo.py:673-706) for N agents at once: mean = actor(obs); a = mean +
 * sqrt(var) * eps; clamp a0 to [0,1], a1 to [-1,1]; log-prob of the CLAMPED action.
 * eps ~ N(0, I) is drawn on device from Philox4x32-10 keyed (seed, agent_id_offset + i) at
 * counter `draw`, unless noise_in[N,2] is given (parity tests).  mu_out[N,2] may be NULL. */
int navppo_act(navppo_t* h, const float* params, const float* obs, int32_t N, double var, uint64_t seed,
               int64_t agent_id_offset, uint32_t draw, const float* noise_in, float* act, float* logp, float* mu_out,
               void* stream);

/* PPO.evaluate (ppo.py:708-737): V[T] = critic(obs), logp[T] = N(actor(obs), var I).log_prob(act). */
int navppo_evaluate(navppo_t* h, const float* params, const float* obs, const float* act, int32_t T, double var,
                    float* v, float* logp, void* stream);

/* The step loop of PPO.rollout (ppo.py:505-549) for all N agents of `sim`, H steps, enqueued
 * back to back with no host round trip: for t < H  { act[t], logp[t] = get_action(obs[t]);
 * obs[t+1], rew[t], flags[t] = sim.step(act[t]) }.  obs[H,N,16] must hold the first
 * observation (Env.reset) in row 0; the observation after the last step goes to next_obs[N,16].
 * act[H,N,2], logp[H,N], rew[H,N], done/arrive/trunc[H,N] are the time-major rollout buffers;
 * ep_return[H,N] / ep_path[H,N] (both or neither, may be NULL) receive, at the step that ends an
 * episode, its return and path length (the per-episode csv of ppo.py:739-746).
 * Action noise: Philox counter draw0 + t (see navppo_act).
 * How it is enqueued depends on the handle: NAVPPO_FP32 -> 2 H launches (policy kernel, step kernel); a tensor-core
 * precision -> ONE persistent launch in which every CTA keeps 128 robots for all H steps (policy forward on tcgen05,
 * sampling, Env.step by the same threads) when the simulator runs the reference's 10-beam sensor on a map of at most
 * 32 walls, else the tensor-core policy kernel and the step kernel chained as programmatic dependent launches.  All
 * forms of one precision return bit-identical buffers.  A handle of a tensor-core precision runs navppo_forward /
 * navppo_act / navppo_evaluate on the tensor cores as well (same products in the same order as the gradient kernel's
 * forward half: the log-prob navppo_grad recomputes for a sample is bit-identical to the one stored here). */
int navppo_rollout(navppo_t* h, navsim_t* sim, const float* params, int32_t H, double var, uint64_t seed,
                   int64_t agent_id_offset, uint32_t draw0, float* obs, float* next_obs, float* act, float* logp, float* rew,
                   uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_return, float* ep_path, void* stream);

/* The same loop with two extras for replaying it as a captured CUDA graph and for device-side episode
 * bookkeeping: ep_len[H,N] (may be NULL) receives, at the step that ends an episode, its length in steps
 * (batch_lens, ppo.py:583); dyn_dev (may be NULL) points to two device words {float bits of the variance,
 * increment of the noise counter} that the sampling epilogue reads at run time INSTEAD of `var` and in
 * addition to `draw0`, so one captured launch sequence serves every rollout. */
int navppo_rollout_ex(navppo_t* h, navsim_t* sim, const float* params, int32_t H, double var, uint64_t seed,
                      int64_t agent_id_offset, uint32_t draw0, float* obs, float* next_obs, float* act, float* logp,
                      float* rew, uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_return, float* ep_path,
                      int32_t* ep_len, const uint32_t* dyn_dev, void* stream);

/* Advantage, ppo.py:277,284, in two halves so a multi-GPU run can all-reduce the three
 * doubles in between: stats[0..2] += (sum, sum of squares, count) of A = rtg - v over T rows;
 * then adv = (A - mean) / (unbiased std + 1e-10). */
int navppo_adv_stats(const float* rtg, const float* v, int32_t T, double* stats, void* stream);
int navppo_adv_normalize(const float* rtg, const float* v, int32_t T, const double* stats, float* adv, void* stream);

/* One epoch body of PPO.learn (ppo.py:307-349,386): forward of both networks, ratio, clipped
 * surrogate and MSE losses, both backward passes.  grad[NAVPPO_FLAT] is OVERWRITTEN with the
 * gradient of (actor_loss, critic_loss) w.r.t. the flat parameters, every per-sample term
 * divided by n_global (= T on one GPU; the global batch when samples are sharded, so that the
 * all-reduced sum equals the reference's .mean() gradient).  metrics[0..3] are overwritten
 * with this rank's share of the four batch means. */
int navppo_grad(navppo_t* h, const float* params, const float* obs, const float* act, const float* logp_old,
                const float* adv, const float* rtg, int32_t T, int64_t n_global, double var, float* grad,
                double* metrics, void* stream);

/* actor_optim.step(); critic_optim.step() (ppo.py:381,392): torch.optim.Adam, step counter
 * `step` (1-based) shared by both networks.  Also writes the squared gradient norms of the two
 * networks into metrics[4], metrics[5] (the clip_grad_norm_(.., inf) measurements). */
int navppo_adam(navppo_t* h, float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t step,
                double* metrics, void* stream);

/* The whole update of one PPO.learn iteration on one GPU (ppo.py:275-397): evaluate ->
 * advantage -> `epochs` x (navppo_grad, navppo_adam), enqueued back to back with no host
 * round trip.  adv_ws[T], v_ws[T]: caller-provided scratch; metrics[epochs, 8] doubles;
 * step0 = Adam steps taken before this call. */
int navppo_update(navppo_t* h, float* params, float* exp_avg, float* exp_avg_sq, int32_t step0, const float* obs,
                  const float* act, const float* logp_old, const float* rtg, int32_t T, double var, int32_t epochs,
                  float* adv_ws, float* v_ws, double* metrics, void* stream);

/* Diagnostic: one CTA computes D[128, N] = A[128, K] * B[N, K]^T on the tcgen05 tensor cores
 * (kind::tf32) with the operand roles of the fused update kernel (a_mode / b_mode: 0 = rows are
 * the M/N index, 1 = rows are K, 2 = the buffer is a verbatim shared-memory image) and
 * caller-supplied descriptor fields {a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_major,
 * b_major} (host array of 8).  Used by the layout self-test. */
int navppo_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                       const uint32_t* strides6, void* stream);

/* Same with BF16 operands (kind::f16): passes = 1 (bf16 x bf16) or 3 (split x = hi + lo:
 * hi*hi + lo*hi + hi*lo, ~16 mantissa bits) — the two arithmetic modes of the fused kernel. */
int navppo_tc_selftest_bf16(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                            int32_t passes, void* stream);

/* Gradient exchange over NVLink peer memory, fused into the optimiser step.  Every rank owns two
 * flat-gradient buffers (epochs alternate between them) and a flag array of 16 words, allocated so that every
 * rank can address every other rank's copy (CUDA VMM / torch symmetric memory).  grad_ptrs[2 * world]: address, on
 * THIS device, of buffer b of rank r at [b * world + r]; flag_ptrs[world]: address of rank r's flag array;
 * multicast_ptrs[2] (may be NULL): NVLS multicast address of buffer b.  The flag arrays must be zero. */
int navppo_peer_setup(navppo_t* h, int32_t rank, int32_t world, const uint64_t* grad_ptrs, const uint64_t* flag_ptrs,
                      const uint64_t* multicast_ptrs);
/* One epoch's optimiser step on every rank: a cross-GPU barrier over the flag arrays (every rank's navppo_grad has
 * written its own buffer `buffer`; `token` = 1, 2, 3, .. must grow by one per call, the same on every rank, for the
 * lifetime of the flag arrays), then Adam with the gradient = sum over the ranks of their buffers, read straight
 * from peer memory in rank order (bit-identical on every rank), or — use_multicast — with one
 * multimem.ld_reduce per element (the NVSwitch adds).  Replaces ncclAllReduce + navppo_adam; the NCCL path stays as
 * the fallback and the correctness reference. */
int navppo_adam_peer(navppo_t* h, float* params, float* exp_avg, float* exp_avg_sq, int32_t step, int32_t buffer,
                     uint32_t token, int32_t use_multicast, double* metrics, void* stream);

/* Diagnostic: while `device_counters` (32 + 4 * 512 int64 on the device) is non-NULL, split-BF16
 * gradient passes run an instrumented build of the tcgen05 kernel whose CTA (0, 0) writes per-role
 * cycle counters there (epilogue warp 0: [0..7], epilogue warp 4: [8..15], MMA thread: [16..23], flush
 * warp: [24..31]) followed by an event trace of its 41st tile, 512 slots per role (slot 0 = start
 * clock, slot 511 = count, entries = category << 56 | clock; categories in tools/tc_ws_profile.py).
 * NULL switches it off.  Process-wide. */
int navppo_tc_profile(long long* device_counters);

#ifdef __cplusplus
}
#endif
#endif /* NAVPPO_H_ */
