// navsim_kernels.cu — batched LiDAR-navigation simulator for sm_100a + its C-ABI (include/navsim.h).
//
// One launch of navsim_step_kernel is Env.step (project_ppo/src/environment_new.py:272-310)
// for N agents at once, with everything that used to be a ROS round trip to gzserver fused in:
//   cmd_vel -> diff-drive pose integration            (row K, navsim_math.h)
//   ray sensor -> LaserScan                           (row R, navsim_math.h)
//   Env.getOdometry   yaw / rel_theta / diff_angle    (:138-181)
//   Env.getState      collision + arrival flags       (:183-207)
//   observation assembly                              (:289-301)
//   Env.setReward                                     (:209-270)
//   PPO.rollout's episode protocol (auto-reset)       (ppo.py:549-593)
//   Env.reset + goal sampling                         (:312-382)
//
// Execution model: G lanes (1..32, a power of two chosen from N and the map) cooperate on one
// agent; a launch runs `nsteps` consecutive steps with the agent state in registers (agents
// never interact, so no grid-wide synchronisation exists) and writes every step's observation,
// reward and flags - into the same [N, .] arrays, or into the time-major [H, N, .] rollout
// buffers.  See navsim_step_kernel below and DESIGN.md section 4.1.
//
// Data layout in HBM: structure-of-arrays, one array per state field, agent index fastest,
// so a warp's agents read/write one contiguous run per field.  Pose, goal and past_distance
// are fp64 because the reference computes in Python floats and quantises (see navsim_math.h);
// the outputs the policy consumes (obs, reward) are fp32 as ppo.py:616 casts them.  The
// obstacle set (wall segments, beam table, spawn-pose scan: a few hundred bytes to a few KB)
// is staged into shared memory once per CTA with one TMA bulk copy (cp.async.bulk + mbarrier),
// waited on only where the first ray is cast, and read as warp-wide broadcasts.  Observation
// rows are assembled in shared memory and leave as 128-bit stores, one warp's agents at a time.
//
// Compile with -fmad=false: the physics must round exactly like the host build of
// navsim_math.h that drives the reference Env in the oracle.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/navsim.h"
#include "nav_common.h"
#include "navsim_math.h"
#include "navsim_device.cuh"

namespace {

thread_local std::string g_err;

}  // namespace

// shared with navppo_kernels.cu (nav_common.h)
int nav_fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

namespace {

int fail(int code, const std::string& msg) { return nav_fail(code, msg); }

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(NAVSIM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

using namespace navsim_dev;

extern __shared__ __align__(16) unsigned char dyn_smem[];

// Programmatic dependent launch (the rollout chains policy and step kernels): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor is still running; it must
// not touch what the predecessor writes before this wait (which returns once the predecessor has completed and
// its writes are visible), and then lets ITS successor start early.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// Env.step for all agents, `nsteps` consecutive steps per launch.
//   G        lanes per agent (1, 2, 4, .. 32): the host picks it from N so that small batches
//            (BASELINE's 8192 agents = 55 per SM) still fill the machine and each agent's
//            step has a short critical path; G = 1 is the plain thread-per-agent form for
//            batches that fill the machine on their own.  Every lane of a group carries the
//            agent's scalar state redundantly (identical arithmetic, identical bits); the wall
//            sweep is the part that is split.
//   KB       10: the reference's 10-beam sensor; kPadBeams: any B <= 36 (beam sweep).
//   SCRIPTED actions drawn on device (benchmark / rollout driver), else read from io.act.
//   CW       two-phase wall sweep with a compacted visible-wall list (maps with > kCompactWalls walls).
// ----------------------------------------------------------------------------------------
template <int G, int KB, bool SCRIPTED, bool CW>
__global__ void __launch_bounds__(kBlock, KB != NAVSIM_LIDAR_FEATS ? 1 : (G == 1 ? NAVSIM_G1_MINBLOCKS : 4)) navsim_step_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                             const uint16_t* __restrict__ rt_tab, StepIO io,
                                                             DevStats* stats, uint64_t action_seed,
                                                             uint32_t script_step0, int nsteps) {
  constexpr int APB = kBlock / G;  // agents per CTA
  constexpr int APW = 32 / G;      // agents per warp
  // shared: [mbarrier 16 B][map blob, padded to 16 B][obs tile: APB x kObsPad floats]
  //         [visible-wall counters: APB ints][visible-wall lists: APB x S uint16]   (S > kCompactWalls only)
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  const uint32_t map_bytes = map_bytes_of(c.B, c.S);
  float* s_obs = reinterpret_cast<float*>(dyn_smem + 16 + map_bytes);
  int* s_cnt = reinterpret_cast<int*>(s_obs + APB * kObsPad);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_cnt + APB);
  stage_map(s_map, g_map, map_bytes, bar);
  const MapView mv = map_view(s_map, c.B, c.S);
  grid_dependency_wait();   // chained launch (rollout): everything above ran under the previous kernel's tail

  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x & (G - 1);
  const int slot = threadIdx.x / G;
  const int i_raw = blockIdx.x * APB + slot;
  const bool valid = i_raw < c.N;
  const int i = valid ? i_raw : c.N - 1;   // surplus lanes shadow the last agent and store nothing
  const bool writer = valid && g == 0;
  const uint64_t agent = (uint64_t)(c.agent_off + i);
  float* my_obs = s_obs + slot * kObsPad;
  Agent a;
  load_agent(st, i, &a);
  if (io.past_act) {   // Env.step(action, past_action): the caller owns the previous action (environment_new.py:272,299)
    const float2 pv = reinterpret_cast<const float2*>(io.past_act)[i];
    a.pa0 = pv.x; a.pa1 = pv.y;
  }
  bool goal_dirty = false;
  uint32_t script_words[4] = {0u, 0u, 0u, 0u};   // one Philox block = the scripted actions of two steps

  for (int t = 0; t < nsteps; ++t) {
    float a0, a1;
    if (SCRIPTED) {
      const uint32_t sstep = script_step0 + (uint32_t)t;
      if (t == 0 || (sstep & 1u) == 0u) nv_scripted_block(action_seed, agent, sstep, script_words);
      nv_scripted_action(script_words, sstep, &a0, &a1);
    } else {
      const float2 av = reinterpret_cast<const float2*>(io.act)[i];
      a0 = av.x; a1 = av.y;
    }
    StepRows rows;
    {
      const long long vo = (long long)t * io.vec_stride;
      rows.rew = io.rew + vo; rows.done = io.done + vo; rows.arrive = io.arrive + vo;
      rows.trunc = io.trunc ? io.trunc + vo : nullptr;
      rows.ep_ret = io.ep_ret ? io.ep_ret + vo : nullptr;
      rows.ep_path = io.ep_path ? io.ep_path + vo : nullptr;
      rows.ep_len = io.ep_len ? io.ep_len + vo : nullptr;
    }
    agent_step<G, KB, CW>(c, mv, rt_tab, stats, a, a0, a1, g, i, agent, valid, writer, my_obs,
                          CW ? s_list + (size_t)slot * c.S : nullptr, CW ? s_cnt + slot : nullptr, rows,
                          t == 0 ? bar : nullptr, goal_dirty);

    // each warp writes out the rows of its own agents as 128-bit stores
    __syncwarp();
    {
      const int wslot = (threadIdx.x >> 5) * APW;
      const int row0 = blockIdx.x * APB + wslot;
      float* dst_base = io.obs + (long long)t * io.obs_stride;
#pragma unroll
      for (int v = lane; v < APW * (NAVSIM_OBS_DIM / 4); v += 32) {
        const int rr = v >> 2, qq = (v & 3) * 4;
        if (row0 + rr < c.N) {
          const float* src = s_obs + (wslot + rr) * kObsPad + qq;
          reinterpret_cast<float4*>(dst_base + (size_t)(row0 + rr) * NAVSIM_OBS_DIM)[v & 3] =
              make_float4(src[0], src[1], src[2], src[3]);
        }
      }
    }
    __syncwarp();
  }
  if (writer) {
    store_agent(st, i, a, goal_dirty);
    if (io.pose_out) {
      double* po = io.pose_out + (size_t)i * 6;
      po[0] = a.x; po[1] = a.y; po[2] = a.th; po[3] = a.gx; po[4] = a.gy; po[5] = a.past;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats->steps, (unsigned long long)c.N * (unsigned long long)nsteps);
}

// Generic beam count (B > kPadBeams, up to NAVSIM_MAX_BEAMS): a warp per agent, one beam at a
// time; every lane sweeps its walls for that beam and the warp combines the maxima.
__global__ void __launch_bounds__(kBlock) navsim_step_anybeam_kernel(SimConst c, SimState st,
                                                                     const float* __restrict__ g_map,
                                                                     const uint16_t* __restrict__ rt_tab, StepIO io,
                                                                     DevStats* stats, uint64_t action_seed,
                                                                     uint32_t script_step0, int nsteps, int scripted) {
  constexpr int APB = kBlock / 32;
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  const uint32_t map_bytes = map_bytes_of(c.B, c.S);
  float* s_obs = reinterpret_cast<float*>(dyn_smem + 16 + map_bytes);
  stage_map(s_map, g_map, map_bytes, bar);
  const MapView mv = map_view(s_map, c.B, c.S);
  grid_dependency_wait();
  int* s_cnt = reinterpret_cast<int*>(s_obs + APB * kObsPad);
  const int g = threadIdx.x & 31, slot = threadIdx.x >> 5;
  uint16_t* my_list = reinterpret_cast<uint16_t*>(s_cnt + APB) + (size_t)slot * c.S;
  const int i_raw = blockIdx.x * APB + slot;
  const bool valid = i_raw < c.N;
  const int i = valid ? i_raw : c.N - 1;
  const bool writer = valid && g == 0;
  const uint64_t agent = (uint64_t)(c.agent_off + i);
  float* my_obs = s_obs + slot * kObsPad;
  Agent a;
  load_agent(st, i, &a);
  if (io.past_act) {
    const float2 pv = reinterpret_cast<const float2*>(io.past_act)[i];
    a.pa0 = pv.x; a.pa1 = pv.y;
  }
  wait_map(bar);
  bool goal_dirty = false;
  double vl = 0.0, vr = 0.0;             // wheel rim speeds (fidelity option wheel_accel)
  if (c.wheel_accel > 0.0) { vl = st.vl[i]; vr = st.vr[i]; }
  for (int t = 0; t < nsteps; ++t) {
    __syncwarp();
    float a0, a1;
    if (scripted) {
      uint32_t o[4];
      const uint32_t sstep = script_step0 + (uint32_t)t;
      nv_scripted_block(action_seed, agent, sstep, o);
      nv_scripted_action(o, sstep, &a0, &a1);
    } else {
      const float2 av = reinterpret_cast<const float2*>(io.act)[i];
      a0 = av.x; a1 = av.y;
    }
    if (a.steps > 0) a.ep_path += a.last_move;
    const double px = a.x, py = a.y;
    if (c.wheel_accel > 0.0)   // fidelity option: the plugin's 30 Hz updates = 6 sub-steps of the 5 Hz LiDAR period
      nv_drive_ramped(&a.x, &a.y, &a.th, &vl, &vr, (double)a0 / 4.0, (double)a1, c.dt, c.wheel_accel, c.wheel_sep, 6);
    else
      nv_drive(&a.x, &a.y, &a.th, (double)a0 / 4.0, (double)a1, c.dt);
    {
      const double mx = a.x - px, my = a.y - py;
      a.last_move = sqrtf((float)(mx * mx + my * my));
    }
    double s_new, c_new;
    nv_sincos(a.th, &s_new, &c_new);
    const float ox = (float)(a.x + c.off_x * c_new), oy = (float)(a.y + c.off_x * s_new);
    const float ch = (float)c_new, sh = (float)s_new, rmin = (float)c.rmin, rmax = (float)c.rmax;
    float mn = NV_INF_F;
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    int pickn = 0;
    // walls this agent can see at all (same two-phase sweep as group_sweep)
    int nvis = c.S;
    const bool compact = c.S > kCompactWalls;
    if (compact) {
      if (g == 0) s_cnt[slot] = 0;
      __syncwarp();
      for (int k = g; k < c.S; k += 32) {
        float wx, wy, ex, ey, tn;
        if (nv_seg_cull(mv.seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &wx, &wy, &ex, &ey, &tn))
          my_list[atomicAdd(s_cnt + slot, 1)] = (uint16_t)k;
      }
      __syncwarp();
      nvis = s_cnt[slot];
    }
    for (int b = 0; b < c.B; ++b) {
      float dxb, dyb, q = 0.0f;
      nv_beam_dir(ch, sh, mv.bc[b], mv.bs[b], &dxb, &dyb);
      for (int idx = g; idx < nvis; idx += 32) {
        const int k = compact ? (int)my_list[idx] : idx;
        nv_seg_view v;
        if (nv_seg_setup(mv.seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &v)) q = nv_ray_q(&v, dxb, dyb, q);
      }
      __syncwarp();
      q = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(q)));
      float rb = nv_range_from_q(q, rmin, rmax);
      if (c.noise_sigma > 0.f) {           // fidelity option: Gaussian range noise, then the sensor's gates again
        if ((b & 3) == 0) nv_scan_noise4(c.seed ^ 0x6e6f697365ull, agent, a.draws, (uint32_t)a.steps, (uint32_t)(b >> 2), nz);
        rb = nv_noisy_range(rb, nz[b & 3], c.noise_sigma, rmin, rmax);
      }
      if (rb == NV_INF_F) rb = 3.5f;
      mn = (rb < mn) ? rb : mn;
      while (pickn < NAVSIM_LIDAR_FEATS && c.pick[pickn] == b) {
        if (writer) my_obs[pickn] = rb * kInvRmax;
        ++pickn;
      }
    }
    const bool done = (c.collide > (double)mn) && (mn > 0.0f);
    const double ddx = a.gx - a.x, ddy = a.gy - a.y;
    const double d = sqrt(ddx * ddx + ddy * ddy);
    const bool arrive = (d <= c.arrive_thr);
    int yaw, rel, diff;
    odom_features(c, rt_tab, a.x, a.y, a.th, a.gx, a.gy, &yaw, &rel, &diff);
    if (writer) write_goal_feats(c, my_obs, a.pa0, a.pa1, d, yaw, rel, diff);
    double reward = c.r_scale * (a.past - d);
    a.past = d;
    if (done) reward = c.r_collide;
    if (arrive) reward = c.r_arrive;
    a.pa0 = a0; a.pa1 = a1;
    a.steps += 1;
    a.ep_ret += (float)reward;
    const bool timeout = a.steps >= c.max_steps;
    if (writer) {
      const long long vo = (long long)t * io.vec_stride + i;
      io.rew[vo] = (float)reward;
      io.done[vo] = done ? 1 : 0;
      io.arrive[vo] = arrive ? 1 : 0;
      if (io.trunc) io.trunc[vo] = (timeout && !done && !arrive) ? 1 : 0;
    }
    if (c.auto_reset) {
      if (done || arrive || timeout) {
        if (arrive && c.n_starts == 0) sample_goal(c, c.respawn_rects, c.n_respawn_rects, agent, &a);
        if (writer) {
          if (io.ep_ret) {
            const long long vo = (long long)t * io.vec_stride + i;
            io.ep_ret[vo] = a.ep_ret;
            io.ep_path[vo] = a.ep_path;
          }
          if (io.ep_len) io.ep_len[(long long)t * io.vec_stride + i] = a.steps;
          atomicAdd(&stats->episodes, 1ull);
          if (arrive) atomicAdd(&stats->successes, 1ull);
          else if (done) atomicAdd(&stats->collisions, 1ull);
          else atomicAdd(&stats->timeouts, 1ull);
          atomicAdd(&stats->return_sum, (double)a.ep_ret);
          atomicAdd(&stats->length_sum, (double)a.steps);
          atomicAdd(&stats->path_sum, (double)a.ep_path);
        }
        reset_agent(c, mv, rt_tab, agent, &a, my_obs, writer, 1, writer ? 0 : -1);
        vl = 0.0; vr = 0.0;                                          // the plugin's Reset() zeroes the wheels
        goal_dirty = true;
      }
    } else if (arrive) {
      respawn_goal(c, agent, &a);
      const double gx = a.gx - a.x, gy = a.gy - a.y;
      a.past = sqrt(gx * gx + gy * gy);
      goal_dirty = true;
    }
    __syncwarp();
    if (valid && g < NAVSIM_OBS_DIM / 4)
      reinterpret_cast<float4*>(io.obs + (long long)t * io.obs_stride + (size_t)i * NAVSIM_OBS_DIM)[g] =
          make_float4(my_obs[4 * g], my_obs[4 * g + 1], my_obs[4 * g + 2], my_obs[4 * g + 3]);
    __syncwarp();
  }
  if (writer) {
    store_agent(st, i, a, goal_dirty);
    if (c.wheel_accel > 0.0) { st.vl[i] = vl; st.vr[i] = vr; }
    if (io.pose_out) {
      double* po = io.pose_out + (size_t)i * 6;
      po[0] = a.x; po[1] = a.y; po[2] = a.th; po[3] = a.gx; po[4] = a.gy; po[5] = a.past;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats->steps, (unsigned long long)c.N * (unsigned long long)nsteps);
}

// Env.reset for the masked agents (thread per agent).
__global__ void __launch_bounds__(kBlock) navsim_reset_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                              const uint16_t* __restrict__ rt_tab,
                                                              const uint8_t* __restrict__ mask, float* __restrict__ obs,
                                                              double* __restrict__ pose_out) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  const uint32_t map_bytes = map_bytes_of(c.B, c.S);
  float* s_obs = reinterpret_cast<float*>(dyn_smem + 16 + map_bytes);
  stage_map(s_map, g_map, map_bytes, bar);
  wait_map(bar);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  if (mask && !mask[i]) return;
  float* my_obs = s_obs + threadIdx.x * kObsPad;
  Agent a;
  load_agent(st, i, &a);
  reset_agent(c, map_view(s_map, c.B, c.S), rt_tab, (uint64_t)(c.agent_off + i), &a, my_obs, true, 1, 0);
  store_agent(st, i, a, true);
  if (c.wheel_accel > 0.0) { st.vl[i] = 0.0; st.vr[i] = 0.0; }   // the diff-drive plugin's Reset() zeroes the wheels
  if (pose_out) {
    double* po = pose_out + (size_t)i * 6;
    po[0] = a.x; po[1] = a.y; po[2] = a.th; po[3] = a.gx; po[4] = a.gy; po[5] = a.past;
  }
  if (obs) {
    float4* dst = reinterpret_cast<float4*>(obs + (size_t)i * NAVSIM_OBS_DIM);
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(my_obs[4 * q], my_obs[4 * q + 1], my_obs[4 * q + 2], my_obs[4 * q + 3]);
  }
}

// LaserScan only (parity tests of row R): ranges[N, B] doubles with the +-inf gates.
__global__ void __launch_bounds__(kBlock) navsim_scan_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                             double* __restrict__ ranges) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  stage_map(s_map, g_map, map_bytes_of(c.B, c.S), bar);
  wait_map(bar);
  const MapView mv = map_view(s_map, c.B, c.S);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  double s, co;
  nv_sincos(st.th[i], &s, &co);
  const float ox = (float)(st.x[i] + c.off_x * co), oy = (float)(st.y[i] + c.off_x * s);
  const float ch = (float)co, sh = (float)s;
  for (int b = 0; b < c.B; ++b) {
    float dx, dy;
    nv_beam_dir(ch, sh, mv.bc[b], mv.bs[b], &dx, &dy);
    ranges[(size_t)i * c.B + b] = (double)nv_range_from_q(nv_beam_q(ox, oy, dx, dy, mv.seg, c.S, c.closed_boxes),
                                                          (float)c.rmin, (float)c.rmax);
  }
}

}  // namespace

// ========================================================================================
// C-ABI
// ========================================================================================
struct navsim {
  navsim_cfg cfg;
  SimConst c;
  SimState st;
  float* d_map = nullptr;        // [8S + 3B] floats: packed wall records, beam cos, beam sin, spawn-pose ranges
  uint16_t* d_rt = nullptr;      // bearing table, (2 rt_R + 1)^2 entries (hundredths of a degree)
  int32_t S = 0;
  DevStats* d_stats = nullptr;
  cudaStream_t own_stream = nullptr;
  // pinned host staging for the *_host entry points
  float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr;
  uint8_t* h_flags = nullptr;    // [3N] done, arrive, trunc  (also the reset mask)
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr;
  uint8_t* d_flags = nullptr;
  int64_t launches = 0;
  uint32_t script_step = 0;
  int lanes = 1;                 // lanes per agent of the step kernel (G)
  const void* pin_seen[2] = {nullptr, nullptr};  // last caller buffers of navsim_step_host (actions, obs)
  const float* pin_alias[2] = {nullptr, nullptr};
  float* h_rew_dev = nullptr;    // device alias of the pinned h_rew block (reward + flags)
  // navsim_*_host_ex: mapped pinned blocks for the caller-owned previous action and the pose read-back
  float *h_past = nullptr, *h_past_dev = nullptr;        // [N,2]
  double *h_pose = nullptr, *h_pose_dev = nullptr;       // [N,6]
  double* reset_pose_out = nullptr;                      // set around a navsim_reset_host_ex launch
  // Mixing device entry points (caller's stream) with host entry points (own_stream): the host side waits for the
  // last caller stream before it touches the agent state; host calls end synchronised (or, for the asynchronous
  // form, leave `async_tail` for the next device call to wait on).
  cudaStream_t last_dev_stream = nullptr;
  bool dev_dirty = false;
  // navsim_step_host_async: pipeline of up to kAsyncDepth steps (kernel on own_stream, observation DMA on copy_stream)
  cudaStream_t copy_stream = nullptr;
  float* d_obs2[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_kernel[4] = {nullptr, nullptr, nullptr, nullptr}, ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t async_issued = 0, async_waited = 0;
  // (one instantiated CUDA graph per slot {actions in -> kernel -> results home}, launched alternately into two streams,
  // was measured too: 20.9 us per step of 8192 robots against 17.5 for the plain form below — graph launches add
  // device-side latency to a chain this short — and is not kept)
  int async_mode = 4;            // NAVSIM_ASYNC_OBS: 4 "dma_ahead" (default) the actions are staged by a host-to-device
                                 // copy on their own stream (it runs under the previous step's kernel) and the step's
                                 // results leave in one copy under the next step's kernel; 1 "dma" the kernel reads
                                 // the actions over PCIe instead; 0 "dma_act" the action copy sits in the kernel's
                                 // stream; 2 "stores" the kernel reads / writes every host buffer itself.  Measured per
                                 // step of 8192 robots (tools/time_async.py): 15.7 / 18.3 / 24.2 / 22.6 us.  In the
                                 // default form the host spends ~10 us issuing and ~5 us waiting per step; the link
                                 // would allow 13.7 us (tools/pcie_d2h.py)
  cudaStream_t h2d_stream = nullptr;                          // mode 4: actions staged one step ahead
  cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr};
  float* d_act2[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<std::pair<const void*, void*>> alias_cache;   // host pointer -> device alias (null: not page-locked)
  // GoalSpawnSampler tables (navsim_set_sampler) and the host copy of the packed map they are cast against
  double *d_starts = nullptr, *d_goals = nullptr;
  float* d_start_scans = nullptr;
  std::vector<float> h_map;      // [8S + 3B] as uploaded
  int closed = 0;
};

namespace {

// variant: 0 = 10-beam register path, 1 = padded (B <= kPadBeams), 2 = any beam count (warp per agent)
// (also the only kernel that implements the fidelity options: range noise, wheel-acceleration ramp)
int variant_of(const navsim* h) {
  if (h->c.noise_sigma > 0.f || h->c.wheel_accel > 0.0) return 2;
  return h->c.B == NAVSIM_LIDAR_FEATS ? 0 : (h->c.B <= kPadBeams ? 1 : 2);
}

// Lanes per agent in force: the request / heuristic, raised for big maps so that the per-agent
// visible-wall lists of a CTA (agents x S x 2 bytes) stay within ~64 KB of shared memory.
int lanes_of(const navsim* h) {
  if (variant_of(h) == 2) return 32;
  int g = h->lanes;
  // measured (tools/lane_sweep.py, house map, 4096 agents, 10 and 36 beams): with hundreds of walls
  // the cull phase dominates and one more doubling pays
  if (h->cfg.lanes_per_agent <= 0 && h->c.S > 64 && g < 32 && (long long)h->c.N * g <= 32768) g *= 2;
  while (g < 32 && (size_t)(kBlock / g) * (size_t)h->c.S * 2 > 65536) g *= 2;
  return g;
}

size_t step_smem_bytes(const navsim* h) {
  const size_t apb = (size_t)(kBlock / lanes_of(h));
  size_t b = 16 + (size_t)map_bytes_of(h->c.B, h->c.S) + apb * kObsPad * sizeof(float);
  if (h->c.S > kCompactWalls) b += apb * sizeof(int) + apb * (size_t)h->c.S * sizeof(uint16_t);
  return b;
}

// reset / scan kernels: thread per agent
size_t aux_smem_bytes(const navsim* h) {
  return 16 + (size_t)map_bytes_of(h->c.B, h->c.S) + (size_t)kBlock * kObsPad * sizeof(float);
}

int aux_grid_of(const navsim* h) { return (h->c.N + kBlock - 1) / kBlock; }

// Lanes per agent when the caller does not say.
int pick_lanes(int n_agents, int requested) {
  if (requested > 0) {
    int g = 1;
    while (g * 2 <= requested && g < 32) g *= 2;
    return g;
  }
  const char* env = getenv("NAVSIM_LANES");
  if (env && atoi(env) > 0) return pick_lanes(n_agents, atoi(env));
  // measured on B200 (tools/lane_sweep.py, stage maps): the fused step is fastest with about
  // 220 lanes per SM in flight -- N = 8192 -> 4 lanes, 16384 -> 2, >= 32768 -> 1
  int g = 32;
  while (g > 1 && (long long)n_agents * g > 32768LL) g /= 2;
  return g;
}

// Device-side alias of a caller's host buffer when it is page-locked and mapped (else null).
// One driver query per new pointer; a training loop passes the same buffers every step.
const float* host_device_alias(navsim* h, int slot, const void* p) {
  if (h->pin_seen[slot] == p) return h->pin_alias[slot];
  cudaPointerAttributes at;
  const float* alias = nullptr;
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess) {
    if (at.type == cudaMemoryTypeHost && at.devicePointer) alias = static_cast<const float*>(at.devicePointer);
  } else {
    cudaGetLastError();
  }
  h->pin_seen[slot] = p;
  h->pin_alias[slot] = alias;
  return alias;
}

int check_ready(const navsim* h) {
  if (!h) return fail(NAVSIM_EINVAL, "null handle");
  if (!h->d_map) return fail(NAVSIM_EINVAL, "navsim_set_map has not been called");
  if (h->cfg.sampler_mode == 1 && h->c.n_starts == 0)
    return fail(NAVSIM_EINVAL, "sampler_mode = 1 but navsim_set_sampler has not been called");
  return NAVSIM_OK;
}

// Device entry point on the caller's stream: remember the stream (the host entry points wait for it) and wait for
// asynchronous host steps still in flight on the handle's own streams.
int begin_device_call(navsim* h, cudaStream_t s) {
  if (h->async_issued > h->async_waited) {
    const int k = (int)((h->async_issued - 1) & 3);
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_kernel[k], 0));    // the agent state is final once the last kernel ran
  }
  h->last_dev_stream = s;
  h->dev_dirty = true;
  return NAVSIM_OK;
}
// Host entry point: device work enqueued on the caller's stream since the last host call must have finished.
int begin_host_call(navsim* h) {
  if (h->dev_dirty) {
    CUDA_TRY(cudaStreamSynchronize(h->last_dev_stream));
    h->dev_dirty = false;
  }
  return NAVSIM_OK;
}

// The per-kernel dynamic shared-memory limit is process-wide: only ever raise it, so that a handle with a
// small map does not lower the limit an earlier handle with a big map still needs.
int raise_smem_limit(const void* func, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> seen;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& kv : seen)
    if (kv.first == func) {
      if (kv.second >= bytes) return NAVSIM_OK;
      CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      kv.second = bytes;
      return NAVSIM_OK;
    }
  CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  seen.emplace_back(func, bytes);
  return NAVSIM_OK;
}

typedef void (*step_kernel_t)(SimConst, SimState, const float*, const uint16_t*, StepIO, DevStats*, uint64_t, uint32_t,
                              int);

template <int KB, bool SCR, bool CW>
step_kernel_t step_kernel_for_lanes(int g) {
  switch (g) {
    case 1: return navsim_step_kernel<1, KB, SCR, CW>;
    case 2: return navsim_step_kernel<2, KB, SCR, CW>;
    case 4: return navsim_step_kernel<4, KB, SCR, CW>;
    case 8: return navsim_step_kernel<8, KB, SCR, CW>;
    case 16: return navsim_step_kernel<16, KB, SCR, CW>;
    default: return navsim_step_kernel<32, KB, SCR, CW>;
  }
}

template <int KB, bool SCR>
step_kernel_t step_kernel_for_map(const navsim* h) {
  const int g = lanes_of(h);
  return h->c.S > kCompactWalls ? step_kernel_for_lanes<KB, SCR, true>(g) : step_kernel_for_lanes<KB, SCR, false>(g);
}

step_kernel_t step_kernel_of(const navsim* h, bool scripted) {
  if (variant_of(h) == 0)
    return scripted ? step_kernel_for_map<NAVSIM_LIDAR_FEATS, true>(h) : step_kernel_for_map<NAVSIM_LIDAR_FEATS, false>(h);
  return scripted ? step_kernel_for_map<kPadBeams, true>(h) : step_kernel_for_map<kPadBeams, false>(h);
}

// One launch = `nsteps` consecutive Env.step calls for every agent.
int launch_step(navsim* h, const StepIO& io, cudaStream_t s, bool scripted, uint64_t action_seed, int nsteps,
                bool chained = false) {
  if (nsteps < 1) return NAVSIM_OK;
  const size_t smem = step_smem_bytes(h);
  const int g = lanes_of(h);
  const int apb = kBlock / g;
  const dim3 grid((h->c.N + apb - 1) / apb), block(kBlock);
  const uint32_t script_step = scripted ? h->script_step : 0u;
  if (scripted) h->script_step += (uint32_t)nsteps;
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = s;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  if (chained) { lc.attrs = &attr; lc.numAttrs = 1; }
  if (variant_of(h) == 2) {
    CUDA_TRY(cudaLaunchKernelEx(&lc, navsim_step_anybeam_kernel, h->c, h->st, (const float*)h->d_map, (const uint16_t*)h->d_rt, io,
                                h->d_stats, action_seed, script_step, nsteps, scripted ? 1 : 0));
  } else {
    CUDA_TRY(cudaLaunchKernelEx(&lc, step_kernel_of(h, scripted), h->c, h->st, (const float*)h->d_map, (const uint16_t*)h->d_rt, io,
                                h->d_stats, action_seed, script_step, nsteps));
  }
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

StepIO make_io(const float* act, float* obs, float* rew, uint8_t* done, uint8_t* arrive, uint8_t* trunc,
               long long obs_stride, long long vec_stride) {
  StepIO io;
  io.act = act; io.obs = obs; io.rew = rew; io.done = done; io.arrive = arrive; io.trunc = trunc;
  io.ep_ret = nullptr; io.ep_path = nullptr; io.ep_len = nullptr; io.past_act = nullptr; io.pose_out = nullptr;
  io.obs_stride = obs_stride; io.vec_stride = vec_stride;
  return io;
}

}  // namespace

extern "C" {

const char* nav_last_error(void) { return g_err.c_str(); }

int navsim_abi_version(void) { return 2; }

int navsim_default_cfg(navsim_cfg* cfg, int32_t num_agents) {
  if (!cfg) return fail(NAVSIM_EINVAL, "cfg is null");
  memset(cfg, 0, sizeof *cfg);
  cfg->num_agents = num_agents;
  cfg->num_beams = 10;                 // gazebo.xacro:111
  cfg->max_episode_steps = 500;        // arguments.py:29
  cfg->auto_reset = 1;
  cfg->device = 0;
  cfg->seed = 0;
  cfg->agent_id_offset = 0;
  cfg->dt = 0.2;                       // 5 Hz, gazebo.xacro:107
  cfg->lidar_offset_x = -0.032;        // urdf.xacro:134-138
  cfg->lidar_min = 0.12;               // gazebo.xacro:118
  cfg->lidar_max = 3.5;                // gazebo.xacro:119
  cfg->fov_min = -1.5707975;           // gazebo.xacro:113
  cfg->fov_max = 1.5707975;            // gazebo.xacro:114
  cfg->collision_range = 0.2;          // environment_new.py:188
  cfg->arrive_threshold = 0.2;         // :45
  cfg->reward_scale = 500.0;           // :213
  cfg->reward_collide = -100.0;        // :217
  cfg->reward_arrive = 120.0;          // :221
  cfg->diag_norm = sqrt(2.0) * (3.8 + 3.8);  // :21
  cfg->goal_lo = -3.6;                 // :337
  cfg->goal_hi = 3.6;
  cfg->start_x = cfg->start_y = cfg->start_theta = 0.0;  // turtlebot3_stage_1.launch:3-5
  const double reset_r[16] = {1.7, 2.3, -1.2, 1.2, -2.3, -1.7, -1.2, 1.2,      // :340-341
                              -1.2, 1.2, 1.7, 2.3, -1.2, 1.2, -2.3, -1.7};     // :342-343
  const double respawn_r[16] = {1.6, 2.4, -1.4, 1.4, -2.4, -1.6, -1.4, 1.4,    // :248-249
                                -1.4, 1.4, 1.6, 2.4, -1.4, 1.4, -2.4, -1.6};   // :250-251
  memcpy(cfg->reset_rects, reset_r, sizeof reset_r);
  memcpy(cfg->respawn_rects, respawn_r, sizeof respawn_r);
  cfg->n_reset_rects = 4;
  cfg->n_respawn_rects = 4;
  cfg->lidar_noise_sigma = 0.0;        // gazebo.xacro:125 has 0.01; off by default (fidelity option)
  cfg->wheel_accel = 0.0;              // gazebo.xacro:67 has 1; off by default (fidelity option)
  cfg->wheel_separation = 0.160;       // gazebo.xacro:65, turtlebot3_fake.cpp:44
  cfg->sampler_min_dist = 1.5;         // spawn_goal_sampler.py:38
  cfg->sampler_max_dist = 6.0;
  cfg->sampler_mode = 0;
  return NAVSIM_OK;
}

int navsim_create(navsim_t** out, const navsim_cfg* cfg) {
  if (!out || !cfg) return fail(NAVSIM_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->num_agents <= 0) return fail(NAVSIM_EINVAL, "num_agents must be positive");
  if (cfg->num_beams < 1 || cfg->num_beams > NAVSIM_MAX_BEAMS) return fail(NAVSIM_EINVAL, "num_beams out of range");
  if (cfg->n_reset_rects < 0 || cfg->n_reset_rects > NAVSIM_MAX_RECTS || cfg->n_respawn_rects < 0 ||
      cfg->n_respawn_rects > NAVSIM_MAX_RECTS)
    return fail(NAVSIM_EINVAL, "too many rejection rectangles");
  if (!(cfg->goal_hi > cfg->goal_lo)) return fail(NAVSIM_EINVAL, "goal_hi must exceed goal_lo");
  if (cfg->lidar_noise_sigma < 0.0 || cfg->wheel_accel < 0.0) return fail(NAVSIM_EINVAL, "negative fidelity option");
  if (cfg->wheel_accel > 0.0 && !(cfg->wheel_separation > 0.0)) return fail(NAVSIM_EINVAL, "wheel_separation must be positive");
  if (cfg->sampler_mode != 0 && cfg->sampler_mode != 1) return fail(NAVSIM_EINVAL, "sampler_mode is 0 or 1");
  if (cfg->sampler_mode == 1 && !(cfg->sampler_max_dist >= cfg->sampler_min_dist))
    return fail(NAVSIM_EINVAL, "sampler_max_dist must not be below sampler_min_dist");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(NAVSIM_ENODEV, "no CUDA device: the simulator has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(NAVSIM_EINVAL, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(cfg->device));
  navsim* h = new (std::nothrow) navsim();
  if (!h) return fail(NAVSIM_ENOMEM, "host allocation failed");
  h->cfg = *cfg;
  SimConst& c = h->c;
  c.N = cfg->num_agents; c.B = cfg->num_beams; c.S = 0; c.max_steps = cfg->max_episode_steps;
  c.auto_reset = cfg->auto_reset; c.n_reset_rects = cfg->n_reset_rects; c.n_respawn_rects = cfg->n_respawn_rects;
  c.seed = cfg->seed; c.agent_off = cfg->agent_id_offset;
  c.dt = cfg->dt; c.off_x = cfg->lidar_offset_x; c.rmin = cfg->lidar_min; c.rmax = cfg->lidar_max;
  c.collide = cfg->collision_range; c.arrive_thr = cfg->arrive_threshold;
  c.r_scale = cfg->reward_scale; c.r_collide = cfg->reward_collide; c.r_arrive = cfg->reward_arrive;
  c.diag = cfg->diag_norm; c.goal_lo = cfg->goal_lo; c.goal_hi = cfg->goal_hi;
  c.sx = cfg->start_x; c.sy = cfg->start_y; c.sth = cfg->start_theta;
  memcpy(c.reset_rects, cfg->reset_rects, sizeof c.reset_rects);
  memcpy(c.respawn_rects, cfg->respawn_rects, sizeof c.respawn_rects);
  c.inv_diag = (float)(1.0 / cfg->diag_norm);
  c.n_starts = 0; c.n_goals = 0; c.starts = nullptr; c.goals = nullptr; c.start_scans = nullptr;
  c.smin = cfg->sampler_min_dist; c.smax = cfg->sampler_max_dist;
  c.noise_sigma = (float)cfg->lidar_noise_sigma; c.wheel_accel = cfg->wheel_accel; c.wheel_sep = cfg->wheel_separation;
  for (int i = 0; i < NAVSIM_LIDAR_FEATS; ++i) c.pick[i] = (int)((double)(i * cfg->num_beams) / 10.0);  // :293
  c.rt_R = 0; c.rt_W = 1;
  h->lanes = pick_lanes(c.N, cfg->lanes_per_agent);
  const size_t N = (size_t)c.N;
  // one slab for the SoA state: 6 doubles, 5 floats, 1 int32, 1 uint32 per agent
  double* dslab = nullptr;
  float* fslab = nullptr;
  auto cleanup = [&]() { navsim_destroy(h); };
#define TRY_OR_CLEAN(expr)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      cleanup();                                                                                   \
      return fail(e__ == cudaErrorMemoryAllocation ? NAVSIM_ENOMEM : NAVSIM_ECUDA,                 \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                            \
    }                                                                                              \
  } while (0)
  TRY_OR_CLEAN(cudaMalloc(&dslab, N * 8 * sizeof(double)));
  h->st.x = dslab; h->st.y = dslab + N; h->st.th = dslab + 2 * N;
  h->st.gx = dslab + 3 * N; h->st.gy = dslab + 4 * N; h->st.past = dslab + 5 * N;
  h->st.vl = dslab + 6 * N; h->st.vr = dslab + 7 * N;
  TRY_OR_CLEAN(cudaMalloc(&fslab, N * 7 * sizeof(float)));
  h->st.pa0 = fslab; h->st.pa1 = fslab + N; h->st.ep_ret = fslab + 2 * N; h->st.ep_path = fslab + 3 * N;
  h->st.last_move = fslab + 4 * N;
  h->st.steps = reinterpret_cast<int32_t*>(fslab + 5 * N);
  h->st.draws = reinterpret_cast<uint32_t*>(fslab + 6 * N);
  TRY_OR_CLEAN(cudaMemset(dslab, 0, N * 8 * sizeof(double)));
  TRY_OR_CLEAN(cudaMemset(fslab, 0, N * 7 * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_stats, sizeof(DevStats)));
  TRY_OR_CLEAN(cudaMemset(h->d_stats, 0, sizeof(DevStats)));
  TRY_OR_CLEAN(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_act, N * 2 * sizeof(float)));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float)));
  // reward + the three flag arrays share one block (one D2H copy): [N floats][3N bytes]
  TRY_OR_CLEAN(cudaMallocHost(&h->h_rew, N * sizeof(float) + N * 3));
  h->h_flags = reinterpret_cast<uint8_t*>(h->h_rew + N);
  {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, h->h_rew, 0) == cudaSuccess) h->h_rew_dev = static_cast<float*>(dp);
    else cudaGetLastError();
  }
  TRY_OR_CLEAN(cudaMallocHost(&h->h_past, N * 2 * sizeof(float)));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_pose, N * 6 * sizeof(double)));
  {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, h->h_past, 0) == cudaSuccess) h->h_past_dev = static_cast<float*>(dp); else cudaGetLastError();
    if (cudaHostGetDevicePointer(&dp, h->h_pose, 0) == cudaSuccess) h->h_pose_dev = static_cast<double*>(dp); else cudaGetLastError();
  }
  TRY_OR_CLEAN(cudaMalloc(&h->d_act, N * 2 * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_obs, N * NAVSIM_OBS_DIM * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_rew, N * sizeof(float) + N * 3));
  h->d_flags = reinterpret_cast<uint8_t*>(h->d_rew + N);
#undef TRY_OR_CLEAN
  *out = h;
  return NAVSIM_OK;
}

int navsim_destroy(navsim_t* h) {
  if (!h) return NAVSIM_OK;
  cudaSetDevice(h->cfg.device);
  if (h->st.x) cudaFree(h->st.x);
  if (h->st.pa0) cudaFree(h->st.pa0);
  if (h->d_map) cudaFree(h->d_map);
  if (h->d_rt) cudaFree(h->d_rt);
  if (h->d_starts) cudaFree(h->d_starts);
  if (h->d_goals) cudaFree(h->d_goals);
  if (h->d_start_scans) cudaFree(h->d_start_scans);
  if (h->d_stats) cudaFree(h->d_stats);
  if (h->h_act) cudaFreeHost(h->h_act);
  if (h->h_obs) cudaFreeHost(h->h_obs);
  if (h->h_rew) cudaFreeHost(h->h_rew);
  if (h->h_past) cudaFreeHost(h->h_past);
  if (h->h_pose) cudaFreeHost(h->h_pose);
  if (h->d_act) cudaFree(h->d_act);
  if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_rew) cudaFree(h->d_rew);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  for (int k = 0; k < 4; ++k) {
    if (h->d_obs2[k]) cudaFree(h->d_obs2[k]);
    if (h->d_act2[k]) cudaFree(h->d_act2[k]);
    if (h->ev_h2d[k]) cudaEventDestroy(h->ev_h2d[k]);
  }
  for (int k = 0; k < 4; ++k) {
    if (h->ev_kernel[k]) cudaEventDestroy(h->ev_kernel[k]);
    if (h->ev_copy[k]) cudaEventDestroy(h->ev_copy[k]);
  }
  delete h;
  return NAVSIM_OK;
}

int navsim_set_map(navsim_t* h, const double* seg_host, int32_t num_segments, int32_t flags) {
  if (!h || !seg_host) return fail(NAVSIM_EINVAL, "null argument");
  if (int rc = navsim_wait(h, 0)) return rc;
  if (num_segments < 1) return fail(NAVSIM_EINVAL, "a map needs at least one segment");
  const int B = h->c.B;
  const size_t bytes = map_bytes_of(B, num_segments);
  if (16 + bytes + (size_t)kBlock * kObsPad * sizeof(float) > 200 * 1024)
    return fail(NAVSIM_EINVAL, "map does not fit in shared memory");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const int closed = (flags & NAVSIM_MAP_CLOSED_BOXES) ? 1 : 0;
  float* host = new (std::nothrow) float[bytes / sizeof(float)]();
  if (!host) return fail(NAVSIM_ENOMEM, "host allocation failed");
  float* h_bc = host + NV_SEG_FLOATS * num_segments;
  float* h_bs = h_bc + B;
  float* h_start = h_bs + B;
  double xmax = 0.0;  // map half-extent, for the bearing table
  for (int k = 0; k < num_segments; ++k) {
    nv_pack_segment(seg_host + 4 * k, h->cfg.lidar_max, host + NV_SEG_FLOATS * k);
    for (int j = 0; j < 4; ++j) xmax = fmax(xmax, fabs(seg_host[4 * k + j]));
  }
  // beam direction table, gazebo.xacro:111-114: B samples over [fov_min, fov_max] inclusive
  for (int i = 0; i < B; ++i) {
    const double a = (B > 1) ? h->cfg.fov_min + (double)i * ((h->cfg.fov_max - h->cfg.fov_min) / (double)(B - 1))
                             : 0.5 * (h->cfg.fov_min + h->cfg.fov_max);
    double sn, cs;
    nv_sincos(a, &sn, &cs);
    h_bc[i] = (float)cs;
    h_bs[i] = (float)sn;
  }
  // LaserScan of the spawn pose (Env.reset always puts the robot there): the host build of the
  // physics header rounds exactly like the device build, so these are the ranges the kernel
  // would cast, sanitised as getState does (:193-194).
  {
    double sn, cs;
    nv_sincos(h->cfg.start_theta, &sn, &cs);
    const float ox = (float)(h->cfg.start_x + h->cfg.lidar_offset_x * cs), oy = (float)(h->cfg.start_y + h->cfg.lidar_offset_x * sn);
    for (int i = 0; i < B; ++i) {
      float dx, dy;
      nv_beam_dir((float)cs, (float)sn, h_bc[i], h_bs[i], &dx, &dy);
      const float r = nv_range_from_q(nv_beam_q(ox, oy, dx, dy, host, num_segments, closed), (float)h->cfg.lidar_min,
                                      (float)h->cfg.lidar_max);
      h_start[i] = (r == NV_INF_F) ? 3.5f : r;
    }
  }
  if (h->d_map) { cudaFree(h->d_map); h->d_map = nullptr; }
  cudaError_t e = cudaMalloc(&h->d_map, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_map, host, bytes, cudaMemcpyHostToDevice);
  h->h_map.assign(host, host + bytes / sizeof(float));
  h->closed = closed;
  delete[] host;
  if (e != cudaSuccess) return fail(NAVSIM_ECUDA, std::string("set_map: ") + cudaGetErrorString(e));
  // bearing table over every goal offset the map allows (capped; larger offsets are computed)
  {
    const double reach = fmax(fabs(h->cfg.goal_lo), fabs(h->cfg.goal_hi)) + fmax(xmax, fmax(fabs(h->cfg.start_x), fabs(h->cfg.start_y)));
    int R = (int)ceil(reach * 10.0) + 2;
    if (R > 256) R = 256;
    if (R < 8) R = 8;
    const int W = 2 * R + 1;
    uint16_t* tab = new (std::nothrow) uint16_t[(size_t)W * W];
    if (!tab) return fail(NAVSIM_ENOMEM, "host allocation failed");
    for (int ny = -R; ny <= R; ++ny)
      for (int nx = -R; nx <= R; ++nx) tab[(size_t)(ny + R) * W + (nx + R)] = (uint16_t)nv_rel_theta_centideg(nx, ny);
    if (h->d_rt) { cudaFree(h->d_rt); h->d_rt = nullptr; }
    e = cudaMalloc(&h->d_rt, (size_t)W * W * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_rt, tab, (size_t)W * W * sizeof(uint16_t), cudaMemcpyHostToDevice);
    delete[] tab;
    if (e != cudaSuccess) return fail(NAVSIM_ECUDA, std::string("set_map: ") + cudaGetErrorString(e));
    h->c.rt_R = R;
    h->c.rt_W = W;
  }
  h->S = num_segments;
  h->c.S = num_segments;
  h->c.closed_boxes = closed;
  if (step_smem_bytes(h) > 200 * 1024 || aux_smem_bytes(h) > 200 * 1024) {
    cudaFree(h->d_map);
    h->d_map = nullptr;
    return fail(NAVSIM_EINVAL, "map does not fit in shared memory");
  }
  const int smem = (int)step_smem_bytes(h);
  if (int rc = raise_smem_limit((const void*)step_kernel_of(h, false), smem)) return rc;
  if (int rc = raise_smem_limit((const void*)step_kernel_of(h, true), smem)) return rc;
  if (int rc = raise_smem_limit((const void*)navsim_step_anybeam_kernel, smem)) return rc;
  const int aux = (int)aux_smem_bytes(h);
  if (int rc = raise_smem_limit((const void*)navsim_reset_kernel, aux)) return rc;
  if (int rc = raise_smem_limit((const void*)navsim_scan_kernel, aux)) return rc;
  return NAVSIM_OK;
}

int navsim_set_sampler(navsim_t* h, const double* starts_host, int32_t n_starts, const double* goals_host, int32_t n_goals) {
  if (!h || !starts_host || !goals_host) return fail(NAVSIM_EINVAL, "null argument");
  if (!h->d_map) return fail(NAVSIM_EINVAL, "navsim_set_map has not been called");
  if (h->cfg.sampler_mode != 1) return fail(NAVSIM_EINVAL, "the handle was not created with sampler_mode = 1");
  if (int rc = navsim_wait(h, 0)) return rc;
  if (n_starts < 1 || n_starts > NAVSIM_MAX_TABLE || n_goals < 1 || n_goals > NAVSIM_MAX_TABLE)
    return fail(NAVSIM_EINVAL, "table sizes must be in 1..NAVSIM_MAX_TABLE");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  // LaserScan of every start pose, cast with the host build of the physics header against the uploaded map
  const int B = h->c.B, S = h->S;
  const float* seg = h->h_map.data();
  const float *h_bc = seg + NV_SEG_FLOATS * S, *h_bs = h_bc + B;
  std::vector<float> scans((size_t)n_starts * B);
  for (int k = 0; k < n_starts; ++k) {
    double sn, cs;
    nv_sincos(starts_host[3 * k + 2], &sn, &cs);
    const float ox = (float)(starts_host[3 * k] + h->cfg.lidar_offset_x * cs), oy = (float)(starts_host[3 * k + 1] + h->cfg.lidar_offset_x * sn);
    for (int i = 0; i < B; ++i) {
      float dx, dy;
      nv_beam_dir((float)cs, (float)sn, h_bc[i], h_bs[i], &dx, &dy);
      const float r = nv_range_from_q(nv_beam_q(ox, oy, dx, dy, seg, S, h->closed), (float)h->cfg.lidar_min, (float)h->cfg.lidar_max);
      scans[(size_t)k * B + i] = (r == NV_INF_F) ? 3.5f : r;
    }
  }
  if (h->d_starts) { cudaFree(h->d_starts); h->d_starts = nullptr; }
  if (h->d_goals) { cudaFree(h->d_goals); h->d_goals = nullptr; }
  if (h->d_start_scans) { cudaFree(h->d_start_scans); h->d_start_scans = nullptr; }
  h->c.n_starts = 0;
  CUDA_TRY(cudaMalloc(&h->d_starts, (size_t)n_starts * 3 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_goals, (size_t)n_goals * 2 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_start_scans, scans.size() * sizeof(float)));
  CUDA_TRY(cudaMemcpy(h->d_starts, starts_host, (size_t)n_starts * 3 * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_goals, goals_host, (size_t)n_goals * 2 * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_start_scans, scans.data(), scans.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->c.starts = h->d_starts; h->c.goals = h->d_goals; h->c.start_scans = h->d_start_scans;
  h->c.n_starts = n_starts; h->c.n_goals = n_goals;
  return NAVSIM_OK;
}

int navsim_reset(navsim_t* h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if ((cudaStream_t)stream != h->own_stream)
    if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  navsim_reset_kernel<<<aux_grid_of(h), kBlock, aux_smem_bytes(h), (cudaStream_t)stream>>>(h->c, h->st, h->d_map, h->d_rt,
                                                                                         mask_dev, obs_dev, h->reset_pose_out);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navsim_step(navsim_t* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev,
                uint8_t* arrive_dev, uint8_t* trunc_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  if (!act_dev || !obs_dev || !rew_dev || !done_dev || !arrive_dev) return fail(NAVSIM_EINVAL, "null buffer");
  return launch_step(h, make_io(act_dev, obs_dev, rew_dev, done_dev, arrive_dev, trunc_dev, 0, 0), (cudaStream_t)stream,
                     false, 0, 1);
}

int navsim_step_ex(navsim_t* h, const float* act_dev, const navsim_step_out* out, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  if (!act_dev || !out || !out->obs || !out->rew || !out->done || !out->arrive) return fail(NAVSIM_EINVAL, "null buffer");
  if ((out->ep_return == nullptr) != (out->ep_path == nullptr))
    return fail(NAVSIM_EINVAL, "ep_return and ep_path must be given together");
  StepIO io = make_io(act_dev, out->obs, out->rew, out->done, out->arrive, out->trunc, 0, 0);
  io.ep_ret = out->ep_return;
  io.ep_path = out->ep_path;
  io.ep_len = out->ep_len;
  return launch_step(h, io, (cudaStream_t)stream, false, 0, 1);
}

}  // extern "C"

// The device side of a handle for the fused rollout kernel (internal: navppo_rollout_ex)
int navsim_device_view(navsim* h, void* stream, navsim_dev::DeviceView* out) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  out->c = h->c; out->st = h->st; out->map = h->d_map; out->rt = h->d_rt; out->stats = h->d_stats;
  out->variant = variant_of(h);
  h->launches++;
  return NAVSIM_OK;
}

// navsim_step_ex as a programmatic dependent launch (internal: navppo_rollout_ex chains it behind the policy kernel)
int navsim_step_chained(navsim_t* h, const float* act_dev, const navsim_step_out* out, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  if (!act_dev || !out || !out->obs || !out->rew || !out->done || !out->arrive) return fail(NAVSIM_EINVAL, "null buffer");
  StepIO io = make_io(act_dev, out->obs, out->rew, out->done, out->arrive, out->trunc, 0, 0);
  io.ep_ret = out->ep_return;
  io.ep_path = out->ep_path;
  io.ep_len = out->ep_len;
  return launch_step(h, io, (cudaStream_t)stream, false, 0, 1, true);
}

extern "C" {

int navsim_step_scripted(navsim_t* h, int32_t num_steps, uint64_t action_seed, float* obs_dev, float* rew_dev,
                         uint8_t* done_dev, uint8_t* arrive_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  if (!obs_dev || !rew_dev || !done_dev || !arrive_dev) return fail(NAVSIM_EINVAL, "null buffer");
  return launch_step(h, make_io(nullptr, obs_dev, rew_dev, done_dev, arrive_dev, nullptr, 0, 0), (cudaStream_t)stream, true,
                     action_seed, num_steps);
}

int navsim_rollout_scripted(navsim_t* h, int32_t num_steps, uint64_t action_seed, float* obs_dev, float* rew_dev,
                            uint8_t* done_dev, uint8_t* arrive_dev, uint8_t* trunc_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  if (!obs_dev || !rew_dev || !done_dev || !arrive_dev) return fail(NAVSIM_EINVAL, "null buffer");
  const long long N = h->c.N;
  return launch_step(h, make_io(nullptr, obs_dev, rew_dev, done_dev, arrive_dev, trunc_dev, N * NAVSIM_OBS_DIM, N),
                     (cudaStream_t)stream, true, action_seed, num_steps);
}

int navsim_reset_host(navsim_t* h, const uint8_t* mask_host, float* obs_host) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_host_call(h)) return rc;
  if (int rc = navsim_wait(h, 0)) return rc;
  if (!obs_host) return fail(NAVSIM_EINVAL, "null buffer");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t N = (size_t)h->c.N;
  cudaStream_t s = h->own_stream;
  if (mask_host) {
    memcpy(h->h_flags, mask_host, N);
    CUDA_TRY(cudaMemcpyAsync(h->d_flags, h->h_flags, N, cudaMemcpyHostToDevice, s));
  }
  if (int rc = navsim_reset(h, mask_host ? h->d_flags : nullptr, h->d_obs, s)) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->h_obs, h->d_obs, N * NAVSIM_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (!mask_host) {
    memcpy(obs_host, h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float));
  } else {
    for (size_t i = 0; i < N; ++i)
      if (mask_host[i]) memcpy(obs_host + i * NAVSIM_OBS_DIM, h->h_obs + i * NAVSIM_OBS_DIM, NAVSIM_OBS_DIM * sizeof(float));
  }
  return NAVSIM_OK;
}

int navsim_reset_host_ex(navsim_t* h, const uint8_t* mask_host, float* obs_host, double* pose_host) {
  if (!h) return fail(NAVSIM_EINVAL, "null handle");
  if (pose_host && !h->h_pose_dev) return fail(NAVSIM_ECUDA, "mapped host memory is not available on this device");
  h->reset_pose_out = pose_host ? h->h_pose_dev : nullptr;
  const int rc = navsim_reset_host(h, mask_host, obs_host);
  h->reset_pose_out = nullptr;
  if (rc == NAVSIM_OK && pose_host) {
    const size_t N = (size_t)h->c.N;
    if (!mask_host) memcpy(pose_host, h->h_pose, N * 6 * sizeof(double));
    else
      for (size_t i = 0; i < N; ++i)
        if (mask_host[i]) memcpy(pose_host + 6 * i, h->h_pose + 6 * i, 6 * sizeof(double));
  }
  return rc;
}

int navsim_step_host_ex(navsim_t* h, const float* act_host, const float* past_act_host, float* obs_host, float* rew_host,
                        uint8_t* done_host, uint8_t* arrive_host, uint8_t* trunc_host, double* pose_host) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_host_call(h)) return rc;
  if (int rc = navsim_wait(h, 0)) return rc;
  if (!act_host || !obs_host || !rew_host || !done_host || !arrive_host) return fail(NAVSIM_EINVAL, "null buffer");
  if (!h->h_rew_dev || !h->h_past_dev || !h->h_pose_dev)
    return fail(NAVSIM_ECUDA, "mapped host memory is not available on this device");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t N = (size_t)h->c.N;
  cudaStream_t s = h->own_stream;
  // everything through mapped pinned blocks: one launch, one synchronisation
  memcpy(h->h_act, act_host, N * 2 * sizeof(float));
  if (past_act_host) memcpy(h->h_past, past_act_host, N * 2 * sizeof(float));
  void* act_alias = nullptr;
  CUDA_TRY(cudaHostGetDevicePointer(&act_alias, h->h_act, 0));
  void* obs_alias = nullptr;
  CUDA_TRY(cudaHostGetDevicePointer(&obs_alias, h->h_obs, 0));
  float* rew_dev = h->h_rew_dev;
  uint8_t* fl_dev = reinterpret_cast<uint8_t*>(rew_dev + N);
  StepIO io = make_io(static_cast<const float*>(act_alias), static_cast<float*>(obs_alias), rew_dev, fl_dev, fl_dev + N,
                      fl_dev + 2 * N, 0, 0);
  io.past_act = past_act_host ? h->h_past_dev : nullptr;
  io.pose_out = pose_host ? h->h_pose_dev : nullptr;
  if (int rc = launch_step(h, io, s, false, 0, 1)) return rc;
  CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(obs_host, h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float));
  memcpy(rew_host, h->h_rew, N * sizeof(float));
  memcpy(done_host, h->h_flags, N);
  memcpy(arrive_host, h->h_flags + N, N);
  if (trunc_host) memcpy(trunc_host, h->h_flags + 2 * N, N);
  if (pose_host) memcpy(pose_host, h->h_pose, N * 6 * sizeof(double));
  return NAVSIM_OK;
}

int navsim_step_host(navsim_t* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host,
                     uint8_t* arrive_host, uint8_t* trunc_host) {
  if (int rc = check_ready(h)) return rc;
  if (int rc = begin_host_call(h)) return rc;
  if (int rc = navsim_wait(h, 0)) return rc;
  if (!act_host || !obs_host || !rew_host || !done_host || !arrive_host) return fail(NAVSIM_EINVAL, "null buffer");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t N = (size_t)h->c.N;
  cudaStream_t s = h->own_stream;
  // Page-locked caller buffers (cudaMallocHost / cudaHostRegister / torch pin_memory) are mapped
  // into the device's address space (unified addressing): the step kernel then reads the actions
  // and writes the observation rows straight over PCIe - no copy engine, no staging, one launch and
  // one stream synchronisation per step.  Reward and flags (7 bytes per agent) go to the handle's
  // own mapped block and are handed over with a small host copy.  Pageable caller buffers take
  // the staged path (pinned bounce buffers + copy engine).
  const float* act_dev = host_device_alias(h, 0, act_host);
  float* obs_dev = const_cast<float*>(host_device_alias(h, 1, obs_host));
  if (act_dev && obs_dev && h->h_rew_dev) {
    float* rew_dev = h->h_rew_dev;
    uint8_t* fl_dev = reinterpret_cast<uint8_t*>(rew_dev + N);
    if (int rc = launch_step(h, make_io(act_dev, obs_dev, rew_dev, fl_dev, fl_dev + N, fl_dev + 2 * N, 0, 0), s, false, 0, 1))
      return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
  } else {
    const float* act_src = act_host;
    if (!act_dev) {
      memcpy(h->h_act, act_host, N * 2 * sizeof(float));
      act_src = h->h_act;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_act, act_src, N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (int rc = launch_step(h, make_io(h->d_act, h->d_obs, h->d_rew, h->d_flags, h->d_flags + N, h->d_flags + 2 * N, 0, 0), s,
                             false, 0, 1))
      return rc;
    CUDA_TRY(cudaMemcpyAsync(obs_dev ? obs_host : h->h_obs, h->d_obs, N * NAVSIM_OBS_DIM * sizeof(float),
                             cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h->h_rew, h->d_rew, N * sizeof(float) + N * 3, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (!obs_dev) memcpy(obs_host, h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float));
  }
  memcpy(rew_host, h->h_rew, N * sizeof(float));
  memcpy(done_host, h->h_flags, N);
  memcpy(arrive_host, h->h_flags + N, N);
  if (trunc_host) memcpy(trunc_host, h->h_flags + 2 * N, N);
  return NAVSIM_OK;
}

int64_t navsim_step_host_async(navsim_t* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host,
                               uint8_t* arrive_host, uint8_t* trunc_host) {
  if (int rc = check_ready(h)) return rc;
  if (!act_host || !obs_host || !rew_host || !done_host || !arrive_host) return fail(NAVSIM_EINVAL, "null buffer");
  if (h->async_issued - h->async_waited >= NAVSIM_ASYNC_DEPTH)
    return fail(NAVSIM_EINVAL, "NAVSIM_ASYNC_DEPTH asynchronous steps are already in flight: navsim_wait for the oldest first");
  if (int rc = begin_host_call(h)) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t N = (size_t)h->c.N;
  if (!h->copy_stream) {   // first use: second stream, second device observation buffer, events
    const char* mode = getenv("NAVSIM_ASYNC_OBS");
    h->async_mode = (mode && std::string(mode) == "stores") ? 2 : (mode && std::string(mode) == "dma_act") ? 0
                    : (mode && std::string(mode) == "dma") ? 1 : 4;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    for (int k2 = 0; k2 < NAVSIM_ASYNC_DEPTH; ++k2) {
      CUDA_TRY(cudaMalloc(&h->d_act2[k2], N * 2 * sizeof(float)));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_h2d[k2], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    // per slot one device block laid out obs | rew | done | arrive | trunc (the layout of VecEnv.alloc_host_buffers)
    for (int k = 0; k < NAVSIM_ASYNC_DEPTH; ++k) CUDA_TRY(cudaMalloc(&h->d_obs2[k], N * (NAVSIM_OBS_DIM * sizeof(float) + 4 + 3)));
    for (int k = 0; k < NAVSIM_ASYNC_DEPTH; ++k) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_kernel[k], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copy[k], cudaEventDisableTiming));
    }
  }
  // every buffer must be page-locked: the kernel reads the actions and writes reward / flags through their device
  // aliases (7 bytes per agent over PCIe), the observations take the copy engine so that the next step's kernel
  // overlaps their transfer
  auto alias = [h](const void* p) -> void* {        // one driver query per new pointer (a loop reuses its buffers)
    for (const auto& kv : h->alias_cache)
      if (kv.first == p) return kv.second;
    cudaPointerAttributes at;
    void* a = nullptr;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else if (at.type == cudaMemoryTypeHost) a = at.devicePointer;
    if (h->alias_cache.size() >= 32) h->alias_cache.clear();
    h->alias_cache.emplace_back(p, a);
    return a;
  };
  const float* act_dev = static_cast<const float*>(alias(act_host));   // (re-pointed at the staged copy in mode 0)
  float* rew_dev = static_cast<float*>(alias(rew_host));
  uint8_t* done_dev = static_cast<uint8_t*>(alias(done_host));
  uint8_t* arrive_dev = static_cast<uint8_t*>(alias(arrive_host));
  uint8_t* trunc_dev = trunc_host ? static_cast<uint8_t*>(alias(trunc_host)) : nullptr;
  if (!act_dev || !rew_dev || !done_dev || !arrive_dev || (trunc_host && !trunc_dev) || !alias(obs_host))
    return fail(NAVSIM_EINVAL, "navsim_step_host_async needs page-locked buffers (cudaHostAlloc / cudaHostRegister / pin_memory)");
  const int k = (int)(h->async_issued & (NAVSIM_ASYNC_DEPTH - 1));
  cudaStream_t s = h->own_stream;
  // copy-engine modes: the kernel leaves the whole step in the slot's device block (a kernel that stores reward and
  // flags into host memory itself issues tens of thousands of 1- and 4-byte PCIe writes and lasts ~20 us for 8192
  // robots); the block goes home in ONE copy when the caller's arrays are laid out the same way, else array by array
  unsigned char* blk = reinterpret_cast<unsigned char*>(h->d_obs2[k]);
  float* b_obs = reinterpret_cast<float*>(blk);
  float* b_rew = reinterpret_cast<float*>(blk + N * NAVSIM_OBS_DIM * sizeof(float));
  uint8_t* b_done = blk + N * (NAVSIM_OBS_DIM * sizeof(float) + 4);
  uint8_t* b_arrive = b_done + N;
  uint8_t* b_trunc = b_arrive + N;
  const bool one_block = reinterpret_cast<unsigned char*>(rew_host) == reinterpret_cast<unsigned char*>(obs_host) + (b_rew - b_obs) * sizeof(float) &&
                         done_host == reinterpret_cast<uint8_t*>(rew_host) + 4 * N && arrive_host == done_host + N &&
                         (!trunc_host || trunc_host == arrive_host + N);
  auto copy_home = [&](cudaStream_t cs) -> cudaError_t {
    if (one_block) return cudaMemcpyAsync(obs_host, blk, N * (NAVSIM_OBS_DIM * sizeof(float) + 4 + (trunc_host ? 3 : 2)), cudaMemcpyDeviceToHost, cs);
    cudaError_t e = cudaMemcpyAsync(obs_host, b_obs, N * NAVSIM_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rew_host, b_rew, N * sizeof(float), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(done_host, b_done, N, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(arrive_host, b_arrive, N, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && trunc_host) e = cudaMemcpyAsync(trunc_host, b_trunc, N, cudaMemcpyDeviceToHost, cs);
    return e;
  };
  if (h->async_mode == 2) {
    // the kernel reads the actions and stores the observation rows in the caller's page-locked buffers itself: one
    // launch + one event per step, but the kernel then lasts as long as its PCIe traffic
    if (int rc = launch_step(h, make_io(act_dev, static_cast<float*>(alias(obs_host)), rew_dev, done_dev, arrive_dev, trunc_dev, 0, 0),
                             s, false, 0, 1))
      return rc;
    CUDA_TRY(cudaEventRecord(h->ev_copy[k], s));
    return ++h->async_issued;
  }
  // (d_obs2[k] is free: the depth check above means the caller has already waited for step t - depth, whose copy read it)
  if (h->async_mode == 0) {
    // actions through the copy engine too
    CUDA_TRY(cudaMemcpyAsync(h->d_act, act_host, N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    act_dev = h->d_act;
  } else if (h->async_mode == 4) {
    // ... on their own stream, so that the copy runs under the previous step's kernel
    CUDA_TRY(cudaMemcpyAsync(h->d_act2[k], act_host, N * 2 * sizeof(float), cudaMemcpyHostToDevice, h->h2d_stream));
    CUDA_TRY(cudaEventRecord(h->ev_h2d[k], h->h2d_stream));
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_h2d[k], 0));
    act_dev = h->d_act2[k];
  }
  if (int rc = launch_step(h, make_io(act_dev, b_obs, b_rew, b_done, b_arrive, b_trunc, 0, 0), s, false, 0, 1))
    return rc;
  CUDA_TRY(cudaEventRecord(h->ev_kernel[k], s));
  CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_kernel[k], 0));
  CUDA_TRY(copy_home(h->copy_stream));
  CUDA_TRY(cudaEventRecord(h->ev_copy[k], h->copy_stream));
  return ++h->async_issued;
}

int64_t navsim_step_host_pipelined(navsim_t* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host,
                                   uint8_t* arrive_host, uint8_t* trunc_host) {
  const int64_t t = navsim_step_host_async(h, act_host, obs_host, rew_host, done_host, arrive_host, trunc_host);
  if (t < 0) return t;
  if (t >= NAVSIM_ASYNC_DEPTH)
    if (int rc = navsim_wait(h, t - (NAVSIM_ASYNC_DEPTH - 1))) return rc;
  return t;
}

int navsim_wait(navsim_t* h, int64_t ticket) {
  if (!h) return fail(NAVSIM_EINVAL, "null handle");
  if (ticket < 0 || ticket > h->async_issued) return fail(NAVSIM_EINVAL, "unknown ticket");
  const int64_t upto = ticket == 0 ? h->async_issued : ticket;
  while (h->async_waited < upto) {
    const int k = (int)(h->async_waited & (NAVSIM_ASYNC_DEPTH - 1));
    CUDA_TRY(cudaEventSynchronize(h->ev_copy[k]));     // kernel done (reward / flags written) and observations copied
    h->async_waited++;
  }
  return NAVSIM_OK;
}

int navsim_scan(navsim_t* h, double* ranges_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (!ranges_dev) return fail(NAVSIM_EINVAL, "null buffer");
  navsim_scan_kernel<<<aux_grid_of(h), kBlock, aux_smem_bytes(h), (cudaStream_t)stream>>>(h->c, h->st, h->d_map, ranges_dev);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

static int field_ptr(navsim_t* h, int32_t field, void** p, size_t* elem) {
  switch (field) {
    case NAVSIM_F_X: *p = h->st.x; *elem = 8; break;
    case NAVSIM_F_Y: *p = h->st.y; *elem = 8; break;
    case NAVSIM_F_THETA: *p = h->st.th; *elem = 8; break;
    case NAVSIM_F_GOAL_X: *p = h->st.gx; *elem = 8; break;
    case NAVSIM_F_GOAL_Y: *p = h->st.gy; *elem = 8; break;
    case NAVSIM_F_PAST_DIST: *p = h->st.past; *elem = 8; break;
    case NAVSIM_F_PREV_A0: *p = h->st.pa0; *elem = 4; break;
    case NAVSIM_F_PREV_A1: *p = h->st.pa1; *elem = 4; break;
    case NAVSIM_F_STEPS: *p = h->st.steps; *elem = 4; break;
    case NAVSIM_F_DRAWS: *p = h->st.draws; *elem = 4; break;
    case NAVSIM_F_EP_RETURN: *p = h->st.ep_ret; *elem = 4; break;
    case NAVSIM_F_EP_PATH: *p = h->st.ep_path; *elem = 4; break;
    case NAVSIM_F_LAST_MOVE: *p = h->st.last_move; *elem = 4; break;
    case NAVSIM_F_WHEEL_L: *p = h->st.vl; *elem = 8; break;
    case NAVSIM_F_WHEEL_R: *p = h->st.vr; *elem = 8; break;
    default: return fail(NAVSIM_EINVAL, "unknown state field");
  }
  return NAVSIM_OK;
}

int navsim_get_state(navsim_t* h, int32_t field, void* host_out) {
  if (!h || !host_out) return fail(NAVSIM_EINVAL, "null argument");
  void* p; size_t elem;
  if (int rc = field_ptr(h, field, &p, &elem)) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(host_out, p, elem * (size_t)h->c.N, cudaMemcpyDeviceToHost));
  return NAVSIM_OK;
}

int navsim_set_state(navsim_t* h, int32_t field, const void* host_in) {
  if (!h || !host_in) return fail(NAVSIM_EINVAL, "null argument");
  void* p; size_t elem;
  if (int rc = field_ptr(h, field, &p, &elem)) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(p, host_in, elem * (size_t)h->c.N, cudaMemcpyHostToDevice));
  return NAVSIM_OK;
}

int navsim_get_stats(navsim_t* h, navsim_stats* out, int32_t clear) {
  if (!h || !out) return fail(NAVSIM_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  DevStats s;
  CUDA_TRY(cudaMemcpy(&s, h->d_stats, sizeof s, cudaMemcpyDeviceToHost));
  out->episodes = s.episodes; out->successes = s.successes; out->collisions = s.collisions;
  out->timeouts = s.timeouts; out->steps = s.steps;
  out->return_sum = s.return_sum; out->length_sum = s.length_sum; out->path_sum = s.path_sum;
  if (clear) CUDA_TRY(cudaMemset(h->d_stats, 0, sizeof(DevStats)));
  return NAVSIM_OK;
}

int navsim_clear_stats(navsim_t* h, void* stream) {
  if (!h) return fail(NAVSIM_EINVAL, "null handle");
  if (int rc = begin_device_call(h, (cudaStream_t)stream)) return rc;
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats), (cudaStream_t)stream));
  return NAVSIM_OK;
}

int navsim_num_agents(const navsim_t* h) { return h ? h->c.N : 0; }

int navsim_lanes_per_agent(const navsim_t* h) { return h ? lanes_of(h) : 0; }

int64_t navsim_launch_count(const navsim_t* h) { return h ? h->launches : 0; }

}  // extern "C"
