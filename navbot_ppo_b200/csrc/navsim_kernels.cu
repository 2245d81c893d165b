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
// Data layout in HBM: structure-of-arrays, one array per state field, agent index fastest,
// so a warp's 32 agents read/write one contiguous 128/256-byte run per field.  Pose, goal
// and past_distance are fp64 because the reference computes in Python floats and quantises
// (see navsim_math.h); the outputs the policy consumes (obs, reward) are fp32 as ppo.py:616
// casts them.  The obstacle set (beam table + wall segments, a few hundred bytes to a few
// KB) is staged into shared memory once per CTA with one TMA bulk copy (cp.async.bulk +
// mbarrier) and read as warp-wide broadcasts.  The observation tile is transposed through
// shared memory so the [N,16] fp32 rows leave as full 128-bit coalesced stores.
//
// Compile with -fmad=false: the physics must round exactly like the host build of
// navsim_math.h that drives the reference Env in the oracle.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>

#include "../../include/navsim.h"
#include "nav_common.h"
#include "navsim_math.h"

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

// Scalars the kernels need, passed by value (fits the 4 KB kernel-parameter window).
struct SimConst {
  int32_t N, B, S, max_steps, auto_reset, n_reset_rects, n_respawn_rects, closed_boxes;
  uint64_t seed;
  int64_t agent_off;
  double dt, off_x, rmin, rmax, collide, arrive_thr, r_scale, r_collide, r_arrive, diag;
  double goal_lo, goal_hi, sx, sy, sth;
  double reset_rects[NAVSIM_MAX_RECTS * 4];
  double respawn_rects[NAVSIM_MAX_RECTS * 4];
};

struct SimState {
  double *x, *y, *th, *gx, *gy, *past;
  float *pa0, *pa1, *ep_ret, *ep_path, *last_move;
  int32_t* steps;
  uint32_t* draws;
};

// Device-side episode statistics (ppo.py:558-580).
struct DevStats {
  unsigned long long episodes, successes, collisions, timeouts, steps;
  double return_sum, length_sum, path_sum;
};

struct Agent {
  double x, y, th, gx, gy, past;
  float pa0, pa1, ep_ret, ep_path, last_move;
  int32_t steps;
  uint32_t draws;
};

constexpr int kObsPad = NAVSIM_OBS_DIM + 1;  // +1 float: conflict-free column access

// Obstacle set as staged into shared memory: S packed wall records (8 floats each), then the
// beam table (B cosines, B sines), fp32, padded to the 16-byte granule of cp.async.bulk.
__host__ __device__ inline uint32_t map_bytes_of(int B, int S) {
  return (uint32_t)(((size_t)(NV_SEG_FLOATS * S + 2 * B) * sizeof(float) + 15) & ~(size_t)15);
}

// ----------------------------------------------------------------------------------------
// TMA bulk copy of the obstacle set (global -> shared), completion on an mbarrier.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void stage_map(float* s_map, const float* g_map, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(s_map)),
        "l"(g_map), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
  }
  // every thread waits for phase 0 of the barrier
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_MAP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra DONE_MAP;\n"
      "bra WAIT_MAP;\n"
      "DONE_MAP:\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// ----------------------------------------------------------------------------------------
// Env.getOdometry (environment_new.py:138-181) from the pose directly.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void odom_features(double x, double y, double th, double gx, double gy, double* yaw_o,
                                              double* rel_theta_o, double* diff_o) {
  // :142 — the quaternion round trip atan2(sin th, cos th) returns th itself for th in (-pi, pi]
  double yaw = nv_pyround0(th * NV_RAD2DEG) + 0.0;
  if (!(yaw >= 0.0)) yaw = yaw + 360.0;            // :144-147
  double rx = nv_pyround1(gx - x);                 // :149
  double ry = nv_pyround1(gy - y);                 // :150
  double theta;
  if (rx > 0.0 && ry > 0.0) theta = nv_atan(ry / rx);                       // :153
  else if (rx > 0.0 && ry < 0.0) theta = 2.0 * NV_PI + nv_atan(ry / rx);    // :155
  else if (rx < 0.0 && ry < 0.0) theta = NV_PI + nv_atan(ry / rx);          // :157
  else if (rx < 0.0 && ry > 0.0) theta = NV_PI + nv_atan(ry / rx);          // :159
  else if (rx == 0.0 && ry > 0.0) theta = 0.5 * NV_PI;                      // :161
  else if (rx == 0.0 && ry < 0.0) theta = 1.5 * NV_PI;                      // :163
  else if (ry == 0.0 && rx > 0.0) theta = 0.0;                              // :165
  else theta = NV_PI;                                                       // :167
  double rel_theta = nv_pyround2(theta * NV_RAD2DEG);                       // :169
  double diff = yaw - rel_theta;                                            // :170
  if ((0.0 <= diff && diff <= 180.0) || (-180.0 <= diff && diff < 0.0)) diff = nv_pyround2(diff);
  else if (diff < -180.0) diff = nv_pyround2(360.0 + diff);
  else diff = nv_pyround2(-360.0 + diff);
  *yaw_o = yaw;
  *rel_theta_o = rel_theta;
  *diff_o = diff;
}

// LaserScan of the pose + getState + observation assembly (:183-207, :289-301).
// s_seg: S packed wall records (navsim_math.h), s_bc/s_bs: beam direction table, fp32, all in
// shared memory.
// KB > 0: beam count known at compile time -> segment-major loop with the per-beam best
// inverse hit distances in registers (the per-wall setup and culling run once per wall);
// KB == 0: any beam count, beam-major.  Both orders visit a beam's segments 0..S-1 in the
// same order with the same arithmetic, so they agree bit for bit with the host build.
template <int KB>
__device__ __forceinline__ void observe(const SimConst& c, const float* s_seg, const float* s_bc, const float* s_bs,
                                        const Agent& a, float pa0, float pa1,
                                        float* obs /* kObsPad row in smem */, bool* done, bool* arrive,
                                        double* dist) {
  double s, co;
  nv_sincos(a.th, &s, &co);
  const float ox = (float)(a.x + c.off_x * co), oy = (float)(a.y + c.off_x * s);
  const float ch = (float)co, sh = (float)s;
  const float rmin = (float)c.rmin, rmax = (float)c.rmax;
  float mn = NV_INF_F;
  if (KB > 0) {
    float dx[KB > 0 ? KB : 1], dy[KB > 0 ? KB : 1], q[KB > 0 ? KB : 1];
#pragma unroll
    for (int b = 0; b < KB; ++b) {
      nv_beam_dir(ch, sh, s_bc[b], s_bs[b], &dx[b], &dy[b]);
      q[b] = 0.0f;
    }
    for (int k = 0; k < c.S; ++k) {
      nv_seg_view v;
      if (!nv_seg_setup(s_seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &v)) continue;
#pragma unroll
      for (int b = 0; b < KB; ++b) q[b] = nv_ray_q(&v, dx[b], dy[b], q[b]);
    }
#pragma unroll
    for (int b = 0; b < KB; ++b) {
      float r = nv_range_from_q(q[b], rmin, rmax);
      if (r == NV_INF_F) r = 3.5f;                  // :193-194
      mn = (r < mn) ? r : mn;
      if (KB == NAVSIM_LIDAR_FEATS) obs[b] = r / 3.5f;  // :289, idx_i == i when L == 10
      else q[b] = r;
    }
    if (KB != NAVSIM_LIDAR_FEATS) {
#pragma unroll
      for (int i = 0; i < NAVSIM_LIDAR_FEATS; ++i) {
        const int idx = (int)((double)(i * KB) / 10.0);  // :293
        float r = 0.f;
#pragma unroll
        for (int b = 0; b < KB; ++b) r = (b == idx) ? q[b] : r;
        obs[i] = r / 3.5f;
      }
    }
  } else {
    int pick = 0;                                   // next lidar feature to emit
    int next_idx = 0;                               // int(pick * L / 10), :293
    for (int b = 0; b < c.B; ++b) {
      float dxb, dyb;
      nv_beam_dir(ch, sh, s_bc[b], s_bs[b], &dxb, &dyb);
      float r = nv_range_from_q(nv_beam_q(ox, oy, dxb, dyb, s_seg, c.S, c.closed_boxes), rmin, rmax);
      if (r == NV_INF_F) r = 3.5f;                  // :193-194
      mn = (r < mn) ? r : mn;
      while (pick < NAVSIM_LIDAR_FEATS && next_idx == b) {
        obs[pick] = r / 3.5f;                       // :289
        ++pick;
        next_idx = (int)((double)(pick * c.B) / 10.0);
      }
    }
  }
  *done = (c.collide > (double)mn) && (mn > 0.0f);  // :200
  const double ddx = a.gx - a.x, ddy = a.gy - a.y;
  const double d = sqrt(ddx * ddx + ddy * ddy);     // :203
  *arrive = (d <= c.arrive_thr);                    // :204
  *dist = d;
  double yaw, rel_theta, diff;
  odom_features(a.x, a.y, a.th, a.gx, a.gy, &yaw, &rel_theta, &diff);
  obs[10] = pa0;                                    // :299-300
  obs[11] = pa1;
  obs[12] = (float)d / (float)c.diag;               // :301
  obs[13] = (float)yaw / 360.0f;
  obs[14] = (float)rel_theta / 360.0f;
  obs[15] = (float)diff / 180.0f;
}

__device__ __forceinline__ bool in_rects(const double* r, int n, double gx, double gy) {
  bool hit = false;
  for (int i = 0; i < n; ++i)
    hit = hit || (r[4 * i] <= gx && gx <= r[4 * i + 1] && r[4 * i + 2] <= gy && gy <= r[4 * i + 3]);
  return hit;
}

// random.uniform(lo, hi) twice + rejection (:337-345 for reset, :245-253 on arrival).
__device__ __forceinline__ void sample_goal(const SimConst& c, const double* rects, int nrects, uint64_t agent,
                                            Agent* a) {
  for (;;) {
    double ux, uy;
    nv_goal_uniforms(c.seed, agent, a->draws, &ux, &uy);
    a->draws += 1u;
    a->gx = c.goal_lo + (c.goal_hi - c.goal_lo) * ux;
    a->gy = c.goal_lo + (c.goal_hi - c.goal_lo) * uy;
    if (!in_rects(rects, nrects, a->gx, a->gy)) return;
  }
}

// Env.reset (:312-382) for one agent; obs row filled with the first observation.
template <int KB>
__device__ __forceinline__ void reset_agent(const SimConst& c, const float* s_seg, const float* s_bc,
                                            const float* s_bs, uint64_t agent, Agent* a, float* obs) {
  a->x = c.sx; a->y = c.sy; a->th = c.sth;                       // reset_world, :325
  sample_goal(c, c.reset_rects, c.n_reset_rects, agent, a);
  const double dx = a->gx - a->x, dy = a->gy - a->y;
  a->past = sqrt(dx * dx + dy * dy);                             // :359 via :116-120
  a->pa0 = 0.f; a->pa1 = 0.f; a->steps = 0;
  a->ep_ret = 0.f; a->ep_path = 0.f; a->last_move = 0.f;
  bool done, arrive; double d;
  observe<KB>(c, s_seg, s_bc, s_bs, *a, 0.f, 0.f, obs, &done, &arrive, &d);
}

__device__ __forceinline__ void load_agent(const SimState& st, int i, Agent* a) {
  a->x = st.x[i]; a->y = st.y[i]; a->th = st.th[i];
  a->gx = st.gx[i]; a->gy = st.gy[i]; a->past = st.past[i];
  a->pa0 = st.pa0[i]; a->pa1 = st.pa1[i];
  a->ep_ret = st.ep_ret[i]; a->ep_path = st.ep_path[i]; a->last_move = st.last_move[i];
  a->steps = st.steps[i]; a->draws = st.draws[i];
}

__device__ __forceinline__ void store_agent(const SimState& st, int i, const Agent& a, bool goal_changed) {
  st.x[i] = a.x; st.y[i] = a.y; st.th[i] = a.th;
  st.past[i] = a.past;
  st.pa0[i] = a.pa0; st.pa1[i] = a.pa1;
  st.ep_ret[i] = a.ep_ret; st.ep_path[i] = a.ep_path; st.last_move[i] = a.last_move;
  st.steps[i] = a.steps;
  if (goal_changed) {  // goal and draw counter only move at episode boundaries
    st.gx[i] = a.gx; st.gy[i] = a.gy; st.draws[i] = a.draws;
  }
}

// Coalesced write-out of a CTA's observation tile: rows [row0, row0+rows) of obs[N,16].
__device__ __forceinline__ void flush_obs_tile(const float* s_obs, float* obs, int row0, int rows) {
  const int vec_total = rows * (NAVSIM_OBS_DIM / 4);
  float4* dst = reinterpret_cast<float4*>(obs + (size_t)row0 * NAVSIM_OBS_DIM);
  for (int v = threadIdx.x; v < vec_total; v += blockDim.x) {
    const int r = v >> 2, q = (v & 3) * 4;
    const float* src = s_obs + r * kObsPad + q;
    dst[v] = make_float4(src[0], src[1], src[2], src[3]);
  }
}

extern __shared__ __align__(16) unsigned char dyn_smem[];

// ----------------------------------------------------------------------------------------
// Env.step for all agents.  SCRIPTED: actions drawn on device (benchmark driver).
// ----------------------------------------------------------------------------------------
template <bool SCRIPTED, int KB>
__global__ void __launch_bounds__(128) navsim_step_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                          const float* __restrict__ act, float* __restrict__ obs,
                                                          float* __restrict__ rew, uint8_t* __restrict__ done_o,
                                                          uint8_t* __restrict__ arrive_o,
                                                          uint8_t* __restrict__ trunc_o, DevStats* stats,
                                                          uint64_t action_seed, uint32_t script_step) {
  // shared: [mbarrier 16 B][map: 8S + 2B floats, padded to 16 B][obs tile: blockDim x kObsPad floats]
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  const uint32_t map_bytes = map_bytes_of(c.B, c.S);
  float* s_obs = reinterpret_cast<float*>(dyn_smem + 16 + map_bytes);
  stage_map(s_map, g_map, map_bytes, bar);
  const float* s_seg = s_map;
  const float* s_bc = s_map + NV_SEG_FLOATS * c.S;
  const float* s_bs = s_bc + c.B;

  const int row0 = blockIdx.x * blockDim.x;
  const int i = row0 + threadIdx.x;
  float* my_obs = s_obs + threadIdx.x * kObsPad;
  if (i < c.N) {
    Agent a;
    load_agent(st, i, &a);
    const uint64_t agent = (uint64_t)(c.agent_off + i);
    float a0, a1;
    if (SCRIPTED) {
      uint32_t o[4];
      nv_philox4x32_10(script_step, 1u, (uint32_t)agent, (uint32_t)(agent >> 32), (uint32_t)action_seed,
                       (uint32_t)(action_seed >> 32), o);
      a0 = (float)(o[0] >> 8) * (1.0f / 16777216.0f);
      a1 = (float)(o[1] >> 8) * (2.0f / 16777216.0f) - 1.0f;
    } else {
      const float2 av = reinterpret_cast<const float2*>(act)[i];
      a0 = av.x; a1 = av.y;
    }
    // ppo.py:535-538 — path length trails the motion by one step
    if (a.steps > 0) a.ep_path += a.last_move;
    const double px = a.x, py = a.y;
    nv_drive(&a.x, &a.y, &a.th, (double)a0 / 4.0, (double)a1, c.dt);  // :276-286
    {
      const double mx = a.x - px, my = a.y - py;
      a.last_move = sqrtf((float)(mx * mx + my * my));
    }
    bool done, arrive; double d;
    observe<KB>(c, s_seg, s_bc, s_bs, a, a.pa0, a.pa1, my_obs, &done, &arrive, &d);  // :288-301
    double reward = c.r_scale * (a.past - d);                        // :211-213
    a.past = d;                                                      // :214
    if (done) reward = c.r_collide;                                  // :216-217
    if (arrive) reward = c.r_arrive;                                 // :220-221
    a.pa0 = a0; a.pa1 = a1;                                          // ppo.py:543
    a.steps += 1;                                                    // ppo.py:549
    a.ep_ret += (float)reward;                                       // ppo.py:544
    const bool timeout = a.steps >= c.max_steps;                     // ppo.py:552
    rew[i] = (float)reward;
    done_o[i] = done ? 1 : 0;
    arrive_o[i] = arrive ? 1 : 0;
    if (trunc_o) trunc_o[i] = (timeout && !done && !arrive) ? 1 : 0;
    bool goal_changed = false;
    if (c.auto_reset) {
      if (done || arrive || timeout) {                               // ppo.py:553-593
        // setReward has already respawned a goal on arrival (:245-253); rollout throws it
        // away by resetting, but the draws it consumed stay consumed
        if (arrive) sample_goal(c, c.respawn_rects, c.n_respawn_rects, agent, &a);
        atomicAdd(&stats->episodes, 1ull);
        if (arrive) atomicAdd(&stats->successes, 1ull);              // ppo.py:558-560
        else if (done) atomicAdd(&stats->collisions, 1ull);
        else atomicAdd(&stats->timeouts, 1ull);
        atomicAdd(&stats->return_sum, (double)a.ep_ret);
        atomicAdd(&stats->length_sum, (double)a.steps);
        atomicAdd(&stats->path_sum, (double)a.ep_path);
        reset_agent<KB>(c, s_seg, s_bc, s_bs, agent, &a, my_obs);
        goal_changed = true;
      }
    } else if (arrive) {                                             // :245-267
      sample_goal(c, c.respawn_rects, c.n_respawn_rects, agent, &a);
      const double gx = a.gx - a.x, gy = a.gy - a.y;
      a.past = sqrt(gx * gx + gy * gy);
      goal_changed = true;
    }
    store_agent(st, i, a, goal_changed);
  }
  __syncthreads();
  const int rows = min((int)blockDim.x, c.N - row0);
  if (rows > 0) flush_obs_tile(s_obs, obs, row0, rows);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats->steps, (unsigned long long)c.N);
}

// Env.reset for the masked agents.
__global__ void __launch_bounds__(128) navsim_reset_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                           const uint8_t* __restrict__ mask,
                                                           float* __restrict__ obs) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  const uint32_t map_bytes = map_bytes_of(c.B, c.S);
  float* s_obs = reinterpret_cast<float*>(dyn_smem + 16 + map_bytes);
  stage_map(s_map, g_map, map_bytes, bar);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  if (mask && !mask[i]) return;
  float* my_obs = s_obs + threadIdx.x * kObsPad;
  Agent a;
  load_agent(st, i, &a);
  reset_agent<0>(c, s_map, s_map + NV_SEG_FLOATS * c.S, s_map + NV_SEG_FLOATS * c.S + c.B,
                 (uint64_t)(c.agent_off + i), &a, my_obs);
  store_agent(st, i, a, true);
  if (obs) {
    float4* dst = reinterpret_cast<float4*>(obs + (size_t)i * NAVSIM_OBS_DIM);
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(my_obs[4 * q], my_obs[4 * q + 1], my_obs[4 * q + 2], my_obs[4 * q + 3]);
  }
}

// LaserScan only (parity tests of row R): ranges[N, B] doubles with the +-inf gates.
__global__ void __launch_bounds__(128) navsim_scan_kernel(SimConst c, SimState st, const float* __restrict__ g_map,
                                                          double* __restrict__ ranges) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(dyn_smem);
  float* s_map = reinterpret_cast<float*>(dyn_smem + 16);
  stage_map(s_map, g_map, map_bytes_of(c.B, c.S), bar);
  const float* s_seg = s_map; const float* s_bc = s_map + NV_SEG_FLOATS * c.S; const float* s_bs = s_bc + c.B;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  double s, co;
  nv_sincos(st.th[i], &s, &co);
  const float ox = (float)(st.x[i] + c.off_x * co), oy = (float)(st.y[i] + c.off_x * s);
  const float ch = (float)co, sh = (float)s;
  for (int b = 0; b < c.B; ++b) {
    float dx, dy;
    nv_beam_dir(ch, sh, s_bc[b], s_bs[b], &dx, &dy);
    ranges[(size_t)i * c.B + b] = (double)nv_range_from_q(nv_beam_q(ox, oy, dx, dy, s_seg, c.S, c.closed_boxes),
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
  float* d_map = nullptr;        // [8S + 2B] floats: packed wall records, beam cos, beam sin
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
  int block = 128;
};

namespace {

size_t smem_bytes(const navsim* h) {
  return 16 + (size_t)map_bytes_of(h->c.B, h->c.S) + (size_t)h->block * kObsPad * sizeof(float);
}

int grid_of(const navsim* h) { return (h->c.N + h->block - 1) / h->block; }

int check_ready(const navsim* h) {
  if (!h) return fail(NAVSIM_EINVAL, "null handle");
  if (!h->d_map) return fail(NAVSIM_EINVAL, "navsim_set_map has not been called");
  return NAVSIM_OK;
}

int launch_step(navsim* h, const float* act, float* obs, float* rew, uint8_t* done, uint8_t* arrive, uint8_t* trunc,
                cudaStream_t s, bool scripted, uint64_t action_seed) {
  const size_t smem = smem_bytes(h);
  const dim3 grid(grid_of(h)), block(h->block);
#define NAVSIM_LAUNCH(SCR, KB)                                                                               \
  navsim_step_kernel<SCR, KB><<<grid, block, smem, s>>>(h->c, h->st, h->d_map, act, obs, rew, done, arrive, trunc, \
                                                        h->d_stats, action_seed, script_step)
  const uint32_t script_step = scripted ? h->script_step++ : 0u;
  const bool ten = (h->c.B == 10);  // the reference's sensor (gazebo.xacro:111) gets the unrolled path
  if (scripted) { if (ten) NAVSIM_LAUNCH(true, 10); else NAVSIM_LAUNCH(true, 0); }
  else          { if (ten) NAVSIM_LAUNCH(false, 10); else NAVSIM_LAUNCH(false, 0); }
#undef NAVSIM_LAUNCH
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

}  // namespace

extern "C" {

const char* nav_last_error(void) { return g_err.c_str(); }

int navsim_abi_version(void) { return 1; }

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
  // a CTA of 64 keeps >= 128 CTAs in flight at N = 8192 (148 SMs); big batches use 128
  h->block = (c.N >= 148 * 128 * 2) ? 128 : 64;
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
  TRY_OR_CLEAN(cudaMalloc(&dslab, N * 6 * sizeof(double)));
  h->st.x = dslab; h->st.y = dslab + N; h->st.th = dslab + 2 * N;
  h->st.gx = dslab + 3 * N; h->st.gy = dslab + 4 * N; h->st.past = dslab + 5 * N;
  TRY_OR_CLEAN(cudaMalloc(&fslab, N * 7 * sizeof(float)));
  h->st.pa0 = fslab; h->st.pa1 = fslab + N; h->st.ep_ret = fslab + 2 * N; h->st.ep_path = fslab + 3 * N;
  h->st.last_move = fslab + 4 * N;
  h->st.steps = reinterpret_cast<int32_t*>(fslab + 5 * N);
  h->st.draws = reinterpret_cast<uint32_t*>(fslab + 6 * N);
  TRY_OR_CLEAN(cudaMemset(dslab, 0, N * 6 * sizeof(double)));
  TRY_OR_CLEAN(cudaMemset(fslab, 0, N * 7 * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_stats, sizeof(DevStats)));
  TRY_OR_CLEAN(cudaMemset(h->d_stats, 0, sizeof(DevStats)));
  TRY_OR_CLEAN(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_act, N * 2 * sizeof(float)));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float)));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_rew, N * sizeof(float)));
  TRY_OR_CLEAN(cudaMallocHost(&h->h_flags, N * 3));
  TRY_OR_CLEAN(cudaMalloc(&h->d_act, N * 2 * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_obs, N * NAVSIM_OBS_DIM * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_rew, N * sizeof(float)));
  TRY_OR_CLEAN(cudaMalloc(&h->d_flags, N * 3));
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
  if (h->d_stats) cudaFree(h->d_stats);
  if (h->h_act) cudaFreeHost(h->h_act);
  if (h->h_obs) cudaFreeHost(h->h_obs);
  if (h->h_rew) cudaFreeHost(h->h_rew);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  if (h->d_act) cudaFree(h->d_act);
  if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_rew) cudaFree(h->d_rew);
  if (h->d_flags) cudaFree(h->d_flags);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return NAVSIM_OK;
}

int navsim_set_map(navsim_t* h, const double* seg_host, int32_t num_segments, int32_t flags) {
  if (!h || !seg_host) return fail(NAVSIM_EINVAL, "null argument");
  if (num_segments < 1) return fail(NAVSIM_EINVAL, "a map needs at least one segment");
  const int B = h->c.B;
  const size_t bytes = map_bytes_of(B, num_segments);
  if (16 + bytes + (size_t)h->block * kObsPad * sizeof(float) > 200 * 1024)
    return fail(NAVSIM_EINVAL, "map does not fit in shared memory");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  float* host = new (std::nothrow) float[bytes / sizeof(float)]();
  if (!host) return fail(NAVSIM_ENOMEM, "host allocation failed");
  for (int k = 0; k < num_segments; ++k) nv_pack_segment(seg_host + 4 * k, h->cfg.lidar_max, host + NV_SEG_FLOATS * k);
  // beam direction table, gazebo.xacro:111-114: B samples over [fov_min, fov_max] inclusive
  for (int i = 0; i < B; ++i) {
    const double a = (B > 1) ? h->cfg.fov_min + (double)i * ((h->cfg.fov_max - h->cfg.fov_min) / (double)(B - 1))
                             : 0.5 * (h->cfg.fov_min + h->cfg.fov_max);
    double sn, cs;
    nv_sincos(a, &sn, &cs);
    host[NV_SEG_FLOATS * num_segments + i] = (float)cs;
    host[NV_SEG_FLOATS * num_segments + B + i] = (float)sn;
  }
  if (h->d_map) { cudaFree(h->d_map); h->d_map = nullptr; }
  cudaError_t e = cudaMalloc(&h->d_map, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_map, host, bytes, cudaMemcpyHostToDevice);
  delete[] host;
  if (e != cudaSuccess) return fail(NAVSIM_ECUDA, std::string("set_map: ") + cudaGetErrorString(e));
  h->S = num_segments;
  h->c.S = num_segments;
  h->c.closed_boxes = (flags & NAVSIM_MAP_CLOSED_BOXES) ? 1 : 0;
  const size_t smem = smem_bytes(h);
  CUDA_TRY(cudaFuncSetAttribute(navsim_step_kernel<false, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(navsim_step_kernel<true, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(navsim_step_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(navsim_step_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(navsim_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(navsim_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return NAVSIM_OK;
}

int navsim_reset(navsim_t* h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  navsim_reset_kernel<<<grid_of(h), h->block, smem_bytes(h), (cudaStream_t)stream>>>(h->c, h->st, h->d_map, mask_dev,
                                                                                    obs_dev);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navsim_step(navsim_t* h, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev,
                uint8_t* arrive_dev, uint8_t* trunc_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (!act_dev || !obs_dev || !rew_dev || !done_dev || !arrive_dev) return fail(NAVSIM_EINVAL, "null buffer");
  return launch_step(h, act_dev, obs_dev, rew_dev, done_dev, arrive_dev, trunc_dev, (cudaStream_t)stream, false, 0);
}

int navsim_step_scripted(navsim_t* h, int32_t num_steps, uint64_t action_seed, float* obs_dev, float* rew_dev,
                         uint8_t* done_dev, uint8_t* arrive_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (!obs_dev || !rew_dev || !done_dev || !arrive_dev) return fail(NAVSIM_EINVAL, "null buffer");
  for (int t = 0; t < num_steps; ++t)
    if (int rc = launch_step(h, nullptr, obs_dev, rew_dev, done_dev, arrive_dev, nullptr, (cudaStream_t)stream, true,
                             action_seed))
      return rc;
  return NAVSIM_OK;
}

int navsim_reset_host(navsim_t* h, const uint8_t* mask_host, float* obs_host) {
  if (int rc = check_ready(h)) return rc;
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

int navsim_step_host(navsim_t* h, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host,
                     uint8_t* arrive_host, uint8_t* trunc_host) {
  if (int rc = check_ready(h)) return rc;
  if (!act_host || !obs_host || !rew_host || !done_host || !arrive_host) return fail(NAVSIM_EINVAL, "null buffer");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t N = (size_t)h->c.N;
  cudaStream_t s = h->own_stream;
  memcpy(h->h_act, act_host, N * 2 * sizeof(float));
  CUDA_TRY(cudaMemcpyAsync(h->d_act, h->h_act, N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  if (int rc = launch_step(h, h->d_act, h->d_obs, h->d_rew, h->d_flags, h->d_flags + N, h->d_flags + 2 * N, s, false, 0))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(h->h_obs, h->d_obs, N * NAVSIM_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_rew, h->d_rew, N * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_flags, h->d_flags, N * 3, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(obs_host, h->h_obs, N * NAVSIM_OBS_DIM * sizeof(float));
  memcpy(rew_host, h->h_rew, N * sizeof(float));
  memcpy(done_host, h->h_flags, N);
  memcpy(arrive_host, h->h_flags + N, N);
  if (trunc_host) memcpy(trunc_host, h->h_flags + 2 * N, N);
  return NAVSIM_OK;
}

int navsim_scan(navsim_t* h, double* ranges_dev, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (!ranges_dev) return fail(NAVSIM_EINVAL, "null buffer");
  navsim_scan_kernel<<<grid_of(h), h->block, smem_bytes(h), (cudaStream_t)stream>>>(h->c, h->st, h->d_map, ranges_dev);
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

int navsim_num_agents(const navsim_t* h) { return h ? h->c.N : 0; }

int64_t navsim_launch_count(const navsim_t* h) { return h ? h->launches : 0; }

}  // extern "C"
