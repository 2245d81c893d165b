// navsim_device.cuh — device-side types and helpers of the simulator kernels (state layout, obstacle-set staging,
// the LaserScan sweep, goal / reset logic), shared by navsim_kernels.cu and by the fused rollout kernel in
// navppo_tcws.cu, which steps its own 128 robots between two policy evaluations.
#ifndef NAVSIM_DEVICE_CUH_
#define NAVSIM_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/navsim.h"
#include "navsim_math.h"

namespace navsim_dev {

// Scalars the kernels need, passed by value (fits the 4 KB kernel-parameter window).
struct SimConst {
  int32_t N, B, S, max_steps, auto_reset, n_reset_rects, n_respawn_rects, closed_boxes;
  int32_t rt_R, rt_W;                 // bearing table: offsets -rt_R..rt_R tenths of a metre, row length
  int32_t pick[NAVSIM_LIDAR_FEATS];   // beam index of lidar feature i = int(i * B / 10), environment_new.py:293
  uint64_t seed;
  int64_t agent_off;
  double dt, off_x, rmin, rmax, collide, arrive_thr, r_scale, r_collide, r_arrive, diag;
  double goal_lo, goal_hi, sx, sy, sth;
  float inv_diag;
  double reset_rects[NAVSIM_MAX_RECTS * 4];
  double respawn_rects[NAVSIM_MAX_RECTS * 4];
  // GoalSpawnSampler tables (device memory; n_starts == 0: the reference Env.reset)
  int32_t n_starts, n_goals;
  const double* starts;        // [n_starts, 3]
  const double* goals;         // [n_goals, 2]
  const float* start_scans;    // [n_starts, B] sanitised ranges seen from every start pose
  double smin, smax;
  // fidelity options (generic kernel only)
  float noise_sigma;
  double wheel_accel, wheel_sep;
};

struct SimState {
  double *x, *y, *th, *gx, *gy, *past;
  double *vl, *vr;             // wheel rim speeds (wheel_accel > 0 only)
  float *pa0, *pa1, *ep_ret, *ep_path, *last_move;
  int32_t* steps;
  uint32_t* draws;
};

// Where one launch reads actions and writes the per-step outputs.  Fused multi-step launches
// advance the output pointers by the strides after every step (time-major [H, N, .] rollout
// layout, ppo.py:476-483) or keep overwriting the same [N, .] arrays (stride 0).
struct StepIO {
  const float* act;   // [N,2], single-step launches with caller-provided actions
  float* obs;         // [N,16]
  float* rew;         // [N]
  uint8_t *done, *arrive, *trunc;  // [N] each, trunc may be null
  float *ep_ret, *ep_path;         // [N] each or null: return / path length of an episode, written at its last step
  int32_t* ep_len;                 // [N] or null: its length in steps (ppo.py:583), written at its last step
  const float* past_act;           // [N,2] or null: the caller's previous action replaces the simulator's copy
  double* pose_out;                // [N,6] or null: x, y, theta, goal x, goal y, past_distance after the step
  long long obs_stride, vec_stride;
};

// Device-side episode statistics (ppo.py:558-580).
struct DevStats {
  unsigned long long episodes, successes, collisions, timeouts, steps;
  double return_sum, length_sum, path_sum;
};

struct Agent {
  double x, y, th, gx, gy, past;
  int32_t start_idx;           // table row of the current start pose (table sampler), else -1
  float pa0, pa1, ep_ret, ep_path, last_move;
  int32_t steps;
  uint32_t draws;
};

constexpr int kBlock = 128;                  // threads per CTA of every simulator kernel
#ifndef NAVSIM_G1_MINBLOCKS
#define NAVSIM_G1_MINBLOCKS 6                // thread-per-agent variant: cap registers for 24 warps / SM
#endif
constexpr int kObsPad = NAVSIM_OBS_DIM + 1;  // +1 float: conflict-free column access
constexpr int kPadBeams = 36;                // register-resident beam count of the padded variant
constexpr float kInvRmax = 1.0f / 3.5f;      // environment_new.py:289 (lidar / 3.5)

// Obstacle set as staged into shared memory: S packed wall records (8 floats each), the beam
// table (B cosines, B sines) and the B sanitised ranges seen from the spawn pose, fp32, padded
// to the 16-byte granule of cp.async.bulk.
__host__ __device__ inline uint32_t map_bytes_of(int B, int S) {
  return (uint32_t)(((size_t)(NV_SEG_FLOATS * S + 3 * B) * sizeof(float) + 15) & ~(size_t)15);
}

struct MapView {
  const float *seg, *bc, *bs, *start_r;
};

__device__ __forceinline__ MapView map_view(const float* s_map, int B, int S) {
  MapView m;
  m.seg = s_map;
  m.bc = s_map + NV_SEG_FLOATS * S;
  m.bs = m.bc + B;
  m.start_r = m.bs + B;
  return m;
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
}

// Wait for phase 0 of the barrier.  Only the first warp polls it (a polling loop costs issue
// slots; bar.sync does not); the CTA barrier then publishes the map to everyone.  Must be
// called by all threads of the CTA.
__device__ __forceinline__ void wait_map(uint64_t* bar) {
  if (threadIdx.x < 32) {
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
  __syncthreads();
}

// ----------------------------------------------------------------------------------------
// Env.getOdometry (environment_new.py:138-181) in integers: yaw in whole degrees, rel_theta
// and diff_angle in hundredths of a degree (navsim_math.h explains why this is exact).  The
// bearing comes from the per-map table rt_tab (built by navsim_set_map with the host build of
// nv_rel_theta_centideg) whenever both goal offsets are within +-rt_R tenths of a metre.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void odom_features(const SimConst& c, const uint16_t* __restrict__ rt_tab, double x, double y,
                                              double th, double gx, double gy, int* yaw_o, int* rel_o, int* diff_o) {
  // :142 — the quaternion round trip atan2(sin th, cos th) returns th itself for th in (-pi, pi]
  int yaw = (int)nv_pyround0(th * NV_RAD2DEG);
  if (yaw < 0) yaw += 360;                                  // :144-147
  const double nxd = nv_round_scaled(gx - x, 10.0);         // :149  rel_dis_x = nx / 10
  const double nyd = nv_round_scaled(gy - y, 10.0);         // :150
  int m;
  const double R = (double)c.rt_R;
  if (fabs(nxd) <= R && fabs(nyd) <= R) {
    m = (int)__ldg(rt_tab + ((int)nyd + c.rt_R) * c.rt_W + ((int)nxd + c.rt_R));
  } else {
    const double lim = 2.0e9;
    const double cx = fmin(fmax(nxd, -lim), lim), cy = fmin(fmax(nyd, -lim), lim);
    m = nv_rel_theta_centideg((int)cx, (int)cy);            // :153-169
  }
  *yaw_o = yaw;
  *rel_o = m;
  *diff_o = nv_diff_angle_centideg(yaw, m);                 // :170-176
}

// obs[10..15] (:299-301) from the integer features; every lane of an agent's group holds the
// same values, lane 0 writes the row.
__device__ __forceinline__ void write_goal_feats(const SimConst& c, float* obs, float pa0, float pa1, double dist, int yaw,
                                                 int rel, int diff) {
  obs[10] = pa0;
  obs[11] = pa1;
  obs[12] = (float)dist * c.inv_diag;
  obs[13] = (float)yaw * (1.0f / 360.0f);
  obs[14] = (float)rel * (1.0f / 36000.0f);
  obs[15] = (float)diff * (1.0f / 18000.0f);
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

// Env.reset (:312-382) for one agent.  The robot always respawns at the same pose
// (turtlebot3_stage_1.launch:3-5), so its first LaserScan is a per-map constant: the B
// sanitised ranges were cast once by navsim_set_map (host build of the same physics) and sit
// behind the beam table in shared memory.  Row ownership follows the step kernel: the lane
// with (beam mod lidar_mod) == lidar_lane writes the lidar features sampling that beam, `feats`
// lanes write obs[10..15].
__device__ __forceinline__ void reset_agent(const SimConst& c, const MapView& mv, const uint16_t* __restrict__ rt_tab,
                                            uint64_t agent, Agent* a, float* obs, bool feats, int lidar_mod,
                                            int lidar_lane) {
  const float* start_r = mv.start_r;
  if (c.n_starts > 0) {
    // spawn_goal_sampler.py:52-63: start pose and goal point from the tables, distance-filtered; the scan of
    // every start pose was cast by navsim_set_sampler
    int is, ig;
    nv_sample_tables(c.seed, agent, &a->draws, c.starts, c.n_starts, c.goals, c.n_goals, c.smin, c.smax, &is, &ig);
    a->x = c.starts[3 * is]; a->y = c.starts[3 * is + 1]; a->th = c.starts[3 * is + 2];
    a->gx = c.goals[2 * ig]; a->gy = c.goals[2 * ig + 1];
    a->start_idx = is;
    start_r = c.start_scans + (size_t)is * c.B;
  } else {
    a->x = c.sx; a->y = c.sy; a->th = c.sth;                     // reset_world, :325
    sample_goal(c, c.reset_rects, c.n_reset_rects, agent, a);
  }
  const double dx = a->gx - a->x, dy = a->gy - a->y;
  a->past = sqrt(dx * dx + dy * dy);                             // :359 via :116-120
  a->pa0 = 0.f; a->pa1 = 0.f; a->steps = 0;
  a->ep_ret = 0.f; a->ep_path = 0.f; a->last_move = 0.f;
  int yaw, rel, diff;
  odom_features(c, rt_tab, a->x, a->y, a->th, a->gx, a->gy, &yaw, &rel, &diff);
#pragma unroll
  for (int i = 0; i < NAVSIM_LIDAR_FEATS; ++i)
    if ((c.pick[i] & (lidar_mod - 1)) == lidar_lane) obs[i] = start_r[c.pick[i]] * kInvRmax;   // :361-369
  if (feats) write_goal_feats(c, obs, 0.f, 0.f, a->past, yaw, rel, diff);                    // :372-376
}

// Goal respawn on arrival when the caller does not reset (environment_new.py:245-267): from the uniform square with
// the wider rejection margins, or - table sampler - a table point at an admissible distance from where the robot stands
__device__ __forceinline__ void respawn_goal(const SimConst& c, uint64_t agent, Agent* a) {
  if (c.n_starts > 0) {
    int ig = 0;
    for (int attempt = 0; attempt < 101; ++attempt) {
      int is;
      nv_table_indices(c.seed, agent, a->draws, c.n_starts, c.n_goals, &is, &ig);
      a->draws += 1u;
      const double dx = a->x - c.goals[2 * ig], dy = a->y - c.goals[2 * ig + 1];
      const double dist = sqrt(dx * dx + dy * dy);
      if (c.smin <= dist && dist <= c.smax) break;
    }
    a->gx = c.goals[2 * ig]; a->gy = c.goals[2 * ig + 1];
  } else {
    sample_goal(c, c.respawn_rects, c.n_respawn_rects, agent, a);
  }
}

__device__ __forceinline__ void load_agent(const SimState& st, int i, Agent* a) {
  a->x = st.x[i]; a->y = st.y[i]; a->th = st.th[i];
  a->gx = st.gx[i]; a->gy = st.gy[i]; a->past = st.past[i];
  a->pa0 = st.pa0[i]; a->pa1 = st.pa1[i];
  a->ep_ret = st.ep_ret[i]; a->ep_path = st.ep_path[i]; a->last_move = st.last_move[i];
  a->steps = st.steps[i]; a->draws = st.draws[i];
  a->start_idx = -1;
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

// ----------------------------------------------------------------------------------------
// LaserScan (row R) of one agent by the G lanes of its group.
//   cull   lane g examines walls g, g + G, ..: is the wall in reach and facing the sensor?
//          S <= 32 (stage maps): each lane's verdicts are bits of a mask, OR-combined over the
//          group with shuffle-xor rounds.  Larger maps (house: 208 walls, CW): the survivors go to
//          the agent's list in shared memory (atomics) - a robot sees a handful of the walls of
//          a house, and without the list a warp would run the beam loop for a wall whenever ANY
//          of its lanes sees it.
//   cast   lane g owns beams g, g + G, ..: it walks the group's visible walls and keeps the
//          largest inverse hit distance of each of its beams.  No reduction is needed (a beam
//          has one owner) and the lanes of a group run in lockstep.
// The maximum over walls does not depend on the order they are visited in, so the result is
// bit-identical to the serial sweep of the host build (nv_beam_q).
// KB = beams the group holds in registers: exactly B when KB == 10 (the reference's sensor),
// else B <= KB.  q[j] belongs to beam g + j G.
// ----------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_min(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w < v) ? w : v;
  }
  return v;
}

constexpr int kCompactWalls = 32;

// The G lanes of a group form a GB x GW grid: lane g owns beams (g mod GB) + j GB and every GW-th
// visible wall.  GB grows with G up to 4 (10-beam sensor) / 8 (padded variant); lanes beyond
// that split the walls, and their per-beam maxima are combined with log2(GW) shuffle rounds.
template <int G, int KB>
struct BeamShare {
  static constexpr int GBMAX = (KB == NAVSIM_LIDAR_FEATS) ? 4 : 8;
  static constexpr int GB = (G < GBMAX) ? G : GBMAX;
  static constexpr int GW = G / GB;
  static constexpr int PER = (KB + GB - 1) / GB;   // beams per lane
};

template <int G, int KB, bool CW>
__device__ __forceinline__ void group_sweep(const SimConst& c, const MapView& mv, float ox, float oy, float ch, float sh,
                                            int g, uint16_t* vis_list, int* vis_count, float* q) {
  constexpr int PER = BeamShare<G, KB>::PER, GB = BeamShare<G, KB>::GB, GW = BeamShare<G, KB>::GW;
  const int B = (KB == NAVSIM_LIDAR_FEATS) ? KB : c.B;
  const int gb = g & (GB - 1), gw = g / GB;
  float dx[PER], dy[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = gb + j * GB;
    if (b < B) nv_beam_dir(ch, sh, mv.bc[b], mv.bs[b], &dx[j], &dy[j]);
    else { dx[j] = 0.f; dy[j] = 0.f; }            // never hits: q stays 0
    q[j] = 0.0f;
  }
  if (!CW) {
    unsigned vis = 0;
    for (int k = g; k < c.S; k += G) {
      float wx, wy, ex, ey, tn;
      if (nv_seg_cull(mv.seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &wx, &wy, &ex, &ey, &tn)) vis |= 1u << k;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) vis |= __shfl_xor_sync(0xffffffffu, vis, o);
    int turn = 0;
    while (vis) {
      const int k = __ffs(vis) - 1;
      vis &= vis - 1;
      if (GW == 1 || ((turn++) & (GW - 1)) == gw) {
        nv_seg_view v;
        nv_seg_setup(mv.seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &v);
#pragma unroll
        for (int j = 0; j < PER; ++j) q[j] = nv_ray_q(&v, dx[j], dy[j], q[j]);
      }
    }
  } else {
    int n = 0;
    if (G > 1) {
      if (g == 0) *vis_count = 0;
      __syncwarp();
    }
    for (int k = g; k < c.S; k += G) {
      float wx, wy, ex, ey, tn;
      if (nv_seg_cull(mv.seg + NV_SEG_FLOATS * k, ox, oy, c.closed_boxes, &wx, &wy, &ex, &ey, &tn)) {
        const int pos = (G == 1) ? n++ : atomicAdd(vis_count, 1);
        vis_list[pos] = (uint16_t)k;
      }
    }
    if (G > 1) {
      __syncwarp();
      n = *vis_count;
    }
    for (int idx = gw; idx < n; idx += GW) {
      nv_seg_view v;
      nv_seg_setup(mv.seg + NV_SEG_FLOATS * (int)vis_list[idx], ox, oy, c.closed_boxes, &v);
#pragma unroll
      for (int j = 0; j < PER; ++j) q[j] = nv_ray_q(&v, dx[j], dy[j], q[j]);
    }
    if (G > 1) __syncwarp();   // the list is rewritten next step
  }
  if (GW > 1) {
#pragma unroll
    for (int j = 0; j < PER; ++j)
#pragma unroll
      for (int o = GB; o < G; o <<= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, q[j], o);
        q[j] = (w > q[j]) ? w : q[j];
      }
  }
}

// Range of one beam with the Gazebo gates and getState's sanitising (+inf -> 3.5, :193-194).
__device__ __forceinline__ float sanitised_range(float q, float rmin, float rmax) {
  const float t = nv_range_from_q(q, rmin, rmax);
  return (t == NV_INF_F) ? 3.5f : t;
}

// ----------------------------------------------------------------------------------------
// One Env.step of one agent, executed by the G lanes of its group (environment_new.py:272-303 + the episode
// protocol of ppo.py:535-593): drive, LaserScan, flags, observation row into `my_obs` (shared memory), reward,
// auto-reset.  `o` points at THIS step's output rows.  Used by navsim_step_kernel and by the fused rollout kernel.
// ----------------------------------------------------------------------------------------
struct StepRows {
  float* rew; uint8_t* done; uint8_t* arrive; uint8_t* trunc;   // [N]; trunc may be null
  float* ep_ret; float* ep_path; int32_t* ep_len;               // [N] or null
};

template <int G, int KB, bool CW>
__device__ __forceinline__ void agent_step(const SimConst& c, const MapView& mv, const uint16_t* __restrict__ rt_tab,
                                           DevStats* stats, Agent& a, float a0, float a1, int g, int i, uint64_t agent,
                                           bool valid, bool writer, float* my_obs, uint16_t* my_list, int* my_cnt,
                                           const StepRows& o, uint64_t* map_bar, bool& goal_dirty) {
  // ppo.py:535-538 — path length trails the motion by one step
  if (a.steps > 0) a.ep_path += a.last_move;

  // cmd_vel -> pose after one LiDAR period (:276-286).  The step needs sin/cos of the
  // midpoint heading (motion) and of the new heading (sensor); with G > 1 even lanes
  // evaluate one, odd lanes the other, and they swap results.
  double ds, th_mid, th_new, s_mid, c_mid, s_new, c_new;
  nv_drive_plan(a.th, (double)a0 / 4.0, (double)a1, c.dt, &ds, &th_mid, &th_new);
  if (G == 1) {
    nv_sincos(th_mid, &s_mid, &c_mid);
    nv_sincos(th_new, &s_new, &c_new);
  } else {
    const bool odd = (g & 1) != 0;
    double sv, cv;
    nv_sincos(odd ? th_new : th_mid, &sv, &cv);
    const double ps = __shfl_xor_sync(0xffffffffu, sv, 1), pc = __shfl_xor_sync(0xffffffffu, cv, 1);
    s_mid = odd ? ps : sv; c_mid = odd ? pc : cv;
    s_new = odd ? sv : ps; c_new = odd ? cv : pc;
  }
  const double px = a.x, py = a.y;
  nv_drive_apply(&a.x, &a.y, ds, s_mid, c_mid);
  a.th = th_new;
  {
    const double mx = a.x - px, my = a.y - py;
    a.last_move = sqrtf((float)(mx * mx + my * my));
  }

  // LaserScan + getState (:183-207).  The map is first needed here: its TMA copy has been
  // in flight behind the state loads and the drive arithmetic.
  if (map_bar) wait_map(map_bar);
  constexpr int PER = BeamShare<G, KB>::PER;
  float q[PER];
  group_sweep<G, KB, CW>(c, mv, (float)(a.x + c.off_x * c_new), (float)(a.y + c.off_x * s_new), (float)c_new, (float)s_new,
                         g, my_list, my_cnt, q);
  const int B = (KB == NAVSIM_LIDAR_FEATS) ? KB : c.B;
  const float rmin = (float)c.rmin, rmax = (float)c.rmax;
  float mn = NV_INF_F;
  // each lane turns its own beams into ranges (one IEEE division each), writes the lidar
  // features that sample them (:289-294), and the group combines the minimum (:200)
  constexpr int GB = BeamShare<G, KB>::GB;
  const int gb = g & (GB - 1);
  const bool lidar_writer = valid && g < GB;           // wall-group 0 of the GB x GW lane grid
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int b = gb + j * GB;
    if (b < B) {
      const float rb = sanitised_range(q[j], rmin, rmax);
      mn = (rb < mn) ? rb : mn;
      if (KB == NAVSIM_LIDAR_FEATS) {
        if (lidar_writer) my_obs[b] = rb * kInvRmax;               // idx_i == i when L == 10
      } else if (lidar_writer) {
#pragma unroll
        for (int i = 0; i < NAVSIM_LIDAR_FEATS; ++i)
          if (c.pick[i] == b) my_obs[i] = rb * kInvRmax;
      }
    }
  }
  if (G > 1) mn = group_min<G>(mn);
  const bool done = (c.collide > (double)mn) && (mn > 0.0f);       // :200
  const double ddx = a.gx - a.x, ddy = a.gy - a.y;
  const double d = sqrt(ddx * ddx + ddy * ddy);                    // :203
  const bool arrive = (d <= c.arrive_thr);                         // :204
  int yaw, rel, diff;
  odom_features(c, rt_tab, a.x, a.y, a.th, a.gx, a.gy, &yaw, &rel, &diff);
  if (writer) write_goal_feats(c, my_obs, a.pa0, a.pa1, d, yaw, rel, diff);   // :299-301

  double reward = c.r_scale * (a.past - d);                        // :211-213
  a.past = d;                                                      // :214
  if (done) reward = c.r_collide;                                  // :216-217
  if (arrive) reward = c.r_arrive;                                 // :220-221
  a.pa0 = a0; a.pa1 = a1;                                          // ppo.py:543
  a.steps += 1;                                                    // ppo.py:549
  a.ep_ret += (float)reward;                                       // ppo.py:544
  const bool timeout = a.steps >= c.max_steps;                     // ppo.py:552
  if (writer) {
    o.rew[i] = (float)reward;
    o.done[i] = done ? 1 : 0;
    o.arrive[i] = arrive ? 1 : 0;
    if (o.trunc) o.trunc[i] = (timeout && !done && !arrive) ? 1 : 0;
  }
  if (c.auto_reset) {
    if (done || arrive || timeout) {                               // ppo.py:553-593
      // setReward has already respawned a goal on arrival (:245-253); rollout throws it
      // away by resetting, but the draws it consumed stay consumed
      if (arrive && c.n_starts == 0) sample_goal(c, c.respawn_rects, c.n_respawn_rects, agent, &a);
      if (writer) {
        if (o.ep_ret) {                                            // ppo.py:739-746: the episode's csv row
          o.ep_ret[i] = a.ep_ret;
          o.ep_path[i] = a.ep_path;
        }
        if (o.ep_len) o.ep_len[i] = a.steps;
        atomicAdd(&stats->episodes, 1ull);
        if (arrive) atomicAdd(&stats->successes, 1ull);            // ppo.py:558-560
        else if (done) atomicAdd(&stats->collisions, 1ull);
        else atomicAdd(&stats->timeouts, 1ull);
        atomicAdd(&stats->return_sum, (double)a.ep_ret);
        atomicAdd(&stats->length_sum, (double)a.steps);
        atomicAdd(&stats->path_sum, (double)a.ep_path);
      }
      reset_agent(c, mv, rt_tab, agent, &a, my_obs, writer, BeamShare<G, KB>::GB,
                  (valid && g < BeamShare<G, KB>::GB) ? g : -1);
      goal_dirty = true;
    }
  } else if (arrive) {                                             // :245-267
    respawn_goal(c, agent, &a);
    const double gx = a.gx - a.x, gy = a.gy - a.y;
    a.past = sqrt(gx * gx + gy * gy);
    goal_dirty = true;
  }

}

// What the fused rollout kernel (navppo_tcws.cu) needs of a simulator handle: filled by navsim_device_view
// (navsim_kernels.cu), which also runs the handle's call checks and counts the launch.
struct DeviceView {
  SimConst c;
  SimState st;
  const float* map;        // packed obstacle set (map_bytes_of(c.B, c.S) bytes)
  const uint16_t* rt;      // bearing table
  DevStats* stats;
  int variant;             // 0: the reference's 10-beam sensor (the only one the fused kernel steps)
};

}  // namespace navsim_dev

struct navsim;
int navsim_device_view(navsim* h, void* stream, navsim_dev::DeviceView* out);

#endif  // NAVSIM_DEVICE_CUH_
