/* navsim.h — C-ABI of the B200 batched LiDAR-navigation simulator.
 *
 * This is the drop-in boundary for the reference's environment: one handle simulates N
 * independent agents, each of which behaves exactly like one instance of
 *     project_ppo/src/environment_new.py  class Env            (:26-382)
 * with the Gazebo/ROS process underneath it (cmd_vel -> diff-drive -> ray sensor -> scan,
 * environment_new.py:279-286) replaced by fused sm_100a kernels.  Each entry point names
 * the reference call it replaces.  The Python classes navbot_ppo_b200.Env / VecEnv bind
 * these symbols through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain C, no exceptions: every call returns 0 or a negative NAVSIM_E* code and
 *    nav_last_error() holds a message for the calling thread;
 *  - "dev" pointers are device memory owned by the caller (torch tensors), fp32/uint8,
 *    contiguous, and must stay alive until `stream` reaches the call; "host" pointers are
 *    ordinary host memory;
 *  - all work is enqueued on the given cudaStream_t (passed as void*), no hidden syncs
 *    except in the *_host entry points, which block until their outputs are in host memory;
 *  - one handle per device; a handle is not thread-safe, separate handles are independent;
 *  - the step/reset calls are CUDA-graph capturable.
 */
#ifndef NAVSIM_H_
#define NAVSIM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAVSIM_OK 0
#define NAVSIM_EINVAL (-22)
#define NAVSIM_ENOMEM (-12)
#define NAVSIM_ECUDA (-5)
#define NAVSIM_ENODEV (-19)

#define NAVSIM_OBS_DIM 16      /* environment_new.py:296-304: 10 lidar + 2 past action + 4 goal */
#define NAVSIM_ACT_DIM 2       /* main.py:19 */
#define NAVSIM_LIDAR_FEATS 10  /* environment_new.py:293 */
#define NAVSIM_MAX_RECTS 8
#define NAVSIM_MAX_BEAMS 360

typedef struct navsim navsim_t;

/* Simulator configuration.  navsim_default_cfg() fills in the reference's numbers. */
typedef struct navsim_cfg {
  int32_t num_agents;        /* N */
  int32_t num_beams;         /* gazebo.xacro:111 <samples>10 */
  int32_t max_episode_steps; /* ppo.py:552 max_timesteps_per_episode (500 on the main.py path) */
  int32_t auto_reset;        /* 1: VecEnv semantics (ppo.py:553-593 reset folded into step) */
  int32_t device;            /* CUDA ordinal */
  int32_t n_reset_rects;     /* goal rejection boxes used by Env.reset  (:340-343) */
  int32_t n_respawn_rects;   /* goal rejection boxes used on arrival    (:248-251) */
  int32_t lanes_per_agent;   /* 0 = chosen from N; else 1,2,..32 GPU lanes cooperating on one agent's step */
  uint64_t seed;             /* Philox key for goal sampling */
  int64_t agent_id_offset;   /* global id of agent 0 (rank * N when sharded) */
  double dt;                 /* 1 / 5 Hz LiDAR, gazebo.xacro:107 */
  double lidar_offset_x;     /* urdf.xacro:134-138 (-0.032) */
  double lidar_min, lidar_max;     /* gazebo.xacro:117-119 */
  double fov_min, fov_max;         /* gazebo.xacro:113-114 */
  double collision_range;    /* environment_new.py:188 min_range */
  double arrive_threshold;   /* :44-47 (0.2 training, 0.4 test) */
  double reward_scale;       /* :213  500 */
  double reward_collide;     /* :217  -100 */
  double reward_arrive;      /* :221  +120 */
  double diag_norm;          /* :21   sqrt(2) * 7.6 */
  double goal_lo, goal_hi;   /* :337  uniform(-3.6, 3.6) */
  double start_x, start_y, start_theta; /* turtlebot3_stage_1.launch:3-5 */
  double reset_rects[NAVSIM_MAX_RECTS * 4];   /* {xlo, xhi, ylo, yhi} each */
  double respawn_rects[NAVSIM_MAX_RECTS * 4];
  /* ---- options of the un-vendored Gazebo plugins, OFF (0) by default and outside the parity contract
   * (SURVEY.md 8 f-4).  Either one routes Env.step through the generic warp-per-agent kernel. */
  double lidar_noise_sigma;  /* gazebo.xacro:122-126 Gaussian range noise; the plugin's value is 0.01 */
  double wheel_accel;        /* gazebo.xacro:67 <wheelAcceleration> applied to the wheels' rim speeds [m/s^2]
                              * at the plugin's 30 Hz update rate (:71); 0 = cmd_vel is reached at once */
  double wheel_separation;   /* gazebo.xacro:65 / turtlebot3_fake.cpp:44: 0.160 m (used by the ramp only) */
  /* ---- start / goal tables of spawn_goal_sampler.py:37-72 (`--use_external_sampler`, arguments.py:41):
   * with sampler_mode = 1 and tables given through navsim_set_sampler, Env.reset places the robot at a
   * table start pose and the goal at a table point whose distance lies in [min, max] */
  double sampler_min_dist, sampler_max_dist;   /* GoalSpawnSampler defaults 1.5, 6.0 */
  int32_t sampler_mode;      /* 0 = Env.reset of environment_new.py (fixed spawn, uniform goal), 1 = tables */
  int32_t reserved0;
} navsim_cfg;

/* Per-agent state fields for navsim_get_state / navsim_set_state (host arrays of N). */
enum navsim_field {
  NAVSIM_F_X = 0,         /* double */
  NAVSIM_F_Y = 1,         /* double */
  NAVSIM_F_THETA = 2,     /* double, wrapped to (-pi, pi] */
  NAVSIM_F_GOAL_X = 3,    /* double */
  NAVSIM_F_GOAL_Y = 4,    /* double */
  NAVSIM_F_PAST_DIST = 5, /* double, Env.past_distance */
  NAVSIM_F_PREV_A0 = 6,   /* float, previous action (environment_new.py:299-300) */
  NAVSIM_F_PREV_A1 = 7,   /* float */
  NAVSIM_F_STEPS = 8,     /* int32, one_round counter of ppo.py:489,549 */
  NAVSIM_F_DRAWS = 9,     /* uint32, goal-sampler draws consumed */
  NAVSIM_F_EP_RETURN = 10,/* float, running episode return (ppo.py:544) */
  NAVSIM_F_EP_PATH = 11,  /* float, running path length (ppo.py:535-538) */
  NAVSIM_F_LAST_MOVE = 12,/* float, displacement of the latest step (not yet in EP_PATH) */
  NAVSIM_F_WHEEL_L = 13,  /* double, left / right rim speed [m/s] (wheel_accel > 0 only; zero after a reset) */
  NAVSIM_F_WHEEL_R = 14
};

/* Episode statistics accumulated on device (ppo.py:558-580 iteration metrics). */
typedef struct navsim_stats {
  uint64_t episodes, successes, collisions, timeouts, steps;
  double return_sum, length_sum, path_sum;
} navsim_stats;

const char* nav_last_error(void);
int navsim_abi_version(void);

/* Fill cfg with the reference constants for `num_agents` agents, 10 beams. */
int navsim_default_cfg(navsim_cfg* cfg, int32_t num_agents);

/* Env.__init__ (environment_new.py:27-47): allocate per-agent state on cfg->device. */
int navsim_create(navsim_t** out, const navsim_cfg* cfg);
int navsim_destroy(navsim_t* h);

/* Static obstacle map = what the .world file gives Gazebo (turtlebot3_stage_1.launch:8).
 * seg_host: S rows of {x0, y0, x1, y1} in metres (host doubles).
 * flags: NAVSIM_MAP_CLOSED_BOXES when the segments are the counter-clockwise edges of closed
 * convex obstacles (SDF collision boxes): walls seen from behind are then skipped.  Pass 0
 * for free-standing two-sided walls. */
#define NAVSIM_MAP_CLOSED_BOXES 1
int navsim_set_map(navsim_t* h, const double* seg_host, int32_t num_segments, int32_t flags);

/* GoalSpawnSampler tables (spawn_goal_sampler.py:5-35): starts_host[n_starts, 3] = (x, y, yaw),
 * goals_host[n_goals, 2] = (x, y), host doubles, at most NAVSIM_MAX_TABLE rows each.  Call after
 * navsim_set_map (the LaserScan seen from every start pose is cast here, once) on a handle created with
 * cfg.sampler_mode = 1; such a handle refuses reset / step until its tables are set. */
#define NAVSIM_MAX_TABLE 64
int navsim_set_sampler(navsim_t* h, const double* starts_host, int32_t n_starts, const double* goals_host,
                       int32_t n_goals);

/* Env.reset (environment_new.py:312-382) for every agent whose mask byte is non-zero
 * (mask_dev == NULL: all agents).  Writes obs[N,16] rows of the agents that were reset. */
int navsim_reset(navsim_t* h, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* Env.step (environment_new.py:272-310) for all agents.
 * act_dev[N,2] float; obs_dev[N,16], rew_dev[N] float; done_dev/arrive_dev/trunc_dev[N]
 * uint8 (trunc_dev may be NULL).  With cfg.auto_reset the episode protocol of
 * PPO.rollout (ppo.py:552-593) is applied in the same launch: on done|arrive|timeout the
 * agent is reset and obs holds the first observation of its next episode. */
int navsim_step(navsim_t* h, const float* act_dev, float* obs_dev, float* rew_dev,
                uint8_t* done_dev, uint8_t* arrive_dev, uint8_t* trunc_dev, void* stream);

/* navsim_step with the optional per-episode outputs: at the step that ends an episode (auto_reset)
 * ep_return[i] / ep_path[i] receive the episode's return and path length - the `return` and
 * `path_length` columns of the reference's per-episode csv (ppo.py:739-746); other entries are
 * left untouched.  trunc, ep_return, ep_path may be NULL (the last two together). */
typedef struct navsim_step_out {
  float* obs;          /* [N,16] */
  float* rew;          /* [N] */
  uint8_t* done;       /* [N] */
  uint8_t* arrive;     /* [N] */
  uint8_t* trunc;      /* [N] or NULL */
  float* ep_return;    /* [N] or NULL */
  float* ep_path;      /* [N] or NULL */
  int32_t* ep_len;     /* [N] or NULL: the episode's length in steps (batch_lens, ppo.py:583), same rule */
} navsim_step_out;
int navsim_step_ex(navsim_t* h, const float* act_dev, const navsim_step_out* out, void* stream);

/* Same two calls with HOST buffers: pinned staging, H2D, kernel, D2H, synchronise. */
int navsim_reset_host(navsim_t* h, const uint8_t* mask_host, float* obs_host);
int navsim_step_host(navsim_t* h, const float* act_host, float* obs_host, float* rew_host,
                     uint8_t* done_host, uint8_t* arrive_host, uint8_t* trunc_host);

/* Asynchronous form of navsim_step_host, a pipeline: the call enqueues one Env.step with HOST buffers and returns
 * a ticket (1, 2, 3, ..; negative = error) at once; navsim_wait(h, ticket) blocks until that step's outputs are in
 * the caller's buffers (ticket 0: every step issued so far).  At most NAVSIM_ASYNC_DEPTH steps may be in flight, so
 * a caller cycles through that many sets of output buffers and prepares later steps while earlier observations
 * cross PCIe.  All buffers must be page-locked.  The actions are staged by a host-to-device copy under the previous
 * step's kernel; the step's results go home in ONE device-to-host copy under the next step's kernel when the output
 * arrays are one block laid out obs[N,16] | rew[N] | done[N] | arrive[N] | trunc[N] (VecEnv.alloc_host_buffers), else
 * array by array.  Stream rules for ALL entry points: device entry points run on the
 * caller's stream, host entry points on streams the handle owns; the library orders the two kinds against each other
 * (a host call first waits for the last caller stream used, a device call waits for asynchronous host steps still
 * in flight), so they can be mixed on one handle without explicit synchronisation. */
#define NAVSIM_ASYNC_DEPTH 4
int64_t navsim_step_host_async(navsim_t* h, const float* act_host, float* obs_host, float* rew_host,
                               uint8_t* done_host, uint8_t* arrive_host, uint8_t* trunc_host);
int navsim_wait(navsim_t* h, int64_t ticket);
/* The two calls of a steady pipeline in one: enqueue a step like navsim_step_host_async, then block until the step
 * issued NAVSIM_ASYNC_DEPTH - 1 calls earlier has delivered its outputs (nothing to wait for during the first calls).
 * A caller cycling through NAVSIM_ASYNC_DEPTH buffer sets therefore finds, when call t returns, the results of step
 * t - (NAVSIM_ASYNC_DEPTH - 1) in the set the NEXT call will overwrite.  Returns the new step's ticket. */
int64_t navsim_step_host_pipelined(navsim_t* h, const float* act_host, float* obs_host, float* rew_host,
                                   uint8_t* done_host, uint8_t* arrive_host, uint8_t* trunc_host);

/* The one-robot calling convention of the reference, Env.step(action, past_action) /
 * Env.reset() followed by reads of env.position / env.goal_position / env.past_distance
 * (environment_new.py:272,299-300; ppo.py:535, main.py:202), in ONE launch and one
 * synchronisation: past_act_host[N,2] (may be NULL) replaces the simulator's own copy of the
 * previous action before the step; pose_host[N,6] doubles (may be NULL) receives x, y, theta,
 * goal x, goal y, past_distance after the step / reset. */
int navsim_step_host_ex(navsim_t* h, const float* act_host, const float* past_act_host, float* obs_host,
                        float* rew_host, uint8_t* done_host, uint8_t* arrive_host, uint8_t* trunc_host,
                        double* pose_host);
int navsim_reset_host_ex(navsim_t* h, const uint8_t* mask_host, float* obs_host, double* pose_host);

/* Scripted-action driver used by benchmarks: `num_steps` consecutive steps in ONE launch with
 * actions a0~U[0,1], a1~U[-1,1] drawn on device (Philox, key action_seed); every step
 * overwrites the [N,.] output arrays, which end up holding the last step's results. */
int navsim_step_scripted(navsim_t* h, int32_t num_steps, uint64_t action_seed, float* obs_dev,
                         float* rew_dev, uint8_t* done_dev, uint8_t* arrive_dev, void* stream);

/* Rollout-layout form of the same driver (PPO.rollout's buffers, ppo.py:476-483, time-major):
 * ONE launch runs `num_steps` steps per agent with the state held in registers and writes
 * obs_dev[H,N,16], rew_dev[H,N], done/arrive/trunc_dev[H,N] (trunc_dev may be NULL); row t holds
 * what navsim_step would have returned at step t. */
int navsim_rollout_scripted(navsim_t* h, int32_t num_steps, uint64_t action_seed, float* obs_dev,
                            float* rew_dev, uint8_t* done_dev, uint8_t* arrive_dev, uint8_t* trunc_dev,
                            void* stream);

/* LiDAR ranges of the current pose, ranges_dev[N, num_beams] doubles (+-inf gated),
 * i.e. the LaserScan message of environment_new.py:284 — exposed for parity tests. */
int navsim_scan(navsim_t* h, double* ranges_dev, void* stream);

/* Parity injection / inspection (blocking copies to/from host arrays of N elements). */
int navsim_get_state(navsim_t* h, int32_t field, void* host_out);
int navsim_set_state(navsim_t* h, int32_t field, const void* host_in);

int navsim_get_stats(navsim_t* h, navsim_stats* out, int32_t clear);
/* Zero the episode statistics in stream order (no read-back, no synchronisation). */
int navsim_clear_stats(navsim_t* h, void* stream);
int navsim_num_agents(const navsim_t* h);
/* Lanes per agent the step kernel runs with (cfg.lanes_per_agent, or the library's choice). */
int navsim_lanes_per_agent(const navsim_t* h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t navsim_launch_count(const navsim_t* h);

#ifdef __cplusplus
}
#endif
#endif /* NAVSIM_H_ */
