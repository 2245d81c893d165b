"""GPU tests of the rows SURVEY.md section 8 marks "next": the GoalSpawnSampler tables (f-3,
spawn_goal_sampler.py:37-72, `--use_external_sampler`) and the fidelity options of the un-vendored Gazebo
plugins (f-4: range noise, wheel-acceleration ramp, the waffle's 360-beam scan), all through the C-ABI."""
import numpy as np
import pytest
import torch

from navbot_ppo_b200 import _capi, maps
from navbot_ppo_b200.env import VecEnv
from oracle import binding
from tests.helpers import OBS_ATOL, REW_ATOL, golden

pytestmark = pytest.mark.gpu


def _rollout_against_oracle(env, sim, n, steps, action_seed, speed_floor=0.0):
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    tot = dict(done=0, arrive=0, trunc=0)
    for t in range(steps):
        act = binding.scripted_actions(action_seed, 0, t, n)
        act[:, 0] = np.maximum(act[:, 0], speed_floor)
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref, err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), tr_ref, err_msg=f"trunc t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
        tot["done"] += int(d_ref.sum()); tot["arrive"] += int(a_ref.sum()); tot["trunc"] += int(tr_ref.sum())
    for k, f in (("x", _capi.F_X), ("y", _capi.F_Y), ("th", _capi.F_THETA), ("gx", _capi.F_GOAL_X), ("gy", _capi.F_GOAL_Y)):
        np.testing.assert_array_equal(env.get_state(f), sim.arr[k], err_msg=k)
    np.testing.assert_array_equal(env.get_state(_capi.F_DRAWS), sim.arr["draws"])
    return tot


@pytest.mark.parametrize("tag", ["house_default", "stage1_default", "stage1_tight"])
def test_table_sampler_on_device_reproduces_the_reference_class(tag):
    """navsim_reset with GoalSpawnSampler tables: start pose, goal point and index pairs consumed of 24 robots
    x 12 resets equal the recording of the reference's own class (oracle/make_golden_sampler.py), including the
    window where most calls run into the 100-attempt cut-off."""
    g = golden("sampler_tables")
    rows, draws = g[tag + "_rows"], g[tag + "_draws"]
    lo, hi, seed = (float(v) for v in g[tag + "_cfg"])
    world = str(g[tag + "_world"])
    agents, episodes = draws.shape
    env = VecEnv(agents, map="house" if world == "small_house" else "stage_1", seed=int(seed), use_external_sampler=world,
                 sampler_min_dist=lo, sampler_max_dist=hi)
    for e in range(episodes):
        env.reset()
        got = np.stack([env.get_state(f) for f in (_capi.F_X, _capi.F_Y, _capi.F_THETA, _capi.F_GOAL_X, _capi.F_GOAL_Y)], 1)
        assert np.array_equal(got, rows[:, e]), (tag, e)
        assert np.array_equal(env.get_state(_capi.F_DRAWS), draws[:, e]), (tag, e)


@pytest.mark.parametrize("map_name,beams,n", [("house", 10, 2048), ("house", 36, 512), ("stage_1", 10, 4096), ("house", 48, 256)])
def test_rollouts_with_the_table_sampler_match_the_oracle(map_name, beams, n):
    """Episodes that start at table poses (any yaw) and aim at table goals, auto-reset inside the step kernel:
    observations (incl. the pre-cast scan of every start pose), rewards, flags, poses, draw counters against
    the C restatement; every kernel variant (10-beam, padded, warp-per-agent)."""
    cfg = _capi.default_cfg(n)
    cfg.seed, cfg.max_episode_steps, cfg.num_beams, cfg.sampler_mode = 21, 40, beams, 1
    if map_name == "house":
        cfg.n_reset_rects = cfg.n_respawn_rects = 0
    seg = maps.get_map(map_name)
    tables = maps.sampler_tables(maps.SAMPLER_FOR_MAP[map_name])
    env = VecEnv(n, map=seg, cfg=cfg)
    env.set_sampler(*tables)
    sim = binding.OracleSim(cfg, seg, nthreads=8, sampler_tables=tables)
    tot = _rollout_against_oracle(env, sim, n, 90, action_seed=4, speed_floor=0.5)
    assert tot["trunc"] > 0 and (tot["done"] > 0 or map_name == "stage_1")   # stage_1: 2 m per episode from the centre
    assert len(np.unique(env.get_state(_capi.F_GOAL_X))) > 4


def test_table_sampler_needs_its_tables_and_valid_sizes():
    cfg = _capi.default_cfg(8)
    cfg.sampler_mode = 1
    env = VecEnv(8, map=maps.get_map("stage_1"), cfg=cfg)
    with pytest.raises(_capi.NavError):
        env.reset()                                             # tables not set
    with pytest.raises(_capi.NavError):
        env.set_sampler(np.zeros((0, 3)), np.zeros((3, 2)))
    with pytest.raises(_capi.NavError):
        env.set_sampler(np.zeros((65, 3)), np.zeros((3, 2)))
    plain = VecEnv(8)
    with pytest.raises(_capi.NavError):
        plain.set_sampler(np.zeros((2, 3)), np.zeros((3, 2)))   # handle not created for tables


def test_wheel_acceleration_ramp_matches_the_host_build_and_ramps():
    """Fidelity option wheel_accel (turtlebot3_burger.gazebo.xacro:67): the device runs the same header function
    as the host build (oracle), so poses / wheel speeds are bit-identical and flags equal; and the robot really
    accelerates: after one step from rest at full command it has covered less ground than the unramped robot."""
    n = 1024
    cfg = _capi.default_cfg(n)
    cfg.seed, cfg.max_episode_steps, cfg.wheel_accel = 9, 60, 0.5
    seg = maps.get_map("stage_2")
    env = VecEnv(n, map=seg, cfg=cfg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    tot = _rollout_against_oracle(env, sim, n, 100, action_seed=2, speed_floor=0.6)
    assert tot["done"] > 0 and tot["trunc"] + tot["arrive"] >= 0
    np.testing.assert_array_equal(env.get_state(_capi.F_WHEEL_L), sim.arr["vl"])
    np.testing.assert_array_equal(env.get_state(_capi.F_WHEEL_R), sim.arr["vr"])
    # physics: one step at v = 0.25 m/s from rest
    ramped, plain = VecEnv(4, wheel_accel=0.5, seed=1), VecEnv(4, seed=1)
    act = torch.tensor([[1.0, 0.0]] * 4, device="cuda")
    for e in (ramped, plain):
        e.reset(); e.step(act)
    xr, xp = ramped.get_state(_capi.F_X)[0], plain.get_state(_capi.F_X)[0]
    assert abs(xp - 0.05) < 1e-12                       # 0.25 m/s x 0.2 s
    # six 30 Hz updates, rim speed += 0.5 * (1/30) each: 0.0167 .. 0.1 m/s -> 0.35 / 30 m
    assert abs(xr - sum(0.5 / 30 * k for k in range(1, 7)) / 30) < 1e-12 and xr < xp
    assert abs(ramped.get_state(_capi.F_WHEEL_L)[0] - 0.1) < 1e-12
    ramped.reset()
    assert ramped.get_state(_capi.F_WHEEL_L)[0] == 0.0  # the plugin's Reset() zeroes the wheels


def test_lidar_range_noise_statistics_and_partition_invariance():
    """Fidelity option lidar_noise_sigma (turtlebot3_burger.gazebo.xacro:122-126, stddev 0.01): the noise on a
    finite range has mean 0 and the requested standard deviation, differs per robot, per beam and per step, and
    is keyed by the global robot id so that a sharded run reproduces the single-device scans."""
    n, sigma = 8192, 0.01
    clean, noisy = VecEnv(n, map="stage_2", seed=3), VecEnv(n, map="stage_2", seed=3, lidar_noise_sigma=sigma)
    act = torch.zeros((n, 2), device="cuda")
    clean.reset(); noisy.reset()
    o0 = clean.step(act)[0].clone()
    o1 = noisy.step(act)[0].clone()                 # (step returns the environment's own buffer)
    r0, r1 = o0[:, :10].cpu().numpy() * 3.5, o1[:, :10].cpu().numpy() * 3.5
    hit = r0[0] < 3.5 - 0.1                         # beams that see a wall from the spawn pose (all robots stand there)
    assert hit.sum() >= 2
    d = (r1 - r0)[:, hit]
    assert abs(d.mean()) < 4 * sigma / np.sqrt(d.size) and abs(d.std() - sigma) < 0.02 * sigma
    assert np.all(np.abs(d) < 6 * sigma)
    c = np.corrcoef(d[:, 0], d[:, 1])[0, 1]
    assert abs(c) < 0.05                            # independent across beams
    o2, *_ = noisy.step(act)
    d2 = (o2[:, :10].cpu().numpy() * 3.5 - r0)[:, hit]
    assert abs(np.corrcoef(d[:, 0], d2[:, 0])[0, 1]) < 0.05    # and across steps
    np.testing.assert_array_equal(o1[:, 10:].cpu().numpy(), o0[:, 10:].cpu().numpy())   # only the ranges move
    # beams without a hit stay at the 3.5 m cap (the plugin's +inf), noise or not
    assert np.all(r1[:, ~hit][r0[:, ~hit] == 3.5] == 3.5)
    # partition invariance: robots 4096.. of the big batch = a shard with agent_id_offset = 4096
    shard = VecEnv(n // 2, map="stage_2", seed=3, lidar_noise_sigma=sigma, agent_id_offset=n // 2)
    shard.reset()
    s1, *_ = shard.step(act[: n // 2])
    assert torch.equal(s1, o1[n // 2:])


def test_waffle_scan_preset_matches_the_oracle():
    """lidar="waffle": 360 beams over the full circle (turtlebot3_waffle.gazebo.xacro:118-125), through the
    reference's subsampling rule idx_i = int(i * 360 / 10) (environment_new.py:292-294); collisions now also come
    from behind the robot."""
    n = 300
    env = VecEnv(n, map="stage_2", seed=13, max_episode_steps=70, lidar="waffle")
    assert int(env.cfg.num_beams) == 360 and env.cfg.fov_min == 0.0 and abs(env.cfg.fov_max - 6.28319) < 1e-12
    sim = binding.OracleSim(env.cfg, maps.get_map("stage_2"), nthreads=8)
    tot = _rollout_against_oracle(env, sim, n, 150, action_seed=8, speed_floor=0.8)
    assert tot["done"] > 0
    np.testing.assert_array_equal(env.scan().cpu().numpy(), sim.scan())
