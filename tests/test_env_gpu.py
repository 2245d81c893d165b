"""GPU parity tests of the simulator, all through the C-ABI (navbot_ppo_b200.VecEnv / Env
are ctypes shims over libnavbot_b200.so).  Checker = the CPU oracle and the golden traces
recorded from the reference's own Env.  Bar: done/arrive/timeout bit-exact, pose/goal state
bit-exact, obs and reward within 1e-5 (they leave the kernel as fp32)."""
import math

import numpy as np
import pytest
import torch

from navbot_ppo_b200 import _capi, maps
from navbot_ppo_b200.env import Env, VecEnv
from oracle import binding
from tests.helpers import OBS_ATOL, REW_ATOL, cfg_from_golden, golden

pytestmark = pytest.mark.gpu


def _vec_from_cfg(cfg, segments):
    return VecEnv(cfg.num_agents, map=segments, device=0, cfg=cfg)


@pytest.mark.parametrize("name", ["env_rollout_stage_1", "env_rollout_stage_2", "env_rollout_stage_1_eval",
                                  "env_rollout_house", "env_rollout_house_36beams"])
def test_step_matches_reference_golden_rollout(name):
    g = golden(name)
    env = _vec_from_cfg(cfg_from_golden(g, auto_reset=1), g["segments"])
    obs0 = env.reset().cpu().numpy()
    np.testing.assert_allclose(obs0, g["obs0"], atol=OBS_ATOL, rtol=0)
    for t in range(g["act"].shape[0]):
        obs, rew, done, arrive = env.step(torch.from_numpy(g["act"][t]).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), g["done"][t].astype(np.uint8), err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), g["arrive"][t].astype(np.uint8), err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), g["trunc"][t].astype(np.uint8), err_msg=f"trunc t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), g["obs_next"][t], atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), g["rew"][t], atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
    # poses integrate bit-identically (shared physics header); goals are the same Philox draws
    for k, f in (("x", _capi.F_X), ("y", _capi.F_Y), ("th", _capi.F_THETA), ("gx", _capi.F_GOAL_X),
                 ("gy", _capi.F_GOAL_Y)):
        np.testing.assert_array_equal(env.get_state(f), g[k][-1], err_msg=k)
    np.testing.assert_array_equal(env.get_state(_capi.F_DRAWS), g["draws"][-1].astype(np.uint32))
    np.testing.assert_allclose(env.get_state(_capi.F_PAST_DIST), g["past"][-1], rtol=1e-15)


def test_raw_env_protocol_matches_reference_golden():
    """auto_reset off: goal respawn inside step on arrival (environment_new.py:245-267),
    explicit masked reset after collisions."""
    g = golden("env_raw_stage_1")
    env = _vec_from_cfg(cfg_from_golden(g, auto_reset=0), g["segments"])
    np.testing.assert_allclose(env.reset().cpu().numpy(), g["obs0"], atol=OBS_ATOL, rtol=0)
    for t in range(g["act"].shape[0]):
        obs, rew, done, arrive = env.step(torch.from_numpy(g["act"][t]).cuda())
        np.testing.assert_allclose(obs.cpu().numpy(), g["obs"][t], atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), g["rew"][t], atol=REW_ATOL, rtol=0)
        np.testing.assert_array_equal(done.cpu().numpy(), g["done"][t].astype(np.uint8))
        np.testing.assert_array_equal(arrive.cpu().numpy(), g["arrive"][t].astype(np.uint8))
        mask = g["reset_mask"][t].astype(np.uint8)
        if mask.any():
            ro = env.reset(torch.from_numpy(mask).cuda()).cpu().numpy()
            sel = mask.astype(bool)
            np.testing.assert_allclose(ro[sel], g["obs_reset"][t][sel], atol=OBS_ATOL, rtol=0)
        np.testing.assert_array_equal(env.get_state(_capi.F_GOAL_X), g["gx"][t])
        np.testing.assert_array_equal(env.get_state(_capi.F_DRAWS), g["draws"][t].astype(np.uint32))


def test_single_env_dropin_matches_reference_golden():
    """navbot_ppo_b200.Env (the class main.py would import, main.py:433) on agent 0's trace."""
    g = golden("env_raw_stage_1")
    env = Env(True, seed=int(g["seed"]), map=g["segments"])
    obs = env.reset()
    assert obs.shape == (16,) and obs.dtype == np.float64
    np.testing.assert_allclose(obs, g["obs0"][0], atol=OBS_ATOL, rtol=0)
    past = np.zeros(2, np.float32)
    for t in range(120):
        act = g["act"][t][0]
        o, r, d, a = env.step(act, past)
        past = act
        assert isinstance(r, float) and isinstance(d, bool) and isinstance(a, bool)
        np.testing.assert_allclose(o, g["obs"][t][0], atol=OBS_ATOL, rtol=0, err_msg=f"t={t}")
        assert abs(r - g["rew"][t][0]) <= REW_ATOL and d == bool(g["done"][t][0]) and a == bool(g["arrive"][t][0])
        assert env.goal_position.position.x == g["gx"][t][0]
        if g["reset_mask"][t][0]:
            np.testing.assert_allclose(env.reset(), g["obs_reset"][t][0], atol=OBS_ATOL, rtol=0)
            past = np.zeros(2, np.float32)


@pytest.mark.parametrize("n,map_name,steps", [(8192, "stage_1", 160), (16384, "stage_2", 60), (1000, "stage_2", 70),
                                              (1, "stage_1", 30)])
def test_step_matches_oracle_at_config_sizes(n, map_name, steps):
    """BASELINE.json configs[1] / configs[2] sizes against the C oracle with scripted actions."""
    cfg = _capi.default_cfg(n)
    cfg.seed = 11
    cfg.max_episode_steps = 50
    seg = maps.get_map(map_name)
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    tot = dict(done=0, arrive=0, trunc=0)
    for t in range(steps):
        act = binding.scripted_actions(3, 0, t, n)
        if t % 2:  # every other step: full speed to provoke collisions
            act[: n // 2, 0] = 1.0
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref, err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), tr_ref, err_msg=f"trunc t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
        tot["done"] += int(d_ref.sum()); tot["arrive"] += int(a_ref.sum()); tot["trunc"] += int(tr_ref.sum())
    for k, f in (("x", _capi.F_X), ("y", _capi.F_Y), ("th", _capi.F_THETA), ("gx", _capi.F_GOAL_X),
                 ("gy", _capi.F_GOAL_Y)):
        np.testing.assert_array_equal(env.get_state(f), sim.arr[k], err_msg=k)
    np.testing.assert_array_equal(env.get_state(_capi.F_DRAWS), sim.arr["draws"])
    np.testing.assert_array_equal(env.get_state(_capi.F_STEPS), sim.arr["steps"])
    np.testing.assert_allclose(env.get_state(_capi.F_EP_RETURN), sim.arr["ep_ret"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(env.get_state(_capi.F_EP_PATH), sim.arr["ep_path"], rtol=1e-5, atol=1e-5)
    if n >= 1000:
        assert tot["trunc"] > 0 and tot["done"] + tot["arrive"] > 0


def test_scripted_device_actions_match_oracle_stream():
    """navsim_step_scripted (the benchmark driver) draws the same actions as the oracle."""
    n = 4096
    cfg = _capi.default_cfg(n)
    cfg.seed = 5
    seg = maps.get_map("stage_1")
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    env.reset(); sim.reset()
    for t in range(25):
        o_ref, r_ref, d_ref, a_ref, _ = sim.step(binding.scripted_actions(99, 0, t, n))
        obs, rew, done, arrive = env.step_scripted(1, action_seed=99)
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref)
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0)
    np.testing.assert_array_equal(env.get_state(_capi.F_X), sim.arr["x"])


def test_bearing_feature_exhaustive_over_offset_grid():
    """rel_theta = round(degrees(atan-with-quadrants), 2) depends only on the two goal
    offsets rounded to 0.1 m (environment_new.py:149-169): check EVERY pair in
    [-12.0, 12.0]^2 — the device atan polynomial vs libm — plus yaw at every whole degree."""
    ks = np.arange(-120, 121)
    gx, gy = np.meshgrid(ks / 10.0, ks / 10.0, indexing="ij")
    n = gx.size
    cfg = _capi.default_cfg(n)
    cfg.auto_reset = 0
    cfg.max_episode_steps = 1 << 30
    cfg.arrive_threshold = -1.0  # never arrive: keep the injected goals
    seg = maps.synthetic_map(4, extent=40.0)
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    env.reset(); sim.reset()
    th = np.deg2rad((np.arange(n) % 721) / 2.0 - 180.0 + 1e-9)
    th = np.where(th <= -math.pi, th + 2 * math.pi, th)
    for f, k, v in ((_capi.F_GOAL_X, "gx", gx.ravel()), (_capi.F_GOAL_Y, "gy", gy.ravel()),
                    (_capi.F_THETA, "th", th)):
        env.set_state(f, v)
        sim.arr[k][:] = v
    act = np.zeros((n, 2), np.float32)
    o_ref, *_ = sim.step(act)
    obs, *_ = env.step(torch.from_numpy(act).cuda())
    got = obs.cpu().numpy()
    # the three quantised features must agree to fp32 rounding, i.e. no 0.01-degree slips
    for col, scale in ((13, 360.0), (14, 360.0), (15, 180.0)):
        err = np.abs(got[:, col].astype(np.float64) - o_ref[:, col]) * scale
        assert err.max() < 2e-4, (col, err.max())


@pytest.mark.parametrize("beams,boxes", [(10, 52), (36, 52), (24, 12)])
def test_scan_bit_exact_on_dense_maps(beams, boxes):
    """Row R alone (house-like segment counts, BASELINE configs[4] beam sweep): the device
    LaserScan equals the host build of the same physics header bit for bit."""
    n = 512
    cfg = _capi.default_cfg(n)
    cfg.num_beams = beams
    cfg.goal_lo, cfg.goal_hi = -6.0, 6.0
    seg = maps.synthetic_map(boxes, seed=3)
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg)
    env.reset(); sim.reset()
    rng = np.random.RandomState(0)
    for f, k, v in ((_capi.F_X, "x", rng.uniform(-6, 6, n)), (_capi.F_Y, "y", rng.uniform(-6, 6, n)),
                    (_capi.F_THETA, "th", rng.uniform(-3.1, 3.1, n))):
        env.set_state(f, v)
        sim.arr[k][:] = v
    got = env.scan().cpu().numpy()
    want = sim.scan()
    np.testing.assert_array_equal(got, want)
    assert np.isinf(want).any() and np.isfinite(want).any()
    act = binding.scripted_actions(1, 0, 0, n)
    o_ref, r_ref, d_ref, a_ref, _ = sim.step(act)
    obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
    np.testing.assert_array_equal(done.cpu().numpy(), d_ref)
    np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0)


def test_host_buffer_entry_points_equal_device_entry_points():
    n = 2048
    a = VecEnv(n, seed=2); b = VecEnv(n, seed=2)
    oa = a.reset().cpu().numpy(); ob = b.reset_host()
    np.testing.assert_array_equal(oa, ob)
    for t in range(10):
        act = binding.scripted_actions(8, 0, t, n)
        obs, rew, done, arrive = a.step(torch.from_numpy(act).cuda())
        hobs, hrew, hdone, harr, htr = b.step_host(act)
        np.testing.assert_array_equal(obs.cpu().numpy(), hobs)
        np.testing.assert_array_equal(rew.cpu().numpy(), hrew)
        np.testing.assert_array_equal(done.cpu().numpy(), hdone)
        np.testing.assert_array_equal(a.trunc.cpu().numpy(), htr)


def test_partition_invariance_and_statistics():
    """Sharding contract: agents [0,N) on one handle == two handles of N/2 with
    agent_id_offset, so multi-GPU runs reproduce the single-GPU episode stream exactly."""
    n = 8192
    whole = VecEnv(n, seed=9, max_episode_steps=40)
    lo = VecEnv(n // 2, seed=9, max_episode_steps=40)
    hi = VecEnv(n // 2, seed=9, max_episode_steps=40, agent_id_offset=n // 2)
    whole.reset(); lo.reset(); hi.reset()
    for t in range(90):
        act = torch.from_numpy(binding.scripted_actions(4, 0, t, n)).cuda()
        o, r, d, a = whole.step(act)
        o1, r1, d1, a1 = lo.step(act[: n // 2].contiguous())
        o2, r2, d2, a2 = hi.step(act[n // 2:].contiguous())
        assert torch.equal(o, torch.cat([o1, o2])) and torch.equal(r, torch.cat([r1, r2]))
        assert torch.equal(d, torch.cat([d1, d2])) and torch.equal(a, torch.cat([a1, a2]))
    s, s1, s2 = whole.stats(), lo.stats(), hi.stats()
    assert s.episodes == s1.episodes + s2.episodes > 0
    assert s.episodes == s.successes + s.collisions + s.timeouts
    assert s.steps == 90 * n
    assert abs(s.return_sum - (s1.return_sum + s2.return_sum)) < 1e-6 * max(1.0, abs(s.return_sum))


def test_million_agent_batch_invariants():
    """Full-scale properties that need no oracle: ranges normalised, flags consistent with
    the observation, auto-reset restores the spawn observation pattern."""
    n = 1 << 20
    env = VecEnv(n, seed=1, max_episode_steps=25)
    env.reset()
    env.step_scripted(29, action_seed=2)
    g = torch.Generator(device="cuda").manual_seed(0)
    act = torch.rand(n, 2, device="cuda", generator=g) * torch.tensor([1.0, 2.0], device="cuda") - \
        torch.tensor([0.0, 1.0], device="cuda")
    obs, rew, done, arrive = env.step(act)
    o = obs.cpu().numpy(); r = rew.cpu().numpy()
    d = done.cpu().numpy().astype(bool); a = arrive.cpu().numpy().astype(bool); tr = env.trunc.cpu().numpy().astype(bool)
    assert np.isfinite(o).all() and (o[:, :10] >= 0).all() and (o[:, :10] <= 1).all()
    assert (o[:, 13] >= 0).all() and (o[:, 13] < 1).all() and (np.abs(o[:, 15]) <= 1).all()
    assert (r[a] == 120.0).all() and (r[d & ~a] == -100.0).all()
    assert not (tr & (d | a)).any()
    term = d | a | tr
    assert term.any() and tr.any()
    # a freshly reset agent sits at the origin with zero past action and yaw 0
    assert (o[term, 10] == 0).all() and (o[term, 11] == 0).all() and (o[term, 13] == 0).all()
    steps = env.get_state(_capi.F_STEPS)
    assert (steps[term] == 0).all() and (steps[~term] > 0).all()
    s = env.stats()
    assert s.steps == 30 * n and s.episodes == s.successes + s.collisions + s.timeouts > 0


def test_error_paths():
    cfg = _capi.default_cfg(0)
    with pytest.raises(_capi.NavError):
        VecEnv(0, cfg=cfg)
    env = VecEnv(8)
    with pytest.raises(ValueError):
        env.step(torch.zeros(7, 2, device="cuda"))
    import ctypes
    h = ctypes.c_void_p()
    cfg = _capi.default_cfg(8)
    _capi.check(_capi.lib().navsim_create(ctypes.byref(h), ctypes.byref(cfg)))
    # step before set_map is refused, not a crash
    rc = _capi.lib().navsim_step(h, env.obs.data_ptr(), env.obs.data_ptr(), env.rew.data_ptr(), env.done.data_ptr(),
                                 env.arrive.data_ptr(), None, None)
    assert rc == -22 and b"set_map" in _capi.lib().nav_last_error()
    _capi.lib().navsim_destroy(h)


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_every_lane_split_is_bit_identical_and_matches_oracle(lanes):
    """The step kernel spreads one agent's wall sweep over G lanes (G picked from N); every G
    must give the oracle's flags/poses and the same observation bits as G = 1."""
    n = 777  # not a multiple of any agents-per-CTA count: exercises the shadow lanes
    cfg = _capi.default_cfg(n)
    cfg.seed = 21
    cfg.max_episode_steps = 30
    cfg.lanes_per_agent = lanes
    seg = maps.get_map("stage_2")
    env = _vec_from_cfg(cfg, seg)
    cfg1 = _capi.default_cfg(n)
    cfg1.seed, cfg1.max_episode_steps, cfg1.lanes_per_agent = 21, 30, 1
    base = _vec_from_cfg(cfg1, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    o0 = env.reset().cpu().numpy()
    np.testing.assert_allclose(o0, sim.reset(), atol=OBS_ATOL, rtol=0)
    np.testing.assert_array_equal(o0, base.reset().cpu().numpy())
    for t in range(80):
        act = binding.scripted_actions(13, 0, t, n)
        if t % 3 == 0:
            act[::2, 0] = 1.0
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        ad = torch.from_numpy(act).cuda()
        obs, rew, done, arrive = env.step(ad)
        ob, rb, db, ab = base.step(ad)
        assert torch.equal(obs, ob) and torch.equal(rew, rb) and torch.equal(done, db) and torch.equal(arrive, ab)
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref, err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), tr_ref, err_msg=f"trunc t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
    for k, f in (("x", _capi.F_X), ("y", _capi.F_Y), ("th", _capi.F_THETA), ("gx", _capi.F_GOAL_X), ("gy", _capi.F_GOAL_Y)):
        np.testing.assert_array_equal(env.get_state(f), sim.arr[k], err_msg=k)
    s = env.stats()
    assert s.episodes == s.successes + s.collisions + s.timeouts > 0 and s.steps == 80 * n


@pytest.mark.parametrize("n,lanes", [(8192, 0), (1500, 4), (300, 1)])
def test_fused_rollout_launch_equals_step_by_step(n, lanes):
    """navsim_rollout_scripted: H steps in ONE launch, state in registers, outputs in the
    [H, N, .] rollout layout == H single-step launches == the oracle."""
    H = 48
    cfg = _capi.default_cfg(n)
    cfg.seed, cfg.max_episode_steps, cfg.lanes_per_agent = 4, 20, lanes
    seg = maps.get_map("stage_1")
    fused, single = _vec_from_cfg(cfg, seg), _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    fused.reset(); single.reset(); sim.reset()
    out = fused.rollout_scripted(H, action_seed=77)
    torch.cuda.synchronize()
    for t in range(H):
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(binding.scripted_actions(77, 0, t, n))
        obs, rew, done, arrive = single.step_scripted(1, action_seed=77)
        assert torch.equal(out["obs"][t], obs) and torch.equal(out["rew"][t], rew), t
        assert torch.equal(out["done"][t], done) and torch.equal(out["arrive"][t], arrive), t
        np.testing.assert_array_equal(out["done"][t].cpu().numpy(), d_ref)
        np.testing.assert_array_equal(out["arrive"][t].cpu().numpy(), a_ref)
        np.testing.assert_array_equal(out["trunc"][t].cpu().numpy(), tr_ref)
        np.testing.assert_allclose(out["obs"][t].cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0)
        np.testing.assert_allclose(out["rew"][t].cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0)
    for f in (_capi.F_X, _capi.F_Y, _capi.F_THETA, _capi.F_GOAL_X, _capi.F_DRAWS, _capi.F_STEPS, _capi.F_EP_RETURN,
              _capi.F_EP_PATH):
        np.testing.assert_array_equal(fused.get_state(f), single.get_state(f))
    np.testing.assert_array_equal(fused.get_state(_capi.F_X), sim.arr["x"])
    assert fused.stats().episodes == single.stats().episodes > 0
    # a second fused call continues the same action stream
    out2 = fused.rollout_scripted(3, action_seed=77)
    for t in range(3):
        obs, *_ = single.step_scripted(1, action_seed=77)
        assert torch.equal(out2["obs"][t], obs)


@pytest.mark.parametrize("beams,lanes", [(12, 0), (18, 8), (24, 1), (36, 32), (48, 0), (360, 0)])
def test_beam_sweep_multi_step_matches_oracle(beams, lanes):
    """BASELINE configs[4] beam sweep on a house-like map (52 boxes = 208 walls): the padded
    register variant (B <= 36) and the warp-per-agent variant (B > 36) over many steps,
    including the lidar subsampling idx_i = int(i * L / 10) (environment_new.py:292-294)."""
    n = 600
    cfg = _capi.default_cfg(n)
    cfg.num_beams, cfg.seed, cfg.max_episode_steps, cfg.lanes_per_agent = beams, 8, 25, lanes
    cfg.goal_lo, cfg.goal_hi = -6.0, 6.0
    seg = maps.synthetic_map(52, seed=5)
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg, nthreads=8)
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    nd = 0
    for t in range(60):
        act = binding.scripted_actions(6, 0, t, n)
        act[:, 0] = np.maximum(act[:, 0], 0.7)
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref, err_msg=f"arrive t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
        nd += int(d_ref.sum())
    assert nd > 0
    np.testing.assert_array_equal(env.get_state(_capi.F_X), sim.arr["x"])
    np.testing.assert_array_equal(env.scan().cpu().numpy(), sim.scan())


@pytest.mark.parametrize("beams,n", [(10, 4096), (36, 1024)])
def test_house_map_matches_oracle(beams, n):
    """BASELINE configs[4]: the vendored turtlebot3_house geometry (52 boxes = 208 walls), spawn
    (-3, 1), 10 and 36 beams, against the C oracle."""
    env = VecEnv(n, map="house", seed=3, max_episode_steps=40, num_beams=beams)
    assert env.segments.shape == (208, 4) and (env.cfg.start_x, env.cfg.start_y) == (-3.0, 1.0)
    sim = binding.OracleSim(env.cfg, env.segments, nthreads=8)
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    nd = 0
    for t in range(70):
        act = binding.scripted_actions(2, 0, t, n)
        act[:, 0] = np.maximum(act[:, 0], 0.6)
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref, err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), tr_ref, err_msg=f"trunc t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0, err_msg=f"rew t={t}")
        nd += int(d_ref.sum())
    assert nd > 0
    for k, f in (("x", _capi.F_X), ("y", _capi.F_Y), ("th", _capi.F_THETA), ("gx", _capi.F_GOAL_X)):
        np.testing.assert_array_equal(env.get_state(f), sim.arr[k], err_msg=k)
    np.testing.assert_array_equal(env.scan().cpu().numpy(), sim.scan())


def test_pinned_host_buffers_take_the_zero_copy_path_and_agree():
    """navsim_step_host with page-locked caller buffers (kernel reads actions / writes observations
    over PCIe directly) == the staged path with pageable buffers == the device entry point."""
    n = 3000
    a = VecEnv(n, seed=6, max_episode_steps=30); b = VecEnv(n, seed=6, max_episode_steps=30)
    c = VecEnv(n, seed=6, max_episode_steps=30)
    a.reset(); b.reset_host(); c.reset_host()
    hb = b.alloc_host_buffers()
    for t in range(45):
        act = binding.scripted_actions(8, 0, t, n)
        obs, rew, done, arrive = a.step(torch.from_numpy(act).cuda())
        hb["act"][:] = act
        pobs, prew, pdone, parr, ptr = b.step_host(hb["act"], out=hb)
        assert pobs is hb["obs"]
        sobs, srew, sdone, sarr, str_ = c.step_host(act)
        for x, y, z in ((obs, pobs, sobs), (rew, prew, srew), (done, pdone, sdone), (arrive, parr, sarr), (a.trunc, ptr, str_)):
            np.testing.assert_array_equal(x.cpu().numpy(), y)
            np.testing.assert_array_equal(y, z)
    assert a.stats().episodes == b.stats().episodes > 0


def test_step_and_policy_calls_are_cuda_graph_capturable():
    """include/navsim.h / navppo.h promise graph-capturable device entry points: capture
    act + step for 4 steps in one CUDA graph, replay it, compare with eager calls."""
    from navbot_ppo_b200.nets import NetActor, _Handles, _stream
    n, steps = 1024, 4
    torch.manual_seed(0)
    actor = NetActor(16, 2)
    flat = actor._ensure_bound(torch.device("cuda:0"))
    h = _Handles.get(torch.device("cuda:0"))
    L = _capi.lib()

    def run(env, obs, act, logp, rew, done, arrive, trunc, stream_ptr):
        for t in range(steps):
            _capi.check(L.navppo_act(h, flat.data_ptr(), obs.data_ptr(), n, 0.8, 0, 0, t, None, act[t].data_ptr(),
                                     logp[t].data_ptr(), None, stream_ptr))
            _capi.check(L.navsim_step(env._h, act[t].data_ptr(), obs.data_ptr(), rew[t].data_ptr(), done[t].data_ptr(),
                                      arrive[t].data_ptr(), trunc[t].data_ptr(), stream_ptr))

    def bufs():
        return (torch.zeros(steps, n, 2, device="cuda"), torch.zeros(steps, n, device="cuda"), torch.zeros(steps, n, device="cuda"),
                torch.zeros(steps, n, dtype=torch.uint8, device="cuda"), torch.zeros(steps, n, dtype=torch.uint8, device="cuda"),
                torch.zeros(steps, n, dtype=torch.uint8, device="cuda"))

    eager, graphed = VecEnv(n, seed=4, max_episode_steps=3), VecEnv(n, seed=4, max_episode_steps=3)
    o1 = eager.reset().clone(); o2 = graphed.reset().clone()
    b1, b2 = bufs(), bufs()
    run(eager, o1, *b1, torch.cuda.current_stream().cuda_stream)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=side):
        run(graphed, o2, *b2, torch.cuda.current_stream().cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    for x, y in zip(b1, b2):
        assert torch.equal(x, y)
    assert b1[3].sum() + b1[4].sum() + b1[5].sum() > 0      # episodes ended (cap 3): auto-reset ran inside the graph


@pytest.mark.parametrize("beams,n,lanes", [(1, 33, 0), (2, 31, 16), (3, 1, 0), (11, 257, 4), (37, 65, 0)])
def test_odd_beam_counts_and_ragged_batches(beams, n, lanes):
    """Edge shapes: 1 / 2 / 3 beams (lidar features repeat beams, idx_i = int(i * L / 10)), 11 and 37
    beams (first counts of the padded and the warp-per-agent variants), batch sizes that do not fill a
    warp or a CTA."""
    cfg = _capi.default_cfg(n)
    cfg.num_beams, cfg.seed, cfg.max_episode_steps, cfg.lanes_per_agent = beams, 31, 15, lanes
    seg = maps.get_map("stage_2")
    env = _vec_from_cfg(cfg, seg)
    sim = binding.OracleSim(cfg, seg)
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    for t in range(40):
        act = binding.scripted_actions(5, 0, t, n)
        act[:, 0] = np.maximum(act[:, 0], 0.8)
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive.cpu().numpy(), a_ref)
        np.testing.assert_array_equal(env.trunc.cpu().numpy(), tr_ref)
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew.cpu().numpy(), r_ref, atol=REW_ATOL, rtol=0)
    np.testing.assert_array_equal(env.get_state(_capi.F_X), sim.arr["x"])
    out = env.rollout_scripted(1, action_seed=3)          # a one-step fused launch
    o_ref, *_ = sim.step(binding.scripted_actions(3, 0, 0, n))
    np.testing.assert_allclose(out["obs"][0].cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0)


def test_open_world_single_wall_and_oversized_map():
    """One free-standing two-sided wall (no closed boxes): beams see it from both sides; robots
    that drive away never collide.  A map that cannot fit in shared memory is refused."""
    n = 200
    cfg = _capi.default_cfg(n)
    cfg.seed, cfg.max_episode_steps = 2, 30
    wall = np.array([[1.0, -2.0, 1.0, 2.0]])
    env = VecEnv(n, map=wall, cfg=cfg, closed_boxes=False)
    sim = binding.OracleSim(cfg, wall, closed_boxes=False)
    np.testing.assert_allclose(env.reset().cpu().numpy(), sim.reset(), atol=OBS_ATOL, rtol=0)
    hits = 0
    for t in range(60):
        act = binding.scripted_actions(9, 0, t, n)
        act[: n // 2, 0] = 1.0; act[: n // 2, 1] = 0.0        # half the robots drive straight at the wall
        o_ref, r_ref, d_ref, a_ref, tr_ref = sim.step(act)
        obs, rew, done, arrive = env.step(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(done.cpu().numpy(), d_ref, err_msg=f"t={t}")
        np.testing.assert_allclose(obs.cpu().numpy(), o_ref, atol=OBS_ATOL, rtol=0, err_msg=f"t={t}")
        hits += int(d_ref.sum())
    assert hits > 0
    np.testing.assert_array_equal(env.scan().cpu().numpy(), sim.scan())
    big = maps.synthetic_map(8000, seed=1, extent=60.0)     # 32,000 walls = 1 MB of wall records
    with pytest.raises(_capi.NavError):
        VecEnv(16, map=big)


def test_async_host_steps_equal_blocking_host_steps_and_mix_with_device_calls():
    """navsim_step_host_async / navsim_wait (pipeline of depth 4, observations through the copy engine) deliver the
    same results as the blocking navsim_step_host, refuse a step when the pipeline is full, and can be mixed with device
    entry points on the same handle without explicit synchronisation (the library orders its own streams against
    the caller's)."""
    n = 3000
    a, b = VecEnv(n, map="stage_2", seed=4, max_episode_steps=30), VecEnv(n, map="stage_2", seed=4, max_episode_steps=30)
    depth = 4                                                        # NAVSIM_ASYNC_DEPTH
    sets = [a.alloc_host_buffers() for _ in range(depth)]
    ref = b.alloc_host_buffers()
    np.testing.assert_array_equal(a.reset_host(), b.reset_host())
    tickets = []
    for t in range(40):
        hb = sets[t % depth]
        if len(tickets) == depth:
            with pytest.raises(_capi.NavError):
                a.step_host_async(hb["act"], hb)                     # the pipeline is full
            tk, t_old = tickets.pop(0)
            a.wait(tk)
            ref["act"][:] = binding.scripted_actions(9, 0, t_old, n)
            b.step_host(ref["act"], out=ref)
            for key in ("obs", "rew", "done", "arrive", "trunc"):
                np.testing.assert_array_equal(sets[t_old % depth][key], ref[key], err_msg=f"{key} t={t_old}")
        hb["act"][:] = binding.scripted_actions(9, 0, t, n)
        tickets.append((a.step_host_async(hb["act"], hb), t))
    a.wait(0)
    for tk, t_old in tickets:
        ref["act"][:] = binding.scripted_actions(9, 0, t_old, n)
        b.step_host(ref["act"], out=ref)
        np.testing.assert_array_equal(sets[t_old % depth]["obs"], ref["obs"])
    # device call right after an asynchronous host step, host call right after a device call: no explicit sync
    hb = sets[0]
    hb["act"][:] = 0.5
    a.step_host_async(hb["act"], hb)
    obs_d, *_ = a.step(torch.full((n, 2), 0.5, device="cuda"))
    o_h, *_ = a.step_host(np.full((n, 2), 0.5, np.float32))
    b.step_host(np.full((n, 2), 0.5, np.float32)); o2 = b.step_host(np.full((n, 2), 0.5, np.float32))[0].copy()
    np.testing.assert_array_equal(obs_d.cpu().numpy(), o2)
    np.testing.assert_array_equal(o_h, b.step_host(np.full((n, 2), 0.5, np.float32))[0])
    np.testing.assert_array_equal(a.get_state(_capi.F_X), b.get_state(_capi.F_X))


def test_pipelined_host_step_delivers_the_step_issued_three_calls_earlier():
    """navsim_step_host_pipelined = step_host_async + the wait of a steady pipeline in one call: when call t returns,
    the results of step t - 3 sit in the buffer set the next call will overwrite, equal to the blocking call's."""
    n = 2500
    a, b = VecEnv(n, map="stage_1", seed=8, max_episode_steps=25), VecEnv(n, map="stage_1", seed=8, max_episode_steps=25)
    depth = 4
    sets = [a.alloc_host_buffers() for _ in range(depth)]
    ref = b.alloc_host_buffers()
    np.testing.assert_array_equal(a.reset_host(), b.reset_host())
    checked = 0
    steps = 30
    for t in range(steps):
        hb = sets[t % depth]
        hb["act"][:] = binding.scripted_actions(5, 0, t, n)
        assert a.step_host_pipelined(hb["act"], hb) == t + 1
        if t >= depth - 1:                                   # step t - 3 is complete: its set is the next one in turn
            t_old = t - (depth - 1)
            ref["act"][:] = binding.scripted_actions(5, 0, t_old, n)
            b.step_host(ref["act"], out=ref)
            for key in ("obs", "rew", "done", "arrive", "trunc"):
                np.testing.assert_array_equal(sets[(t + 1) % depth][key], ref[key], err_msg=f"{key} t={t_old}")
            checked += 1
    a.wait(0)
    for t_old in range(steps - (depth - 1), steps):
        ref["act"][:] = binding.scripted_actions(5, 0, t_old, n)
        b.step_host(ref["act"], out=ref)
        np.testing.assert_array_equal(sets[t_old % depth]["obs"], ref["obs"])
    assert checked == steps - (depth - 1)
