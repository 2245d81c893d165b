"""CPU: the C restatement (oracle/navsim_oracle.c) against golden traces produced by the
reference's own Env code (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import binding
from tests.helpers import cfg_from_golden, golden


@pytest.mark.parametrize("name", ["env_rollout_stage_1", "env_rollout_stage_2", "env_rollout_stage_1_eval",
                                  "env_rollout_house", "env_rollout_house_36beams"])
def test_oracle_reproduces_reference_rollout(name):
    g = golden(name)
    cfg = cfg_from_golden(g, auto_reset=1)
    sim = binding.OracleSim(cfg, g["segments"])
    obs0 = sim.reset()
    np.testing.assert_array_equal(obs0, g["obs0"])
    T = g["act"].shape[0]
    for t in range(T):
        obs, rew, done, arrive, trunc = sim.step(g["act"][t])
        np.testing.assert_array_equal(done, g["done"][t].astype(np.uint8), err_msg=f"done t={t}")
        np.testing.assert_array_equal(arrive, g["arrive"][t].astype(np.uint8), err_msg=f"arrive t={t}")
        np.testing.assert_array_equal(trunc, g["trunc"][t].astype(np.uint8), err_msg=f"trunc t={t}")
        # same libm, same physics header: the restatement matches the reference bit for bit
        np.testing.assert_array_equal(obs, g["obs_next"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(rew, g["rew"][t], err_msg=f"rew t={t}")
        for k in ("x", "y", "th", "gx", "gy", "past"):
            np.testing.assert_array_equal(sim.arr[k], g[k][t], err_msg=f"{k} t={t}")
        np.testing.assert_array_equal(sim.arr["draws"], g["draws"][t])


def test_oracle_reproduces_reference_raw_env():
    """auto_reset off: arrival respawns the goal inside step (environment_new.py:245-267),
    collisions are followed by an explicit Env.reset()."""
    g = golden("env_raw_stage_1")
    cfg = cfg_from_golden(g, auto_reset=0)
    sim = binding.OracleSim(cfg, g["segments"])
    np.testing.assert_array_equal(sim.reset(), g["obs0"])
    for t in range(g["act"].shape[0]):
        obs, rew, done, arrive, _ = sim.step(g["act"][t])
        np.testing.assert_array_equal(obs, g["obs"][t])
        np.testing.assert_array_equal(rew, g["rew"][t])
        np.testing.assert_array_equal(done, g["done"][t].astype(np.uint8))
        np.testing.assert_array_equal(arrive, g["arrive"][t].astype(np.uint8))
        mask = g["reset_mask"][t].astype(np.uint8)
        if mask.any():
            ro = sim.reset(mask)
            sel = mask.astype(bool)
            np.testing.assert_array_equal(ro[sel], g["obs_reset"][t][sel])
        for k in ("gx", "gy", "past"):
            np.testing.assert_array_equal(sim.arr[k], g[k][t], err_msg=f"{k} t={t}")
        np.testing.assert_array_equal(sim.arr["draws"], g["draws"][t])


def test_golden_covers_the_edge_cases():
    g = golden("env_rollout_stage_2")
    assert g["done"].sum() > 10 and g["arrive"].sum() > 3 and g["trunc"].sum() > 10
    g1 = golden("env_rollout_stage_1")
    assert g1["done"].sum() > 10 and g1["arrive"].sum() > 10 and g1["trunc"].sum() > 10
    # all four quadrants + wrap of the bearing feature were visited
    rel = g1["obs_next"][..., 14]
    assert rel.min() < 0.1 and rel.max() > 0.9
    diff = g1["obs_next"][..., 15]
    assert diff.min() < -0.9 and diff.max() > 0.9


def test_table_sampler_restatement_reproduces_the_reference_class():
    """spawn_goal_sampler.py:52-63 restated in C (oracle: ref_sample_start_and_goal): start pose / goal point
    sequences and index pairs consumed, for 24 robots x 12 resets, against the recording of the reference's own
    GoalSpawnSampler driven by the same index stream (oracle/make_golden_sampler.py) — default window on both
    table sets, and a 1.9 .. 2.0 m window where most calls run into the 100-attempt cut-off."""
    from navbot_ppo_b200 import maps
    g = golden("sampler_tables")
    for tag in ("house_default", "stage1_default", "stage1_tight"):
        rows, draws = g[tag + "_rows"], g[tag + "_draws"]
        lo, hi, seed = (float(v) for v in g[tag + "_cfg"])
        world = str(g[tag + "_world"])
        agents, episodes = draws.shape
        cfg = binding.default_cfg(agents)
        cfg.seed = int(seed)
        cfg.sampler_mode, cfg.sampler_min_dist, cfg.sampler_max_dist = 1, lo, hi
        sim = binding.OracleSim(cfg, maps.get_map("house" if world == "small_house" else "stage_1"),
                                sampler_tables=maps.sampler_tables(world))
        for e in range(episodes):
            sim.reset()
            got = np.stack([sim.arr[k] for k in ("x", "y", "th", "gx", "gy")], 1)
            assert np.array_equal(got, rows[:, e]), (tag, e)
            assert np.array_equal(sim.arr["draws"], draws[:, e]), (tag, e)
        if tag == "stage1_tight":
            per_call = np.diff(np.concatenate([np.zeros((agents, 1), np.int64), draws], 1), axis=1)
            assert (per_call == 101).any() and (per_call < 101).any()      # both the cut-off and early exits occur


def test_oracle_cfg_mirror_matches_the_public_struct():
    """oracle/binding.OracleCfg restates include/navsim.h's navsim_cfg so that bench.py's CPU arm never loads
    the product library: same fields, same defaults, same stage_1 geometry."""
    import ctypes
    from navbot_ppo_b200 import _capi, maps
    assert [f[0] for f in binding.OracleCfg._fields_] == [f[0] for f in _capi.NavsimCfg._fields_]
    assert ctypes.sizeof(binding.OracleCfg) == ctypes.sizeof(_capi.NavsimCfg)
    assert bytes(binding.default_cfg(77)) == bytes(_capi.default_cfg(77))
    assert np.array_equal(binding.stage_1_segments(), maps.get_map("stage_1"))
