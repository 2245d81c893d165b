"""make_golden_sampler.py — TEST INFRASTRUCTURE.  Records tests/golden/sampler_tables.npz by running the reference's
own `GoalSpawnSampler` (project_ppo/src/spawn_goal_sampler.py:37-72, imported unmodified from /root/reference).

    python oracle/make_golden_sampler.py      # needs /root/reference; run in the build container

The class draws table indices with `self.rng.randint(len(table))` from numpy's Mersenne Twister; the simulator
draws them from its per-agent Philox stream (navsim_math.h: nv_table_indices).  To pin the LOGIC the class applies
to those indices — distance window, 100 attempts, unconditional fallback, draws consumed — the harness replaces
`sampler.rng` by an object whose randint() hands out the simulator's index stream of (seed, agent), and records
the (start pose, goal point) sequence each agent gets, plus the number of index pairs consumed after every call.
Cases: both world types with the class defaults (1.5 .. 6.0 m), and a 1.9 .. 2.0 m window on stage1 (1.4 % of the pairs, one of them exactly 2.0 m apart) that most
attempts fail, so that the 100-attempt cut-off and the fallback are exercised."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import binding, fake_ros  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class PhiloxIndexStream:
    """randint(n) -> the simulator's table indices: call 2k = start index, call 2k + 1 = goal index of draw k."""

    def __init__(self, seed, agent, n_starts, n_goals):
        self.seed, self.agent, self.n = seed, agent, (n_starts, n_goals)
        self.draws, self._pending = 0, None

    def randint(self, n):
        if self._pending is None:
            i_s, i_g = ctypes.c_int(), ctypes.c_int()
            binding.lib().oracle_table_indices(self.seed, self.agent, self.draws, self.n[0], self.n[1], ctypes.byref(i_s),
                                               ctypes.byref(i_g))
            self.draws += 1
            assert n == self.n[0], "the class draws the start index first (spawn_goal_sampler.py:55)"
            self._pending = i_g.value
            return i_s.value
        assert n == self.n[1]
        v, self._pending = self._pending, None
        return v


def record(ref_mod, world, lo, hi, seed, agents, episodes):
    rows = np.zeros((agents, episodes, 5))
    draws = np.zeros((agents, episodes), np.int64)
    for a in range(agents):
        s = ref_mod.GoalSpawnSampler(world_type=world, min_dist=lo, max_dist=hi, seed=0)
        s.rng = PhiloxIndexStream(seed, a, len(s.start_poses), len(s.goal_points))
        for e in range(episodes):
            start, goal = s.sample_start_and_goal()
            rows[a, e] = (*start, *goal)
            draws[a, e] = s.rng.draws
    return rows, draws


def main():
    if not fake_ros.reference_available():
        raise SystemExit("reference sources not found")
    sys.path.insert(0, fake_ros.REF_SRC)
    import spawn_goal_sampler as ref_mod     # the reference's file, unmodified
    out = {}
    for tag, world, lo, hi, seed in (("house_default", "small_house", 1.5, 6.0, 5), ("stage1_default", "stage1", 1.5, 6.0, 6),
                                     ("stage1_tight", "stage1", 1.9, 2.0, 7)):
        rows, draws = record(ref_mod, world, lo, hi, seed, agents=24, episodes=12)
        out[tag + "_rows"], out[tag + "_draws"] = rows, draws
        out[tag + "_cfg"] = np.asarray([lo, hi, seed], np.float64)
        out[tag + "_world"] = np.asarray(world)
        per_call = np.diff(np.concatenate([np.zeros((24, 1), np.int64), draws], 1), axis=1)
        print(tag, "index pairs per call: min", per_call.min(), "max", per_call.max(), "fallbacks", int((per_call == 101).sum()))
    np.savez_compressed(os.path.join(GOLD, "sampler_tables.npz"), **out)
    print("sampler_tables.npz", os.path.getsize(os.path.join(GOLD, "sampler_tables.npz")))


if __name__ == "__main__":
    main()
