"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

from navbot_ppo_b200 import _capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# obs/reward tolerance of the parity contract (BASELINE.json north_star: "within 1e-5 for
# float state/reward"); flags are compared exactly.
OBS_ATOL = 1e-5
# reward is 500 * (delta distance): same absolute tolerance scaled by its magnitude range
REW_ATOL = 1e-4


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def cfg_from_golden(g, auto_reset=1):
    n = g["act"].shape[1]
    cfg = _capi.default_cfg(n)
    cfg.seed = int(g["seed"])
    cfg.auto_reset = auto_reset
    if "max_episode_steps" in g.files:
        cfg.max_episode_steps = int(g["max_episode_steps"])
    else:
        cfg.max_episode_steps = 1 << 30
    if "arrive_threshold" in g.files:
        cfg.arrive_threshold = float(g["arrive_threshold"])
    if "start" in g.files:
        cfg.start_x, cfg.start_y, cfg.start_theta = (float(v) for v in g["start"])
    if "num_beams" in g.files:
        cfg.num_beams = int(g["num_beams"])
    return cfg
