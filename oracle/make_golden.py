"""make_golden.py — TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by EXECUTING THE
REFERENCE'S OWN CODE (project_ppo/src/environment_new.py, ppo.py, net_actor.py,
net_critic.py, imported unmodified from /root/reference) over oracle/fake_ros.py.

    python oracle/make_golden.py            # needs /root/reference; run in the build container

The fixtures travel to the GPU box, where /root/reference does not exist.

  env_rollout_<map>.npz   A agents x T steps under PPO.rollout's episode protocol
                          (ppo.py:486-593): obs fed to the policy, reward, done, arrive,
                          timeout, pose/goal after every step.  Actions: a mix of random,
                          wall-seeking, goal-seeking and spinning controllers so that
                          collisions, arrivals, timeouts, all quadrants of the bearing
                          computation and many rounding boundaries occur.
  env_rollout_house*.npz  the same on the 208-wall house map (10 beams, and 36 beams through the
                          reference's lidar subsampling rule)
  env_raw_stage_1.npz     Env.step used without resets (DDPG/TD3-style callers):
                          arrival respawns the goal inside step (environment_new.py:245-267).
  ppo_*.npz               see make_golden_ppo() — nets, reward-to-go, evaluate, update epochs.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from navbot_ppo_b200 import maps  # noqa: E402
from oracle import fake_ros  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def controller(kind, rng, st, t):
    """Scripted action generators (float32, already inside the clamp box of ppo.py:700-703)."""
    if kind == 0:  # uniform random
        return np.array([rng.uniform(0, 1), rng.uniform(-1, 1)], np.float32)
    if kind == 1:  # full speed ahead, slight drift: ends in a wall
        return np.array([1.0, rng.uniform(-0.05, 0.05)], np.float32)
    if kind == 2:  # steer to the goal: ends in an arrival
        bearing = math.atan2(st["gy"] - st["y"], st["gx"] - st["x"])
        err = (bearing - st["th"] + math.pi) % (2 * math.pi) - math.pi
        return np.array([min(1.0, 0.2 + 0.8 * max(0.0, math.cos(err))), float(np.clip(2.0 * err, -1, 1))], np.float32)
    if kind == 3:  # spin in place / crawl: ends in a timeout
        return np.array([0.02 * (t % 3), rng.choice([-1.0, 1.0, 0.37])], np.float32)
    raise ValueError(kind)


def gen_env_rollout(map_name, agents, steps, max_ep, seed, is_training=True, num_beams=10):
    seg = maps.get_map(map_name)
    start = maps.SPAWN.get(map_name, (0.0, 0.0, 0.0))
    rng = np.random.RandomState(1234 + seed)
    envs = [fake_ros.RefEnv(seg, seed=seed, agent=a, is_training=is_training, start=start, num_beams=num_beams)
            for a in range(agents)]
    rec = {k: [] for k in ("act", "obs_next", "obs_step", "rew", "done", "arrive", "trunc", "x", "y", "th", "gx", "gy",
                           "past", "draws")}
    obs0 = np.stack([e.reset() for e in envs])
    past = [np.zeros(2, np.float32) for _ in envs]
    rounds = [0] * agents
    kinds = [a % 4 for a in range(agents)]
    for t in range(steps):
        row = {k: [] for k in rec}
        for a, e in enumerate(envs):
            act = controller(kinds[a], rng, e.state(), rounds[a])
            obs, rew, done, arrive = e.step(act, past[a])      # ppo.py:541
            past[a] = act                                        # ppo.py:543
            rounds[a] += 1                                       # ppo.py:549
            timeout = rounds[a] >= max_ep                        # ppo.py:552
            obs_next = obs
            if done or arrive or timeout:                        # ppo.py:553
                rounds[a] = 0
                past[a] = np.zeros(2, np.float32)                # ppo.py:591
                obs_next = e.reset()                             # ppo.py:593
                kinds[a] = int(rng.randint(0, 4))
            st = e.state()
            row["act"].append(act); row["obs_next"].append(obs_next); row["obs_step"].append(obs)
            row["rew"].append(rew); row["done"].append(done); row["arrive"].append(arrive)
            row["trunc"].append(timeout and not done and not arrive)
            for k in ("x", "y", "th", "gx", "gy", "past", "draws"):
                row[k].append(st[k])
        for k in rec:
            rec[k].append(np.asarray(row[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(obs0=obs0, segments=seg, seed=np.int64(seed), max_episode_steps=np.int32(max_ep),
               arrive_threshold=np.float64(0.2 if is_training else 0.4), start=np.asarray(start, np.float64),
               num_beams=np.int32(num_beams))
    return out


def gen_env_raw(map_name, agents, steps, seed):
    """No resets on arrival: the goal respawns inside Env.step; collisions get an explicit
    reset() call (recorded as a mask) on the following step boundary."""
    seg = maps.get_map(map_name)
    rng = np.random.RandomState(99 + seed)
    envs = [fake_ros.RefEnv(seg, seed=seed, agent=a) for a in range(agents)]
    obs0 = np.stack([e.reset() for e in envs])
    past = [np.zeros(2, np.float32) for _ in envs]
    rec = {k: [] for k in ("act", "obs", "rew", "done", "arrive", "reset_mask", "obs_reset", "gx", "gy", "past", "draws")}
    for t in range(steps):
        row = {k: [] for k in rec}
        for a, e in enumerate(envs):
            act = controller(2 if a % 2 == 0 else 1, rng, e.state(), t)
            obs, rew, done, arrive = e.step(act, past[a])
            past[a] = act
            row["act"].append(act); row["obs"].append(obs); row["rew"].append(rew)
            row["done"].append(done); row["arrive"].append(arrive)
            if done:
                row["reset_mask"].append(1)
                row["obs_reset"].append(e.reset())
                past[a] = np.zeros(2, np.float32)
            else:
                row["reset_mask"].append(0)
                row["obs_reset"].append(np.zeros(16))
            st = e.state()
            for k in ("gx", "gy", "past", "draws"):
                row[k].append(st[k])
        for k in rec:
            rec[k].append(np.asarray(row[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(obs0=obs0, segments=seg, seed=np.int64(seed))
    return out


def main():
    if not fake_ros.reference_available():
        raise SystemExit("reference sources not found; golden vectors can only be generated where /root/reference exists")
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "env_rollout_stage_1.npz"), **gen_env_rollout("stage_1", 32, 300, 90, seed=0))
    np.savez_compressed(os.path.join(GOLD, "env_rollout_stage_2.npz"), **gen_env_rollout("stage_2", 24, 200, 50, seed=7))
    np.savez_compressed(os.path.join(GOLD, "env_rollout_stage_1_eval.npz"),
                        **gen_env_rollout("stage_1", 8, 200, 90, seed=3, is_training=False))
    np.savez_compressed(os.path.join(GOLD, "env_raw_stage_1.npz"), **gen_env_raw("stage_1", 12, 260, seed=5))
    # house geometry (208 walls, spawn (-3, 1)) under the reference Env's own hard-coded goal sampler
    # (environment_new.py:337-345), and a 36-beam scan through its subsampling rule idx_i = int(i*L/10)
    # (:292-294)
    np.savez_compressed(os.path.join(GOLD, "env_rollout_house.npz"), **gen_env_rollout("house", 16, 160, 60, seed=11))
    np.savez_compressed(os.path.join(GOLD, "env_rollout_house_36beams.npz"),
                        **gen_env_rollout("house", 12, 120, 50, seed=12, num_beams=36))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
    if "--env-only" not in sys.argv:
        try:
            from oracle.make_golden_ppo import main as ppo_main
        except ImportError:
            return
        ppo_main()


if __name__ == "__main__":
    main()
