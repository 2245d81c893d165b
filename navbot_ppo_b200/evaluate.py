"""Evaluation mode of the reference's main.py (project_ppo/src/main.py:135-252).

  evaluate(env, hyperparameters, actor_model, critic_model, num_episodes)
      the reference's function on the one-robot Env: latest `actor_iter*_step*.pth`, deterministic
      policy (mean action, main.py:197-199), one episode after another, per-episode rows in
      `<method>_eval_episodes.csv` with the reference's columns, summary on stdout.
  evaluate_vec(actor, num_episodes, ...)
      the same protocol for `num_episodes` robots at once on the GPU: episode e is robot e of a
      VecEnv; every step is one policy forward + one simulator launch for all robots still running.
      `is_training` picks the arrival threshold like environment_new.py:44-47 (0.2 / 0.4); the default is
      True because main.py builds its Env with the module constant is_training = True (main.py:22,442) for
      `--eval` runs too.

Both return the metrics dict the reference accumulates (success / collision / timeout counts and the
per-episode lists).
"""
from __future__ import annotations

import csv
import glob
import os
import sys
import time

import numpy as np
import torch

from . import _capi
from .env import VecEnv
from .nets import NetActor

EVAL_COLUMNS = ["episode", "success", "collision", "timeout", "length", "return", "path_length", "time"]


def _latest_actor(output_dir, method_name):
    ckpt_dir = os.path.join(output_dir, method_name, "checkpoints")
    paths = sorted(glob.glob(os.path.join(ckpt_dir, "actor_iter*_step*.pth")))       # main.py:158
    if not paths:
        paths = sorted(glob.glob(os.path.join(ckpt_dir, "actor_step*.pth")))         # main.py:161
    return paths[-1] if paths else ""


def _print_summary(method_name, num_episodes, m, path):
    n = max(num_episodes, 1)
    print("\n" + "=" * 60 + "\nEVALUATION SUMMARY\n" + "=" * 60, flush=True)
    print(f"Method: {method_name}\nEpisodes: {num_episodes}", flush=True)
    print(f"Success Rate: {m['success'] / n * 100:.2f}%\nCollision Rate: {m['collision'] / n * 100:.2f}%\n"
          f"Timeout Rate: {m['timeout'] / n * 100:.2f}%", flush=True)
    for label, key in (("Episode Length", "lengths"), ("Return", "returns"), ("Path Length", "path_lengths")):
        print(f"Mean {label}: {np.mean(m[key]):.2f} ± {np.std(m[key]):.2f}", flush=True)
    print(f"Mean Episode Time: {np.mean(m['times']):.2f}s ± {np.std(m['times']):.2f}s", flush=True)
    print(f"Results saved to: {path}\n" + "=" * 60, flush=True)


def _open_csv(output_dir, method_name):
    log_dir = os.path.join(output_dir, method_name, "logs")
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"{method_name}_eval_episodes.csv")
    with open(path, "w", newline="") as f:
        csv.writer(f).writerow(EVAL_COLUMNS)                                          # main.py:178-180
    return path


def evaluate(env, hyperparameters, actor_model, critic_model, num_episodes, verbose=True):
    """main.py:135-252 on a one-robot Env (navbot_ppo_b200.Env or anything with its surface)."""
    method_name = hyperparameters.get("method_name", "baseline")
    state_dim = hyperparameters.get("state_dim", _capi.OBS_DIM)
    output_dir = hyperparameters.get("output_dir") or "runs"
    if actor_model == "":
        actor_model = _latest_actor(output_dir, method_name)
        if not actor_model:
            print("No checkpoint found for evaluation. Exiting.", flush=True)
            sys.exit(0)
    if verbose:
        print(f"Loading actor: {actor_model}", flush=True)
    policy = NetActor(state_dim, _capi.ACT_DIM)
    policy.load_state_dict(torch.load(actor_model))
    policy.eval()
    path = _open_csv(output_dir, method_name)
    m = {"success": 0, "collision": 0, "timeout": 0, "lengths": [], "returns": [], "path_lengths": [], "times": []}
    max_len = hyperparameters["max_timesteps_per_episode"]
    for ep in range(num_episodes):
        t0 = time.time()
        obs = env.reset()
        done = arrive = False
        ep_return, ep_length, path_length = 0, 0, 0.0
        past_action = np.array([0.0, 0.0])
        prev_pos = None
        while ep_length < max_len:
            with torch.no_grad():
                action = policy(torch.as_tensor(np.asarray(obs, dtype=np.float32))).cpu().numpy()   # deterministic, :197-199
            curr_pos = np.array([env.position.x, env.position.y])                                   # :202-205
            if prev_pos is not None:
                path_length += np.linalg.norm(curr_pos - prev_pos)
            prev_pos = curr_pos
            obs, rew, done, arrive = env.step(action, past_action)
            past_action = action
            ep_return += rew
            ep_length += 1
            if done or arrive:
                break
        ep_time = time.time() - t0
        success = 1 if arrive else 0                                                                # :218-220
        collision = 1 if done and not arrive else 0
        timeout = 1 if (not done and not arrive and ep_length >= max_len) else 0
        m["success"] += success; m["collision"] += collision; m["timeout"] += timeout
        m["lengths"].append(ep_length); m["returns"].append(ep_return)
        m["path_lengths"].append(path_length); m["times"].append(ep_time)
        with open(path, "a", newline="") as f:
            csv.writer(f).writerow([ep, success, collision, timeout, ep_length, ep_return, path_length, ep_time])
    if verbose:
        _print_summary(method_name, num_episodes, m, path)
    return m


def evaluate_vec(actor: NetActor, num_episodes: int, map="stage_1", device=0, seed=0, max_timesteps_per_episode=500,
                 output_dir=None, method_name="baseline", is_training=True, agent_id_offset=0, verbose=False):
    """`num_episodes` evaluation episodes at once: robot e of a VecEnv plays episode e with the
    deterministic policy until done | arrive | timeout (main.py:195-213)."""
    n = int(num_episodes)
    env = VecEnv(n, map=map, device=device, seed=seed, max_episode_steps=max_timesteps_per_episode, auto_reset=False,
                 is_training=is_training, agent_id_offset=agent_id_offset)
    dev = env.device
    actor._ensure_bound(dev)
    t0 = time.time()
    obs = env.reset()
    running = np.ones(n, bool)
    out = {k: np.zeros(n, dt) for k, dt in (("success", np.int64), ("collision", np.int64), ("timeout", np.int64),
                                             ("lengths", np.int64), ("returns", np.float64), ("path_lengths", np.float64),
                                             ("times", np.float64))}
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
    live = torch.ones(n, dtype=torch.float64, device=dev)
    for t in range(max_timesteps_per_episode):
        with torch.no_grad():
            mu = actor(obs)                                         # mean action, main.py:197-199
        obs, rew, done, arrive = env.step(mu)
        ret += rew.double() * live
        d = done.cpu().numpy().astype(bool); a = arrive.cpu().numpy().astype(bool); tr = env.trunc.cpu().numpy().astype(bool)
        ended = running & (d | a | tr)
        if ended.any():
            idx = np.nonzero(ended)[0]
            path = env.get_state(_capi.F_EP_PATH)
            r_host = ret.cpu().numpy()
            out["success"][idx] = a[idx]                            # main.py:218-220
            out["collision"][idx] = d[idx] & ~a[idx]
            out["timeout"][idx] = tr[idx] & ~d[idx] & ~a[idx]
            out["lengths"][idx] = t + 1
            out["returns"][idx] = r_host[idx]
            out["path_lengths"][idx] = path[idx]
            out["times"][idx] = time.time() - t0
            running &= ~ended
            live[torch.from_numpy(idx).to(dev)] = 0.0
        if not running.any():
            break
    m = {"success": int(out["success"].sum()), "collision": int(out["collision"].sum()), "timeout": int(out["timeout"].sum()),
         "lengths": out["lengths"].tolist(), "returns": out["returns"].tolist(), "path_lengths": out["path_lengths"].tolist(),
         "times": out["times"].tolist(), "per_episode": out}
    if output_dir is not None:
        path = _open_csv(output_dir, method_name)
        with open(path, "a", newline="") as f:
            w = csv.writer(f)
            for e in range(n):
                w.writerow([e, out["success"][e], out["collision"][e], out["timeout"][e], out["lengths"][e], out["returns"][e],
                            out["path_lengths"][e], out["times"][e]])
        if verbose:
            _print_summary(method_name, n, m, path)
    env.close()
    return m
