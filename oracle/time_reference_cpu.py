"""time_reference_cpu.py — TEST INFRASTRUCTURE.  BASELINE.json configs[0]: the reference's own
Gazebo-free CPU path (unmodified environment_new.Env over the fake-ROS shim + unmodified ppo.PPO),
seed 0, timed in the container that has /root/reference (it cannot travel to the GPU box):

  * Env.step only, scripted actions               -> env-steps/s per core
  * PPO.rollout (batch-1 actor + MultivariateNormal + Env.step)
  * PPO update, timesteps_per_batch = 5000, 50 epochs (arguments.py:31, main.py:471)

Writes one JSON object (profiles/r01_reference_cpu_container.json when run with --out)."""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import random
import sys
import tempfile
import time

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from navbot_ppo_b200 import maps  # noqa: E402
from oracle import fake_ros  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env-steps", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=5000)
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--out", default="")
    ap.add_argument("--procs", type=int, default=0,
                    help="also run P independent processes of the Env.step-only loop side by side (one robot each, pinned "
                         "to one core each) and report their aggregate env-steps/s")
    ap.add_argument("--env-only", action="store_true", help="(internal) only the Env.step loop; prints its seconds")
    args = ap.parse_args()
    if not fake_ros.reference_available():
        raise SystemExit("reference sources not found")
    fake_ros.install_stubs()
    import torch
    import net_actor
    import net_critic
    import ppo as ppo_mod
    torch.set_num_threads(1)
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    seg = maps.get_map("stage_1")
    ref = fake_ros.RefEnv(seg, seed=0)
    rng = np.random.RandomState(0)
    # ---- Env.step alone
    ref.reset()
    past = np.zeros(2)
    acts = np.stack([rng.uniform(0, 1, 1000), rng.uniform(-1, 1, 1000)], axis=1)
    for t in range(1000):
        _, _, d, a = ref.step(acts[t], past); past = acts[t]
        if d or a:
            ref.reset(); past = np.zeros(2)
    t0 = time.perf_counter()
    n_ep = 0
    for t in range(args.env_steps):
        _, _, d, a = ref.step(acts[t % 1000], past); past = acts[t % 1000]
        if d or a or (t % 500 == 499):
            ref.reset(); past = np.zeros(2); n_ep += 1
    env_dt = time.perf_counter() - t0
    if args.env_only:
        print(f"ENV_ONLY {env_dt:.6f}")
        return
    multi = None
    if args.procs > 1:
        import subprocess
        cpus = sorted(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        ps = [subprocess.Popen(["taskset", "-c", str(cpus[i % len(cpus)]), sys.executable, os.path.abspath(__file__), "--env-only",
                                "--env-steps", str(args.env_steps)], stdout=subprocess.PIPE, text=True) for i in range(args.procs)]
        secs = [float([l for l in p_.communicate()[0].splitlines() if l.startswith("ENV_ONLY")][0].split()[1]) for p_ in ps]
        wall = time.perf_counter() - t0
        multi = {"processes": args.procs, "cpus_available": len(cpus), "steps_per_process": args.env_steps,
                 "loop_seconds_min_max": [min(secs), max(secs)], "wall_seconds_including_imports": wall,
                 "aggregate_env_steps_per_s": sum(args.env_steps / s_ for s_ in secs)}
    # ---- PPO.rollout + the update of one learn() iteration
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        agent = ppo_mod.PPO(net_actor.NetActor, net_critic.NetCritic, fake_ros.RefEnv(seg, seed=0).env, 16, 2,
                            timesteps_per_batch=args.batch, max_timesteps_per_episode=500, gamma=0.99,
                            n_updates_per_iteration=args.epochs, lr=3e-4, clip=0.2, render=False, save_freq=1000, seed=0,
                            method_name="timing", output_dir=tmp)
        t0 = time.perf_counter()
        agent.learn(total_timesteps=1, past_action=[0, 0])      # exactly one iteration
        learn_dt = time.perf_counter() - t0
        lg = agent.logger
        rollout_t = float(lg.get("rollout_time", float("nan")))
        update_t = float(lg.get("update_time", float("nan")))
    out = {
        "what": "reference project_ppo code (unmodified) over oracle/fake_ros.py, 1 thread, build container (no GPU)",
        "cores": 1, "torch_threads": 1, "host_cpus": os.cpu_count(),
        "env_step_only": {"steps": args.env_steps, "seconds": env_dt, "env_steps_per_s": args.env_steps / env_dt,
                          "us_per_step": 1e6 * env_dt / args.env_steps},
        "env_step_only_P_processes": multi,
        "learn_iteration": {"timesteps_per_batch": args.batch, "epochs": args.epochs, "seconds": learn_dt,
                            "rollout_seconds": rollout_t, "update_seconds": update_t,
                            "rollout_env_steps_per_s": args.batch / rollout_t if rollout_t == rollout_t else None,
                            "train_env_steps_per_s": args.batch / learn_dt},
    }
    s = json.dumps(out, indent=1)
    print(s)
    if args.out:
        open(args.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
