"""make_golden_eval.py — TEST INFRASTRUCTURE.  Records tests/golden/eval_*.npz by running the reference's own
`main.evaluate` (project_ppo/src/main.py:135-252, imported unmodified from /root/reference) over
oracle/fake_ros.py with fixed checkpoints.

    python oracle/make_golden_eval.py        # needs /root/reference; run in the build container

Two checkpoints: the reference NetActor as seeded-initialised (episodes end in collisions and timeouts) and a
hand-set "steer to the goal" actor (residual paths only, heads reading the goal features; episodes end in
arrivals) and a "full speed ahead" one (collisions), each in the reference's state_dict layout.  Per checkpoint:
  seq_*   ONE reference Env (main.py:442: is_training=True, so threshold_arrive 0.2) playing `n_seq` episodes one
          after another, the way `python main.py --eval` does -> navbot_ppo_b200.evaluate.evaluate on the one-robot
          drop-in Env must reproduce the csv rows;
  vec_*   `n_vec` separate reference Envs (robot ids 0..n_vec-1, one episode each) -> evaluate_vec's episode e.
"""
from __future__ import annotations

import csv
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from navbot_ppo_b200 import maps  # noqa: E402
from oracle import fake_ros  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
COLS = ("success", "collision", "timeout", "length", "return", "path_length")


def steering_state_dict(NetActor, seed):
    """A NetActor whose blocks pass their input through (fc2 = 0, so y = lrelu(x)) and whose heads read the goal
    features: forward speed high, turn rate against the heading error obs[15] (environment_new.py:170-176)."""
    torch.manual_seed(seed)
    net = NetActor(16, 2)
    sd = net.state_dict()
    with torch.no_grad():
        for k in ("rb1.fc2.weight", "rb1.fc2.bias", "rb2.fc2.weight", "rb2.fc2.bias"):
            sd[k].zero_()
        sd["out1.weight"].zero_(); sd["out1.bias"].fill_(1.5)        # sigmoid(1.5) = 0.82 of full speed
        sd["out2.weight"].zero_(); sd["out2.bias"].zero_()
        sd["out2.weight"][0, 15] = -6.0                              # cat[X0, X][15] = diff_angle / 180
        sd["out2.weight"][0, 31] = -6.0
    return sd


class FlatActionEnv:
    """The reference's evaluate hands Env.step the actor's [1, 2] output as it is (main.py:197-207), and
    Env.step indexes action[1] (environment_new.py:274): as shipped, `python main.py --eval` stops with an
    IndexError at the first step.  The harness therefore flattens `action` / `past_action` on the way in —
    the one adaptation made to get a recording; everything else is the reference's code."""

    def __init__(self, env):
        self._env = env

    def __getattr__(self, name):
        return getattr(self._env, name)

    def step(self, action, past_action):
        return self._env.step(np.asarray(action).reshape(-1), np.asarray(past_action).reshape(-1))


def ramming_state_dict(NetActor, seed):
    """Full speed straight ahead: every episode ends in the wall the robot faces (collision outcome)."""
    sd = steering_state_dict(NetActor, seed)
    with torch.no_grad():
        sd["out1.bias"].fill_(4.0)
        sd["out2.weight"].zero_()
    return sd


def run_reference_evaluate(main_mod, ref_env, ckpt, episodes, max_len, tmp, tag):
    hp = {"exp_id": "golden", "method_name": tag, "output_dir": tmp, "max_timesteps_per_episode": max_len, "state_dim": 16}
    ref_env._enter()
    main_mod.evaluate(env=FlatActionEnv(ref_env.env), hyperparameters=hp, actor_model=ckpt, critic_model="", num_episodes=episodes)
    rows = list(csv.reader(open(os.path.join(tmp, tag, "logs", f"{tag}_eval_episodes.csv"))))
    assert rows[0] == ["episode", "success", "collision", "timeout", "length", "return", "path_length", "time"]
    return np.asarray([[float(v) for v in r[1:7]] for r in rows[1:]], np.float64)


def main():
    if not fake_ros.reference_available():
        raise SystemExit("reference sources not found; golden vectors can only be generated where /root/reference exists")
    fake_ros.install_stubs()
    import main as ref_main          # the reference's main.py, unmodified
    from net_actor import NetActor   # the reference's class
    seg = maps.get_map("stage_1")
    seed, n_seq, n_vec, max_len = 29, 6, 10, 120
    out = dict(seed=np.int64(seed), max_len=np.int64(max_len), arrive_threshold=np.float64(0.2), columns=np.asarray(COLS))
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(11)
        cks = {"init": NetActor(16, 2).state_dict(), "steer": steering_state_dict(NetActor, 12),
               "ram": ramming_state_dict(NetActor, 13)}
        for name, sd in cks.items():
            path = os.path.join(tmp, f"actor_{name}.pth")
            torch.save(sd, path)
            for k, v in sd.items():
                out[f"{name}__sd__{k}"] = v.detach().cpu().numpy()
            env = fake_ros.RefEnv(seg, seed=seed, agent=0, is_training=ref_main.is_training)
            out[f"{name}_seq"] = run_reference_evaluate(ref_main, env, path, n_seq, max_len, tmp, f"{name}_seq")
            rows = []
            for e in range(n_vec):
                env = fake_ros.RefEnv(seg, seed=seed, agent=e, is_training=ref_main.is_training)
                rows.append(run_reference_evaluate(ref_main, env, path, 1, max_len, tmp, f"{name}_vec{e}")[0])
            out[f"{name}_vec"] = np.stack(rows)
            print(name, "seq outcomes (success, collision, timeout, length):\n", out[f"{name}_seq"][:, :4])
            print(name, "vec outcomes:\n", out[f"{name}_vec"][:, :4])
    np.savez_compressed(os.path.join(GOLD, "eval_stage_1.npz"), **out)
    print("eval_stage_1.npz", os.path.getsize(os.path.join(GOLD, "eval_stage_1.npz")))


if __name__ == "__main__":
    main()
