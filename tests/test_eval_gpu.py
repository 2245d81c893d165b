"""Evaluation mode (main.py:135-252): the one-robot loop on navbot_ppo_b200.Env and the batched form reproduce
the episodes the reference's own main.evaluate played over the CPU shim (tests/golden/eval_stage_1.npz,
recorded by oracle/make_golden_eval.py), and agree with each other and with a hand-rolled loop over the C-ABI."""
import csv
import os

import numpy as np
import pytest
import torch

from navbot_ppo_b200 import _capi
from navbot_ppo_b200.env import Env, VecEnv
from navbot_ppo_b200.evaluate import EVAL_COLUMNS, evaluate, evaluate_vec
from navbot_ppo_b200.nets import NetActor
from tests.helpers import golden

pytestmark = pytest.mark.gpu


def _actor(seed=3):
    torch.manual_seed(seed)
    return NetActor(16, 2)


def _manual_episode(actor, seed, agent, max_len):
    env = VecEnv(1, seed=seed, max_episode_steps=max_len, auto_reset=False, is_training=False, agent_id_offset=agent)
    obs = env.reset()
    ret, path, prev = 0.0, 0.0, None
    for t in range(max_len):
        pos = np.array([env.get_state(_capi.F_X)[0], env.get_state(_capi.F_Y)[0]])
        if prev is not None:
            path += np.linalg.norm(pos - prev)
        prev = pos
        obs, rew, done, arrive = env.step(actor(obs))
        ret += float(rew[0])
        if bool(done[0]) or bool(arrive[0]):
            break
    return dict(length=t + 1, ret=ret, path=path, success=bool(arrive[0]), collision=bool(done[0]) and not bool(arrive[0]))


def test_batched_evaluation_equals_one_robot_at_a_time(tmp_path):
    actor = _actor()
    n, max_len = 6, 60
    m = evaluate_vec(actor, n, seed=17, max_timesteps_per_episode=max_len, output_dir=str(tmp_path), method_name="ev",
                     is_training=False)
    assert m["success"] + m["collision"] + m["timeout"] == n
    for e in range(n):
        ref = _manual_episode(actor, 17, e, max_len)
        assert m["lengths"][e] == ref["length"]
        assert bool(m["per_episode"]["success"][e]) == ref["success"] and bool(m["per_episode"]["collision"][e]) == ref["collision"]
        assert abs(m["returns"][e] - ref["ret"]) < 1e-2
        assert abs(m["path_lengths"][e] - ref["path"]) < 1e-4
    rows = list(csv.reader(open(os.path.join(tmp_path, "ev", "logs", "ev_eval_episodes.csv"))))
    assert rows[0] == EVAL_COLUMNS and len(rows) == n + 1


def test_reference_style_evaluate_loads_latest_checkpoint(tmp_path):
    actor = _actor(5)
    ck = tmp_path / "m1" / "checkpoints"
    ck.mkdir(parents=True)
    torch.save({k: v.detach().cpu().clone() for k, v in _actor(4).state_dict().items()}, ck / "actor_iter0001_step00000100.pth")
    torch.save({k: v.detach().cpu().clone() for k, v in actor.state_dict().items()}, ck / "actor_iter0002_step00000200.pth")
    hp = dict(method_name="m1", exp_id="x", max_timesteps_per_episode=40, output_dir=str(tmp_path))
    env = Env(False, seed=23)
    m = evaluate(env, hp, "", "", 2, verbose=False)
    assert len(m["lengths"]) == 2 and m["success"] + m["collision"] + m["timeout"] == 2
    # episode 0 is robot 0 of the batched evaluation with the same seed and the LATEST checkpoint's weights
    mv = evaluate_vec(actor, 1, seed=23, max_timesteps_per_episode=40, is_training=False)
    assert m["lengths"][0] == mv["lengths"][0]
    assert abs(m["returns"][0] - mv["returns"][0]) < 1e-2 and abs(m["path_lengths"][0] - mv["path_lengths"][0]) < 1e-4
    rows = list(csv.reader(open(os.path.join(tmp_path, "m1", "logs", "m1_eval_episodes.csv"))))
    assert rows[0] == EVAL_COLUMNS and len(rows) == 3


def _golden_actor(g, name):
    sd = {k.split("__sd__")[1]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith(name + "__sd__")}
    net = NetActor(16, 2)
    net.load_state_dict(sd)
    return net, sd


def _check_rows(m, want, tag):
    """want[e] = (success, collision, timeout, length, return, path_length) as the reference's csv has them."""
    n = len(want)
    assert len(m["lengths"]) == n
    got_out = np.stack([np.asarray(m["per_episode"][k])[:n] for k in ("success", "collision", "timeout")], 1)
    for e in range(n):
        assert m["lengths"][e] == int(want[e, 3]), (tag, e, m["lengths"][e], want[e, 3])
        # rewards leave the simulator as fp32 (500 x a distance difference, +-100 / 120): 1e-4 per step
        assert abs(m["returns"][e] - want[e, 4]) <= 1e-4 * want[e, 3] + 1e-3, (tag, e)
        assert abs(m["path_lengths"][e] - want[e, 5]) <= 1e-5 * max(1.0, want[e, 5]) + 2e-5, (tag, e)
        assert tuple(int(v) for v in got_out[e]) == tuple(int(v) for v in want[e, :3]), (tag, e)
    assert (m["success"], m["collision"], m["timeout"]) == tuple(int(v) for v in want[:, :3].sum(0)), tag


@pytest.mark.parametrize("name", ["init", "steer", "ram"])
def test_evaluation_reproduces_the_reference_main_evaluate(name, tmp_path):
    """Lengths, outcomes, returns and path lengths of the episodes `main.evaluate` played with three fixed
    checkpoints (timeouts with the seeded-initialised actor, arrivals with a goal-steering one, collisions with a
    full-speed one): `seq` = one robot, episode after episode (python main.py --eval) vs evaluate() on the one-robot
    Env loading the checkpoint file; `vec` = robot e plays episode e vs evaluate_vec()."""
    g = golden("eval_stage_1")
    seed, max_len = int(g["seed"]), int(g["max_len"])
    actor, sd = _golden_actor(g, name)
    ck = tmp_path / name / "checkpoints"
    ck.mkdir(parents=True)
    torch.save(sd, ck / "actor_iter0001_step00000001.pth")
    hp = dict(method_name=name, exp_id="golden", max_timesteps_per_episode=max_len, output_dir=str(tmp_path), state_dim=16)
    seq = np.asarray(g[name + "_seq"])
    m = evaluate(Env(True, seed=seed), hp, "", "", len(seq), verbose=False)       # main.py:22,442: is_training = True
    rows = list(csv.reader(open(os.path.join(tmp_path, name, "logs", f"{name}_eval_episodes.csv"))))
    m["per_episode"] = {k: np.asarray([int(r[i]) for r in rows[1:]]) for i, k in ((1, "success"), (2, "collision"), (3, "timeout"))}
    _check_rows(m, seq, name + "/seq")
    vec = np.asarray(g[name + "_vec"])
    mv = evaluate_vec(actor, len(vec), seed=seed, max_timesteps_per_episode=max_len, is_training=True)
    _check_rows(mv, vec, name + "/vec")
    assert float(g["arrive_threshold"]) == 0.2
