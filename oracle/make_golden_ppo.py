"""make_golden_ppo.py — TEST INFRASTRUCTURE.  Records tests/golden/ppo_*.npz by EXECUTING THE
REFERENCE'S OWN net_actor.py / net_critic.py / ppo.py (imported unmodified from
/root/reference/project_ppo/src) with its environment_new.Env running over oracle/fake_ros.py.

    python oracle/make_golden_ppo.py        # needs /root/reference; run in the build container

  ppo_nets.npz        NetActor / NetCritic forward on a fixed batch (weights as the reference
                      initialises them under torch.manual_seed), the state_dict key list.
  ppo_rtgs.npz        PPO.compute_rtgs on ragged episodes.
  ppo_get_action.npz  PPO.get_action: mean, the N(0, I) draw, clamped action, log-prob.
  ppo_learn_<tag>.npz one full PPO.learn iteration (rollout over the reference Env + the
                      inline update loop, ppo.py:245-457): the batch the rollout returned, the
                      flat parameters before / after, per-epoch losses and gradients (taken
                      from the parameters' .grad just before each optimizer step), and the
                      logger's means of approx_kl / clip_frac / grad norms.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from navbot_ppo_b200 import layout, maps  # noqa: E402
from oracle import fake_ros  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _flat(module, kind):
    return layout.flatten_state_dict(module.state_dict(), kind)


def _flat_grad(module, kind):
    named = dict(module.named_parameters())
    return np.concatenate([named[n].grad.detach().cpu().numpy().astype(np.float32).ravel() for n, _ in layout.spec(kind)])


def gen_nets(torch, net_actor, net_critic):
    torch.manual_seed(0)
    actor = net_actor.NetActor(16, 2)
    critic = net_critic.NetCritic(16, 1)
    rng = np.random.RandomState(0)
    obs = rng.uniform(-1, 1, size=(96, 16)).astype(np.float32)
    obs[:, :10] = rng.uniform(0, 1, size=(96, 10))
    with torch.no_grad():
        mu = actor(torch.from_numpy(obs)).numpy()
        v = critic(torch.from_numpy(obs)).numpy()
        mu1 = actor(obs[0]).numpy()          # 1-D numpy input path, net_actor.py:105-112
    return dict(obs=obs, mu=mu, v=v, mu_single=mu1, actor=_flat(actor, "actor"), critic=_flat(critic, "critic"),
                actor_keys=np.array(list(actor.state_dict().keys())),
                critic_keys=np.array(list(critic.state_dict().keys())),
                actor_shapes=np.array([str(tuple(t.shape)) for t in actor.state_dict().values()]),
                critic_shapes=np.array([str(tuple(t.shape)) for t in critic.state_dict().values()]))


def _make_agent(ppo_mod, net_actor, net_critic, env, tmp, **hp):
    base = dict(timesteps_per_batch=96, max_timesteps_per_episode=20, gamma=0.99, n_updates_per_iteration=4, lr=3e-4,
                clip=0.2, render=False, save_freq=1000, seed=0, method_name="golden", output_dir=tmp)
    base.update(hp)
    with contextlib.redirect_stdout(io.StringIO()):
        return ppo_mod.PPO(net_actor.NetActor, net_critic.NetCritic, env, 16, 2, **base)


def gen_rtgs(agent):
    rng = np.random.RandomState(5)
    eps = [list(rng.uniform(-3, 3, size=n)) for n in (1, 7, 20, 3, 12)]
    eps[1][-1] = 120.0
    eps[2][-1] = -100.0
    out = agent.compute_rtgs(eps).numpy()
    return dict(lens=np.array([len(e) for e in eps]), rews=np.concatenate(eps), rtgs=out, gamma=np.float64(agent.gamma))


def gen_get_action(torch, agent):
    rng = np.random.RandomState(9)
    obs = rng.uniform(0, 1, size=(40, 16)).astype(np.float32)
    rec = dict(obs=obs, mean=[], eps=[], act=[], logp=[])
    for i in range(len(obs)):
        with torch.no_grad():
            rec["mean"].append(agent.actor(obs[i]).numpy()[0])
        torch.manual_seed(100 + i)
        rec["eps"].append(torch.empty(1, 2).normal_().numpy()[0])   # what dist.sample() will draw
        torch.manual_seed(100 + i)
        a, lp = agent.get_action(obs[i], t_so_far=0, one_round=1)
        rec["act"].append(a); rec["logp"].append(lp)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["var"] = np.float64(agent.cov_var[0].item())
    # sanity: where the clamp was inactive the sample is mean + sqrt(var) * eps
    pred = out["mean"] + np.sqrt(out["var"]) * out["eps"]
    free = (pred[:, 0] > 0) & (pred[:, 0] < 1) & (np.abs(pred[:, 1]) < 1)
    assert free.any() and np.allclose(pred[free], out["act"][free], atol=1e-6), "noise replay does not match dist.sample()"
    return out


def gen_learn(torch, ppo_mod, net_actor, net_critic, tag, **hp):
    seg = maps.get_map("stage_1")
    ref_env = fake_ros.RefEnv(seg, seed=21, agent=0)

    class EnvFacade:  # PPO only touches these members of Env (ppo.py:83,486,535,541,593)
        use_vision = False

        def __init__(self, r):
            self._r = r
            self.position = r.env.position

        def reset(self):
            o = self._r.reset()
            self.position = self._r.env.position
            return o

        def step(self, action, past_action):
            out = self._r.step(action, past_action)
            self.position = self._r.env.position
            return out

    with tempfile.TemporaryDirectory() as tmp:
        agent = _make_agent(ppo_mod, net_actor, net_critic, EnvFacade(ref_env), tmp, **hp)
        before_a, before_c = _flat(agent.actor, "actor"), _flat(agent.critic, "critic")
        captured = {}
        orig_rollout = agent.rollout

        def rollout(past_action, t_so_far):
            out = orig_rollout(past_action=past_action, t_so_far=t_so_far)
            captured["batch"] = [o.clone() if hasattr(o, "clone") else o for o in out[:5]]
            captured["rews"] = [list(map(float, ep)) for ep in agent.logger["batch_rews"]]
            return out

        agent.rollout = rollout
        grads_a, grads_c = [], []
        a_step, c_step = agent.actor_optim.step, agent.critic_optim.step
        agent.actor_optim.step = lambda *a, **k: (grads_a.append(_flat_grad(agent.actor, "actor")), a_step(*a, **k))[1]
        agent.critic_optim.step = lambda *a, **k: (grads_c.append(_flat_grad(agent.critic, "critic")), c_step(*a, **k))[1]
        orig_summary = agent._log_summary

        def summary():  # _log_summary clears the per-iteration loss lists (ppo.py:813-946)
            captured["actor_losses"] = [float(x) for x in agent.logger["actor_losses"]]
            captured["critic_losses"] = [float(x) for x in agent.logger["critic_losses"]]
            return orig_summary()

        agent._log_summary = summary
        with contextlib.redirect_stdout(io.StringIO()):
            agent.learn(total_timesteps=1, past_action=[0, 0])
        obs, acts, logp, rtgs, lens = captured["batch"]
        lg = agent.logger
        out = dict(
            obs=obs.numpy(), acts=acts.numpy(), logp=logp.numpy(), rtgs=rtgs.numpy(), lens=np.asarray(lens),
            ep_lens_all=np.array([len(e) for e in captured["rews"]]), rews=np.concatenate(captured["rews"]),
            actor_before=before_a, critic_before=before_c, actor_after=_flat(agent.actor, "actor"),
            critic_after=_flat(agent.critic, "critic"), actor_grads=np.stack(grads_a), critic_grads=np.stack(grads_c),
            actor_losses=np.array(captured["actor_losses"]), critic_losses=np.array(captured["critic_losses"]),
            approx_kl=np.float64(lg["approx_kl"]), clip_frac=np.float64(lg["clip_frac"]), entropy=np.float64(lg["entropy"]),
            actor_grad_norm=np.float64(lg["actor_grad_norm"]), critic_grad_norm=np.float64(lg["critic_grad_norm"]),
            actor_param_delta=np.float64(lg["actor_param_delta"]), critic_param_delta=np.float64(lg["critic_param_delta"]),
            var=np.float64(agent.cov_var[0].item()), lr=np.float64(agent.lr), clip=np.float64(agent.clip),
            gamma=np.float64(agent.gamma), epochs=np.int32(agent.n_updates_per_iteration),
            max_episode_steps=np.int32(agent.max_timesteps_per_episode), env_seed=np.int64(21), segments=seg,
            v_mean_last_epoch=np.float64(float(agent.V.detach().mean())))
    return out


def main():
    if not fake_ros.reference_available():
        raise SystemExit("reference sources not found; golden vectors can only be generated where /root/reference exists")
    fake_ros.install_stubs()
    import torch
    import net_actor
    import net_critic
    import ppo as ppo_mod
    torch.set_num_threads(1)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "ppo_nets.npz"), **gen_nets(torch, net_actor, net_critic))
    with tempfile.TemporaryDirectory() as tmp:
        seg = maps.get_map("stage_1")
        agent = _make_agent(ppo_mod, net_actor, net_critic, fake_ros.RefEnv(seg, seed=0).env, tmp)
        np.savez_compressed(os.path.join(GOLD, "ppo_rtgs.npz"), **gen_rtgs(agent))
        np.savez_compressed(os.path.join(GOLD, "ppo_get_action.npz"), **gen_get_action(torch, agent))
    np.savez_compressed(os.path.join(GOLD, "ppo_learn_a.npz"), **gen_learn(torch, ppo_mod, net_actor, net_critic, "a"))
    # larger step size and more epochs so that ratios leave the clip interval
    np.savez_compressed(os.path.join(GOLD, "ppo_learn_b.npz"),
                        **gen_learn(torch, ppo_mod, net_actor, net_critic, "b", lr=2e-3, n_updates_per_iteration=10,
                                    timesteps_per_batch=160, max_timesteps_per_episode=30, seed=3))
    for f in sorted(os.listdir(GOLD)):
        if f.startswith("ppo_"):
            print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
