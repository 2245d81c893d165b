"""One tiny PPO iteration on cuda:0 checked against the float64 oracle (called by
__graft_entry__.smoke(); the oracle import makes this test infrastructure, not product path)."""
from __future__ import annotations

import tempfile

import numpy as np
import torch

from . import _capi, layout
from .env import VecEnv
from .nets import NetActor, NetCritic
from .ppo import PPO


def run() -> None:
    from oracle import ppo_oracle as po

    n, h, epochs = 256, 16, 2
    env = VecEnv(n, map="stage_1", device=0, seed=1, max_episode_steps=12)
    with tempfile.TemporaryDirectory() as tmp:
        agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=n * h, max_timesteps_per_episode=12,
                    n_updates_per_iteration=epochs, seed=0, output_dir=tmp, method_name="smoke", verbose=False)
        obs, acts, logp, rtgs, lens, it, _ = agent.rollout([0, 0], 0)
        torch.cuda.synchronize()
        assert obs.shape == (n * h, 16) and it["ep_count"] > 0 and len(lens) == it["ep_count"]
        flat0 = agent.flat.cpu().numpy().astype(np.float64)
        a0 = flat0[:layout.ACTOR_PARAMS]
        c0 = flat0[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS]
        o, a_, lp, rt = (x.cpu().numpy().astype(np.float64) for x in (obs, acts, logp, rtgs))
        # rollout: log-probs of the sampled actions and reward-to-go against the oracle
        v_ref, lp_ref = po.evaluate(a0, c0, o, a_, agent.var)
        np.testing.assert_allclose(lp, lp_ref, atol=2e-5, rtol=0)
        term = agent._b_term.cpu().numpy()
        np.testing.assert_allclose(rt.reshape(h, n), po.rtg_scan(agent._b_rew.cpu().numpy(), term, agent.gamma), rtol=1e-6,
                                   atol=1e-4)
        # update: 2 full-batch epochs against the oracle's Adam trajectory
        res = agent.update(obs, acts, logp, rtgs)
        a1, c1, _adv, hist = po.update(a0, c0, o, a_, lp, rt, agent.var, agent.clip, agent.lr, epochs)
        flat1 = agent.flat.cpu().numpy()
        np.testing.assert_allclose(flat1[:layout.ACTOR_PARAMS], a1, atol=2e-5, rtol=0)
        np.testing.assert_allclose(flat1[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS], c1, atol=2e-5,
                                   rtol=0)
        np.testing.assert_allclose(res["actor_losses"], [hh["actor_loss"] for hh in hist], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(res["critic_losses"], [hh["critic_loss"] for hh in hist], rtol=1e-4)
    print(f"ppo smoke ok: rollout {n}x{h}, {epochs} epochs, actor loss {res['actor_losses'][-1]:.5f}, "
          f"critic loss {res['critic_losses'][-1]:.3f}")
