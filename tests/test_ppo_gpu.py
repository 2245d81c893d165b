"""GPU parity tests of the PPO kernels, all through the C-ABI (navbot_ppo_b200.PPO / NetActor /
NetCritic are ctypes shims over libnavbot_b200.so).  Checkers: golden vectors recorded from the
reference's torch code (tests/golden/ppo_*.npz) and the float64 oracle pinned to them.

Tolerances: MLP forward / log-prob <= 1e-5 absolute against the fp32
reference; losses and gradients <= 1e-4 relative to the largest gradient entry (reduction
order); parameters after k Adam steps <= 2e-5 absolute."""
import ctypes
import os

import numpy as np
import pytest
import torch

from navbot_ppo_b200 import _capi, layout
from navbot_ppo_b200.env import Env, VecEnv
from navbot_ppo_b200.nets import NetActor, NetCritic, _Handles
from navbot_ppo_b200.ppo import PPO
from oracle import ppo_oracle as po
from tests.helpers import golden

pytestmark = pytest.mark.gpu

FWD_ATOL = 1e-5
GRAD_RTOL = 1e-4
DEV = "cuda:0"


def _flat_from(actor, critic):
    f = np.zeros(_capi.PPO_FLAT, np.float32)
    f[:layout.ACTOR_PARAMS] = actor
    f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS] = critic
    return torch.from_numpy(f).to(DEV)


def _split(flat):
    f = flat.cpu().numpy()
    return f[:layout.ACTOR_PARAMS], f[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS]


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=DEV, dtype=dtype).contiguous()


def _handle(T=1 << 16, clip=0.2, lr=3e-4):
    return _Handles.get(torch.device(DEV), T, clip, lr)


def _sp():
    return torch.cuda.current_stream().cuda_stream


def test_rtg_scan_matches_reference_and_oracle():
    g = golden("ppo_rtgs")
    term = np.zeros(len(g["rews"]), np.uint8)
    term[np.cumsum(g["lens"]) - 1] = 1
    rew = _t(g["rews"]); out = torch.empty_like(rew); term_d = _t(term, torch.uint8)
    _capi.check(_capi.lib().navppo_rtg_scan(rew.data_ptr(), term_d.data_ptr(), None, None, float(g["gamma"]),
                                            1.0, out.data_ptr(), len(term), 1, _sp()))
    np.testing.assert_allclose(out.cpu().numpy(), g["rtgs"], rtol=2e-6, atol=1e-5)
    # [H, N] layout at rollout size against the oracle scan, incl. GAE form
    rng = np.random.RandomState(1)
    H, N = 128, 8192
    r = rng.normal(size=(H, N)).astype(np.float32) * 10
    t = (rng.uniform(size=(H, N)) < 0.03).astype(np.uint8)
    v = rng.normal(size=(H, N)).astype(np.float32)
    lv = rng.normal(size=N).astype(np.float32)
    rd, td_, vd, lvd = _t(r), _t(t, torch.uint8), _t(v), _t(lv)
    out = torch.empty_like(rd)
    _capi.check(_capi.lib().navppo_rtg_scan(rd.data_ptr(), td_.data_ptr(), None, None, 0.99, 1.0, out.data_ptr(), H, N, _sp()))
    np.testing.assert_allclose(out.cpu().numpy(), po.rtg_scan(r, t, 0.99), rtol=1e-6, atol=1e-5)
    # bootstrapped reward-to-go (PPO's bootstrap_value option)
    _capi.check(_capi.lib().navppo_rtg_scan(rd.data_ptr(), td_.data_ptr(), None, lvd.data_ptr(), 0.99, 1.0, out.data_ptr(), H, N, _sp()))
    np.testing.assert_allclose(out.cpu().numpy(), po.rtg_scan(r, t, 0.99, last_value=lv), rtol=1e-6, atol=1e-5)
    _capi.check(_capi.lib().navppo_rtg_scan(rd.data_ptr(), td_.data_ptr(), vd.data_ptr(), lvd.data_ptr(), 0.99, 0.95,
                                            out.data_ptr(), H, N, _sp()))
    np.testing.assert_allclose(out.cpu().numpy(), po.gae_scan(r, t, v, lv, 0.99, 0.95), rtol=1e-5, atol=1e-4)


def test_forward_matches_reference_nets():
    g = golden("ppo_nets")
    flat = _flat_from(g["actor"], g["critic"])
    obs = _t(g["obs"])
    T = obs.shape[0]
    mu = torch.empty((T, 2), device=DEV); v = torch.empty(T, device=DEV)
    _capi.check(_capi.lib().navppo_forward(_handle(), flat.data_ptr(), obs.data_ptr(), T, mu.data_ptr(), v.data_ptr(), _sp()))
    np.testing.assert_allclose(mu.cpu().numpy(), g["mu"], atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(v.cpu().numpy(), g["v"][:, 0], atol=FWD_ATOL, rtol=0)
    # large, ragged batch (persistent-loop path) against the oracle
    rng = np.random.RandomState(2)
    big = rng.uniform(-1, 1, size=(70001, 16)).astype(np.float32)
    bo = _t(big); T = big.shape[0]
    mu = torch.empty((T, 2), device=DEV); v = torch.empty(T, device=DEV)
    _capi.check(_capi.lib().navppo_forward(_handle(), flat.data_ptr(), bo.data_ptr(), T, mu.data_ptr(), v.data_ptr(), _sp()))
    sel = np.r_[0:300, T - 300:T, rng.randint(0, T, 400)]
    np.testing.assert_allclose(mu.cpu().numpy()[sel], po.actor_forward(g["actor"], big[sel]), atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(v.cpu().numpy()[sel], po.critic_forward(g["critic"], big[sel]), atol=FWD_ATOL, rtol=0)


def test_net_modules_have_reference_surface():
    """Same torch seed -> same initial weights as the reference's constructors; same state_dict."""
    g = golden("ppo_nets")
    torch.manual_seed(0)
    actor = NetActor(16, 2); critic = NetCritic(16, 1)
    assert list(actor.state_dict().keys()) == g["actor_keys"].tolist()
    assert list(critic.state_dict().keys()) == g["critic_keys"].tolist()
    np.testing.assert_array_equal(layout.flatten_state_dict(actor.state_dict(), "actor"), g["actor"])
    np.testing.assert_array_equal(layout.flatten_state_dict(critic.state_dict(), "critic"), g["critic"])
    mu = actor(g["obs"])                       # numpy batch
    np.testing.assert_allclose(mu.cpu().numpy(), g["mu"], atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(actor(g["obs"][0]).cpu().numpy(), g["mu_single"], atol=FWD_ATOL, rtol=0)   # 1-D input
    np.testing.assert_allclose(critic(torch.from_numpy(g["obs"])).cpu().numpy(), g["v"], atol=FWD_ATOL, rtol=0)
    with pytest.raises(ValueError):
        actor(np.zeros((3, 15), np.float32))


def test_act_matches_reference_get_action():
    g = golden("ppo_get_action")
    actor = golden("ppo_learn_a")["actor_before"]
    flat = _flat_from(actor, np.zeros(layout.CRITIC_PARAMS, np.float32))
    n = len(g["obs"])
    act = torch.empty((n, 2), device=DEV); logp = torch.empty(n, device=DEV); mu = torch.empty((n, 2), device=DEV)
    obs, eps = _t(g["obs"]), _t(g["eps"])       # keep the inputs alive across the launch
    _capi.check(_capi.lib().navppo_act(_handle(), flat.data_ptr(), obs.data_ptr(), n, float(g["var"]), 0, 0, 0,
                                       eps.data_ptr(), act.data_ptr(), logp.data_ptr(), mu.data_ptr(), _sp()))
    np.testing.assert_allclose(mu.cpu().numpy(), g["mean"], atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(act.cpu().numpy(), g["act"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logp.cpu().numpy(), g["logp"], atol=1e-5, rtol=1e-6)
    # clamped entries are exactly on the bounds, like the reference's
    np.testing.assert_array_equal(act.cpu().numpy()[:, 0] == 0, g["act"][:, 0] == 0)
    np.testing.assert_array_equal(np.abs(act.cpu().numpy()[:, 1]) == 1, np.abs(g["act"][:, 1]) == 1)


def test_act_device_noise_is_standard_normal_and_partition_invariant():
    g = golden("ppo_nets")
    flat = _flat_from(g["actor"], g["critic"])
    n = 1 << 17
    obs = torch.rand((n, 16), device=DEV)
    var = 1e-4   # tiny exploration noise: nothing clamps on the tanh head, so eps is recoverable

    def run(lo, hi, draw, seed=5):
        m = hi - lo
        act = torch.empty((m, 2), device=DEV); logp = torch.empty(m, device=DEV); mu = torch.empty((m, 2), device=DEV)
        o = obs[lo:hi].contiguous()
        _capi.check(_capi.lib().navppo_act(_handle(), flat.data_ptr(), o.data_ptr(), m, var, seed, lo,
                                           draw, None, act.data_ptr(), logp.data_ptr(), mu.data_ptr(), _sp()))
        return act, logp, mu

    act, logp, mu = run(0, n, 3)
    eps = ((act - mu) / var ** 0.5).cpu().numpy()[:, 1].astype(np.float64)
    assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1.0) < 0.02
    assert abs((np.abs(eps) > 1.96).mean() - 0.05) < 0.005
    a2, _, _ = run(0, n, 3)
    assert torch.equal(act, a2)                                  # deterministic in (seed, agent, draw)
    a3, _, _ = run(0, n, 4)
    assert not torch.equal(act, a3)                              # a new draw index gives new noise
    lo_a, _, _ = run(0, n // 2, 3); hi_a, _, _ = run(n // 2, n, 3)
    assert torch.equal(act, torch.cat([lo_a, hi_a]))             # keyed by GLOBAL agent id
    # log-prob is that of the emitted action
    ref = po.gaussian_logp(mu.cpu().numpy().astype(np.float64), act.cpu().numpy().astype(np.float64), var)
    np.testing.assert_allclose(logp.cpu().numpy(), ref, atol=2e-4, rtol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_evaluate_advantage_grad_match_reference(tag):
    g = golden("ppo_learn_" + tag)
    var, clip = float(g["var"]), float(g["clip"])
    flat = _flat_from(g["actor_before"], g["critic_before"])
    obs, act, lp, rtg = _t(g["obs"]), _t(g["acts"]), _t(g["logp"]), _t(g["rtgs"])
    T = obs.shape[0]
    L, h = _capi.lib(), _handle(1 << 16, clip, float(g["lr"]))
    v = torch.empty(T, device=DEV); logp = torch.empty(T, device=DEV); adv = torch.empty(T, device=DEV)
    _capi.check(L.navppo_evaluate(h, flat.data_ptr(), obs.data_ptr(), act.data_ptr(), T, var, v.data_ptr(), logp.data_ptr(), _sp()))
    v_ref, lp_ref = po.evaluate(g["actor_before"], g["critic_before"], g["obs"], g["acts"], var)
    np.testing.assert_allclose(v.cpu().numpy(), v_ref, atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(logp.cpu().numpy(), lp_ref, atol=2e-5, rtol=0)
    np.testing.assert_allclose(logp.cpu().numpy(), g["logp"], atol=3e-5, rtol=0)     # what the reference's rollout stored
    stats = torch.zeros(3, dtype=torch.float64, device=DEV)
    _capi.check(L.navppo_adv_stats(rtg.data_ptr(), v.data_ptr(), T, stats.data_ptr(), _sp()))
    _capi.check(L.navppo_adv_normalize(rtg.data_ptr(), v.data_ptr(), T, stats.data_ptr(), adv.data_ptr(), _sp()))
    adv_ref = po.advantage(g["rtgs"], v_ref)
    np.testing.assert_allclose(adv.cpu().numpy(), adv_ref, atol=2e-5, rtol=1e-5)
    grad = torch.empty(_capi.PPO_FLAT, device=DEV)
    met = torch.zeros(_capi.PPO_NUM_METRICS, dtype=torch.float64, device=DEV)
    _capi.check(L.navppo_grad(h, flat.data_ptr(), obs.data_ptr(), act.data_ptr(), lp.data_ptr(), adv.data_ptr(),
                              rtg.data_ptr(), T, T, var, grad.data_ptr(), met.data_ptr(), _sp()))
    ga, gc = _split(grad)
    for got, want in ((ga, g["actor_grads"][0]), (gc, g["critic_grads"][0])):      # the reference's .grad, epoch 0
        np.testing.assert_allclose(got, want, atol=GRAD_RTOL * np.abs(want).max(), rtol=0)
    m = met.cpu().numpy()
    assert abs(m[_capi.M_ACTOR_LOSS] - g["actor_losses"][0]) <= 1e-5 + GRAD_RTOL * abs(g["actor_losses"][0])
    assert abs(m[_capi.M_CRITIC_LOSS] - g["critic_losses"][0]) <= GRAD_RTOL * abs(g["critic_losses"][0])
    # padding of the flat layout carries no gradient
    assert float(grad[layout.ACTOR_PARAMS:_capi.PPO_CRITIC_OFFSET].abs().max()) == 0.0


@pytest.mark.parametrize("tag", ["a", "b"])
def test_update_matches_reference_learn_iteration(tag):
    """navppo_update == the inline update loop of PPO.learn (ppo.py:275-397), epoch by epoch."""
    g = golden("ppo_learn_" + tag)
    epochs = int(g["epochs"])
    flat = _flat_from(g["actor_before"], g["critic_before"])
    m1 = torch.zeros_like(flat); m2 = torch.zeros_like(flat)
    obs, act, lp, rtg = _t(g["obs"]), _t(g["acts"]), _t(g["logp"]), _t(g["rtgs"])
    T = obs.shape[0]
    adv = torch.empty(T, device=DEV); v = torch.empty(T, device=DEV)
    met = torch.zeros((epochs, _capi.PPO_NUM_METRICS), dtype=torch.float64, device=DEV)
    h = _handle(1 << 16, float(g["clip"]), float(g["lr"]))
    _capi.check(_capi.lib().navppo_update(h, flat.data_ptr(), m1.data_ptr(), m2.data_ptr(), 0, obs.data_ptr(), act.data_ptr(),
                                          lp.data_ptr(), rtg.data_ptr(), T, float(g["var"]), epochs, adv.data_ptr(),
                                          v.data_ptr(), met.data_ptr(), _sp()))
    a, c = _split(flat)
    # one Adam step moves a parameter by up to lr whatever the gradient's size, so entries whose
    # gradient is rounding noise may differ by a fraction of lr: bound = lr / 15 (2e-5 at the
    # reference's lr = 3e-4), and all but a handful must sit within 2e-5 regardless
    atol = float(g["lr"]) / 15
    for got, want in ((a, g["actor_after"]), (c, g["critic_after"])):
        np.testing.assert_allclose(got, want, atol=atol, rtol=0)
        assert (np.abs(got - want) > 2e-5).mean() < 1e-3
    m = met.cpu().numpy()
    np.testing.assert_allclose(m[:, _capi.M_ACTOR_LOSS], g["actor_losses"], atol=2e-5, rtol=2e-4)
    np.testing.assert_allclose(m[:, _capi.M_CRITIC_LOSS], g["critic_losses"], rtol=2e-4)
    assert abs(m[:, _capi.M_APPROX_KL].mean() - float(g["approx_kl"])) < 1e-5 + 2e-3 * abs(float(g["approx_kl"]))
    assert abs(m[:, _capi.M_CLIP_FRAC].mean() - float(g["clip_frac"])) <= 2.0 / T
    assert abs(np.sqrt(m[:, _capi.M_ACTOR_GRAD_SQ]).mean() - float(g["actor_grad_norm"])) < 2e-4 * float(g["actor_grad_norm"])
    assert abs(np.sqrt(m[:, _capi.M_CRITIC_GRAD_SQ]).mean() - float(g["critic_grad_norm"])) < 2e-4 * float(g["critic_grad_norm"])


def test_grad_is_additive_over_shards_and_deterministic():
    """Data-parallel contract: with the 1/n_global factor, the gradients of two disjoint shards
    sum to the full-batch gradient (what the per-epoch all-reduce computes).  Also: bit-identical
    results run to run (fixed-order reductions), at a size with many tiles per CTA."""
    g = golden("ppo_learn_b")
    rng = np.random.RandomState(4)
    T = 50000
    idx = rng.randint(0, len(g["obs"]), T)
    obs = g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16)).astype(np.float32)
    act, lp, rtg = g["acts"][idx], g["logp"][idx] + rng.normal(scale=0.1, size=T).astype(np.float32), g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    flat = _flat_from(g["actor_after"], g["critic_after"])
    L, h = _capi.lib(), _handle(1 << 16)

    def grad(lo, hi, n_global):
        out = torch.empty(_capi.PPO_FLAT, device=DEV); met = torch.zeros(8, dtype=torch.float64, device=DEV)
        o, a_, l_, ad, rt = _t(obs[lo:hi]), _t(act[lo:hi]), _t(lp[lo:hi]), _t(adv[lo:hi]), _t(rtg[lo:hi])
        _capi.check(L.navppo_grad(h, flat.data_ptr(), o.data_ptr(), a_.data_ptr(), l_.data_ptr(), ad.data_ptr(),
                                  rt.data_ptr(), hi - lo, n_global, float(g["var"]), out.data_ptr(), met.data_ptr(), _sp()))
        torch.cuda.synchronize()
        return out, met

    full, mfull = grad(0, T, T)
    again, _ = grad(0, T, T)
    assert torch.equal(full, again)
    a, ma = grad(0, 20001, T); b, mb = grad(20001, T, T)
    scale = float(full.abs().max())
    assert float((a + b - full).abs().max()) <= 2e-5 * scale
    np.testing.assert_allclose((ma + mb).cpu().numpy()[:4], mfull.cpu().numpy()[:4], rtol=1e-6, atol=1e-9)
    # and the full-batch gradient agrees with the float64 oracle
    sub = slice(0, 4000)
    gs, ms = grad(0, 4000, 4000)
    m_ref, ga_ref, gc_ref = po.losses_and_grads(g["actor_after"], g["critic_after"], obs[sub], act[sub], lp[sub], adv[sub],
                                                rtg[sub], float(g["var"]), 0.2)
    ga, gc = _split(gs)
    np.testing.assert_allclose(ga, ga_ref, atol=GRAD_RTOL * np.abs(ga_ref).max(), rtol=0)
    np.testing.assert_allclose(gc, gc_ref, atol=GRAD_RTOL * np.abs(gc_ref).max(), rtol=0)
    assert abs(ms.cpu().numpy()[_capi.M_CLIP_FRAC] - m_ref["clip_frac"]) <= 2.0 / 4000 and m_ref["clip_frac"] > 0.05


def test_single_env_rollout_protocol_reproduces_reference_batch(tmp_path):
    """The drop-in pair (Env, PPO) fed the actions the reference sampled reproduces the batch
    the reference's rollout returned: observations, reward-to-go, episode lengths."""
    g = golden("ppo_learn_a")
    env = Env(True, seed=int(g["env_seed"]), map=g["segments"])
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=len(g["obs"]),
                max_timesteps_per_episode=int(g["max_episode_steps"]), output_dir=str(tmp_path), verbose=False)
    obs = env.reset()
    past, one_round, ep_rews, rews, lens, seen = [0, 0], 0, [], [], [], []
    for t in range(len(g["obs"])):                      # ppo.py:505-593 with the recorded actions
        seen.append(obs)
        a = g["acts"][t]
        obs, r, done, arrive = env.step(a, past)
        past = a; ep_rews.append(r); one_round += 1
        if done or arrive or one_round >= int(g["max_episode_steps"]):
            lens.append(one_round); rews.append(ep_rews); ep_rews = []; one_round = 0; past = [0, 0]
            obs = env.reset()
    rews.append(ep_rews)
    np.testing.assert_allclose(np.asarray(seen, np.float32), g["obs"], atol=1e-5, rtol=0)
    assert lens == g["lens"].tolist()
    rtgs = agent.compute_rtgs(rews)
    assert rtgs.dtype == torch.float32 and rtgs.device.type == "cpu"
    np.testing.assert_allclose(rtgs.numpy(), g["rtgs"], rtol=1e-5, atol=2e-3)
    # get_action with injected N(0, I) draws; evaluate through the class
    ga = golden("ppo_get_action")
    with torch.no_grad():
        agent.flat[:layout.ACTOR_PARAMS].copy_(torch.from_numpy(g["actor_before"]))
    a, lp = agent.get_action(ga["obs"][0], 0, 1, noise=ga["eps"][0])
    assert a.shape == (2,) and a.dtype == np.float32
    np.testing.assert_allclose(a, ga["act"][0], atol=2e-6); assert abs(float(lp) - float(ga["logp"][0])) < 1e-5
    with torch.no_grad():
        agent.flat[_capi.PPO_CRITIC_OFFSET:_capi.PPO_CRITIC_OFFSET + layout.CRITIC_PARAMS].copy_(torch.from_numpy(g["critic_before"]))
    V, logp = agent.evaluate(torch.from_numpy(g["obs"]), torch.from_numpy(g["acts"]))
    np.testing.assert_allclose(logp.cpu().numpy(), g["logp"], atol=3e-5, rtol=0)
    res = agent.update(torch.from_numpy(g["obs"]), torch.from_numpy(g["acts"]), torch.from_numpy(g["logp"]),
                       torch.from_numpy(g["rtgs"]), epochs=int(g["epochs"]))
    np.testing.assert_allclose(agent.flat.cpu().numpy()[:layout.ACTOR_PARAMS], g["actor_after"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(res["actor_losses"], g["actor_losses"], atol=2e-5, rtol=2e-4)
    assert abs(res["actor_param_delta"] - float(g["actor_param_delta"])) < 1e-3 * float(g["actor_param_delta"])


def test_learn_on_vecenv_trains_and_writes_reference_checkpoints(tmp_path):
    torch.manual_seed(0)
    env = VecEnv(512, map="stage_1", seed=1, max_episode_steps=40)
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=512 * 48, max_timesteps_per_episode=40,
                n_updates_per_iteration=5, output_dir=str(tmp_path), method_name="t", save_freq=1, seed=0, verbose=False)
    before = agent.flat.clone()
    obs, acts, logp, rtgs, lens, it, _ = agent.rollout([0, 0], 0)
    assert obs.shape == (512 * 48, 16) and acts.shape == (512 * 48, 2) and logp.shape == rtgs.shape == (512 * 48,)
    a = acts.cpu().numpy()
    assert a[:, 0].min() >= 0 and a[:, 0].max() <= 1 and np.abs(a[:, 1]).max() <= 1
    assert it["ep_count"] == len(lens) > 0 and lens.max() <= 40 and it["ep_count"] == it["successes"] + it["collisions"] + it["timeouts"]
    assert int(lens.sum()) == int(it["length_sum"])
    # rollout buffers against the oracle scan of the recorded rewards / flags
    np.testing.assert_allclose(agent._b_rtg.cpu().numpy(), po.rtg_scan(agent._b_rew.cpu().numpy(), agent._b_term.cpu().numpy(), 0.99),
                               rtol=1e-6, atol=1e-4)
    # per-episode csv (ppo.py:739-746): one row per completed episode, reference columns; the return
    # column is the sum of the rewards the rollout buffer holds for that episode
    import csv as _csv
    rows = list(_csv.reader(open(agent.episode_csv_path)))
    assert rows[0] == ["episode", "timestep", "success", "collision", "timeout", "length", "return", "path_length", "time"]
    assert len(rows) - 1 == it["ep_count"] == agent.episode_count
    rew = agent._b_rew.cpu().numpy(); term = agent._b_term.cpu().numpy()
    first_end = int(np.nonzero(term[:, 0])[0][0])            # agent 0's first episode is row 0
    r0 = rows[1]
    assert int(r0[5]) == first_end + 1 and abs(float(r0[6]) - rew[:first_end + 1, 0].sum()) < 1e-2
    assert int(r0[2]) + int(r0[3]) + int(r0[4]) == 1 and float(r0[7]) >= 0.0
    assert sum(int(r[2]) for r in rows[1:]) == it["successes"] and sum(int(r[3]) for r in rows[1:]) == it["collisions"]
    assert abs(sum(float(r[6]) for r in rows[1:]) - it["return_sum"]) < 1e-3 * max(1.0, abs(it["return_sum"]))
    assert abs(sum(float(r[7]) for r in rows[1:]) - it["path_sum"]) < 1e-3 * max(1.0, abs(it["path_sum"]))
    # stored log-probs are those of the stored actions under the rollout policy
    V, lp2 = agent.evaluate(obs, acts)
    np.testing.assert_allclose(lp2.cpu().numpy(), logp.cpu().numpy(), atol=1e-4, rtol=1e-5)
    total = agent.learn(total_timesteps=1)
    assert total > 0 and not torch.equal(before, agent.flat) and torch.isfinite(agent.flat).all()
    s = agent.logger["summary"]
    assert np.isfinite(list(s.values())).all() and 0 <= s["ppo/clip_frac"] <= 1
    ck = sorted(os.listdir(agent.checkpoint_dir))
    assert len(ck) == 2 and ck[0].startswith("actor_iter0001_step") and ck[1].startswith("critic_iter0001_step")
    sd = torch.load(os.path.join(agent.checkpoint_dir, ck[0]))
    assert list(sd.keys()) == golden("ppo_nets")["actor_keys"].tolist()      # loads into the reference's NetActor
    assert sd["rb1.fc1.weight"].shape[1] == 16                               # main.py:64 resume check
    fresh = NetActor(16, 2)
    fresh.load_state_dict(sd)
    np.testing.assert_array_equal(layout.flatten_state_dict(fresh.state_dict(), "actor"),
                                  agent.flat.cpu().numpy()[:layout.ACTOR_PARAMS])


def test_ppo_error_paths():
    L = _capi.lib()
    cfg = _capi.default_ppo_cfg()
    cfg.max_samples = 0
    h = ctypes.c_void_p()
    assert L.navppo_create(ctypes.byref(h), ctypes.byref(cfg)) == -22
    hh = _handle(1024)
    x = torch.zeros(64, device=DEV)
    assert L.navppo_forward(hh, x.data_ptr(), x.data_ptr(), 4, None, None, None) == -22
    assert L.navppo_grad(hh, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 2048, 2048,
                         0.8, x.data_ptr(), x.data_ptr(), None) == -22 and b"max_samples" in L.nav_last_error()
    assert L.navppo_act(hh, x.data_ptr(), x.data_ptr(), 4, 0.0, 0, 0, 0, None, x.data_ptr(), x.data_ptr(), None, None) == -22


def test_graph_replayed_rollout_equals_the_launch_by_launch_rollout(tmp_path):
    """PPO.rollout replays the 2 H launches of ppo.py:505-549 as ONE captured CUDA graph whose sampling
    epilogue reads the variance and the noise counter from device words: every rollout buffer is
    bit-identical to navppo_rollout enqueued launch by launch, rollout after rollout (the counter and the
    decayed variance move), and the episode lengths the simulator writes at episode ends equal the ones
    reconstructed from the flags."""
    def make(graph):
        env = VecEnv(384, map="stage_2", seed=3, max_episode_steps=25)
        return PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=384 * 40, max_timesteps_per_episode=25,
                   n_updates_per_iteration=1, output_dir=str(tmp_path / ("g" if graph else "l")), method_name="r", seed=5,
                   verbose=False, graph_rollout=graph)
    a, b = make(True), make(False)
    b.flat.copy_(a.flat)
    for it in range(3):
        if it == 2:                      # a decayed exploration variance must reach the replayed graph
            a._decay_cov(); b._decay_cov()
        ra, rb = a.rollout([0, 0], 0), b.rollout([0, 0], 0)
        torch.cuda.synchronize()
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y)
        for name in ("_b_rew", "_b_flags", "_b_epret", "_b_eppath", "_b_eplen", "_b_term", "_next_obs"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        assert np.array_equal(ra[4], rb[4]) and ra[5] == rb[5]
        # lengths from the flags: every agent starts a fresh episode at t = 0 (ppo.py:486)
        term = a._b_term.cpu().numpy()
        nn_, tt = np.nonzero(term.T)
        prev = np.where(np.r_[True, nn_[1:] != nn_[:-1]], -1, np.r_[0, tt[:-1]])
        assert np.array_equal(ra[4], tt - prev) and len(ra[4]) == ra[5]["ep_count"] > 0
    assert a._rollout_graph is not None and b._rollout_graph is None


def test_vectorised_learn_counts_every_simulated_step_and_checks_the_episode_cap(tmp_path):
    """learn() on a VecEnv advances t_so_far by H x N per iteration (episodes longer than the horizon would
    otherwise never be counted and the loop would not terminate), and a trainer whose
    max_timesteps_per_episode disagrees with the simulator's cap is refused."""
    env = VecEnv(256, map="stage_1", seed=1, max_episode_steps=500)
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=256 * 8, max_timesteps_per_episode=500,
                n_updates_per_iteration=1, output_dir=str(tmp_path), method_name="c", seed=0, verbose=False)
    assert agent.learn(total_timesteps=256 * 8 * 2) == 256 * 8 * 2      # two iterations, no episode needs to end
    with pytest.raises(ValueError):
        PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=256 * 8, max_timesteps_per_episode=800,
            output_dir=str(tmp_path), method_name="d", verbose=False)


def test_sharded_update_equals_single_gpu_update_on_two_ranks():
    """Two NCCL ranks, each on half of a fixed batch, land on the parameters one GPU reaches on the whole batch, for
    the NCCL all-reduce and for the peer-memory exchange fused into Adam (tools/dist_check.py; skipped on a box with
    one GPU)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert "DIST_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
