"""GPU parity tests of the tcgen05 tensor-core gradient path (navppo_grad with
precision = NAVPPO_BF16X3 / NAVPPO_BF16), through the C-ABI, against the reference's own
gradients (tests/golden/ppo_learn_*.npz) and the float64 oracle.

Tolerances, relative to the largest entry of each network's gradient:
  BF16X3 (split operands, ~16 mantissa bits): 1e-4 — the same bar as the fp32 CUDA-core path;
  BF16   (plain bf16 operands, fp32 accumulate): 1e-2 (measured ~1e-3)."""
import numpy as np
import pytest
import torch

from navbot_ppo_b200 import _capi, layout
from navbot_ppo_b200.env import VecEnv
from navbot_ppo_b200.nets import NetActor, NetCritic, _Handles
from navbot_ppo_b200.ppo import PPO
from oracle import ppo_oracle as po
from tests.helpers import golden
from tests.test_ppo_gpu import DEV, _flat_from, _split, _sp, _t

pytestmark = pytest.mark.gpu

MODES = [(_capi.PREC_BF16X3, 1e-4), (_capi.PREC_BF16, 1e-2)]


def _grad(h, flat, obs, act, lp, adv, rtg, n_global, var):
    T = obs.shape[0]
    grad = torch.empty(_capi.PPO_FLAT, device=DEV)
    met = torch.zeros(_capi.PPO_NUM_METRICS, dtype=torch.float64, device=DEV)
    _capi.check(_capi.lib().navppo_grad(h, flat.data_ptr(), obs.data_ptr(), act.data_ptr(), lp.data_ptr(), adv.data_ptr(),
                                        rtg.data_ptr(), T, n_global, var, grad.data_ptr(), met.data_ptr(), _sp()))
    torch.cuda.synchronize()
    return grad, met.cpu().numpy()


@pytest.mark.parametrize("prec,rtol", MODES)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_tc_grad_matches_reference_gradients(tag, prec, rtol):
    g = golden("ppo_learn_" + tag)
    var, clip = float(g["var"]), float(g["clip"])
    h = _Handles.get(torch.device(DEV), 1 << 16, clip, float(g["lr"]), prec)
    flat = _flat_from(g["actor_before"], g["critic_before"])
    obs, act, lp, rtg = _t(g["obs"]), _t(g["acts"]), _t(g["logp"]), _t(g["rtgs"])
    v0, _ = po.evaluate(g["actor_before"], g["critic_before"], g["obs"], g["acts"], var)
    adv = _t(po.advantage(g["rtgs"], v0).astype(np.float32))
    grad, m = _grad(h, flat, obs, act, lp, adv, rtg, obs.shape[0], var)
    ga, gc = _split(grad)
    for got, want in ((ga, g["actor_grads"][0]), (gc, g["critic_grads"][0])):      # the reference's .grad, epoch 0
        np.testing.assert_allclose(got, want, atol=rtol * np.abs(want).max(), rtol=0)
    assert abs(m[_capi.M_ACTOR_LOSS] - g["actor_losses"][0]) <= (1e-5 if prec == _capi.PREC_BF16X3 else 1e-4) + 10 * rtol * abs(g["actor_losses"][0])
    assert abs(m[_capi.M_CRITIC_LOSS] - g["critic_losses"][0]) <= 10 * rtol * abs(g["critic_losses"][0])
    assert float(grad[layout.ACTOR_PARAMS:_capi.PPO_CRITIC_OFFSET].abs().max()) == 0.0


@pytest.mark.parametrize("prec,rtol", MODES)
def test_tc_grad_many_tiles_additive_deterministic(prec, rtol):
    """Ragged multi-tile batch (persistent loop, partial last tile), clipped ratios: against the
    float64 oracle on a sub-batch, additivity over shards, bit-reproducibility."""
    g = golden("ppo_learn_b")
    rng = np.random.RandomState(4)
    T = 33333
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    act, lp, rtg = g["acts"][idx], (g["logp"][idx] + rng.normal(scale=0.1, size=T)).astype(np.float32), g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    var = float(g["var"])
    flat = _flat_from(g["actor_after"], g["critic_after"])
    h = _Handles.get(torch.device(DEV), 1 << 16, 0.2, 3e-4, prec)
    o, a_, l_, ad, rt = _t(obs), _t(act), _t(lp), _t(adv), _t(rtg)
    full, mfull = _grad(h, flat, o, a_, l_, ad, rt, T, var)
    again, _ = _grad(h, flat, o, a_, l_, ad, rt, T, var)
    assert torch.equal(full, again)
    cut = 12345
    ga, ma = _grad(h, flat, o[:cut].contiguous(), a_[:cut].contiguous(), l_[:cut].contiguous(), ad[:cut].contiguous(),
                   rt[:cut].contiguous(), T, var)
    gb, mb = _grad(h, flat, o[cut:].contiguous(), a_[cut:].contiguous(), l_[cut:].contiguous(), ad[cut:].contiguous(),
                   rt[cut:].contiguous(), T, var)
    fa, fc = _split(full); sa, sc = _split(ga + gb)
    assert np.abs(sa - fa).max() <= 10 * rtol * np.abs(fa).max() and np.abs(sc - fc).max() <= 10 * rtol * np.abs(fc).max()
    np.testing.assert_allclose((ma + mb)[:4], mfull[:4], rtol=1e-6, atol=1e-9)
    n = 3000
    gs, ms = _grad(h, flat, o[:n].contiguous(), a_[:n].contiguous(), l_[:n].contiguous(), ad[:n].contiguous(), rt[:n].contiguous(),
                   n, var)
    m_ref, ga_ref, gc_ref = po.losses_and_grads(g["actor_after"], g["critic_after"], obs[:n], act[:n], lp[:n], adv[:n], rtg[:n],
                                                var, 0.2)
    xa, xc = _split(gs)
    np.testing.assert_allclose(xa, ga_ref, atol=rtol * np.abs(ga_ref).max(), rtol=0)
    np.testing.assert_allclose(xc, gc_ref, atol=rtol * np.abs(gc_ref).max(), rtol=0)
    assert abs(ms[_capi.M_CLIP_FRAC] - m_ref["clip_frac"]) <= (2.0 if prec == _capi.PREC_BF16X3 else 12.0) / n
    assert m_ref["clip_frac"] > 0.05


def test_tc_update_tracks_reference_learn_iteration():
    """Whole update (10 epochs, lr 2e-3) in split-BF16 arithmetic stays on the reference's
    parameter trajectory."""
    g = golden("ppo_learn_b")
    epochs = int(g["epochs"])
    flat = _flat_from(g["actor_before"], g["critic_before"])
    m1 = torch.zeros_like(flat); m2 = torch.zeros_like(flat)
    obs, act, lp, rtg = _t(g["obs"]), _t(g["acts"]), _t(g["logp"]), _t(g["rtgs"])
    T = obs.shape[0]
    adv = torch.empty(T, device=DEV); v = torch.empty(T, device=DEV)
    met = torch.zeros((epochs, _capi.PPO_NUM_METRICS), dtype=torch.float64, device=DEV)
    h = _Handles.get(torch.device(DEV), 1 << 16, float(g["clip"]), float(g["lr"]), _capi.PREC_BF16X3)
    _capi.check(_capi.lib().navppo_update(h, flat.data_ptr(), m1.data_ptr(), m2.data_ptr(), 0, obs.data_ptr(), act.data_ptr(),
                                          lp.data_ptr(), rtg.data_ptr(), T, float(g["var"]), epochs, adv.data_ptr(),
                                          v.data_ptr(), met.data_ptr(), _sp()))
    a, c = _split(flat)
    for got, want in ((a, g["actor_after"]), (c, g["critic_after"])):
        np.testing.assert_allclose(got, want, atol=float(g["lr"]) / 4, rtol=0)   # a quarter of one Adam step
        assert (np.abs(got - want) > 1e-4).mean() < 1e-2
    m = met.cpu().numpy()
    np.testing.assert_allclose(m[:, _capi.M_ACTOR_LOSS], g["actor_losses"], atol=1e-4, rtol=2e-3)
    np.testing.assert_allclose(m[:, _capi.M_CRITIC_LOSS], g["critic_losses"], rtol=2e-3)


def test_learn_on_vecenv_in_bf16_mode(tmp_path):
    env = VecEnv(1024, map="stage_1", seed=2, max_episode_steps=60)
    agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=1024 * 32, max_timesteps_per_episode=60,
                n_updates_per_iteration=4, output_dir=str(tmp_path), method_name="tc", seed=0, verbose=False,
                precision=_capi.PREC_BF16)
    before = agent.flat.clone()
    agent.learn(total_timesteps=1)
    assert torch.isfinite(agent.flat).all() and not torch.equal(before, agent.flat)
    s = agent.logger["summary"]
    assert np.isfinite(list(s.values())).all()


def _raw_handle(prec, max_samples, single_role):
    """A navppo handle outside the cache; `single_role` picks the first tcgen05 kernel (navppo_tc.cu)."""
    import ctypes
    import os
    if single_role:
        os.environ["NAVPPO_TC_KERNEL"] = "single"
    try:
        cfg = _capi.default_ppo_cfg()
        cfg.device = 0
        cfg.max_samples = int(max_samples)
        cfg.precision = int(prec)
        h = ctypes.c_void_p()
        _capi.check(_capi.lib().navppo_create(ctypes.byref(h), ctypes.byref(cfg)))
    finally:
        os.environ.pop("NAVPPO_TC_KERNEL", None)
    return h


@pytest.mark.parametrize("prec", [_capi.PREC_BF16X3, _capi.PREC_BF16])
@pytest.mark.parametrize("T", [1, 127, 128, 129, 5000, 74 * 128 * 3 + 17])
def test_warp_specialised_kernel_is_bit_identical_to_the_single_role_kernel(T, prec):
    """The warp-specialised kernel (transposed backward pass, operands in tensor memory, mbarrier
    pipeline) issues every product in the first kernel's k-order and pass order: same bits, for
    ragged sizes from one sample to several tiles per CTA."""
    g = golden("ppo_learn_b")
    rng = np.random.RandomState(T)
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    act, lp, rtg = g["acts"][idx], (g["logp"][idx] + rng.normal(scale=0.1, size=T)).astype(np.float32), g["rtgs"][idx]
    adv = rng.normal(size=T).astype(np.float32)
    flat = _flat_from(g["actor_after"], g["critic_after"])
    o, a_, l_, ad, rt = _t(obs), _t(act), _t(lp), _t(adv), _t(rtg)
    out = []
    for single in (True, False):
        h = _raw_handle(prec, max(T, 1024), single)
        grad, met = _grad(h, flat, o, a_, l_, ad, rt, T, float(g["var"]))
        out.append((grad.clone(), met.copy()))
        _capi.lib().navppo_destroy(h)
    assert torch.equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert float(out[1][0].abs().max()) > 0


# ---------------------------------------------------------------------------------------- tensor-core inference
def _infer_all(h, flat, obs, act_in, noise, var):
    L = _capi.lib()
    T = obs.shape[0]
    z = lambda *s: torch.zeros(*s, device=DEV)  # noqa: E731
    out = dict(mu=z(T, 2), v=z(T), act=z(T, 2), lp=z(T), mu_act=z(T, 2), v_ev=z(T), lp_ev=z(T))
    _capi.check(L.navppo_forward(h, flat.data_ptr(), obs.data_ptr(), T, out["mu"].data_ptr(), out["v"].data_ptr(), _sp()))
    _capi.check(L.navppo_act(h, flat.data_ptr(), obs.data_ptr(), T, var, 7, 0, 3, noise.data_ptr(), out["act"].data_ptr(),
                             out["lp"].data_ptr(), out["mu_act"].data_ptr(), _sp()))
    _capi.check(L.navppo_evaluate(h, flat.data_ptr(), obs.data_ptr(), act_in.data_ptr(), T, var, out["v_ev"].data_ptr(),
                                  out["lp_ev"].data_ptr(), _sp()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("prec,tol", [(_capi.PREC_BF16X3, 2e-5), (_capi.PREC_BF16, 2e-2)])
@pytest.mark.parametrize("T", [1, 127, 128, 129, 5000])
def test_tc_inference_matches_the_float64_oracle(T, prec, tol):
    """forward / get_action / evaluate of a tensor-core handle (mlp_infer_ws_kernel) against the float64 restatement
    of net_actor.py / net_critic.py / ppo.py:690-737, ragged sizes.  Tolerance relative to 1 + max |reference|:
    2e-5 for split-bf16 operands (measured ~5e-6), 2e-2 for plain bf16 (measured ~2e-3)."""
    g = golden("ppo_learn_b")
    rng = np.random.RandomState(T)
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    acts = np.ascontiguousarray(g["acts"][idx])
    noise = rng.normal(size=(T, 2)).astype(np.float32)
    var = float(g["var"])
    h = _Handles.get(torch.device(DEV), 1 << 16, 0.2, 3e-4, prec)
    flat = _flat_from(g["actor_after"], g["critic_after"])
    out = _infer_all(h, flat, _t(obs), _t(acts), _t(noise), var)
    v_ref, lp_ref = po.evaluate(g["actor_after"], g["critic_after"], obs, acts, var)
    mu_ref = po.actor_forward(g["actor_after"], obs)
    close = lambda got, want: np.abs(got.cpu().numpy() - want).max() <= tol * (1.0 + np.abs(want).max())  # noqa: E731
    assert close(out["mu"], mu_ref) and close(out["v"], v_ref) and close(out["v_ev"], v_ref) and close(out["lp_ev"], lp_ref)
    assert torch.equal(out["mu"], out["mu_act"]) and torch.equal(out["v"], out["v_ev"])
    # get_action: a = clamp(mu + sqrt(var) eps), log-prob at the clamped action (ppo.py:698-704)
    a_ref = mu_ref + np.sqrt(var) * noise
    a_ref[:, 0] = np.clip(a_ref[:, 0], 0.0, 1.0); a_ref[:, 1] = np.clip(a_ref[:, 1], -1.0, 1.0)
    assert close(out["act"], a_ref)
    _, lp_act_ref = po.evaluate(g["actor_after"], g["critic_after"], obs, out["act"].cpu().numpy(), var)
    assert close(out["lp"], lp_act_ref)


@pytest.mark.parametrize("prec", [_capi.PREC_BF16X3, _capi.PREC_BF16])
def test_rollout_log_probs_are_bit_identical_to_what_the_update_recomputes(prec):
    """The inference kernel is the forward half of the gradient kernel: with unchanged weights the update's
    ratio exp(logp - logp_old) is exactly 1 for every sample — approx-KL and clip fraction are exactly 0
    (ppo.py:316,326,335) — wherever a sample sits in its tile."""
    g = golden("ppo_learn_b")
    T = 3 * 128 + 77
    rng = np.random.RandomState(5)
    idx = rng.randint(0, len(g["obs"]), T)
    obs = _t((g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32))
    var = float(g["var"])
    h = _Handles.get(torch.device(DEV), 1 << 16, 0.2, 3e-4, prec)
    flat = _flat_from(g["actor_after"], g["critic_after"])
    act = torch.zeros(T, 2, device=DEV); lp = torch.zeros(T, device=DEV)
    _capi.check(_capi.lib().navppo_act(h, flat.data_ptr(), obs.data_ptr(), T, var, 11, 0, 0, None, act.data_ptr(), lp.data_ptr(),
                                       None, _sp()))
    perm = torch.from_numpy(rng.permutation(T)).to(DEV)            # the update sees the samples at other tile rows
    adv = _t(rng.normal(size=T).astype(np.float32)); rtg = _t(rng.normal(size=T).astype(np.float32))
    _, m = _grad(h, flat, obs[perm].contiguous(), act[perm].contiguous(), lp[perm].contiguous(), adv, rtg, T, var)
    assert m[_capi.M_APPROX_KL] == 0.0 and m[_capi.M_CLIP_FRAC] == 0.0
    assert abs(m[_capi.M_ACTOR_LOSS] + float(adv.double().mean())) < 1e-6        # every surrogate is exactly 1 * A


def test_one_launch_fused_rollout_equals_the_kernel_by_kernel_rollout(tmp_path):
    """A tensor-core handle runs PPO.rollout's whole step loop as ONE launch (each CTA keeps 128 robots for all H
    steps: policy forward on tcgen05, sampling, Env.step by the same threads).  Every rollout buffer, the simulator
    state it leaves behind and the episode statistics are bit-identical to the chain of policy / step kernels
    (NAVPPO_ROLLOUT_FUSED=0), for a robot count that leaves the last tile ragged, episodes that end inside the
    horizon, rollout after rollout."""
    import os

    def make(fused, lr):
        if not fused:
            os.environ["NAVPPO_ROLLOUT_FUSED"] = "0"
        try:
            env = VecEnv(434, map="stage_2", seed=3, max_episode_steps=25)
            return PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=434 * 40, max_timesteps_per_episode=25,
                       n_updates_per_iteration=1, lr=lr, output_dir=str(tmp_path / ("f" if fused else "c")), method_name="r",
                       seed=5, verbose=False, precision=_capi.PREC_BF16X3)
        finally:
            os.environ.pop("NAVPPO_ROLLOUT_FUSED", None)
    a, b = make(True, 3e-4), make(False, 3.0001e-4)       # (the learning rate keys the handle cache; a rollout ignores it)
    b.flat.copy_(a.flat)
    for it in range(3):
        if it == 2:
            a._decay_cov(); b._decay_cov()
        ra, rb = a.rollout([0, 0], 0), b.rollout([0, 0], 0)
        torch.cuda.synchronize()
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y)
        for name in ("_b_rew", "_b_flags", "_b_epret", "_b_eppath", "_b_eplen", "_b_term", "_next_obs"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        assert np.array_equal(ra[4], rb[4]) and ra[5] == rb[5] and ra[5]["ep_count"] > 0
        for f in (_capi.F_X, _capi.F_Y):
            assert np.array_equal(a.env.get_state(f), b.env.get_state(f))
    assert a.env.launch_count < b.env.launch_count        # one simulator-side launch per rollout instead of H


def test_rollout_on_a_map_the_fused_kernel_does_not_step_falls_back_to_the_kernel_chain(tmp_path):
    """House map (208 walls: compacted wall lists, 16 lanes per robot): the fused rollout kernel declines, and the
    tensor-core policy kernel + step kernel chained as programmatic dependent launches return the same batch as the
    two kernels launched plainly (NAVPPO_ROLLOUT_CHAIN=0)."""
    import os

    def make(chain, lr):
        if not chain:
            os.environ["NAVPPO_ROLLOUT_CHAIN"] = "0"
        try:
            env = VecEnv(300, map="house", seed=4, max_episode_steps=30)
            return PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=300 * 24, max_timesteps_per_episode=30,
                       n_updates_per_iteration=1, lr=lr, output_dir=str(tmp_path / ("p" if chain else "q")), method_name="r",
                       seed=6, verbose=False, precision=_capi.PREC_BF16X3)
        finally:
            os.environ.pop("NAVPPO_ROLLOUT_CHAIN", None)
    a, b = make(True, 3.0002e-4), make(False, 3.0003e-4)
    b.flat.copy_(a.flat)
    for _ in range(2):
        ra, rb = a.rollout([0, 0], 0), b.rollout([0, 0], 0)
        torch.cuda.synchronize()
        for x, y in zip(ra[:4], rb[:4]):
            assert torch.equal(x, y)
        assert torch.equal(a._b_rew, b._b_rew) and torch.equal(a._b_flags, b._b_flags)
    assert a.env.launch_count == b.env.launch_count >= 24         # a step launch per env step was captured: not the fused kernel
