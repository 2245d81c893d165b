"""CPU: the float64 restatement of the PPO side (oracle/ppo_oracle.py) against golden vectors
recorded from the reference's own torch code (oracle/make_golden_ppo.py).  This pins the
oracle the GPU tests check the kernels with."""
import numpy as np
import pytest

from navbot_ppo_b200 import layout
from oracle import ppo_oracle as po
from tests.helpers import golden

FWD_ATOL = 1e-5      # MLP forward <= 1e-5 against the fp32 reference
GRAD_RTOL = 1e-4     # loss / gradient relative (reduction order)


def test_layout_matches_reference_state_dict():
    g = golden("ppo_nets")
    assert g["actor"].size == layout.ACTOR_PARAMS == 50290 and g["critic"].size == layout.CRITIC_PARAMS == 50257
    for kind, keys, shapes in (("actor", g["actor_keys"], g["actor_shapes"]), ("critic", g["critic_keys"], g["critic_shapes"])):
        ref = dict(zip(keys.tolist(), shapes.tolist()))
        for name, shape in layout.spec(kind):
            assert ref[name] == str(tuple(shape)), name
        # everything else in the reference's state_dict is a dead BatchNorm entry
        live = {n for n, _ in layout.spec(kind)}
        for k in ref:
            assert k in live or any(k.startswith(bn + ".") for bn, _ in layout.DEAD_BATCHNORMS), k


def test_forward_matches_reference_nets():
    g = golden("ppo_nets")
    np.testing.assert_allclose(po.actor_forward(g["actor"], g["obs"]), g["mu"], atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(po.critic_forward(g["critic"], g["obs"]), g["v"][:, 0], atol=FWD_ATOL, rtol=0)
    np.testing.assert_allclose(po.actor_forward(g["actor"], g["obs"][0])[0], g["mu_single"][0], atol=FWD_ATOL, rtol=0)


def test_rtgs_match_reference_compute_rtgs():
    g = golden("ppo_rtgs")
    eps, off = [], 0
    for n in g["lens"]:
        eps.append(g["rews"][off:off + n].tolist()); off += n
    np.testing.assert_array_equal(po.compute_rtgs(eps, float(g["gamma"])), g["rtgs"])
    # the vectorised [H, N] scan on the same data laid out as one column
    term = np.zeros(len(g["rews"]), np.uint8)
    term[np.cumsum(g["lens"]) - 1] = 1
    np.testing.assert_array_equal(po.rtg_scan(g["rews"][:, None], term[:, None], float(g["gamma"]))[:, 0], g["rtgs"])


def test_get_action_matches_reference():
    g = golden("ppo_get_action")
    # the actor that drew these actions is the seed-0 agent of ppo_learn_a before training
    actor = golden("ppo_learn_a")["actor_before"]
    mu = po.actor_forward(actor, g["obs"])
    np.testing.assert_allclose(mu, g["mean"], atol=FWD_ATOL, rtol=0)
    act, logp = po.get_action(actor, g["obs"], g["eps"], float(g["var"]))
    np.testing.assert_allclose(act, g["act"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logp, g["logp"], atol=1e-5, rtol=1e-6)
    assert ((g["act"][:, 0] == 0) | (g["act"][:, 0] == 1) | (np.abs(g["act"][:, 1]) == 1)).any(), "no clamped sample in the fixture"


@pytest.mark.parametrize("tag", ["a", "b"])
def test_learn_iteration_matches_reference(tag):
    g = golden("ppo_learn_" + tag)
    var, clip, lr, epochs = float(g["var"]), float(g["clip"]), float(g["lr"]), int(g["epochs"])
    # reward-to-go of the recorded batch, incl. the trailing partial episode (ppo.py:601)
    eps, off = [], 0
    for n in g["ep_lens_all"]:
        eps.append(g["rews"][off:off + n].tolist()); off += n
    np.testing.assert_array_equal(po.compute_rtgs(eps, float(g["gamma"])), g["rtgs"])
    assert g["lens"].sum() <= len(g["rtgs"]) == len(g["obs"])       # batch_lens omits the partial episode
    # first evaluate: V mean as logged (ppo.py:276)
    v0, logp0 = po.evaluate(g["actor_before"], g["critic_before"], g["obs"], g["acts"], var)
    np.testing.assert_allclose(logp0, g["logp"], atol=2e-5, rtol=0)  # same actor -> ratio == 1 in epoch 0
    # epoch-by-epoch: losses and gradients
    a = g["actor_before"].astype(np.float64); c = g["critic_before"].astype(np.float64)
    adv = po.advantage(g["rtgs"], v0)
    adam_a, adam_c = po.Adam(a.size, lr), po.Adam(c.size, lr)
    kl, cf, gna, gnc = [], [], [], []
    for e in range(epochs):
        if e == epochs - 1:  # self.V left behind by the last evaluate (ppo.py:307)
            assert abs(po.critic_forward(c, g["obs"]).mean() - float(g["v_mean_last_epoch"])) < 1e-5
        m, ga, gc = po.losses_and_grads(a, c, g["obs"], g["acts"], g["logp"], adv, g["rtgs"], var, clip)
        assert abs(m["actor_loss"] - g["actor_losses"][e]) <= 1e-5 + GRAD_RTOL * abs(g["actor_losses"][e]), e
        assert abs(m["critic_loss"] - g["critic_losses"][e]) <= GRAD_RTOL * abs(g["critic_losses"][e]), e
        for got, want in ((ga, g["actor_grads"][e]), (gc, g["critic_grads"][e])):
            scale = np.abs(want).max()
            np.testing.assert_allclose(got, want, atol=GRAD_RTOL * scale, rtol=0, err_msg=f"epoch {e}")
        a = adam_a.step(a, ga); c = adam_c.step(c, gc)
        kl.append(m["approx_kl"]); cf.append(m["clip_frac"]); gna.append(m["actor_grad_norm"]); gnc.append(m["critic_grad_norm"])
    np.testing.assert_allclose(a, g["actor_after"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(c, g["critic_after"], atol=2e-5, rtol=0)
    assert abs(np.mean(kl) - float(g["approx_kl"])) < 1e-5 + 1e-3 * abs(float(g["approx_kl"]))
    assert abs(np.mean(cf) - float(g["clip_frac"])) < 2.0 / len(g["obs"])
    assert abs(np.mean(gna) - float(g["actor_grad_norm"])) < GRAD_RTOL * float(g["actor_grad_norm"])
    assert abs(np.mean(gnc) - float(g["critic_grad_norm"])) < GRAD_RTOL * float(g["critic_grad_norm"])
    assert abs(np.linalg.norm(a - g["actor_before"]) - float(g["actor_param_delta"])) < 1e-3 * float(g["actor_param_delta"])
    if tag == "b":
        assert float(g["clip_frac"]) > 0.01, "fixture b is meant to exercise the clipped branch"
    # the one-call form used by the GPU tests
    a2, c2, _, _ = po.update(g["actor_before"], g["critic_before"], g["obs"], g["acts"], g["logp"], g["rtgs"], var, clip,
                             lr, epochs)
    np.testing.assert_array_equal(a2, a); np.testing.assert_array_equal(c2, c)


def test_gae_reduces_to_rtg_minus_value():
    rng = np.random.RandomState(0)
    H, N = 40, 6
    rew = rng.normal(size=(H, N)); v = rng.normal(size=(H, N))
    term = (rng.uniform(size=(H, N)) < 0.1).astype(np.uint8)
    term[-1] = 1
    adv = po.gae_scan(rew, term, v, np.zeros(N), 0.99, 1.0)
    np.testing.assert_allclose(adv, po.rtg_scan(rew, term, 0.99) - v, atol=1e-5)
