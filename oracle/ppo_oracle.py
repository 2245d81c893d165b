"""ppo_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

float64 numpy restatement of the PPO side of the reference's hot path, used only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg to CHECK the CUDA kernels
(navbot_ppo_b200/csrc/navppo_kernels.cu).  Every function cites the reference lines it
follows.  Pinned against tests/golden/ppo_*.npz, which oracle/make_golden_ppo.py records by
executing the reference's own net_actor.py / net_critic.py / ppo.py (torch fp32): the
restatement computes in float64, so it agrees with the reference to fp32 rounding (1e-5).
"""
from __future__ import annotations

import math

import numpy as np

from navbot_ppo_b200 import layout

LEAK = 0.2  # nn.LeakyReLU(negative_slope=0.2), net_actor.py:37


def _lrelu(x):
    return np.where(x > 0, x, LEAK * x)


def _dlrelu(x):
    # torch leaky_relu_backward: grad * (x > 0 ? 1 : slope)
    return np.where(x > 0, 1.0, LEAK)


def trunk_forward(p, x0):
    """ResBlock x2 (net_actor.py:39-53, 138-139): returns every intermediate for backward."""
    z1 = x0 @ p["rb1.fc1.weight"].T + p["rb1.fc1.bias"]          # :43
    h1 = _lrelu(z1)                                              # :45
    u1 = x0 + h1 @ p["rb1.fc2.weight"].T + p["rb1.fc2.bias"]     # :47-49 (Fin == Fout: Xin = x)
    y1 = _lrelu(u1)                                              # :52
    x1 = np.concatenate([x0, y1], axis=-1)                       # :139 cat[X0, X]
    z2 = x1 @ p["rb2.fc1.weight"].T + p["rb2.fc1.bias"]
    h2 = _lrelu(z2)
    u2 = x1 + h2 @ p["rb2.fc2.weight"].T + p["rb2.fc2.bias"]
    y2 = _lrelu(u2)
    return dict(x0=x0, z1=z1, h1=h1, u1=u1, y1=y1, x1=x1, z2=z2, h2=h2, u2=u2, y2=y2)


def _params(flat, kind):
    return {k: np.asarray(v, dtype=np.float64) for k, v in layout.unflatten(np.asarray(flat), kind).items()}


def actor_forward(flat, obs, want_cache=False):
    """NetActor.forward, net_actor.py:94-144: mu = [sigmoid(out1 X), tanh(out2 X)]."""
    p = _params(flat, "actor")
    c = trunk_forward(p, np.asarray(obs, dtype=np.float64).reshape(-1, layout.OBS_DIM))
    o1 = c["y2"] @ p["out1.weight"].T + p["out1.bias"]
    o2 = c["y2"] @ p["out2.weight"].T + p["out2.bias"]
    mu = np.concatenate([1.0 / (1.0 + np.exp(-o1)), np.tanh(o2)], axis=-1)   # :141-143
    return (mu, c, p) if want_cache else mu


def critic_forward(flat, obs, want_cache=False):
    """NetCritic.forward, net_critic.py:83-130: V = out(X), squeezed by ppo.py:725."""
    p = _params(flat, "critic")
    c = trunk_forward(p, np.asarray(obs, dtype=np.float64).reshape(-1, layout.OBS_DIM))
    v = (c["y2"] @ p["out.weight"].T + p["out.bias"])[:, 0]
    return (v, c, p) if want_cache else v


def gaussian_logp(mu, act, var):
    """MultivariateNormal(mu, var*I).log_prob(act) for k = 2 (ppo.py:698,704,734-735)."""
    d = np.asarray(act, dtype=np.float64) - mu
    return -0.5 * (d * d).sum(-1) / var - math.log(2.0 * math.pi) - math.log(var)


def get_action(actor_flat, obs, noise, var):
    """PPO.get_action, ppo.py:673-706, with the N(0, I) draw passed in: a = mu + sqrt(var) * eps
    (MultivariateNormal.rsample: loc + scale_tril @ eps), clamp (:700-703), log-prob AT THE
    CLAMPED action (:704)."""
    mu = actor_forward(actor_flat, obs)
    a = mu + math.sqrt(var) * np.asarray(noise, dtype=np.float64)
    a = np.stack([np.clip(a[:, 0], 0.0, 1.0), np.clip(a[:, 1], -1.0, 1.0)], axis=1)
    # the reference holds the sample in fp32 before clamping and scoring it
    a = a.astype(np.float32).astype(np.float64)
    return a, gaussian_logp(mu, a, var)


def evaluate(actor_flat, critic_flat, obs, act, var):
    """PPO.evaluate, ppo.py:708-737 (the clamp of the mean at :730-733 is the identity after
    sigmoid / tanh)."""
    v = critic_forward(critic_flat, obs)
    mu = actor_forward(actor_flat, obs)
    return v, gaussian_logp(mu, act, var)


def compute_rtgs(batch_rews, gamma):
    """PPO.compute_rtgs, ppo.py:643-671: per episode, reverse, R = r + gamma * R, no bootstrap;
    accumulated in Python floats, cast to fp32 at the end (:669)."""
    out = []
    for ep in reversed(batch_rews):
        r = 0.0
        for rew in reversed(ep):
            r = float(rew) + r * gamma
            out.insert(0, r)
    return np.asarray(out, dtype=np.float32)


def rtg_scan(rew, term, gamma, last_value=None):
    """The same recurrence on the vectorised [H, N] layout: `term[t, n]` marks the last step
    of an episode (done | arrive | timeout, ppo.py:552-553); the end of the horizon is a
    boundary too (the trailing partial episode is appended as it is, ppo.py:601)."""
    rew = np.asarray(rew, dtype=np.float64)
    H = rew.shape[0]
    out = np.zeros_like(rew)
    acc = np.zeros(rew.shape[1:], dtype=np.float64) if last_value is None else np.asarray(last_value, dtype=np.float64).copy()
    for t in range(H - 1, -1, -1):
        acc = np.where(np.asarray(term[t]).astype(bool), 0.0, acc)
        acc = rew[t] + gamma * acc
        out[t] = acc
    return out.astype(np.float32)


def gae_scan(rew, term, values, last_value, gamma, lam):
    """Generalised advantage estimation on the [H, N] layout (not in the reference, which is
    the lam = 1, zero-bootstrap special case; kept so the kernel's general form has a check)."""
    rew = np.asarray(rew, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    H = rew.shape[0]
    adv = np.zeros_like(rew)
    acc = np.zeros(rew.shape[1:], dtype=np.float64)
    nxt = np.asarray(last_value, dtype=np.float64)
    for t in range(H - 1, -1, -1):
        live = 1.0 - np.asarray(term[t]).astype(np.float64)
        delta = rew[t] + gamma * nxt * live - v[t]
        acc = delta + gamma * lam * live * acc
        adv[t] = acc
        nxt = v[t]
    return adv.astype(np.float32)


def advantage(rtgs, v):
    """ppo.py:277,284: A = rtg - V; A = (A - mean) / (std + 1e-10), torch.std is unbiased."""
    a = np.asarray(rtgs, dtype=np.float64) - np.asarray(v, dtype=np.float64)
    return (a - a.mean()) / (a.std(ddof=1) + 1e-10)


def _trunk_backward(p, c, g_y2):
    """Backward through the two ResBlocks; returns {key: grad} for the trunk parameters."""
    g = {}
    g_u2 = g_y2 * _dlrelu(c["u2"])
    g["rb2.fc2.weight"] = g_u2.T @ c["h2"]
    g["rb2.fc2.bias"] = g_u2.sum(0)
    g_z2 = (g_u2 @ p["rb2.fc2.weight"]) * _dlrelu(c["z2"])
    g["rb2.fc1.weight"] = g_z2.T @ c["x1"]
    g["rb2.fc1.bias"] = g_z2.sum(0)
    g_x1 = g_u2 + g_z2 @ p["rb2.fc1.weight"]                    # skip connection + fc1
    g_y1 = g_x1[:, layout.OBS_DIM:]
    g_u1 = g_y1 * _dlrelu(c["u1"])
    g["rb1.fc2.weight"] = g_u1.T @ c["h1"]
    g["rb1.fc2.bias"] = g_u1.sum(0)
    g_z1 = (g_u1 @ p["rb1.fc2.weight"]) * _dlrelu(c["z1"])
    g["rb1.fc1.weight"] = g_z1.T @ c["x0"]
    g["rb1.fc1.bias"] = g_z1.sum(0)
    return g


def _flat_grad(g, kind):
    return np.concatenate([np.asarray(g[name], dtype=np.float64).ravel() for name, _ in layout.spec(kind)])


def losses_and_grads(actor_flat, critic_flat, obs, act, logp_old, adv, rtgs, var, clip, n_global=None):
    """One epoch body of PPO.learn, ppo.py:305-343 + the two backward() calls (:349,386).
    Returns (metrics, actor_grad_flat, critic_grad_flat).  `n_global` = the divisor of the
    .mean() losses (the global batch when samples are sharded over ranks)."""
    obs = np.asarray(obs, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    adv = np.asarray(adv, dtype=np.float64)
    rtgs = np.asarray(rtgs, dtype=np.float64)
    logp_old = np.asarray(logp_old, dtype=np.float64)
    T = obs.shape[0]
    n = float(n_global if n_global is not None else T)
    mu, ca, pa = actor_forward(actor_flat, obs, want_cache=True)
    v, cc, pc = critic_forward(critic_flat, obs, want_cache=True)
    logp = gaussian_logp(mu, act, var)
    log_ratio = logp - logp_old
    ratio = np.exp(log_ratio)                                      # :316
    surr1 = ratio * adv                                            # :319
    surr2 = np.clip(ratio, 1 - clip, 1 + clip) * adv               # :320
    actor_loss = (-np.minimum(surr1, surr2)).sum() / n             # :342
    critic_loss = ((v - rtgs) ** 2).sum() / n                      # :343 nn.MSELoss
    metrics = dict(actor_loss=actor_loss, critic_loss=critic_loss,
                   approx_kl=((ratio - 1) - log_ratio).sum() / n,              # :326
                   clip_frac=(np.abs(ratio - 1.0) > clip).sum() / n)           # :335
    # d(-min(s1, s2))/d ratio: torch.min splits ties, and clamp passes gradient on its closed
    # interval, so the derivative is -adv wherever s1 <= s2 and 0 where the clipped branch wins
    g_ratio = np.where(surr1 <= surr2, -adv, 0.0) / n
    g_logp = g_ratio * ratio
    g_mu = g_logp[:, None] * (act - mu) / var
    g_o1 = g_mu[:, 0:1] * mu[:, 0:1] * (1.0 - mu[:, 0:1])          # sigmoid'
    g_o2 = g_mu[:, 1:2] * (1.0 - mu[:, 1:2] ** 2)                  # tanh'
    ga = {"out1.weight": g_o1.T @ ca["y2"], "out1.bias": g_o1.sum(0),
          "out2.weight": g_o2.T @ ca["y2"], "out2.bias": g_o2.sum(0)}
    ga.update(_trunk_backward(pa, ca, g_o1 @ pa["out1.weight"] + g_o2 @ pa["out2.weight"]))
    g_v = (2.0 * (v - rtgs) / n)[:, None]
    gc = {"out.weight": g_v.T @ cc["y2"], "out.bias": g_v.sum(0)}
    gc.update(_trunk_backward(pc, cc, g_v @ pc["out.weight"]))
    fa, fc = _flat_grad(ga, "actor"), _flat_grad(gc, "critic")
    metrics["actor_grad_norm"] = float(np.sqrt((fa ** 2).sum()))   # :352 clip_grad_norm_(.., inf)
    metrics["critic_grad_norm"] = float(np.sqrt((fc ** 2).sum()))  # :389
    return metrics, fa, fc


class Adam:
    """torch.optim.Adam(params, lr) with its defaults (ppo.py:116-117): betas (0.9, 0.999),
    eps 1e-8, no weight decay, no amsgrad."""

    def __init__(self, n, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8):
        self.m, self.v, self.t = np.zeros(n), np.zeros(n), 0
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def step(self, p, g):
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        denom = np.sqrt(self.v) / math.sqrt(bc2) + self.eps
        return p - (self.lr / bc1) * self.m / denom


def update(actor_flat, critic_flat, obs, act, logp_old, rtgs, var, clip, lr, epochs, adam_a=None, adam_c=None):
    """The update part of one PPO.learn iteration (ppo.py:275-397): evaluate -> advantage ->
    `epochs` full-batch epochs, actor then critic step per epoch."""
    a = np.asarray(actor_flat, dtype=np.float64).copy()
    c = np.asarray(critic_flat, dtype=np.float64).copy()
    v0, _ = evaluate(a, c, obs, act, var)                           # :275
    adv = advantage(rtgs, v0)                                       # :277-284
    adam_a = adam_a or Adam(a.size, lr)
    adam_c = adam_c or Adam(c.size, lr)
    hist = []
    for _ in range(epochs):
        m, ga, gc = losses_and_grads(a, c, obs, act, logp_old, adv, rtgs, var, clip)
        a = adam_a.step(a, ga)                                      # :381
        c = adam_c.step(c, gc)                                      # :392
        hist.append(m)
    return a, c, adv, hist
