"""CPU, world_size 2 over gloo: the host-side data-parallel logic of the N > 1 path
(navbot_ppo_b200/dist.py) — agent sharding, the 3-double advantage-statistics exchange and the
per-epoch gradient all-reduce with the 1/n_global convention — checked with the float64 oracle
standing in for the kernels (the kernels themselves are checked on the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from navbot_ppo_b200 import dist as navdist
from oracle import ppo_oracle as po
from tests.helpers import golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = golden("ppo_learn_b")
        T = len(g["obs"])
        lo, hi = navdist.shard_range(T, rank, world)
        var, clip = float(g["var"]), float(g["clip"])
        a = torch.from_numpy(g["actor_before"].astype(np.float64))
        c = torch.from_numpy(g["critic_before"].astype(np.float64))
        if rank != 0:                       # every rank starts from rank 0's weights
            a.zero_(); c.zero_()
        navdist.broadcast_(a); navdist.broadcast_(c)
        a, c = a.numpy(), c.numpy()
        # advantage statistics: local (sum, sum^2, n) -> global mean / unbiased std
        v, _ = po.evaluate(a, c, g["obs"][lo:hi], g["acts"][lo:hi], var)
        A = g["rtgs"][lo:hi].astype(np.float64) - v
        mean, std, n = navdist.advantage_moments(float(A.sum()), float((A * A).sum()), float(A.size))
        adv = (A - mean) / (std + 1e-10)
        # gradient: each rank's share with the global divisor, then ONE all-reduce (sum)
        m, ga, gc = po.losses_and_grads(a, c, g["obs"][lo:hi], g["acts"][lo:hi], g["logp"][lo:hi], adv, g["rtgs"][lo:hi], var,
                                        clip, n_global=int(n))
        flat = torch.from_numpy(np.concatenate([ga, gc]))
        navdist.all_reduce_sum_(flat)
        met = torch.tensor([m["actor_loss"], m["critic_loss"], m["approx_kl"], m["clip_frac"]], dtype=torch.float64)
        navdist.all_reduce_sum_(met)
        if rank == 0:
            np.savez(out, grad=flat.numpy(), met=met.numpy(), n=n, mean=mean, std=std, world=navdist.world_size())
    finally:
        td.destroy_process_group()


def test_two_rank_update_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    g = golden("ppo_learn_b")
    var, clip = float(g["var"]), float(g["clip"])
    v, _ = po.evaluate(g["actor_before"], g["critic_before"], g["obs"], g["acts"], var)
    A = g["rtgs"].astype(np.float64) - v
    assert int(r["world"]) == 2 and int(r["n"]) == len(A)
    assert abs(float(r["mean"]) - A.mean()) < 1e-9 and abs(float(r["std"]) - A.std(ddof=1)) < 1e-9
    m, ga, gc = po.losses_and_grads(g["actor_before"], g["critic_before"], g["obs"], g["acts"], g["logp"], po.advantage(g["rtgs"], v),
                                    g["rtgs"], var, clip)
    np.testing.assert_allclose(r["grad"], np.concatenate([ga, gc]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["met"], [m["actor_loss"], m["critic_loss"], m["approx_kl"], m["clip_frac"]], rtol=1e-9)
    # and that IS the reference's gradient (fixture recorded from ppo.py)
    np.testing.assert_allclose(r["grad"][:len(ga)], g["actor_grads"][0], atol=1e-4 * np.abs(g["actor_grads"][0]).max())


def test_shard_ranges_partition_the_agents():
    for n, w in ((8192, 8), (65536, 8), (10, 3), (7, 8), (32768, 4)):
        spans = [navdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_helpers_are_no_ops():
    t = torch.arange(4.0)
    assert navdist.world_size() == 1 and navdist.rank() == 0
    assert torch.equal(navdist.all_reduce_sum_(t.clone()), t) and torch.equal(navdist.broadcast_(t.clone()), t)
    mean, std, n = navdist.advantage_moments(10.0, 30.0, 4.0)
    assert (mean, n) == (2.5, 4.0) and abs(std - np.std([1, 2, 3, 4], ddof=1)) < 1e-12
