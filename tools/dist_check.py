"""torchrun --nproc-per-node 2 tools/dist_check.py
The ranks each update on their share of a fixed batch (3-double all-reduce of the advantage statistics, per-epoch
exchange of the flat gradient) and must land on the parameters a single GPU reaches on the whole batch — for every
gradient-exchange mode: "nccl" (all-reduce + Adam), "peer" (Adam adds the ranks' peer-mapped buffers itself, in rank
order) and "nvls" (multimem.ld_reduce, where the platform has a multicast address).  Every rank must end with
bit-identical weights in the nccl and peer modes."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navbot_ppo_b200 import _capi, dist as navdist, layout  # noqa: E402
from navbot_ppo_b200.env import VecEnv  # noqa: E402
from navbot_ppo_b200.nets import NetActor, NetCritic  # noqa: E402
from navbot_ppo_b200.ppo import PPO  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ppo_learn_b.npz"))
    rng = np.random.RandomState(0)
    T = 20000
    idx = rng.randint(0, len(g["obs"]), T)
    obs = (g["obs"][idx] + rng.normal(scale=0.01, size=(T, 16))).astype(np.float32)
    act, lp, rtg = g["acts"][idx], g["logp"][idx], g["rtgs"][idx]
    ok = True
    for prec, tol, exchange in ((_capi.PREC_FP32, 2e-5, "nccl"), (_capi.PREC_FP32, 2e-5, "peer"), (_capi.PREC_BF16X3, 5e-5, "peer"),
                                (_capi.PREC_BF16X3, 5e-5, "nvls"), (_capi.PREC_BF16X3, 5e-5, "nccl")):
        results = {}
        used = exchange
        for mode in ("sharded", "single"):
            with tempfile.TemporaryDirectory() as tmp:
                torch.manual_seed(0)
                env = VecEnv(64, device=local, seed=0, agent_id_offset=rank * 64)
                agent = PPO(NetActor, NetCritic, env, 16, 2, timesteps_per_batch=64, n_updates_per_iteration=3, lr=3e-4,
                            output_dir=tmp, method_name=f"r{rank}", verbose=False, precision=prec,
                            gradient_exchange=exchange)
                if mode == "sharded":
                    used = "nccl" if agent._peer is None else ("nvls" if agent._peer["nvls"] else "peer")
                if mode == "single":
                    agent.world = 1
                    lo, hi = 0, T
                else:
                    lo, hi = navdist.shard_range(T, rank, world)
                t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
                res = agent.update(t(obs[lo:hi]), t(act[lo:hi]), t(lp[lo:hi]), t(rtg[lo:hi]), epochs=3)
                results[mode] = (agent.flat.clone(), res)
        d = float((results["sharded"][0] - results["single"][0]).abs().max())
        la = np.abs(results["sharded"][1]["actor_losses"] - results["single"][1]["actor_losses"]).max()
        lc = np.abs(results["sharded"][1]["critic_losses"] / results["single"][1]["critic_losses"] - 1).max()
        # every rank must hold identical weights after the update
        w = results["sharded"][0].clone()
        dist.broadcast(w, src=0)
        same = bool(torch.equal(w, results["sharded"][0]))
        if rank == 0:
            print(f"precision={prec} world={world} exchange={exchange} (in force: {used}): max |param(sharded) - param(single)| = {d:.3e} (tol {tol:.0e}); "
                  f"actor-loss diff {la:.2e}; critic-loss rel diff {lc:.2e}", flush=True)
        flag = torch.tensor([1.0 if ((same or used == "nvls") and d <= tol) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item() == 1.0)
        if rank == 0:
            print(f"  ranks hold identical weights: {bool(flag.item() == 1.0)}", flush=True)
    if rank == 0:
        print("DIST_CHECK", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
