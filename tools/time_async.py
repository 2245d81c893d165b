"""Host-buffer env steps per second through navsim_step_host_async / navsim_wait (4 steps in flight), with the
three ways the library can move the host buffers (NAVSIM_ASYNC_OBS):  python tools/time_async.py [agents] [steps]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(N, steps):
    import numpy as np

    from navbot_ppo_b200.env import VecEnv
    env = VecEnv(N, map="stage_1", device=0, seed=0, max_episode_steps=500)
    depth = 4
    sets = [env.alloc_host_buffers() for _ in range(depth)]
    rng = np.random.RandomState(0)
    for sb in sets:
        sb["act"][:] = rng.uniform(0, 1, sb["act"].shape).astype(np.float32)
    env.reset_host()

    acc = [0.0, 0.0]

    def run_one_call(n):
        for i in range(n):
            sb = sets[i % depth]
            env.step_host_pipelined(sb["act"], sb)
        env.wait(0)
    run_one_call(50)
    t0 = time.perf_counter(); run_one_call(steps); dt1 = time.perf_counter() - t0
    print(f"  one library call per step (step_host_pipelined): {N * steps / dt1:.4e} env-steps/s, {dt1 / steps * 1e6:.2f} us per step", flush=True)

    def run(n):
        tickets = []
        for i in range(n):
            if len(tickets) == depth:
                ta = time.perf_counter()
                env.wait(tickets.pop(0))
                acc[1] += time.perf_counter() - ta
            sb = sets[i % depth]
            ta = time.perf_counter()
            tickets.append(env.step_host_async(sb["act"], sb))
            acc[0] += time.perf_counter() - ta
        env.wait(0)
    run(50)
    acc[0] = acc[1] = 0.0
    t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
    print(f"  host time per step: {acc[0] / steps * 1e6:.2f} us inside step_host_async, {acc[1] / steps * 1e6:.2f} us inside wait", flush=True)
    chk = float(sum(np.asarray(sb["obs"], dtype=np.float64).sum() for sb in sets))
    print(f"RESULT {N * steps / dt:.4e} {dt / steps * 1e6:.2f} {chk:.6f}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]), int(sys.argv[3]))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    for name, env in (("actions staged one step ahead, results home in one copy (default)", {}),
                      ("kernel reads host actions, results home in one copy", {"NAVSIM_ASYNC_OBS": "dma"}),
                      ("copy engines for actions + obs", {"NAVSIM_ASYNC_OBS": "dma_act"}),
                      ("kernel reads / stores host buffers", {"NAVSIM_ASYNC_OBS": "stores"})):
        r = subprocess.run([sys.executable, __file__, "--child", str(N), str(steps)], env={**os.environ, **env}, capture_output=True,
                           text=True, timeout=600)
        print("\n".join(l for l in r.stdout.splitlines() if l.startswith("  ")))
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        if not line:
            print(name, "FAILED", r.stdout[-1500:], r.stderr[-1500:])
            return 1
        _, v, us, chk = line[0].split()
        print(f"{name:68s} {N} robots: {v} env-steps/s, {us} us per step, checksum {chk}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
