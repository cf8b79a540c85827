"""Host-side cost of one recording step on a SHORT shard (what an 8-GPU split of the 50 k dataset leaves per GPU): device time per
step (CUDA events) of the resident and of the end-to-end step beside the wall time the host needs to ISSUE one step, and a
cProfile of the issuing loop.
    python tools/bench_host_overhead.py [--n 6250] [--steps 50] [--profile 1]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import engine, synthetic  # noqa: E402
from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=6250)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--profile", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    base = {k: v.to(dev) for k, v in synthetic.sngan_state_dict(32, seed=7).items()}
    sets = [synthetic.perturb_(base, 35000 + 100 * i, 1e-3, device=dev) for i in range(8)]
    host = synthetic.uniform_images_u8(a.n, 32, seed=1, pin=True)
    rec = LogitRecorder(ResidentDataset(host.to(dev)), dev, precision="fp16", inplace_relu=True, keep_snapshots=False, batch=4)
    snap = torch.zeros(a.n, dtype=torch.float32, device=dev)
    t_conf = 0.3
    out = [(torch.empty(a.n, dtype=torch.float64).pin_memory(), torch.empty(100, dtype=torch.int64).pin_memory(), torch.cuda.Event())
           for _ in range(2)]

    def finish():
        full = rec.stats.score(t_conf, eps=1e-6)
        return full, engine.top_indices(full, 100, True)

    def step_resident(i):
        rec.record(sets[i % 8], step=i, out=snap, range_check="deferred")
        return finish()

    def step_host(i):
        rec.record_from_host(sets[i % 8], host, step=i, chunk=12544, first_chunk=2048, range_check="deferred")
        full, top = finish()
        hf, ht, ev = out[i & 1]
        hf.copy_(full, non_blocking=True)
        ht.copy_(top, non_blocking=True)
        ev.record()
        out[(i + 1) & 1][2].synchronize()
        return hf, ht

    for name, fn in (("resident", step_resident), ("end to end", step_host)):
        rec.stats = None
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(a.steps):
            fn(5 + i)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"{name:10s} n={a.n}: device {e0.elapsed_time(e1) / a.steps:.3f} ms per step, host issue loop "
              f"{1e3 * t_issue / a.steps:.3f} ms per step")
        if a.profile:
            pr = cProfile.Profile()
            pr.enable()
            for i in range(a.steps):
                fn(5 + i)
            pr.disable()
            torch.cuda.synchronize()
            st = pstats.Stats(pr)
            st.sort_stats("cumulative")
            print(f"---- cProfile of {a.steps} {name} steps (top 28 by cumulative time)")
            st.print_stats(28)


if __name__ == "__main__":
    main()
