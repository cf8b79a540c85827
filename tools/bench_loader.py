"""Throughput of the DataLoader-fed recording pass: what a user of the UNMODIFIED scripts gets after
``diagan_b200.patch.install()`` without building a ResidentDataset (LogTrainer._get_logit -> LogitRecorder.record_from_loader):
shuffled batches of 64 with the reference's item contract (data, target, weight, index) (predefined.py:22-24;
train_mimicry_phase1.py:18-24 uses batch_size 64), forward in the CUDA engine, logits scattered by dataset index.
    python tools/bench_loader.py [--n 50000] [--batch 64] [--workers 0]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import synthetic  # noqa: E402
from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset  # noqa: E402


class Items(torch.utils.data.Dataset):
    def __init__(self, x):
        self.x = x

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], 0, 1.0, i


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--workers", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    x_u8 = synthetic.uniform_images_u8(a.n, 32, seed=1)
    xf = ((x_u8.permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5).contiguous()       # what the reference transform yields
    sd = {k: v.to(dev) for k, v in synthetic.sngan_state_dict(32, seed=1).items()}
    loader = torch.utils.data.DataLoader(Items(xf), batch_size=a.batch, shuffle=True, num_workers=a.workers)
    rec = LogitRecorder(None, dev)
    for group in (1, 16):
        rec.record_from_loader(sd, loader, group=group)                              # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec.record_from_loader(sd, loader, group=group)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"record_from_loader  n={a.n} batch={a.batch} workers={a.workers} group={group:2d}: {dt * 1e3:8.1f} ms  "
              f"{a.n / dt:12,.0f} samples/s")
    t0 = time.perf_counter()
    for _ in loader:
        pass
    dt = time.perf_counter() - t0
    print(f"the DataLoader alone (collate of {a.batch}-sample float32 batches, no GPU work):   {dt * 1e3:8.1f} ms  {a.n / dt:12,.0f} samples/s")
    res = LogitRecorder(ResidentDataset(x_u8.to(dev)), dev)
    res.record(sd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res.record(sd)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"resident uint8 dataset (LogitRecorder.record), same engine:                  {dt * 1e3:8.1f} ms  {a.n / dt:12,.0f} samples/s")


if __name__ == "__main__":
    main()
