"""BASELINE configs[3]: throughput of the DRS acceptance pass with the SNGAN-64 discriminator in the CUDA engine.
    python tools/bench_drs.py [--batch 256] [--images 50000]
The generator is outside the diagnosis path (SURVEY 8(d) item 4 allows a stand-in): candidates are tanh(randn) images made
on the device.  Timed: DRS.generate_images(images) = per batch of `batch` candidates the generator stand-in, netD(x) through
EngineNetD (float32 NCHW in, [B,1] out -- the contract of drs.py:24-28), the acceptance kernel (running max, F, percentile
gamma, sigmoid, compare with psi from the global NumPy stream, compaction) and the device-side gather of accepted images."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diagan_b200 import synthetic  # noqa: E402
from diagan_b200.models.drs import DRS  # noqa: E402
from diagan_b200.models.engine_netd import EngineNetD  # noqa: E402


class StandInG:
    def __init__(self, dev):
        self.gen = torch.Generator(device=dev).manual_seed(11)

    def generate_images(self, n, device=None):
        return torch.randn(n, 3, 64, 64, generator=self.gen, device=device).tanh()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--images", type=int, default=50000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    np.random.seed(1)
    netD = EngineNetD(synthetic.sngan_state_dict(64, 1), dev)
    drs = DRS(StandInG(dev), netD, dev, batch_size=a.batch)          # includes the 50 burn-in batches (drs.py:31-36)
    drs.generate_images(a.batch, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = drs.generate_images(a.images, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # candidates scored = accepted / acceptance rate; count them by re-running the bookkeeping cheaply
    x = StandInG(dev).generate_images(a.batch, dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ldr = netD(x)
    e1.record()
    torch.cuda.synchronize()
    d_ms = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20):
        drs.accept(ldr.view(-1))
    e1.record()
    torch.cuda.synchronize()
    a_ms = e0.elapsed_time(e1) / 20
    print(f"DRS SNGAN-64 batch {a.batch}: {out.shape[0]} accepted images in {dt:.2f} s = {out.shape[0] / dt:,.0f} accepted/s "
          f"(wall clock; up to 32 candidate batches per host round trip)")
    print(f"  per batch of {a.batch} candidates: netD forward {d_ms * 1e3:.0f} us ({a.batch / d_ms * 1e3:,.0f} candidates/s), "
          f"acceptance pass {a_ms * 1e3:.0f} us")


if __name__ == "__main__":
    main()
