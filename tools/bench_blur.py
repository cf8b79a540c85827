"""StyleGAN2 Blur (16-bit NHWC) against the HBM roofline: the TMA-fed kernel (csrc/blur_tma.cu, variants 1..3) beside the
register-sliding kernel it replaces (csrc/sg2_fp32.cu, SDG_BLUR_TMA=0), on the shapes of the StyleGAN2-256 discriminator.
    python tools/bench_blur.py [--n 56]
Algorithmic bytes: every input element read once + every output element written once (2 B each)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import _lib, engine  # noqa: E402,F401
from diagan_b200._lib import check, ptr, stream_ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=56)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    engine.DiscriminatorEngine(dev)                 # per-device initialisation (tensor-map encoder)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    print(f"n = {a.n} images per launch, fp16, HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json)")
    print("| H x W x C, pad, stride | MB | " + " | ".join(f"SDG_BLUR_TMA={v}: us (frac)" for v in (0, 1, 2, 3)) + " | bit-equal |")
    print("|---|---:|" + "---:|" * 5)
    for (H, C) in ((256, 128), (128, 256), (64, 512), (32, 512)):
        n = a.n * (256 // H)
        x = torch.randn(n, H, H, C, device=dev, dtype=torch.float16)
        for pad, st in ((2, 1), (1, 2)):
            ho = (H + 2 * pad - 4) // st + 1
            outs, cells = [], []
            nbytes = 2.0 * (x.numel() + n * ho * ho * C)
            for v in (0, 1, 2, 3):
                os.environ["SDG_BLUR_TMA"] = str(v)
                out = torch.full((n, ho, ho, C), float("nan"), device=dev, dtype=torch.float16)
                fn = lambda: check(lib.sdg_blur_h16(ptr(x), ptr(out), n, H, H, C, pad, st, _lib.PREC_FP16, stream_ptr(dev)), "blur")
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / a.reps * 1e-3
                cells.append(f"{t * 1e6:.0f} ({nbytes / t / 1e9 / peak:.2f})")
                outs.append(out)
            same = all(torch.equal(outs[0].view(torch.int16), o.view(torch.int16)) for o in outs[1:])
            print(f"| {H} x {H} x {C}, {pad}, {st} | {nbytes / 1e6:.0f} | " + " | ".join(cells) + f" | {same} |")
            del outs
        del x
    os.environ.pop("SDG_BLUR_TMA", None)


if __name__ == "__main__":
    main()
