"""Recording-pass throughput of one discriminator architecture on one GPU (CUDA events), for the configs that are
parity-test cases rather than the headline bench line (BASELINE.json configs[2..4]).
    python tools/bench_arch.py --arch sngan64 [--n 8192] [--precision fp16]
    python tools/bench_arch.py --arch dcgan32 | sngan32 | stylegan2 --size 256 --n 64 --batch 4"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import engine, synthetic  # noqa: E402

FLOP = {"sngan32": 2 * 272_072_832, "sngan64": 2 * 644_809_728, "dcgan32": 2 * 30_789_632}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="sngan64")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--pair", type=int, default=1, help="0: force the single-CTA conv kernel for Cout = 128 layers")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = engine.DiscriminatorEngine(dev)
    eng.lib.sdg_set_conv_pair(a.pair)
    if a.arch.startswith("sngan"):
        size = int(a.arch[5:])
        eng.load_sngan(synthetic.sngan_state_dict(size, 1), size, a.precision, True)
        flop = FLOP[a.arch]
    elif a.arch == "dcgan32":
        size = 32
        eng.load_dcgan(synthetic.dcgan_state_dict(1), a.precision)
        flop = FLOP[a.arch]
    else:
        size = a.size
        eng.load_stylegan2(synthetic.stylegan2_state_dict(size, 1), a.precision, batch=a.batch)
        flop = synthetic.stylegan2_flops(size)
    if a.chunk:
        eng.set_chunk(a.chunk)
    x = synthetic.uniform_images_u8(a.n, size, seed=1).to(dev)
    out = torch.empty(a.n, dtype=torch.float32, device=dev)
    for _ in range(2):
        eng.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        eng.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    rate = a.n / ms * 1e3
    extra = f", {rate * flop / 1e12:.1f} TFLOP/s (reference-formulation FLOPs)" if flop else ""
    print(f"{a.arch} size={size} n={a.n} {a.precision}: {ms:.2f} ms/pass, "
          f"{rate:,.0f} samples/s{extra}")


if __name__ == "__main__":
    main()
