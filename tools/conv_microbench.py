"""Kernel-level timing of one fused conv stage through the C ABI (CUDA events on the launching stream).
    python tools/conv_microbench.py [--n 4096] [--hw 32] [--cin 128] [--cout 128] [--pool 0] [--pair 1] [--iters 20]
Environment switches of libsdg apply: SDG_SWAP, SDG_SWAP_STAGES, SDG_PAIR_STAGES, SDG_PAIR_STREAM128 (kernel selection /
pipeline depth, results stay correct); SDG_DEBUG_SKIP_A / SDG_DEBUG_SKIP_EPI (loads or epilogues skipped: timing only, wrong
results) only in a library built with `make -C self-diagnosing-gan_b200/csrc EXTRA=-DSDG_TIMING_EXPERIMENTS`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import _lib  # noqa: E402
from diagan_b200._lib import check, ptr, stream_ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--cin", type=int, default=128)
    ap.add_argument("--cout", type=int, default=128)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--pair", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--first", type=int, default=0, help="time the first conv (from uint8 bytes) instead")
    ap.add_argument("--b1fused", type=int, default=0, help="time SNGAN-32 block 1 as one launch (conv_b1fused.cu) and as the two "
                    "launches it replaces")
    ap.add_argument("--fused-only", type=int, default=0)
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    if a.b1fused:
        S, ch = (64, 64) if a.b1fused == 64 else (32, 128)      # --b1fused 64: SNGAN-64 (per image quadrant); else SNGAN-32
        img = torch.randint(0, 256, (a.n, S, S, 3), dtype=torch.uint8, device=dev)
        w1 = torch.zeros(ch, 64, device=dev).half()
        w1[:, :27] = (torch.randn(ch, 27, device=dev) / 5).half()
        w2 = (torch.randn(ch, 16 * ch, device=dev) / (16 * ch) ** 0.5).half()
        b1, b2 = torch.zeros(ch, device=dev), torch.zeros(ch, device=dev)
        w3 = torch.randn(ch, 3, device=dev)
        T = torch.empty(a.n, S, S, ch, device=dev, dtype=torch.float16)
        out = torch.empty(a.n, S // 2, S // 2, ch, device=dev, dtype=torch.float16)
        st = stream_ptr(dev)
        fn = lib.sdg_sngan64_block1_fused_h16 if ch == 64 else lib.sdg_sngan32_block1_fused_h16

        def fused():
            check(fn(ptr(img), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(out), None, a.n, _lib.PREC_FP16, st), "fused")

        def two():
            check(lib.sdg_first_conv_h16(ptr(img), _lib.LAYOUT_U8_NHWC, ptr(w1), ptr(b1), ptr(T), a.n, S, ch, _lib.PREC_FP16, st), "c1")
            check(lib.sdg_conv2d_h16(ptr(T), ptr(w2), ptr(b2), a.n, S, S, ch, ch, 3, None, 0, 2, None, 0, ptr(img),
                                     _lib.LAYOUT_U8_NHWC, ptr(w3), ptr(out), None, None, _lib.PREC_FP16, st), "c2")
        flops = 2.0 * a.n * S * S * ch * (16 * ch * 0.25 + 27)
        runs = (("fused", fused), ("fused", fused)) if a.fused_only else (("two launches", two), ("fused", fused), ("two launches", two), ("fused", fused))
        for name, run in runs:
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
            ev[0].record()
            for i in range(a.iters):
                run()
                ev[i + 1].record()
            torch.cuda.synchronize()
            t = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(a.iters))
            ms = sum(t) / len(t) / 1e3
            print(f"block1 n={a.n} {name}: mean {ms * 1e3:.1f} us (min {t[0]:.1f} median {t[len(t) // 2]:.1f} max {t[-1]:.1f})  "
                  f"{flops / ms / 1e9:.1f} TFLOP/s (useful, 4x4 stride-2 form)")
        return
    if a.first:
        img = torch.randint(0, 256, (a.n, a.hw, a.hw, 3), dtype=torch.uint8, device=dev)
        wb = (torch.randn(a.cout, 64, device=dev) / 5).half()
        bias = torch.zeros(a.cout, device=dev)
        out = torch.empty(a.n, a.hw, a.hw, a.cout, device=dev, dtype=torch.float16)
        run = lambda: check(lib.sdg_first_conv_h16(ptr(img), _lib.LAYOUT_U8_NHWC, ptr(wb), ptr(bias), ptr(out), a.n, a.hw,
                                                   a.cout, _lib.PREC_FP16, stream_ptr(dev)), "first")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        gb = out.numel() * 2 / 1e9
        print(f"first conv n={a.n} S={a.hw} ->{a.cout}: {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s written")
        return
    x = torch.randn(a.n, a.hw, a.hw, a.cin, device=dev).half()
    w = (torch.randn(a.cout, 9 * a.cin, device=dev) / (9 * a.cin) ** 0.5).half()
    b = torch.zeros(a.cout, device=dev)
    ho = a.hw // 2 if a.pool else a.hw
    out = torch.empty(a.n, ho, ho, a.cout, device=dev, dtype=torch.float16)
    lib.sdg_set_conv_pair(a.pair)

    def run():
        check(lib.sdg_conv2d_h16(ptr(x), ptr(w), ptr(b), a.n, a.hw, a.hw, a.cin, a.cout, 3, None, 0, a.pool, None, 0, None, 0,
                                 None, ptr(out), None, None, _lib.PREC_FP16, stream_ptr(dev)), "conv")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    flops = 2.0 * a.n * a.hw * a.hw * a.cout * 9 * a.cin
    env = {k: v for k, v in os.environ.items() if k.startswith("SDG_")}
    print(f"n={a.n} hw={a.hw} {a.cin}->{a.cout} pool={a.pool} pair={a.pair} {env}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
