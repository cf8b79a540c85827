"""HBM roofline of the statistics / score / selection kernels (north_star: "achieved HBM GB/s for the statistics and score
kernels"; SURVEY 8(d): "at N = 50 k these are launch-latency-bound; report GB/s also at N = 16 M synthetic").
    python tools/bench_stats.py [--n 16777216] [--T 50]
Algorithmic bytes per sample: stats_update 68 B; window_moments 4*T (8*T for float64 snapshots) + 32 B;
score_floor_min 24 B; score_clip 16 B; top-k radix select 8 B per sweep, the sweeps the launch actually ran (digit passes
until the early exit + one compaction sweep; read back from the kernel's histogram block)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import engine  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6544.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    n, T = a.n, a.T
    gen = torch.Generator(device="cuda").manual_seed(0)
    snap = (1.0 + 1.5 * torch.randn(n, generator=gen, device=dev)).float()
    st = engine.RunningStats(n, dev)
    st.update(snap)
    st.update(snap)
    rows = []
    t = timed(lambda: st.update(snap), a.reps)
    rows.append(("stats_update (Welford + last + sum|d|)", 68.0 * n, t))
    Tn = min(T, max(2, (8 << 30) // (4 * n)))          # bound the window buffer at 8 GiB
    snaps = (1.0 + 1.5 * torch.randn(Tn, n, generator=gen, device=dev)).float()
    t = timed(lambda: engine.window_moments(snaps), max(2, a.reps // 4))
    rows.append((f"window_moments f32 (T={Tn}, 4 outputs)", (4.0 * Tn + 32.0) * n, t))
    mom = engine.window_moments(snaps)
    del snaps
    Td = min(T, max(2, (8 << 30) // (8 * n)))
    snaps = (1.0 + 1.5 * torch.randn(Td, n, generator=gen, device=dev, dtype=torch.float64))
    t = timed(lambda: engine.window_moments(snaps), max(2, a.reps // 4))
    rows.append((f"window_moments f64 (T={Td}, 4 outputs; the pickle's dtype)", (8.0 * Td + 32.0) * n, t))
    del snaps
    t03 = engine.conf_from_key("ldr_conf_0.3_ratio_50")
    t = timed(lambda: engine.scores_from_moments(mom["mean"], mom["var"], [t03], eps=1e-6), a.reps)
    rows.append(("score: floor+min then clip+eps (1 key)", 40.0 * n, t))
    s = engine.scores_from_moments(mom["mean"], mom["var"], [t03], eps=1e-6)[0]
    t = timed(lambda: engine.top_indices(s, 100, True), a.reps)
    # FusedState (csrc/select.cu): unsigned hist[9][2048] first; a digit pass ran iff its histogram row is non-zero
    hist = engine._topk_ws[s.device][:9 * 2048 * 4].view(torch.int32).view(9, 2048)
    sweeps = int((hist != 0).any(dim=1).sum().item()) + 1
    rows.append((f"top-100 radix select ({sweeps - 1} digit passes + compaction, one launch)", 8.0 * sweeps * n, t))
    print(f"N = {n:,} samples, HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json)")
    print("| kernel | algorithmic bytes | time | GB/s | frac of HBM peak |")
    print("|---|---:|---:|---:|---:|")
    for name, b, t in rows:
        print(f"| {name} | {b / 1e6:.1f} MB | {t * 1e6:.1f} us | {b / t / 1e9:.0f} | {b / t / 1e9 / peak:.2f} |")


if __name__ == "__main__":
    main()
