"""Where does one bench step go?  CUDA-event timing of each stage of the SNGAN-32 recording step on one GPU, at the full
dataset and at the shard sizes of the strong-scaling run (50 000 / N samples per GPU).
    python tools/step_breakdown.py [--n 50000 25000 12500 6250] [--chunk 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import engine, synthetic  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--n", type=int, nargs="+", default=[50000, 25000, 12500, 6250])
    ap.add_argument("--n-total", type=int, default=50000, help="length of the gathered score vector the top-k runs over")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    base = {k: v.to(dev) for k, v in synthetic.sngan_state_dict(32, seed=1).items()}
    t_conf = engine.conf_from_key("ldr_conf_0.3_ratio_50")
    print(f"{'stage':30s}" + "".join(f"{('n=' + str(n)):>12s}" for n in a.n) + "   (ms per call, CUDA events, 20 reps back to back)")
    rows = {}
    for n in a.n:
        x = synthetic.uniform_images_u8(n, 32, seed=1).to(dev)
        eng = engine.DiscriminatorEngine(dev)
        if a.chunk:
            eng.set_chunk(a.chunk)
        eng.load_sngan(base, 32, "fp16", True)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        st = engine.RunningStats(n, dev)
        st.update(out.zero_())
        st.update(out)
        full = torch.rand(a.n_total, dtype=torch.float64, device=dev)
        res = {
            "load: sigma + pack": timed(lambda: eng.load_sngan(base, 32, "fp16", True)),
            "forward (shard)": timed(lambda: eng.forward(x, out=out)),
            "stats update (shard)": timed(lambda: st.update(out)),
            "score floor+min+clip (shard)": timed(lambda: st.score(t_conf, eps=1e-6)),
            f"top-100 of {a.n_total}": timed(lambda: engine.top_indices(full, 100, True)),
        }

        def step():
            eng.load_sngan(base, 32, "fp16", True)
            eng.forward(x, out=out)
            st.update(out)
            st.score(t_conf, eps=1e-6)
            engine.top_indices(full, 100, True)
        res["whole step, back to back"] = timed(step)
        res["sum of the stages"] = sum(v for k, v in res.items() if k != "whole step, back to back")
        for k, v in res.items():
            rows.setdefault(k, []).append(v)
        del eng, x
    for k, vs in rows.items():
        print(f"{k:30s}" + "".join(f"{v:12.3f}" for v in vs))


if __name__ == "__main__":
    main()
