"""Where does one bench step go?  CUDA-event timing of each stage of the SNGAN-32 recording step on one GPU.
    python tools/step_breakdown.py [--chunk 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-diagnosing-gan_b200"))
import torch  # noqa: E402

from diagan_b200 import engine, synthetic  # noqa: E402


def timed(fn, reps=10):
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
    ap.add_argument("--n", type=int, default=50000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    x = synthetic.uniform_images_u8(a.n, 32, seed=1).to(dev)
    base = {k: v.to(dev) for k, v in synthetic.sngan_state_dict(32, seed=1).items()}
    eng = engine.DiscriminatorEngine(dev)
    if a.chunk:
        eng.set_chunk(a.chunk)
    eng.load_sngan(base, 32, "fp16", True)
    out = torch.empty(a.n, dtype=torch.float32, device=dev)
    st = engine.RunningStats(a.n, dev)
    st.update(out.zero_())
    st.update(out)
    t_conf = engine.conf_from_key("ldr_conf_0.3_ratio_50")
    res = {
        "perturb weights (torch)": timed(lambda: synthetic.perturb_(base, 35000, 1e-3, device=dev)),
        "load: sigma + pack": timed(lambda: eng.load_sngan(base, 32, "fp16", True)),
        "forward 50k": timed(lambda: eng.forward(x, out=out)),
        "stats update": timed(lambda: st.update(out)),
        "score + floor": timed(lambda: st.score(t_conf, eps=1e-6)),
    }
    w = st.score(t_conf, eps=1e-6)
    res["top-100"] = timed(lambda: engine.top_indices(w, 100, True))
    for k, v in res.items():
        print(f"{k:28s} {v:8.3f} ms")
    print(f"{'sum':28s} {sum(res.values()):8.3f} ms   (chunk={a.chunk or 'default'})")


if __name__ == "__main__":
    main()
