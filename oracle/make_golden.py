"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle; test infrastructure only).

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference has no tests, fixtures or golden vectors for the diagnosis path, so
these files are the pins: inputs made from fixed seeds, outputs produced by the reference's own
``calculate_scores`` (plot.py:220-249), ``DRS`` (models/drs.py:10-69) and
``MNIST_DCGAN_Discriminator`` (mnist.py:155-223) imported from where they lie.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dcgan as dcgan_oracle   # noqa: E402
from oracle import stylegan2 as sg2_oracle # noqa: E402
from oracle import ref_loader              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def synth_logits(seed, steps, n, mode):
    """fp32-valued float64 snapshots, like trainer.py:144,154 stores them."""
    rng = np.random.RandomState(seed)
    base = rng.normal(1.0, 1.5, n)
    if mode == "narrow":          # random-init-D-like: tiny spread, some samples below the floor
        base = rng.normal(0.02, 0.03, n)
        noise = 0.005
    elif mode == "wide":          # both clip bounds active
        noise = 0.6
    else:                         # "ties": quantised so many equal scores
        base = np.round(base * 2) / 2
        noise = 0.0
    out = {}
    for s in steps:
        x = base + noise * rng.standard_normal(n)
        out[int(s)] = x.astype(np.float32).astype(np.float64)
    return out


def make_scores():
    calc = ref_loader.reference_calculate_scores()
    cases = [
        ("scores_cifar_window", 11, list(range(35000, 40001, 100)), 384, "wide", 35000, 40000),   # T=50 of 51
        ("scores_ffhq_window", 12, list(range(195000, 200001, 100)), 257, "narrow", 195000, 200001),  # T=51
        ("scores_ties", 13, list(range(0, 1000, 100)), 200, "ties", 200, 900),                    # T=7
    ]
    for name, seed, steps, n, mode, lo, hi in cases:
        logits = synth_logits(seed, steps, n, mode)
        ref = _quiet(calc, logits, start_epoch=lo, end_epoch=hi)
        keys = list(ref.keys())
        assert len(keys) == 103
        w = ref["ldr_conf_0.3_ratio_50"]
        wl = [1e-6 if i < 1e-6 else i for i in w]              # train_mimicry_phase2.py:23
        torch.manual_seed(7)
        sampler = torch.utils.data.WeightedRandomSampler(wl, len(wl), replacement=True)
        stream = np.array(list(iter(sampler)), dtype=np.int64)
        top = np.argsort(w, kind="stable")
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            steps=np.array(steps, dtype=np.int64),
            logits=np.stack([logits[s] for s in steps]),
            start=np.int64(lo), end=np.int64(hi),
            keys=np.array(keys),
            scores=np.stack([ref[k] for k in keys]),
            stream_seed=np.int64(7), stream=stream,
            argsort_stable=top,
        )
        print(name, "T_window", sum(lo <= s < hi for s in steps), "N", n)


class _IndexG:
    """Stand-in generator: image i of every batch is filled with a seeded normal draw."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)

    def generate_images(self, n, device=None):
        return torch.randn(n, 3, 4, 4, generator=self.gen)


class _LinD(torch.nn.Module):
    """Stand-in discriminator: a fixed affine map of the image mean, float32 [n,1]."""

    def forward(self, x):
        return (x.mean(dim=(1, 2, 3), keepdim=False).view(-1, 1) * 9.0 + 0.25).float()


def make_drs():
    DRS = ref_loader.reference_drs_class()
    for name, batch in (("drs_b256", 256), ("drs_b128", 128)):
        g, d = _IndexG(5), _LinD()
        burn = []
        orig = DRS.get_fake_samples_and_ldr

        def rec(self, num):           # record the ldr of every burn-in batch as the reference saw it
            imgs, ldr = orig(self, num)
            burn.append(ldr.copy())
            return imgs, ldr

        DRS.get_fake_samples_and_ldr = rec
        try:
            drs = DRS(g, d, "cpu")
        finally:
            DRS.get_fake_samples_and_ldr = orig
        max_after_burn = np.float32(drs.maximum)
        np.random.seed(1)
        ldrs, masks, maxes = [], [], []
        for _ in range(8):
            imgs, ldr = drs.get_fake_samples_and_ldr(batch)
            tagged = torch.arange(batch, dtype=torch.float32).view(-1, 1, 1, 1).expand(batch, 3, 2, 2).contiguous()
            acc = drs.sub_rejection_sampler(tagged, ldr)
            idx = acc[:, 0, 0, 0].numpy().astype(np.int64) if acc.numel() else np.zeros(0, np.int64)
            m = np.zeros(batch, dtype=bool)
            m[idx] = True
            ldrs.append(ldr.copy()); masks.append(m); maxes.append(np.float32(drs.maximum))
        np.random.seed(1)
        psi = np.random.rand(8 * batch).reshape(8, batch)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            burn_ldr=np.stack(burn).astype(np.float32),          # [50, 256, 1]
            max_after_burn=max_after_burn,
            ldr=np.stack(ldrs).astype(np.float32),               # [8, batch, 1]
            psi=psi, accept=np.stack(masks), max_after=np.array(maxes, dtype=np.float32),
            percentile=np.int64(80),
        )
        print(name, "accepted per batch", [int(m.sum()) for m in masks])


def make_dcgan():
    D = ref_loader.reference_dcgan_discriminator()
    net = _quiet(D)
    params = dcgan_oracle.init_params(seed=1)
    sd = net.state_dict()
    for k, v in params.items():
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        sd[k] = v.clone()
    net.load_state_dict(sd)
    net.eval()
    rng = np.random.RandomState(3)
    x_u8 = rng.randint(0, 256, (24, 32, 32, 3)).astype(np.uint8)
    x = ((torch.from_numpy(x_u8).permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5).contiguous()
    with torch.no_grad():
        y = net(x).view(-1).numpy()
    chk = np.float64(sum(float(v.double().sum()) for v in params.values()))
    np.savez_compressed(os.path.join(OUT, "dcgan_eval.npz"), param_seed=np.int64(1), param_checksum=chk,
                        x_u8=x_u8, logits=y.astype(np.float32))
    print("dcgan_eval logits[:4]", y[:4])


def make_stylegan2():
    """StyleGANDiscriminator (diagan/models/stylegan2.py:619-677) on the CPU fall-back ops: weights from
    oracle.stylegan2.init_params(seed) (regenerated by the tests), whole reference batches (minibatch-stddev)."""
    D = ref_loader.reference_stylegan2_discriminator()
    for size, n, batch in ((32, 16, 8), (128, 4, 4)):
        net = D(size=size)
        params = sg2_oracle.init_params(size, seed=1)
        sd = net.state_dict()
        for k, v in params.items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k] = v.clone()
        net.load_state_dict(sd)
        net.eval()
        rng = np.random.RandomState(5)
        x_u8 = rng.randint(0, 256, (n, size, size, 3)).astype(np.uint8)
        x = ((torch.from_numpy(x_u8).permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5).contiguous()
        with torch.no_grad():
            y = torch.cat([net(x[s:s + batch]) for s in range(0, n, batch)]).view(-1).numpy()
        chk = np.float64(sum(float(np.sum(v.numpy().astype(np.float64))) for v in params.values()))   # NumPy: thread-independent
        np.savez_compressed(os.path.join(OUT, f"stylegan2_d{size}.npz"), size=np.int64(size), batch=np.int64(batch),
                            param_seed=np.int64(1), param_checksum=chk, x_u8=x_u8, logits=y.astype(np.float32))
        print(f"stylegan2_d{size} logits[:4]", y[:4])


def make_stylegan2_256():
    """BASELINE configs[4] at its own size: StyleGANDiscriminator(256), two reference batches of 4 (train_ffhq.py:318).  The
    input bytes are regenerated by the tests from ``x_seed`` (RandomState(seed).randint, stable across NumPy versions), so the
    fixture holds only the eight logits."""
    D = ref_loader.reference_stylegan2_discriminator()
    size, n, batch, x_seed = 256, 8, 4, 7
    net = D(size=size)
    params = sg2_oracle.init_params(size, seed=1)
    sd = net.state_dict()
    for k, v in params.items():
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        sd[k] = v.clone()
    net.load_state_dict(sd)
    net.eval()
    x_u8 = np.random.RandomState(x_seed).randint(0, 256, (n, size, size, 3)).astype(np.uint8)
    x = ((torch.from_numpy(x_u8).permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5).contiguous()
    with torch.no_grad():
        y = torch.cat([net(x[s:s + batch]) for s in range(0, n, batch)]).view(-1).numpy()
    chk = np.float64(sum(float(np.sum(v.numpy().astype(np.float64))) for v in params.values()))
    np.savez_compressed(os.path.join(OUT, "stylegan2_d256.npz"), size=np.int64(size), batch=np.int64(batch),
                        param_seed=np.int64(1), param_checksum=chk, x_seed=np.int64(x_seed), n=np.int64(n),
                        x_checksum=np.int64(int(x_u8.astype(np.int64).sum())), logits=y.astype(np.float32))
    print("stylegan2_d256 logits", y)


def make_resize():
    """The reference's own transform (datasets/transform.py: Resize, CenterCrop, ToTensor, Normalize) applied to PIL images
    built like color_mnist.py:92 does; the uint8 image is recovered exactly from the normalised tensor."""
    from PIL import Image
    get_transform = ref_loader.reference_get_transform()
    rng = np.random.RandomState(11)
    out = {}
    for tag, name, n, h, w in (("mnist28", "color_mnist", 6, 28, 28), ("celeba", "celeba", 3, 218, 178)):
        tf = get_transform(name)
        imgs = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
        res = []
        for im in imgs:
            t = tf(Image.fromarray(im, mode="RGB"))                     # float [3,s,s] in [-1,1]
            u8 = torch.round((t * 0.5 + 0.5) * 255.0).to(torch.uint8).permute(1, 2, 0).numpy()
            res.append(u8)
        out[f"{tag}_in"] = imgs
        out[f"{tag}_out"] = np.stack(res)
        out[f"{tag}_size"] = np.int64(res[0].shape[0])
    np.savez_compressed(os.path.join(OUT, "resize_pil.npz"), **out)
    print("resize_pil", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)      # single-thread convs: deterministic accumulation order
    if len(sys.argv) > 1:         # python oracle/make_golden.py make_stylegan2_256  -> only that fixture
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    make_scores()
    make_drs()
    make_dcgan()
    make_stylegan2()
    make_stylegan2_256()
    make_resize()
