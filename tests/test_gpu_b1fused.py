"""SNGAN-32 block 1 as one launch (conv_b1fused.cu) against (a) the two-kernel path it replaces -- sdg_first_conv_h16 then
sdg_conv2d_h16 in the 4x4 stride-2 form with the 3-FMA image shortcut -- and (b) torch fp32 on the same 16-bit-rounded
operands.  The intermediate tensor relu(c1(x)) must be BIT-identical to the first conv's (same operands, same two K = 16
tcgen05.mma, same cvt.rn.relu); the block output may differ from the two-kernel path only by the fp32 summation order of
the 128 MMAs (taps are visited odd rows first), i.e. by at most one 16-bit ulp of the output scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sngan as sngan_oracle      # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _tdt(prec):
    return torch.float16 if prec == "fp16" else torch.bfloat16


def _operands(n, prec, seed):
    tdt = _tdt(prec)
    gen = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (n, 32, 32, 3), generator=gen, dtype=torch.uint8)
    w1 = (torch.randn(128, 3, 3, 3, generator=gen) / np.sqrt(27)).to(tdt)
    b1 = torch.randn(128, generator=gen) * 0.1
    w2 = (torch.randn(128, 128, 3, 3, generator=gen) / np.sqrt(128 * 9)).to(tdt)
    b2 = torch.randn(128, generator=gen) * 0.1
    w3 = torch.randn(128, 3, generator=gen) * 0.5
    w1p = torch.zeros(128, 64, dtype=tdt)
    w1p[:, :27] = w1.permute(0, 2, 3, 1).reshape(128, 27)
    w4 = torch.zeros(128, 4, 4, 128)
    wf = w2.float()
    for a in range(4):
        for b in range(4):
            for ky in (a - 1, a):
                for kx in (b - 1, b):
                    if 0 <= ky <= 2 and 0 <= kx <= 2:
                        w4[:, a, b, :] += wf[:, :, ky, kx]
    w2p = (0.25 * w4).reshape(128, 2048).to(tdt)
    return img, w1, b1, w2p, b2, w3, w1p


def _run(dev, n, prec, seed, want_dbg=True):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    tdt = _tdt(prec)
    pc = _lib.PREC_FP16 if prec == "fp16" else _lib.PREC_BF16
    img, w1, b1, w2p, b2, w3, w1p = _operands(n, prec, seed)
    d = lambda t: t.contiguous().to(dev)
    imgd, w1d, b1d, w2d, b2d, w3d = d(img), d(w1p), d(b1), d(w2p), d(b2), d(w3)
    nan = float("nan")
    t_ref = torch.full((n, 32, 32, 128), nan, dtype=tdt, device=dev)
    o_ref = torch.full((n, 16, 16, 128), nan, dtype=tdt, device=dev)
    check(lib.sdg_first_conv_h16(ptr(imgd), _lib.LAYOUT_U8_NHWC, ptr(w1d), ptr(b1d), ptr(t_ref), n, 32, 128, pc, stream_ptr(dev)),
          "sdg_first_conv_h16")
    check(lib.sdg_conv2d_h16(ptr(t_ref), ptr(w2d), ptr(b2d), n, 32, 32, 128, 128, 3, None, 0, 2, None, 0, ptr(imgd),
                             _lib.LAYOUT_U8_NHWC, ptr(w3d), ptr(o_ref), None, None, pc, stream_ptr(dev)), "sdg_conv2d_h16")
    t_dbg = torch.full((n, 32, 32, 128), nan, dtype=tdt, device=dev) if want_dbg else None
    o_fus = torch.full((n, 16, 16, 128), nan, dtype=tdt, device=dev)
    check(lib.sdg_sngan32_block1_fused_h16(ptr(imgd), ptr(w1d), ptr(b1d), ptr(w2d), ptr(b2d), ptr(w3d), ptr(o_fus), ptr(t_dbg), n,
                                           pc, stream_ptr(dev)), "sdg_sngan32_block1_fused_h16")
    torch.cuda.synchronize()
    # torch fp32 on the same rounded operands: T rounded to 16 bits, c2 as the 4x4 stride-2 conv with the re-rounded weights
    xn = sngan_oracle.normalise_u8(img)
    T = F.conv2d(xn.to(tdt).float(), w1.float(), b1, padding=1).relu().to(tdt).float()
    w4t = w2p.float().reshape(128, 4, 4, 128).permute(0, 3, 1, 2).contiguous()
    v = F.conv2d(T, w4t, b2, stride=2, padding=1) + F.conv2d(F.avg_pool2d(xn, 2), w3.view(128, 3, 1, 1))
    want = v.relu()
    return dict(t_ref=t_ref, t_dbg=t_dbg, o_ref=o_ref, o_fus=o_fus, want=want, T=T)


def _report(r, tag):
    """Where do the differences sit?  (printed on failure: which rows / columns / channels / images)"""
    a, b = r["o_fus"].float().cpu(), r["o_ref"].float().cpu()
    bad = ~torch.isclose(a, b, rtol=0, atol=4e-3 * float(b.abs().max())) | torch.isnan(a)
    msg = [f"{tag}: {int(bad.sum())} of {bad.numel()} outputs differ; nan {int(torch.isnan(a).sum())}"]
    if bad.any():
        msg.append("per image: " + str(bad.sum(dim=(1, 2, 3)).tolist()[:8]))
        msg.append("per row:   " + str(bad.sum(dim=(0, 2, 3)).tolist()))
        msg.append("per col:   " + str(bad.sum(dim=(0, 1, 3)).tolist()))
        msg.append("per ch/8:  " + str(bad.sum(dim=(0, 1, 2)).view(16, 8).sum(1).tolist()))
    if r["t_dbg"] is not None:
        ta, tb = r["t_dbg"].float().cpu(), r["t_ref"].float().cpu()
        tbad = (ta != tb) | torch.isnan(ta)
        msg.append(f"T: {int(tbad.sum())} of {tbad.numel()} differ; nan {int(torch.isnan(ta).sum())}")
        if tbad.any():
            msg.append("T per row:  " + str(tbad.sum(dim=(0, 2, 3)).tolist()))
            msg.append("T per col:  " + str(tbad.sum(dim=(0, 1, 3)).tolist()))
            msg.append("T per ch/8: " + str(tbad.sum(dim=(0, 1, 2)).view(16, 8).sum(1).tolist()))
    return "\n".join(msg)


@pytest.mark.parametrize("n", [1, 2, 5, 74, 149, 300])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_block1_fused_vs_two_kernels(n, prec, dev):
    r = _run(dev, n, prec, seed=100 + n)
    rep = _report(r, f"block1 fused {prec} n={n}")
    print(rep)
    ulp = 2.0 ** -10 if prec == "fp16" else 2.0 ** -7
    assert torch.equal(r["t_dbg"].view(torch.int16), r["t_ref"].view(torch.int16)), rep          # relu(c1(x)): bit-identical
    want = r["want"].permute(0, 2, 3, 1)
    scale = float(want.abs().max())
    e_fus = float((r["o_fus"].float().cpu() - want).abs().max())
    e_ref = float((r["o_ref"].float().cpu() - want).abs().max())
    e_pair = float((r["o_fus"].float() - r["o_ref"].float()).abs().max())
    print(f"  vs torch: fused {e_fus:.2e} two-kernel {e_ref:.2e} (scale {scale:.2f}); fused vs two-kernel {e_pair:.2e}")
    assert np.isfinite(e_fus) and e_fus <= 2 * ulp * scale, rep
    assert e_pair <= 2 * ulp * scale, rep


def test_block1_fused_without_debug_copy_and_empty(dev):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, stream_ptr
    r = _run(dev, 33, "fp16", seed=7, want_dbg=False)
    assert float((r["o_fus"].float() - r["o_ref"].float()).abs().max()) <= 2 * 2.0 ** -10 * float(r["want"].abs().max())
    lib = _lib.load()
    x = torch.zeros(16, dtype=torch.uint8, device=dev)
    p = lambda: _lib.ptr(x)
    check(lib.sdg_sngan32_block1_fused_h16(p(), p(), p(), p(), p(), p(), p(), None, 0, _lib.PREC_FP16, stream_ptr(dev)), "n = 0")


@pytest.mark.parametrize("seed", [1, 2])
def test_engine_fused_block1_equals_unfused_pass(seed, dev, monkeypatch):
    """The whole SNGAN-32 forward with and without the fused block 1: logits equal up to the 16-bit ulps that fp32 summation order flips."""
    import subprocess
    import sys
    import os
    code = (
        "import sys, torch, numpy as np\n"
        "sys.path[:0] = [%r, %r]\n"
        "from diagan_b200 import engine, synthetic\n"
        "dev = torch.device('cuda', 0)\n"
        "x = synthetic.uniform_images_u8(777, 32, seed=%d).to(dev)\n"
        "sd = synthetic.sngan_state_dict(32, seed=%d)\n"
        "eng = engine.DiscriminatorEngine(dev).load_sngan(sd, 32, 'fp16', True)\n"
        "np.save(sys.argv[1], eng.forward(x).cpu().numpy())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
         os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "self-diagnosing-gan_b200"), seed, seed)
    outs = []
    for flag in ("1", "0"):
        path = f"/tmp/b1fused_{seed}_{flag}.npy"
        env = dict(os.environ, SDG_FUSE_B1=flag)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        outs.append(np.load(path))
    d = np.abs(outs[0] - outs[1]).max()
    spread = outs[1].std()
    print(f"fused vs unfused logits: max |d| {d:.3e}, logit std {spread:.3e}, mean {outs[1].mean():.3f}")
    assert np.isfinite(outs[0]).all() and d <= 0.05 * spread      # the fp16 error itself is ~0.04 of the spread


# ---------------------------------------------------------------------------------------------------
# SNGAN-64: the same kernel per image quadrant (CH = 64)
# ---------------------------------------------------------------------------------------------------
def _operands64(n, prec, seed):
    tdt = _tdt(prec)
    gen = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (n, 64, 64, 3), generator=gen, dtype=torch.uint8)
    w1 = (torch.randn(64, 3, 3, 3, generator=gen) / np.sqrt(27)).to(tdt)
    b1 = torch.randn(64, generator=gen) * 0.1
    w2 = (torch.randn(64, 64, 3, 3, generator=gen) / np.sqrt(64 * 9)).to(tdt)
    b2 = torch.randn(64, generator=gen) * 0.1
    w3 = torch.randn(64, 3, generator=gen) * 0.5
    w1p = torch.zeros(64, 64, dtype=tdt)
    w1p[:, :27] = w1.permute(0, 2, 3, 1).reshape(64, 27)
    w4 = torch.zeros(64, 4, 4, 64)
    wf = w2.float()
    for a in range(4):
        for b in range(4):
            for ky in (a - 1, a):
                for kx in (b - 1, b):
                    if 0 <= ky <= 2 and 0 <= kx <= 2:
                        w4[:, a, b, :] += wf[:, :, ky, kx]
    w2p = (0.25 * w4).reshape(64, 1024).to(tdt)
    return img, w1, b1, w2p, b2, w3, w1p


def _run64(dev, n, prec, seed, want_dbg=True):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    tdt = _tdt(prec)
    pc = _lib.PREC_FP16 if prec == "fp16" else _lib.PREC_BF16
    img, w1, b1, w2p, b2, w3, w1p = _operands64(n, prec, seed)
    d = lambda t: t.contiguous().to(dev)
    imgd, w1d, b1d, w2d, b2d, w3d = d(img), d(w1p), d(b1), d(w2p), d(b2), d(w3)
    nan = float("nan")
    t_ref = torch.full((n, 64, 64, 64), nan, dtype=tdt, device=dev)
    check(lib.sdg_first_conv_h16(ptr(imgd), _lib.LAYOUT_U8_NHWC, ptr(w1d), ptr(b1d), ptr(t_ref), n, 64, 64, pc, stream_ptr(dev)),
          "sdg_first_conv_h16")
    t_dbg = torch.full((n, 64, 64, 64), nan, dtype=tdt, device=dev) if want_dbg else None
    o_fus = torch.full((n, 32, 32, 64), nan, dtype=tdt, device=dev)
    check(lib.sdg_sngan64_block1_fused_h16(ptr(imgd), ptr(w1d), ptr(b1d), ptr(w2d), ptr(b2d), ptr(w3d), ptr(o_fus), ptr(t_dbg), n,
                                           pc, stream_ptr(dev)), "sdg_sngan64_block1_fused_h16")
    torch.cuda.synchronize()
    xn = sngan_oracle.normalise_u8(img)
    T = F.conv2d(xn.to(tdt).float(), w1.float(), b1, padding=1).relu().to(tdt).float()
    w4t = w2p.float().reshape(64, 4, 4, 64).permute(0, 3, 1, 2).contiguous()
    v = F.conv2d(T, w4t, b2, stride=2, padding=1) + F.conv2d(F.avg_pool2d(xn, 2), w3.view(64, 3, 1, 1))
    return dict(t_ref=t_ref, t_dbg=t_dbg, o_fus=o_fus, want=v.relu(), T=T)


def _report64(r, tag):
    want = r["want"].permute(0, 2, 3, 1)
    a = r["o_fus"].float().cpu()
    bad = ~torch.isclose(a, want, rtol=0, atol=4e-3 * float(want.abs().max())) | torch.isnan(a)
    msg = [f"{tag}: {int(bad.sum())} of {bad.numel()} outputs differ from torch; nan {int(torch.isnan(a).sum())}"]
    if bad.any():
        msg.append("per image: " + str(bad.sum(dim=(1, 2, 3)).tolist()[:8]))
        msg.append("per row:   " + str(bad.sum(dim=(0, 2, 3)).tolist()))
        msg.append("per col:   " + str(bad.sum(dim=(0, 1, 3)).tolist()))
        msg.append("per ch/8:  " + str(bad.sum(dim=(0, 1, 2)).view(8, 8).sum(1).tolist()))
    if r["t_dbg"] is not None:
        ta, tb = r["t_dbg"].float().cpu(), r["t_ref"].float().cpu()
        tbad = (ta != tb) | torch.isnan(ta)
        msg.append(f"T: {int(tbad.sum())} of {tbad.numel()} differ; nan {int(torch.isnan(ta).sum())}")
        if tbad.any():
            msg.append("T per row:  " + str(tbad.sum(dim=(0, 2, 3)).tolist()))
            msg.append("T per col:  " + str(tbad.sum(dim=(0, 1, 3)).tolist()))
            msg.append("T per ch/8: " + str(tbad.sum(dim=(0, 1, 2)).view(8, 8).sum(1).tolist()))
    return "\n".join(msg)


@pytest.mark.parametrize("n", [1, 2, 5, 19, 75, 150])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_block1_fused64_vs_first_conv_and_torch(n, prec, dev):
    """relu(c1(x)) bit-identical to sdg_first_conv_h16's on every pixel of every quadrant (incl. the recomputed halos, which
    are written by two tiles with the same bits); block output within two 16-bit ulps of torch fp32 on the same operands."""
    r = _run64(dev, n, prec, seed=200 + n)
    rep = _report64(r, f"block1 fused64 {prec} n={n}")
    print(rep)
    ulp = 2.0 ** -10 if prec == "fp16" else 2.0 ** -7
    assert torch.equal(r["t_dbg"].view(torch.int16), r["t_ref"].view(torch.int16)), rep
    want = r["want"].permute(0, 2, 3, 1)
    scale = float(want.abs().max())
    e_fus = float((r["o_fus"].float().cpu() - want).abs().max())
    print(f"  vs torch: fused {e_fus:.2e} (scale {scale:.2f})")
    assert np.isfinite(e_fus) and e_fus <= 2 * ulp * scale, rep


def test_block1_fused64_without_debug_copy_and_empty(dev):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, stream_ptr
    r = _run64(dev, 9, "fp16", seed=7, want_dbg=False)
    want = r["want"].permute(0, 2, 3, 1)
    assert float((r["o_fus"].float().cpu() - want).abs().max()) <= 2 * 2.0 ** -10 * float(want.abs().max())
    lib = _lib.load()
    x = torch.zeros(16, dtype=torch.uint8, device=dev)
    p = lambda: _lib.ptr(x)
    check(lib.sdg_sngan64_block1_fused_h16(p(), p(), p(), p(), p(), p(), p(), None, 0, _lib.PREC_FP16, stream_ptr(dev)), "n = 0")


def test_engine_fused_block1_sngan64_equals_unfused_pass(dev):
    """The whole SNGAN-64 forward with the fused block 1 and with the two-kernel super-pixel path (SDG_FUSE_B1=32 keeps the
    fusion for SNGAN-32 only): logits equal up to the 16-bit ulps that the fp32 summation order flips."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch, numpy as np\n"
        "sys.path[:0] = [%r, %r]\n"
        "from diagan_b200 import engine, synthetic\n"
        "dev = torch.device('cuda', 0)\n"
        "x = synthetic.uniform_images_u8(333, 64, seed=3).to(dev)\n"
        "sd = synthetic.sngan_state_dict(64, seed=3)\n"
        "eng = engine.DiscriminatorEngine(dev).load_sngan(sd, 64, 'fp16', True)\n"
        "np.save(sys.argv[1], eng.forward(x).cpu().numpy())\n"
        "print('launches', engine.launch_count())\n"
    ) % (root, os.path.join(root, "self-diagnosing-gan_b200"))
    outs = []
    for flag in ("1", "32"):
        path = f"/tmp/b1fused64_{flag}.npy"
        env = dict(os.environ, SDG_FUSE_B1=flag)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        outs.append(np.load(path))
    d = np.abs(outs[0] - outs[1]).max()
    spread = outs[1].std()
    print(f"fused vs unfused SNGAN-64 logits: max |d| {d:.3e}, logit std {spread:.3e}, mean {outs[1].mean():.4f}")
    assert np.isfinite(outs[0]).all() and d <= 0.15 * spread      # the fp16 error itself is ~0.1 of the spread
