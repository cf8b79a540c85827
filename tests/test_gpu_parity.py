"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the golden vectors the
reference itself produced.  Run on the B200 box: ``python -m pytest tests -m gpu``.

Tolerances (BASELINE.json north_star / BASELINE.md section 4):
  * score stage fed identical logits: BIT-EXACT float64 against the reference (snapshot-window path);
    <= 1e-12 relative for the streaming Welford accumulator; resampled index stream and selected index
    sets bit-exact;
  * discriminator logits, error measured as |a-b| / max(|b|, mean|b|) (relative to the logit scale of
    the batch) against the oracle evaluated in float64:
      - fp16 tensor-core engine (the throughput mode): measure B = |a-b| / (L1 mass of the head's dot
        product: sum_j |w_j h_j| + |b|, the magnitude of the terms a logit sums) <= 2e-4, asserted
        unconditionally; the measure above ("A") <= 1e-3 is asserted where the logits are O(1) and
        reported where a random-init network's logits cancel to ~0 (measured up to 1.4e-2 there, the same as
        PyTorch eager with cuDNN TF32 -- the reference's own default GPU arithmetic, SURVEY 0.1 item 11 --
        which is printed beside it as information, not as a bar);
      - fp32 engine: <= 1e-5, or no further from exact than twice the reference's own fp32 CPU path is
        (that path itself sits ~8e-6 absolute from the float64 result, so 1e-5 of a near-zero logit is
        below the noise floor of the arithmetic being compared against);
      - bf16 tensor-core engine: reported, bounded at 1.5e-2 -- it does NOT meet the 1e-3 bar (measured
        4e-3..7e-3), which is why fp16 operands are the default;
  * DRS: acceptance decisions identical under the same psi except where |p - psi| < 1e-5 (fp32
    exp/log differ from NumPy's in the last ulp), running maximum identical.
"""
import math
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import dcgan as dcgan_oracle      # noqa: E402
from oracle import drs as drs_oracle          # noqa: E402
from oracle import scores as so               # noqa: E402
from oracle import sngan as sngan_oracle      # noqa: E402

SCORE_CASES = ["scores_cifar_window", "scores_ffhq_window", "scores_ties"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _logit_close(a, b, tol=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.maximum(np.abs(b), np.abs(b).mean())
    err = np.abs(a - b) / scale
    return err.max(), err.mean()


def _fp32_ok(gpu, cpu32, exact):
    """fp32 criterion of the module docstring -> (ok, gpu_err, cpu_err)."""
    e_gpu = _logit_close(gpu, exact)[0]
    e_cpu = _logit_close(cpu32, exact)[0]
    return e_gpu <= max(1e-5, 2.0 * e_cpu), e_gpu, e_cpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


# ---------------------------------------------------------------------------------------------------
# scoring stage
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCORE_CASES)
def test_calculate_scores_bit_exact_vs_reference(golden_dir, name):
    from diagan_b200.utils.plot import calculate_scores
    g = _load(golden_dir, name)
    logits = {int(s): g["logits"][i] for i, s in enumerate(g["steps"])}
    out = calculate_scores(logits, start_epoch=int(g["start"]), end_epoch=int(g["end"]))
    keys = [str(k) for k in g["keys"]]
    assert list(out.keys()) == keys
    for i, k in enumerate(keys):
        assert out[k].dtype == np.float64
        assert np.array_equal(out[k], g["scores"][i]), f"{name}:{k} differs from the reference"


@pytest.mark.parametrize("name", SCORE_CASES)
def test_fp32_snapshots_give_same_scores(golden_dir, name, dev):
    """The recorder keeps fp32 snapshots on the device; widening inside the kernel == the float64 pickle."""
    from diagan_b200.utils.plot import calculate_scores_device
    g = _load(golden_dir, name)
    logits = {int(s): torch.from_numpy(g["logits"][i].astype(np.float32)).to(dev) for i, s in enumerate(g["steps"])}
    out = calculate_scores_device(logits, int(g["start"]), int(g["end"]), keys=["ldr_conf_0.3_ratio_50"])
    keys = [str(k) for k in g["keys"]]
    for k in ("ldrm", "ldrv", "ldrd", "ldr", "ldr_conf_0.3_ratio_50"):
        assert np.array_equal(out[k].cpu().numpy(), g["scores"][keys.index(k)]), k


@pytest.mark.parametrize("name", SCORE_CASES)
def test_running_stats_and_resample_stream(golden_dir, name, dev):
    from diagan_b200 import engine
    g = _load(golden_dir, name)
    steps = g["steps"]
    sel = (steps >= g["start"]) & (steps < g["end"])
    arr = g["logits"][sel]
    T, n = arr.shape
    st = engine.RunningStats(n, dev)
    for t in range(T):
        st.update(torch.from_numpy(arr[t].astype(np.float32)).to(dev))
    keys = [str(k) for k in g["keys"]]
    ref = lambda k: g["scores"][keys.index(k)]
    np.testing.assert_allclose(st.mean.cpu().numpy(), ref("ldrm"), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(st.ldrv().cpu().numpy(), ref("ldrv"), rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(st.ldrd().cpu().numpy(), ref("ldrd"), rtol=1e-12, atol=1e-15)
    assert np.array_equal(st.ldr().cpu().numpy(), ref("ldr"))
    # same update as the oracle's streaming restatement, bit for bit
    om, oq, ol, osad = so.welford(arr)
    assert np.array_equal(st.mean.cpu().numpy(), om) and np.array_equal(st.m2.cpu().numpy(), oq)
    # weights -> WeightedRandomSampler stream identical to the reference's (train_mimicry_phase2.py:21-24)
    w = st.score(engine.conf_from_key("ldr_conf_0.3_ratio_50"), eps=1e-6).cpu().numpy()
    np.testing.assert_allclose(w, ref("ldr_conf_0.3_ratio_50"), rtol=1e-11)
    assert np.array_equal(so.resample_stream(w, int(g["stream_seed"])), g["stream"])


def test_odd_sizes_and_unaligned_shards(dev):
    """ragged / tiny inputs: N = 1, 2, 3, odd N, shard views starting at odd offsets."""
    from diagan_b200 import engine
    rng = np.random.RandomState(0)
    for n in (1, 2, 3, 255, 1001):
        arr = rng.normal(0.3, 1.0, (6, n)).astype(np.float32)
        st = engine.RunningStats(n, dev)
        for t in range(6):
            st.update(torch.from_numpy(arr[t]).to(dev))
        om, oq, ol, osad = so.welford(arr.astype(np.float64))
        assert np.array_equal(st.mean.cpu().numpy(), om)
        assert np.array_equal(st.state[3].cpu().numpy(), osad)
        mom = engine.window_moments(torch.from_numpy(arr).to(dev))
        mean, var = so.moments(arr.astype(np.float64))
        assert np.array_equal(mom["mean"].cpu().numpy(), mean) and np.array_equal(mom["var"].cpu().numpy(), var)
    # unaligned: statistics over a slice [3:] of a bigger buffer
    full = torch.from_numpy(rng.normal(size=(4, 40)).astype(np.float32)).to(dev)
    st = engine.RunningStats(37, dev)
    for t in range(4):
        st.update(full[t, 3:].contiguous()[0:37])
    om, _, _, _ = so.welford(full[:, 3:].cpu().numpy().astype(np.float64))
    assert np.array_equal(st.mean.cpu().numpy(), om)
    mom = engine.window_moments(full[:, 3:])       # strided rows (ld = 40)
    assert np.array_equal(mom["mean"].cpu().numpy(), so.moments(full[:, 3:].cpu().numpy().astype(np.float64))[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("T", [2, 33, 50, 51, 64, 160, 161, 200])
def test_window_moments_every_dispatch_branch(dev, T, dtype):
    """the shared-memory-staged kernel (aligned 16-byte copies, element copies for ragged N) and the generic two-pass kernel
    (windows too long for the tile) give NumPy's bits:
    mean / var as plot.py:243-246 (np.mean / np.var(ddof=1) along axis 0), ldr = last row, ldrd = mean |row diff|."""
    from diagan_b200 import engine
    rng = np.random.RandomState(T)
    for n in (1, 129, 5000, 40000):
        arr = rng.normal(1.0, 1.5, (T, n)).astype(dtype)
        mom = engine.window_moments(torch.from_numpy(arr).to(dev))
        a64 = arr.astype(np.float64)
        mean, var = so.moments(a64)
        assert np.array_equal(mom["mean"].cpu().numpy(), mean)
        assert np.array_equal(mom["var"].cpu().numpy(), var)
        assert np.array_equal(mom["ldr"].cpu().numpy(), a64[-1])
        sad = np.zeros(n)
        for t in range(1, T):                       # row order, like the kernels and oracle.scores.welford
            sad = sad + np.abs(a64[t] - a64[t - 1])
        assert np.array_equal(mom["ldrd"].cpu().numpy(), sad / (T - 1))


@pytest.mark.parametrize("n_conf", [1, 3])
def test_score_vectorised_and_scalar_paths(dev, n_conf):
    """odd / even N, aligned and unaligned bases: the 16-byte path and the scalar path of the score kernels agree with NumPy"""
    from diagan_b200 import engine
    rng = np.random.RandomState(3)
    confs = [0.3, 1.0, 5.0][:n_conf]
    for n in (1, 2, 7, 1024, 4097, 100_001):
        big_m = torch.from_numpy(rng.normal(1.0, 1.5, n + 1)).to(dev)
        big_v = torch.from_numpy(rng.gamma(2.0, 1.0, n + 1)).to(dev)
        for off in (0, 1):                          # off = 1: 8-byte aligned only
            m, v = big_m[off:off + n], big_v[off:off + n]
            got = engine.scores_from_moments(m, v, confs, eps=1e-6).cpu().numpy()
            mh, vh = m.cpu().numpy(), v.cpu().numpy()
            for j, c in enumerate(confs):
                sc = np.clip(mh + c * np.sqrt(vh), 1e-2, None)
                want = np.maximum(np.minimum(sc, sc.min() * 50), 1e-6)
                assert np.array_equal(got[j], want), (n, off, c)


def test_top_indices_sizes_around_the_sweep_unroll(dev):
    from diagan_b200 import engine
    rng = np.random.RandomState(5)
    for n in (1, 31, 33, 2047, 2049, 70_001, 700_001):
        w = rng.normal(0.0, 2.0, n)
        w[rng.randint(0, n, max(1, n // 50))] = w[0]          # some exact ties
        wd = torch.from_numpy(w).to(dev)
        order = np.argsort(w, kind="stable")
        for k in sorted({1, min(n, 100), min(n, 4096)}):
            assert np.array_equal(engine.top_indices(wd, k, True).cpu().numpy(), order[-k:])
            assert np.array_equal(engine.top_indices(wd, k, False).cpu().numpy(), order[:k])


@pytest.mark.parametrize("name", SCORE_CASES)
def test_top_indices_match_stable_argsort(golden_dir, name, dev):
    from diagan_b200 import engine
    g = _load(golden_dir, name)
    keys = [str(k) for k in g["keys"]]
    w = g["scores"][keys.index("ldr_conf_0.3_ratio_50")]
    wd = torch.from_numpy(w).to(dev)
    order = g["argsort_stable"]
    for k in (1, 20, 100, len(w)):
        assert np.array_equal(engine.top_indices(wd, k, True).cpu().numpy(), order[-k:])
        assert np.array_equal(engine.top_indices(wd, k, False).cpu().numpy(), order[:k])


def test_top_indices_heavy_ties_large(dev):
    """After clipping most samples tie at a bound ('Ties'); tie-break = ascending index."""
    from diagan_b200 import engine
    rng = np.random.RandomState(1)
    n = 200_000
    w = np.clip(rng.normal(1.0, 1.0, n), 0.01, 0.5)
    w[rng.randint(0, n, 50)] *= -1.0          # a few negatives: sign handling of the key transform
    wd = torch.from_numpy(w).to(dev)
    order = np.argsort(w, kind="stable")
    for k in (100, 4096):
        assert np.array_equal(engine.top_indices(wd, k, True).cpu().numpy(), order[-k:])
        assert np.array_equal(engine.top_indices(wd, k, False).cpu().numpy(), order[:k])


# ---------------------------------------------------------------------------------------------------
# DRS
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["drs_b256", "drs_b128"])
def test_drs_accept_vs_reference(golden_dir, name, dev):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    g = _load(golden_dir, name)
    mx = torch.full((1,), -100000.0, dtype=torch.float32, device=dev)
    for b in range(g["burn_ldr"].shape[0]):
        l = torch.from_numpy(g["burn_ldr"][b].reshape(-1)).to(dev)
        check(lib.sdg_drs_update_max(ptr(l), l.numel(), ptr(mx), stream_ptr(dev)))
    assert np.float32(mx.item()) == g["max_after_burn"]
    oracle = drs_oracle.DRSOracle(int(g["percentile"]))
    oracle.maximum = g["max_after_burn"]
    for b in range(g["ldr"].shape[0]):
        l = torch.from_numpy(g["ldr"][b].reshape(-1)).to(dev)
        n = l.numel()
        psi = torch.from_numpy(g["psi"][b]).to(dev)
        p = torch.empty(n, dtype=torch.float32, device=dev)
        acc = torch.empty(n, dtype=torch.uint8, device=dev)
        idx = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.sdg_drs_accept(ptr(l), n, ptr(mx), 1e-6, float(g["percentile"]), 0, 0.0, ptr(psi), ptr(p), ptr(acc),
                                 ptr(idx), ptr(cnt), stream_ptr(dev)))
        assert np.float32(mx.item()) == g["max_after"][b]
        p_ref, acc_ref = oracle.accept(g["ldr"][b], g["psi"][b])
        assert np.array_equal(acc_ref, g["accept"][b])
        p_gpu, acc_gpu = p.cpu().numpy(), acc.cpu().numpy().astype(bool)
        np.testing.assert_allclose(p_gpu, p_ref, rtol=0, atol=2e-6)
        diff = acc_gpu != g["accept"][b]
        assert np.all(np.abs(p_ref[diff] - g["psi"][b][diff]) < 1e-5), "acceptance differs away from the boundary"
        k = int(cnt.item())
        assert k == int(acc_gpu.sum())
        assert np.array_equal(idx[:k].cpu().numpy(), np.nonzero(acc_gpu)[0])


def test_drs_module_contract(dev):
    """diagan_b200.models.drs.DRS with stand-in G/D: reference method contracts (drs.py:21-69)."""
    from diagan_b200.models.drs import DRS
    from diagan_b200.trainer.evaluate import DRS as EvalDRS

    class G:
        def __init__(self):
            self.gen = torch.Generator(device="cuda").manual_seed(5)

        def generate_images(self, n, device=None):
            return torch.randn(n, 3, 4, 4, generator=self.gen, device=device)

    class D(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=(1, 2, 3)).view(-1, 1) * 9.0 + 0.25

    drs = DRS(G(), D(), dev)
    assert isinstance(drs.maximum, np.float32) and drs.maximum > -100000
    imgs, ldr = drs.get_fake_samples_and_ldr(256)
    assert imgs.is_cuda and ldr.dtype == np.float32 and ldr.shape == (256, 1)
    np.random.seed(1)
    acc = drs.sub_rejection_sampler(imgs, ldr)
    assert acc.device.type == "cpu" and acc.dtype == torch.float32 and acc.shape[1:] == (3, 4, 4)
    # same decisions as the oracle under the same psi
    o = drs_oracle.DRSOracle(80)
    o.maximum = drs.maximum
    np.random.seed(1)
    p_ref, acc_ref = o.accept(ldr, np.random.rand(256))
    assert abs(int(acc_ref.sum()) - acc.shape[0]) <= 1
    out = drs.generate_images(300)
    assert out.shape == (300, 3, 4, 4)
    e = EvalDRS(G(), D(), dev, batch_size=128)
    assert e.generate_images(50).is_cuda and e.batch_size == 128


def test_drs_generate_images_batched_equals_sequential(dev):
    """generate_images scores up to 32 candidate batches per host round trip; the accepted images, their
    order, the running maximum and BOTH random streams (NumPy psi, generator) must be those of the one-batch-at-a-time loop
    of drs.py:59-69."""
    from diagan_b200.models.drs import DRS

    class G:
        def __init__(self):
            self.gen = torch.Generator(device="cuda").manual_seed(5)
            self.calls = 0

        def generate_images(self, n, device=None):
            self.calls += 1
            return torch.randn(n, 3, 8, 8, generator=self.gen, device=device)

    class D(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=(1, 2, 3)).view(-1, 1) * 9.0 + 0.25

    outs = []
    for in_flight in (1, 32):
        g = G()
        np.random.seed(7)
        drs = DRS(g, D(), dev, batch_size=64)
        drs.max_batches_in_flight = in_flight
        imgs = drs.generate_images(1000, device=dev)
        outs.append((imgs, drs.maximum, np.random.rand(), g.calls, torch.randn(4, generator=g.gen, device=dev)))
    a, b = outs
    assert a[0].shape == (1000, 3, 8, 8) and torch.equal(a[0], b[0])
    assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3] and torch.equal(a[4], b[4])


def test_drs_with_engine_discriminator_sngan64(dev):
    """BASELINE config 4: DRS acceptance pass with the SNGAN-64 discriminator running in the CUDA engine
    (EngineNetD keeps the netD(x) -> [B,1] contract of drs.py:24-28); logits checked against the oracle."""
    from diagan_b200.models.drs import DRS
    from diagan_b200.models.engine_netd import EngineNetD
    params = sngan_oracle.init_params(64, seed=1)

    class G:
        def __init__(self):
            self.gen = torch.Generator(device="cuda").manual_seed(11)

        def generate_images(self, n, device=None):
            return torch.randn(n, 3, 64, 64, generator=self.gen, device=device).tanh()

    netD = EngineNetD(params, dev, precision="fp32")
    drs = DRS(G(), netD, dev, batch_size=64)
    imgs, ldr = drs.get_fake_samples_and_ldr(64)
    with torch.no_grad():
        want = sngan_oracle.forward(params, imgs.cpu(), 64).numpy()
        exact = sngan_oracle.forward({k: v.double() for k, v in params.items()}, imgs.cpu().double(), 64).numpy()
    assert ldr.shape == (64, 1) and ldr.dtype == np.float32
    assert _fp32_ok(ldr.reshape(-1), want.reshape(-1), exact.reshape(-1))[0]
    np.random.seed(3)
    out = drs.generate_images(40)
    assert out.shape == (40, 3, 64, 64)
    # tensor-core engine through the same wrapper
    tc = EngineNetD(params, dev)          # fp16 operands by default
    y = tc(imgs).cpu().numpy().reshape(-1)
    assert _logit_close(y, exact.reshape(-1))[0] <= 2e-3


# ---------------------------------------------------------------------------------------------------
# discriminator forward
# ---------------------------------------------------------------------------------------------------
def _u8(n, size, seed):
    return torch.from_numpy(np.random.RandomState(seed).randint(0, 256, (n, size, size, 3)).astype(np.uint8))


def test_dcgan_fp32_vs_reference_golden(golden_dir, dev):
    from diagan_b200 import engine
    g = _load(golden_dir, "dcgan_eval")
    params = dcgan_oracle.init_params(int(g["param_seed"]))
    eng = engine.DiscriminatorEngine(dev).load(params, "fp32")
    assert eng.arch == "dcgan32"
    x = torch.from_numpy(g["x_u8"]).to(dev)
    y = eng.forward(x).cpu().numpy()
    emax, _ = _logit_close(y, g["logits"], 1e-5)
    print(f"dcgan fp32 vs reference golden: max rel err {emax:.2e}")
    assert emax <= 1e-5, emax
    # float32 NCHW entry (what netD(x) receives in trainer.py:150)
    xf = sngan_oracle.normalise_u8(torch.from_numpy(g["x_u8"])).contiguous().to(dev)
    y2 = eng.forward(xf).cpu().numpy()
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("arch,n", [(32, 70), (64, 20)])
@pytest.mark.parametrize("inplace", [True, False])
def test_sngan_fp32_vs_oracle(arch, n, inplace, dev):
    from diagan_b200 import engine
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    params = sngan_oracle.init_params(arch, seed=1)
    x = _u8(n, arch, 2)
    want = sngan_oracle.logits_pass(params, x, arch, inplace_relu=inplace)
    exact = sngan_oracle.logits_pass(params, x, arch, inplace_relu=inplace, dtype=torch.float64)
    eng = engine.DiscriminatorEngine(dev).load_sngan(params, arch, "fp32", inplace)
    got = eng.forward(x.to(dev)).cpu().numpy()
    ok, e_gpu, e_cpu = _fp32_ok(got, want, exact)
    print(f"sngan{arch} fp32 inplace={inplace}: GPU vs f64 {e_gpu:.2e}, CPU fp32 vs f64 {e_cpu:.2e}, "
          f"GPU vs CPU fp32 {_logit_close(got, want)[0]:.2e}")
    assert ok
    # sigma of every layer against the oracle's power iteration
    sig = eng.sigmas().cpu().numpy()
    ref = np.array([float(sngan_oracle.sigma_eval(params[f"{k}.weight"], params[f"{k}.sn_u"]))
                    for k in engine.sngan_layer_keys(arch)])
    np.testing.assert_allclose(sig, ref, rtol=2e-6)
    # chunking must not change results
    eng.set_chunk(7)
    assert np.array_equal(eng.forward(x.to(dev)).cpu().numpy(), got)


@pytest.mark.parametrize("size", [32, 128])
def test_stylegan2_fp32_vs_reference_golden(golden_dir, size, dev):
    """StyleGANDiscriminator (BASELINE config 5 architecture) in the fp32 engine against logits produced by the
    reference module itself; whole reference batches (minibatch-stddev), both input layouts, chunking invariance."""
    from diagan_b200 import engine
    from oracle import stylegan2 as sg2_oracle
    g = _load(golden_dir, f"stylegan2_d{size}")
    batch = int(g["batch"])
    params = sg2_oracle.init_params(size, int(g["param_seed"]))
    eng = engine.DiscriminatorEngine(dev).load_stylegan2(params, "fp32", batch=batch)
    assert eng.arch == "stylegan2" and eng.size == size
    x = torch.from_numpy(g["x_u8"]).to(dev)
    y = eng.forward(x).cpu().numpy()
    emax, _ = _logit_close(y, g["logits"])
    print(f"stylegan2 D{size} fp32 vs reference golden: max rel err {emax:.2e}")
    assert emax <= 1e-5
    xf = sngan_oracle.normalise_u8(torch.from_numpy(g["x_u8"])).contiguous().to(dev)
    assert np.array_equal(eng.forward(xf).cpu().numpy(), y)
    if x.shape[0] > batch:            # one reference batch per internal sweep gives the same logits
        eng.set_chunk(batch)
        assert np.array_equal(eng.forward(x).cpu().numpy(), y)
        # a different batch composition changes the stddev channel, hence the logits
        eng.set_chunk(0); eng.set_batch(batch // 2)
        assert not np.array_equal(eng.forward(x).cpu().numpy(), y)
    with pytest.raises(Exception):
        eng.set_batch(batch)
        eng.forward(x[:batch - 1].contiguous())


def _tdt(prec):
    return torch.float16 if prec == "fp16" else torch.bfloat16


def _conv_fused(dev, n, hw, cin, cout, ks, prec, seed, sc_c=0, pool=0, res=False, res_relu=0, img=False, pair=1):
    """Run one fused conv stage through the C ABI and return {name: (max abs err, scale)} per output against
    torch fp32 on the same 16-bit-rounded operands."""
    import torch.nn.functional as F
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    tdt = _tdt(prec)
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, hw, hw, generator=gen).to(tdt)
    w = (torch.randn(cout, cin, ks, ks, generator=gen) / np.sqrt(cin * ks * ks)).to(tdt)
    b = torch.randn(cout, generator=gen)
    v = F.conv2d(x.float(), w.float(), b, padding=ks // 2)
    if pool == 2:       # 4x4 stride-2 form: 0.25 * sums of the (already 16-bit) 3x3 weights, rounded to 16 bits again
        w4 = torch.zeros(cout, 4, 4, cin)
        wf = w.float()
        for a in range(4):
            for bb in range(4):
                for ky in (a - 1, a):
                    for kx in (bb - 1, bb):
                        if 0 <= ky <= 2 and 0 <= kx <= 2:
                            w4[:, a, bb, :] += wf[:, :, ky, kx]
        packs = [(0.25 * w4).reshape(cout, 16 * cin).to(tdt)]
    else:
        packs = [w.permute(0, 2, 3, 1).reshape(cout, ks * ks * cin)]
    sc_x = None
    if sc_c:
        sc_x = torch.randn(n, sc_c, hw, hw, generator=gen).to(tdt)
        w_sc = (torch.randn(cout, sc_c, 1, 1, generator=gen) / np.sqrt(sc_c)).to(tdt)
        v = v + F.conv2d(sc_x.float(), w_sc.float())
        if pool == 2:
            packs.append((0.25 * w_sc.float()).reshape(cout, 1, sc_c).expand(cout, 4, sc_c).reshape(cout, 4 * sc_c).to(tdt))
        else:
            packs.append(w_sc.reshape(cout, sc_c))
    if pool:
        v = F.avg_pool2d(v, 2)
    img_t = w3 = None
    if img:
        img_t = torch.randint(0, 256, (n, hw, hw, 3), generator=gen, dtype=torch.uint8)
        w3 = torch.randn(cout, 3, generator=gen) * 0.5
        v = v + F.conv2d(F.avg_pool2d(sngan_oracle.normalise_u8(img_t), 2), w3.view(cout, 3, 1, 1))
    ho = hw // 2 if pool else hw
    res_t = None
    if res:
        res_t = torch.randn(n, cout, ho, ho, generator=gen)
        v = v + (res_t.relu() if res_relu else res_t)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(dev)
    wb = torch.cat(packs, dim=1).contiguous().to(dev)
    o_relu = torch.full((n, ho, ho, cout), float("nan"), dtype=tdt, device=dev)
    o_raw = torch.full((n, ho, ho, cout), float("nan"), dtype=tdt, device=dev)
    o_f32 = torch.full((n, ho, ho, cout), float("nan"), dtype=torch.float32, device=dev)
    xd, bd = nhwc(x), b.to(dev)
    scd = nhwc(sc_x) if sc_c else None
    resd = nhwc(res_t) if res else None
    imgd = img_t.to(dev) if img else None
    w3d = w3.contiguous().to(dev) if img else None
    lib.sdg_set_conv_pair(pair)
    try:
        check(lib.sdg_conv2d_h16(ptr(xd), ptr(wb), ptr(bd), n, hw, hw, cin, cout, ks, ptr(scd), sc_c, pool, ptr(resd),
                                 res_relu, ptr(imgd), _lib.LAYOUT_U8_NHWC, ptr(w3d), ptr(o_relu), ptr(o_raw), ptr(o_f32),
                                 _lib.PREC_FP16 if prec == "fp16" else _lib.PREC_BF16, stream_ptr(dev)), "sdg_conv2d_h16")
        torch.cuda.synchronize()
    finally:
        lib.sdg_set_conv_pair(1)
    back = lambda t: t.float().cpu().permute(0, 3, 1, 2)
    scale = v.abs().max().item()
    return {"f32": ((back(o_f32) - v).abs().max().item(), scale),
            "raw": ((back(o_raw) - v).abs().max().item(), scale),
            "relu": ((back(o_relu) - v.relu()).abs().max().item(), scale)}


def _check_conv(errs, prec, tag, slack=1.0):
    ulp = slack * (2.0 ** -10 if prec == "fp16" else 2.0 ** -7)
    print(tag, {k: f"{e:.2e}" for k, (e, _) in errs.items()}, f"scale {errs['f32'][1]:.2f}")
    for k, (e, scale) in errs.items():
        assert np.isfinite(e), (tag, k)
        assert e <= scale * ((2e-5 if slack == 1.0 else ulp) if k == "f32" else ulp), (tag, k, e, scale)


@pytest.mark.parametrize("n,hw,cin,cout,ks", [
    (3, 32, 128, 128, 3),      # SNGAN-32 block1.c2 shape (55% of the FLOPs)
    (5, 16, 128, 128, 3),      # block2 convs
    (7, 8, 128, 128, 3),       # blocks 3/4: two images per 128-pixel tile, ragged last tile
    (3, 16, 128, 128, 1),      # a plain 1x1 conv
    (2, 64, 64, 64, 3),        # SNGAN-64 block1.c2 (N tile 64, two rows per tile)
    (9, 4, 512, 1024, 3),      # SNGAN-64 block5.c2: 8 images per tile, 8 N tiles, 8 K chunks
    (300, 32, 128, 128, 3),    # more tiles than SMs: persistent loop + TMEM double buffering
    (37, 8, 128, 128, 3),      # odd number of M tiles (19): the pair kernel's last M=256 tile is half empty
])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("pair", [0, 1, 2])
def test_conv2d_h16_tcgen05_vs_torch(n, hw, cin, cout, ks, prec, pair, dev):
    """The tcgen05 implicit-GEMM kernels alone (no fusion) against F.conv2d on the same 16-bit-rounded operands
    (fp32 accumulate both sides; 16-bit outputs within one output ulp of the scale, fp32 output within 2e-5).
    pair=1: default kernel selection (role-swapped kernel for Cout = 128, streamed CTA pairs for Cout % 256 == 0);
    2: CTA-pair kernels without the role swap (resident weights for Cout = 128); 0: single-CTA kernel."""
    errs = _conv_fused(dev, n, hw, cin, cout, ks, prec, seed=n * 1000 + hw, pair=pair)
    _check_conv(errs, prec, f"conv {prec} pair={pair} n={n} hw={hw} {cin}->{cout} k{ks}")


@pytest.mark.parametrize("tag,n,hw,cin,cout,kw", [
    ("sngan32 b1.c2: pool + 3-FMA image shortcut (quarter tiles, W=32)", 3, 32, 128, 128, dict(pool=1, img=True)),
    ("sngan32 b2.c2: pool + folded 1x1 shortcut", 5, 16, 128, 128, dict(pool=1, sc_c=128)),
    ("sngan32 b3.c2: identity residual, rectified", 7, 8, 128, 128, dict(res=True, res_relu=1)),
    ("sngan32 b4.c2: identity residual, raw", 6, 8, 128, 128, dict(res=True, res_relu=0)),
    ("sngan64 b1.c2: pool + image shortcut (quarter tiles, W=64, N=64)", 2, 64, 64, 64, dict(pool=1, img=True)),
    ("sngan64 b2.c2: 64->128 pool + folded shortcut 64", 3, 32, 64, 128, dict(pool=1, sc_c=64)),
    ("sngan64 b3.c2: 128->256 pool + folded shortcut 128", 3, 16, 128, 256, dict(pool=1, sc_c=128)),
    ("sngan64 b4.c2: 256->512 @8 pool + folded shortcut 256", 5, 8, 256, 512, dict(pool=1, sc_c=256)),
    ("sngan64 b5.c2: 512->1024 @4 pool + folded shortcut 512", 9, 4, 512, 1024, dict(pool=1, sc_c=512)),
    ("many tiles, pooled", 200, 32, 128, 128, dict(pool=1, img=True)),
])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("pair", [0, 1, 2])
def test_conv2d_fused_block_stage_vs_torch(tag, n, hw, cin, cout, kw, prec, pair, dev):
    """Fused epilogues: pooling by warp shuffles, shortcut conv as extra K columns, image shortcut FMAs, identity
    residual from the fp32 stream, the three output forms -- on both kernel variants."""
    errs = _conv_fused(dev, n, hw, cin, cout, 3, prec, seed=hw * 7 + n, pair=pair, **kw)
    _check_conv(errs, prec, f"{prec} pair={pair} {tag}")
    if kw.get("pool") and pair >= 1:
        # the same stage as a 4x4 stride-2 conv (TMA traversal stride 2); weights are re-rounded after summing,
        # so allow two output ulps
        kw4 = dict(kw, pool=2)
        errs = _conv_fused(dev, n, hw, cin, cout, 3, prec, seed=hw * 7 + n, pair=pair, **kw4)
        _check_conv(errs, prec, f"{prec} pool4 {tag}", slack=2.0)


@pytest.mark.parametrize("S,cout,n", [(32, 128, 5), (64, 64, 3), (32, 128, 333)])
@pytest.mark.parametrize("layout", ["u8", "f32"])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_first_conv_from_bytes_vs_torch(S, cout, n, layout, prec, dev):
    """relu(conv3x3(normalise(x)) + b) straight from uint8 NHWC / fp32 NCHW input (builder warps + tcgen05)."""
    import torch.nn.functional as F
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    tdt = _tdt(prec)
    gen = torch.Generator().manual_seed(S + n)
    img = torch.randint(0, 256, (n, S, S, 3), generator=gen, dtype=torch.uint8)
    xn = sngan_oracle.normalise_u8(img).contiguous()
    w = (torch.randn(cout, 3, 3, 3, generator=gen) / np.sqrt(27)).to(tdt)
    b = torch.randn(cout, generator=gen) * 0.1
    want = F.conv2d(xn.to(tdt).float(), w.float(), b, padding=1).relu()
    wb = torch.zeros(cout, 64, dtype=tdt)
    wb[:, :27] = w.permute(0, 2, 3, 1).reshape(cout, 27)
    out = torch.full((n, S, S, cout), float("nan"), dtype=tdt, device=dev)
    xd = img.to(dev) if layout == "u8" else xn.to(dev)
    wbd, bd = wb.to(dev), b.to(dev)          # keep the device copies alive until the kernel has run
    check(lib.sdg_first_conv_h16(ptr(xd), _lib.LAYOUT_U8_NHWC if layout == "u8" else _lib.LAYOUT_F32_NCHW, ptr(wbd),
                                 ptr(bd), ptr(out), n, S, cout, _lib.PREC_FP16 if prec == "fp16" else _lib.PREC_BF16,
                                 stream_ptr(dev)), "sdg_first_conv_h16")
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    err, scale = (got - want).abs().max().item(), want.abs().max().item()
    print(f"first conv {prec} {layout} S={S} n={n}: max abs err {err:.2e} scale {scale:.2f}")
    assert np.isfinite(err) and err <= scale * (2.0 ** -10 if prec == "fp16" else 2.0 ** -7)


@pytest.mark.parametrize("arch,n,seed", [(32, 300, 1), (32, 128, 2), (64, 40, 1)])
@pytest.mark.parametrize("inplace", [True, False])
@pytest.mark.parametrize("prec,tol", [("fp16", 1e-3), ("bf16", 1.5e-2)])
def test_sngan_tensorcore_vs_oracle(arch, n, seed, inplace, prec, tol, dev):
    from diagan_b200 import engine
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    params = sngan_oracle.init_params(arch, seed=seed)
    x = _u8(n, arch, 3)
    want, l1 = sngan_oracle.logits_pass(params, x, arch, inplace_relu=inplace, dtype=torch.float64, with_head_l1=True)
    eng = engine.DiscriminatorEngine(dev).load_sngan(params, arch, prec, inplace)
    got = eng.forward(x.to(dev)).cpu().numpy()
    emax, emean = _logit_close(got, want)
    eb = float((np.abs(got - want) / l1).max())
    # yardstick: the reference network in PyTorch eager on this GPU with its default TF32 convolutions
    torch.backends.cudnn.allow_tf32 = True
    pd = {k: v.to(dev) for k, v in params.items()}
    with torch.no_grad():
        ytf = sngan_oracle.forward(pd, sngan_oracle.normalise_u8(x).to(dev), arch, inplace).view(-1).cpu().numpy()
    etf = _logit_close(ytf, want)[0]
    print(f"sngan{arch} {prec} seed={seed} inplace={inplace}: A (rel. logit scale) max {emax:.2e} mean {emean:.2e}, B (rel. head "
          f"L1 mass) {eb:.2e}, max abs {np.abs(got - want).max():.2e} | torch-eager TF32 on this GPU: A {etf:.2e} "
          f"(logit mean {want.mean():.4f} std {want.std():.4f})")
    # B is the bar (no TF32 escape); A is asserted when the logits are O(1).  bf16 is reported, not a parity mode.
    assert eb <= (2e-4 if prec == "fp16" else 3e-3)
    if abs(want.mean()) >= 1.0:
        assert emax <= tol
    eng.set_chunk(64)
    assert np.array_equal(eng.forward(x.to(dev)).cpu().numpy(), got)
    # float32 NCHW input path gives the same logits as the uint8 path: bit for bit where both run the same kernels; with
    # mimicry's in-place ReLU the byte dataset runs block 1 as ONE kernel (conv_b1fused.cu) whose fp32 summation order over the
    # taps differs from the two-kernel path the fp32 input takes, which flips a few 16-bit ulps of relu(h1): equal to a small
    # fraction of the fp16 error itself
    xf = sngan_oracle.normalise_u8(x).contiguous().to(dev)
    gf = eng.forward(xf).cpu().numpy()
    if inplace and os.environ.get("SDG_FUSE_B1", "1") in ("1", str(arch)):
        assert np.abs(gf - got).max() <= (0.05 if prec == "fp16" else 0.4) * max(float(np.std(want)), 1e-3)   # bf16: 8x the ulp
    else:
        assert np.array_equal(gf, got)


# ---------------------------------------------------------------------------------------------------
# StyleGAN2 discriminator on the tensor-core path (BASELINE config 5)
# ---------------------------------------------------------------------------------------------------
def _sg2_conv(dev, n, hin, cin, cout, ks, stride, pad, act, res, prec, seed, skip=False):
    """One ConvLayer stage through sdg_conv2d_sg2_h16 vs torch fp32 on the same 16-bit-rounded operands."""
    import ctypes as C
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    dt = _tdt(prec)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, hin, hin, generator=g).to(dt)
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).to(dt)
    b = 0.1 * torch.randn(cout, generator=g)
    y = F.conv2d(x.float(), w.float(), stride=stride, padding=(ks // 2 if pad else 0))
    ho = y.shape[-1]
    r = torch.randn(n, cout, ho, ho, generator=g) if res else None
    y = y + b.view(1, -1, 1, 1)
    if act:
        y = F.leaky_relu(y, 0.2) * math.sqrt(2.0)
    scale = 1.0 / math.sqrt(2.0) if (res or skip) else 1.0
    sk = wk = None
    if skip:                                  # the ResBlock skip: a 1x1 conv of an output-resolution tensor, second accumulator
        sk = torch.randn(n, cin, ho, ho, generator=g).to(dt)
        wk = (torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)).to(dt)
        y = y + F.conv2d(sk.float(), wk.float())
    if res:
        y = y + r
    y = y * scale
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    wd = w.permute(0, 2, 3, 1).reshape(cout, -1)
    if skip:
        wd = torch.cat([wd, wk.view(cout, cin)], 1)
    wd = wd.contiguous().to(dev)
    skd = sk.permute(0, 2, 3, 1).contiguous().to(dev) if skip else None
    bd = b.to(dev)
    rd = r.permute(0, 2, 3, 1).contiguous().to(dev) if res else None
    o16 = torch.empty(n, ho, ho, cout, dtype=dt, device=dev)
    o32 = torch.empty(n, ho, ho, cout, dtype=torch.float32, device=dev)
    pc = {"fp16": _lib.PREC_FP16, "bf16": _lib.PREC_BF16}[prec]
    check(lib.sdg_conv2d_sg2_h16(ptr(xd), ptr(wd), ptr(bd), n, ho, ho, hin, hin, cin, cout, ks, stride, 1 if pad else 0,
                                 1 if act else 0, ptr(skd), cin if skip else 0, ptr(rd), C.c_float(scale), ptr(o16), ptr(o32), pc,
                                 stream_ptr(dev)),
          "sdg_conv2d_sg2_h16")
    torch.cuda.synchronize()
    want = y.permute(0, 2, 3, 1)
    e32 = (o32.cpu() - want).abs().max().item()
    e16 = (o16.float().cpu() - want).abs().max().item()
    return e32, e16, want.abs().max().item()


@pytest.mark.parametrize("tag,n,hin,cin,cout,ks,stride,pad,act,res", [
    ("conv1_256_wide_pair", 1, 256, 128, 128, 3, 1, 1, 1, False),     # 256-pixel rows: two tiles per row, CTA-pair kernel
    ("conv1_64", 2, 64, 256, 256, 3, 1, 1, 1, False),
    ("conv1_4", 9, 4, 512, 512, 3, 1, 1, 1, False),
    ("conv2_s2_from_257", 1, 257, 128, 256, 3, 2, 0, 1, True),        # blurred (hw+1)^2 input, stride 2, no padding, + skip
    ("conv2_s2_from_33", 3, 33, 512, 512, 3, 2, 0, 1, True),
    ("conv2_s2_from_9", 5, 9, 512, 512, 3, 2, 0, 1, True),
    ("skip_1x1", 3, 16, 256, 512, 1, 1, 0, 0, False),
    ("conv2_s2_skipfold_257", 1, 257, 128, 256, 3, 2, 0, 1, "skip"),  # skip conv folded in: second TMEM accumulator
    ("conv2_s2_skipfold_17", 6, 17, 512, 512, 3, 2, 0, 1, "skip"),
    ("conv2_s2_skipfold_9", 7, 9, 512, 512, 3, 2, 0, 1, "skip"),
])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_stylegan2_conv_stage_vs_torch(tag, n, hin, cin, cout, ks, stride, pad, act, res, prec, dev):
    e32, e16, mag = _sg2_conv(dev, n, hin, cin, cout, ks, stride, pad, act, res is True, prec, seed=len(tag),
                              skip=res == "skip")
    print(f"{tag} {prec}: fp32-out err {e32:.2e}, 16-bit-out err {e16:.2e}, |y|max {mag:.2f}")
    # operands are identical 16-bit values on both sides: only the accumulation order (fp32) and the output rounding differ
    assert e32 <= 2e-4 * max(1.0, mag)
    assert e16 <= mag * (2.0 ** -10 if prec == "fp16" else 2.0 ** -7) + 2e-4


@pytest.mark.parametrize("H,C,pad,stride,n", [(16, 64, 2, 1, 3), (16, 64, 1, 2, 3), (8, 512, 2, 1, 2), (256, 128, 1, 2, 1),
                                              (4, 512, 2, 1, 5)])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_blur_h16_vs_upfirdn2d_oracle(H, C, pad, stride, n, prec, dev):
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    from oracle import stylegan2 as sg2_oracle
    lib = _lib.load()
    dt = _tdt(prec)
    x = torch.randn(n, C, H, H, generator=torch.Generator().manual_seed(H + C)).to(dt)
    want = sg2_oracle.blur(x.float(), pad, pad)[:, :, ::stride, ::stride].permute(0, 2, 3, 1)
    ho = want.shape[1]
    out = torch.empty(n, ho, ho, C, dtype=dt, device=dev)
    pc = {"fp16": _lib.PREC_FP16, "bf16": _lib.PREC_BF16}[prec]
    check(lib.sdg_blur_h16(ptr(x.permute(0, 2, 3, 1).contiguous().to(dev)), ptr(out), n, H, H, C, pad, stride, pc,
                           stream_ptr(dev)), "sdg_blur_h16")
    err = (out.float().cpu() - want).abs().max().item()
    assert err <= want.abs().max().item() * (2.0 ** -10 if prec == "fp16" else 2.0 ** -7)


@pytest.mark.parametrize("H,W,C,n", [(32, 32, 64, 3), (64, 64, 128, 2), (100, 36, 64, 2), (256, 256, 128, 1), (33, 70, 256, 2)])
@pytest.mark.parametrize("pad,stride", [(2, 1), (1, 2), (1, 1), (0, 1), (3, 2)])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_blur_tma_bit_equal_to_register_kernel(H, W, C, n, pad, stride, prec, dev):
    """every variant of the TMA-fed blur (csrc/blur_tma.cu) == the register-sliding kernel it replaces, bit for bit, incl.
    ragged strips / segments, non-square images and every padding; the latter is pinned to upfirdn2d above"""
    import os
    from diagan_b200 import _lib
    from diagan_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    dt = _tdt(prec)
    pc = {"fp16": _lib.PREC_FP16, "bf16": _lib.PREC_BF16}[prec]
    x = torch.randn(n, H, W, C, generator=torch.Generator().manual_seed(H * W + C)).to(dt).to(dev)
    ho, wo = (H + 2 * pad - 4) // stride + 1, (W + 2 * pad - 4) // stride + 1
    outs = []
    try:
        for v in (0, 1, 2, 3):
            os.environ["SDG_BLUR_TMA"] = str(v)
            out = torch.full((n, ho, wo, C), float("nan"), dtype=dt, device=dev)
            check(lib.sdg_blur_h16(ptr(x), ptr(out), n, H, W, C, pad, stride, pc, stream_ptr(dev)), "sdg_blur_h16")
            outs.append(out.view(torch.int16).cpu())
    finally:
        os.environ.pop("SDG_BLUR_TMA", None)
    assert not torch.isnan(outs[0].view(dt).float()).any()
    for v in (1, 2, 3):
        assert torch.equal(outs[0], outs[v]), f"variant {v} differs from the register kernel"


@pytest.mark.parametrize("size", [32, 128])
@pytest.mark.parametrize("prec,tol", [("fp16", 1e-3), ("bf16", 1.5e-2)])
def test_stylegan2_tensorcore_vs_reference_golden(golden_dir, size, prec, tol, dev):
    """The tcgen05 StyleGAN2 discriminator against logits of the reference module itself (golden fixture) and the
    float64 oracle; yardstick = the fp32 oracle network in torch eager with TF32 convolutions on this GPU."""
    from diagan_b200 import engine
    from oracle import stylegan2 as sg2_oracle
    g = _load(golden_dir, f"stylegan2_d{size}")
    batch = int(g["batch"])
    params = sg2_oracle.init_params(size, int(g["param_seed"]))
    x = torch.from_numpy(g["x_u8"])
    eng = engine.DiscriminatorEngine(dev).load_stylegan2(params, prec, batch=batch)
    got = eng.forward(x.to(dev)).cpu().numpy()
    want, l1 = sg2_oracle.logits_pass(params, x, size, batch, dtype=torch.float64, with_head_l1=True)
    emax, emean = _logit_close(got, want)
    eb = float((np.abs(got - want) / l1).max())
    egold = _logit_close(got, g["logits"])[0]
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    pd = {k: v.to(dev) for k, v in params.items()}
    ytf = np.zeros(x.shape[0])
    with torch.no_grad():
        for s0 in range(0, x.shape[0], batch):
            ytf[s0:s0 + batch] = sg2_oracle.forward(pd, sngan_oracle.normalise_u8(x[s0:s0 + batch]).to(dev), size).view(-1).cpu().numpy()
    etf = _logit_close(ytf, want)[0]
    print(f"stylegan2 D{size} {prec}: vs float64 oracle A {emax:.2e} (mean {emean:.2e}) B {eb:.2e}; vs reference golden A {egold:.2e} | "
          f"torch-eager TF32 on this GPU: A {etf:.2e} (logit mean {want.mean():.4f} std {want.std():.4f})")
    assert eb <= (2e-4 if prec == "fp16" else 3e-3)
    assert emax <= (1.5e-3 if prec == "fp16" else 1.5e-2)        # O(0.25) logits: measured 4.0e-4 (D32) / 1.03e-3 (D128) in fp16
    xf = sngan_oracle.normalise_u8(x).contiguous().to(dev)
    assert np.array_equal(eng.forward(xf).cpu().numpy(), got)
    if x.shape[0] > batch:
        eng.set_chunk(batch)
        assert np.array_equal(eng.forward(x.to(dev)).cpu().numpy(), got)


# ---------------------------------------------------------------------------------------------------
# input pipeline: Resize + CenterCrop on the GPU, bit-exact with the reference's PIL transform
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,size,n", [(28, 28, 32, 37), (218, 178, 64, 9), (45, 37, 32, 5), (37, 45, 32, 5), (100, 300, 64, 3),
                                         (32, 32, 32, 4), (64, 48, 64, 3), (20, 20, 64, 2), (500, 333, 64, 2)])
def test_resize_center_crop_bit_exact_vs_pil_and_oracle(h, w, size, n, dev):
    from diagan_b200.datasets.transform import DeviceTransform
    from oracle import resize as resize_oracle
    rng = np.random.RandomState(h * 1000 + w)
    imgs = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    got = DeviceTransform(size)(torch.from_numpy(imgs).to(dev)).cpu().numpy()
    assert got.shape == (n, size, size, 3)
    assert np.array_equal(got, resize_oracle.resize_center_crop_u8(imgs, size))
    try:                                                      # the real thing, when the box has it (same image: it does)
        from PIL import Image
        import torchvision.transforms as T
        tf = T.Compose([T.Resize(size), T.CenterCrop(size)])
        want = np.stack([np.asarray(tf(Image.fromarray(im, mode="RGB"))) for im in imgs])
        assert np.array_equal(got, want)
    except ImportError:
        pass


def test_resize_edge_cases(dev):
    """Empty input, single-channel images (mnist_fmnist transform, transform.py:25-33), and a non-CUDA tensor is refused."""
    from diagan_b200 import _lib
    from diagan_b200.datasets.transform import DeviceTransform, get_transform
    from oracle import resize as resize_oracle
    tf = get_transform("mnist_fmnist")
    assert tf.img_size == 32
    assert tf(torch.empty(0, 28, 28, 1, dtype=torch.uint8, device=dev)).shape == (0, 32, 32, 1)
    g = np.random.RandomState(3).randint(0, 256, (11, 28, 28, 1)).astype(np.uint8)
    got = tf(torch.from_numpy(g).to(dev)).cpu().numpy()
    assert np.array_equal(got, resize_oracle.resize_center_crop_u8(g, 32))
    try:
        from PIL import Image
        want = np.stack([np.asarray(Image.fromarray(im[..., 0], mode="L").resize((32, 32), Image.BILINEAR))[..., None] for im in g])
        assert np.array_equal(got, want)
    except ImportError:
        pass
    with pytest.raises(_lib.SdgError):
        DeviceTransform(32)(torch.zeros(2, 28, 28, 3, dtype=torch.uint8))
    with pytest.raises(_lib.SdgError):
        DeviceTransform(32)(torch.zeros(2, 28, 28, 3, dtype=torch.float32, device=dev))


def test_resident_dataset_from_raw_colour_mnist_shape(dev):
    """Colour-MNIST items are 28x28 uint8 RGB, resized to 32 (color_mnist.py:90-100 + transform.py:23-31): the resident
    dataset built on the GPU equals the per-item PIL path, normalisation included (first-conv LUT == ToTensor + Normalize)."""
    from diagan_b200.trainer.trainer import ResidentDataset
    from oracle import resize as resize_oracle
    rng = np.random.RandomState(5)
    raw = (rng.rand(64, 28, 28, 1) < 0.19).astype(np.uint8) * np.array([255, 0, 0], np.uint8)      # red digits-like masks
    ds = ResidentDataset.from_raw_images(raw, "color_mnist", dev)
    assert ds.data.shape == (64, 32, 32, 3) and ds.data.dtype == torch.uint8
    assert np.array_equal(ds.data.cpu().numpy(), resize_oracle.resize_center_crop_u8(raw, 32))


@pytest.mark.parametrize("size,batch,n", [(8, 4, 8), (16, 2, 6), (64, 4, 4)])
def test_stylegan2_tensorcore_small_sizes_vs_oracle(size, batch, n, dev):
    """Sizes without a golden fixture (one to four ResBlocks; a single reference batch; batch 2 = stddev group of 2) against
    the float64 oracle, which the 32 / 128 fixtures pin to the reference module."""
    from diagan_b200 import engine, synthetic
    from oracle import stylegan2 as sg2_oracle
    params = synthetic.stylegan2_state_dict(size, seed=size)
    x = _u8(n, size, 5)
    want = sg2_oracle.logits_pass(params, x, size, batch, dtype=torch.float64)
    got = engine.DiscriminatorEngine(dev).load_stylegan2(params, "fp16", batch=batch).forward(x.to(dev)).cpu().numpy()
    exact = engine.DiscriminatorEngine(dev).load_stylegan2(params, "fp32", batch=batch).forward(x.to(dev)).cpu().numpy()
    e16, e32 = _logit_close(got, want)[0], _logit_close(exact, want)[0]
    print(f"stylegan2 D{size} batch {batch}: fp16 {e16:.2e}, fp32 engine {e32:.2e} (logit mean {want.mean():.3f} std {want.std():.3f})")
    assert e32 <= 1e-5 and e16 <= 2e-3


@pytest.mark.parametrize("arch", [32, 64])
def test_sngan_tensorcore_tiny_batches(arch, dev):
    """1, 2, 3 and 5 samples: fewer pixels than one 256-pixel tile in the late blocks, so the kernel selection falls back
    from the role-swapped / CTA-pair kernels to the single-CTA one (fused head by atomics); logits must match the oracle and
    agree with the same samples' logits inside a large pass up to fp32 summation order."""
    from diagan_b200 import engine, synthetic
    sd = synthetic.sngan_state_dict(arch, seed=3)
    x = _u8(261, arch, 9).to(dev)
    eng = engine.DiscriminatorEngine(dev).load_sngan(sd, arch, "fp16", True)
    full = eng.forward(x)
    want = sngan_oracle.logits_pass(sd, x[:5].cpu(), arch, dtype=torch.float64)
    assert _logit_close(full[:5].cpu().numpy(), want)[0] <= 2e-3
    for n in (1, 2, 3, 5):
        got = eng.forward(x[:n].contiguous())
        err = _logit_close(got.cpu().numpy(), want[:n])[0]
        assert err <= 2e-3, (n, err)
        # different kernels may sum in a different order: equal to fp32 rounding, not necessarily bit for bit
        torch.testing.assert_close(got, full[:n], rtol=2e-4, atol=2e-5)


# ---------------------------------------------------------------------------------------------------
# the recorder end to end: pass -> snapshots -> pickle -> calculate_scores -> weights
# ---------------------------------------------------------------------------------------------------
def test_recorder_end_to_end(tmp_path, dev):
    from diagan_b200.trainer.trainer import LogitRecorder, LogTrainer, ResidentDataset
    from diagan_b200.utils.plot import calculate_scores

    class Net:                                   # stands in for the mimicry module: state_dict + train()
        def __init__(self, p): self.p, self.mode = p, "eval"
        def state_dict(self): return self.p
        def train(self): self.mode = "train"
        def eval(self): self.mode = "eval"

    n = 257
    data = _u8(n, 32, 9)
    ds = ResidentDataset(data.to(dev))
    base = sngan_oracle.init_params(32, seed=1)
    rec = LogitRecorder(ds, dev, precision="fp32")
    tr = LogTrainer(tmp_path, Net(base), recorder=rec, device=dev, logit_save_steps=100, save_logit_after=300,
                    stop_save_logit_after=800, save_steps=400)
    oracle_logits = {}
    for step in range(0, 1001, 100):
        tr.netD.p = sngan_oracle.perturb_params(base, step, 2e-2)
        tr.on_step(step)
        if tr.should_record(step):
            oracle_logits[step] = sngan_oracle.logits_pass(tr.netD.p, data, 32)
    assert tr.netD.mode == "train"                                   # trainer.py:155
    got = tr.logit_results["netD_eval"]
    assert list(got.keys()) == [300, 400, 500, 600, 700, 800]
    for s in got:
        assert got[s].dtype == np.float64 and got[s].shape == (n,)
        exact = sngan_oracle.logits_pass(sngan_oracle.perturb_params(base, s, 2e-2), data, 32, dtype=torch.float64)
        assert _fp32_ok(got[s], oracle_logits[s], exact)[0]
    saved = pickle.load(open(tmp_path / "logits_netD_eval.pkl", "rb"))   # written at step 400 and 800
    assert list(saved.keys()) == [300, 400, 500, 600, 700, 800]
    sc = calculate_scores(saved, start_epoch=300, end_epoch=800)          # 5 snapshots, 800 excluded
    want = so.calculate_scores(saved, 300, 800)
    for k in want:
        assert np.array_equal(sc[k], want[k]), k


def test_config0_colour_mnist_dcgan_chain(tmp_path, dev):
    """BASELINE configs[0] end to end at reduced size: raw 28x28 Colour-MNIST-shaped images (major_ratio 0.99) -> the
    reference transform (Resize 32 + CenterCrop, bit-exact on the GPU) -> MNIST_DCGAN discriminator in eval mode (exact
    fp32 engine) over six weight snapshots -> logits_netD_eval.pkl -> calculate_scores -> ldr_conf_1.0_ratio_50 weights
    (train_mimicry_color_mnist_phase2.py:24-27 passes them to the sampler un-floored) -> top-100 indices."""
    from diagan_b200 import engine
    from diagan_b200.trainer.trainer import LogitRecorder, LogTrainer, ResidentDataset
    from diagan_b200.utils.plot import calculate_scores
    from oracle import resize as resize_oracle

    class Net:
        def __init__(self, p): self.p = p
        def state_dict(self): return self.p
        def train(self): pass

    n = 600
    rng = np.random.RandomState(1)
    mask = (rng.rand(n, 28, 28, 1) < 0.19).astype(np.uint8)
    colour = np.where((np.arange(n) < int(n * 0.99))[:, None, None, None], np.array([255, 0, 0], np.uint8),
                      np.array([0, 255, 0], np.uint8))
    raw = (mask * colour).astype(np.uint8)[rng.permutation(n)]
    ds = ResidentDataset.from_raw_images(raw, "color_mnist", dev)
    x32 = resize_oracle.resize_center_crop_u8(raw, 32)
    assert np.array_equal(ds.data.cpu().numpy(), x32)                       # the loader's images, bit for bit

    base = dcgan_oracle.init_params(seed=4)
    tr = LogTrainer(tmp_path, Net(base), recorder=LogitRecorder(ds, dev, precision="fp32"), device=dev, logit_save_steps=100,
                    save_logit_after=0, stop_save_logit_after=500, save_steps=500, save_eval_logits=True)
    want_logits = {}
    for step in range(0, 501, 100):
        gen = torch.Generator().manual_seed(step)
        tr.netD.p = {k: (v + 2e-2 * torch.randn(v.shape, generator=gen) if k.endswith("weight") and v.dim() == 4 else v)
                     for k, v in base.items()}
        tr.on_step(step)
        want_logits[step] = dcgan_oracle.logits_pass(tr.netD.p, torch.from_numpy(x32))
    saved = pickle.load(open(tmp_path / "logits_netD_eval.pkl", "rb"))
    assert list(saved.keys()) == [0, 100, 200, 300, 400, 500]
    for s_ in saved:
        assert _logit_close(saved[s_], want_logits[s_])[0] <= 1e-5
    got = calculate_scores(saved, start_epoch=0, end_epoch=501)
    assert np.array_equal(got["ldr_conf_1.0_ratio_50"], so.calculate_scores(saved, 0, 501)["ldr_conf_1.0_ratio_50"])
    # same chain on the oracle's own logits: scores agree to the logit tolerance, and so do the selected samples
    ref = so.calculate_scores(want_logits, 0, 501)["ldr_conf_1.0_ratio_50"]
    w = got["ldr_conf_1.0_ratio_50"]
    assert np.abs(w - ref).max() <= 1e-4 * np.abs(ref).max()
    top = engine.top_indices(torch.from_numpy(w).to(dev), 100, True).cpu().numpy()
    assert np.array_equal(top, np.argsort(w, kind="stable")[-100:])
    assert len(set(top.tolist()) & set(np.argsort(ref, kind="stable")[-100:].tolist())) >= 95


def test_get_logit_from_dataloader_contract(dev):
    """The generic path: (data, target, weight, index) batches from a shuffled DataLoader, scattered by index."""
    from diagan_b200.trainer.trainer import LogTrainer

    class DS(torch.utils.data.Dataset):
        def __init__(self, x): self.x = x
        def __len__(self): return self.x.shape[0]
        def __getitem__(self, i): return sngan_oracle.normalise_u8(self.x[i:i + 1])[0], 0, 1.0, i

    class Net:
        def __init__(self, p): self.p = p
        def state_dict(self): return self.p
        def train(self): pass

    data = _u8(100, 32, 4)
    params = sngan_oracle.init_params(32, seed=2)
    loader = torch.utils.data.DataLoader(DS(data), batch_size=64, shuffle=True)
    tr = LogTrainer("/tmp", Net(params), dataloader=loader, device=dev)
    tr.recorder = None
    from diagan_b200.trainer.trainer import LogitRecorder
    tr.recorder = LogitRecorder(None, dev, precision="fp32")
    got = tr._get_logit(tr.netD, eval_mode=True)
    want = sngan_oracle.logits_pass(params, data, 32)
    exact = sngan_oracle.logits_pass(params, data, 32, dtype=torch.float64)
    ok, e_gpu, e_cpu = _fp32_ok(got, want, exact)
    print(f"dataloader path fp32: GPU vs f64 {e_gpu:.2e}, CPU fp32 vs f64 {e_cpu:.2e}")
    assert got.dtype == np.float64 and ok


def test_stylegan2_get_logit_reference_signature(golden_dir, dev):
    """``get_logit(dataloader, netD, device)`` exactly as stylegan2/train_ffhq.py:320 calls it: items ``(img, idx)``, every
    loader batch is one minibatch-stddev batch, result np.float64 [N] by dataset index, netD back in train mode."""
    from diagan_b200 import distributed as D
    from oracle import stylegan2 as sg2_oracle
    g = _load(golden_dir, "stylegan2_d32")
    batch = int(g["batch"])
    params = sg2_oracle.init_params(32, int(g["param_seed"]))
    x = sngan_oracle.normalise_u8(torch.from_numpy(g["x_u8"]))

    class DS(torch.utils.data.Dataset):
        def __len__(self): return x.shape[0]
        def __getitem__(self, i): return x[i], i

    class Net:
        mode = "eval"
        def state_dict(self): return params
        def train(self): self.mode = "train"

    net = Net()
    loader = torch.utils.data.DataLoader(DS(), batch_size=batch, shuffle=False, drop_last=True)
    got = D.get_logit(dataloader=loader, netD=net, device=dev)
    n_full = x.shape[0] // batch * batch
    assert got.dtype == np.float64 and got.shape == (x.shape[0],) and net.mode == "train"
    assert _logit_close(got[:n_full], g["logits"][:n_full])[0] <= 2e-3 and np.all(got[n_full:] == 0.0)
    # SNGAN through the same call, shuffled loader with a ragged last batch: per-sample logits land at their index
    sd = sngan_oracle.init_params(32, seed=1)
    xs = _u8(75, 32, 4)
    xn = sngan_oracle.normalise_u8(xs)

    class DS2(torch.utils.data.Dataset):
        def __len__(self): return 75
        def __getitem__(self, i): return xn[i], i

    class Net2(Net):
        def state_dict(self): return sd

    got2 = D.get_logit(torch.utils.data.DataLoader(DS2(), batch_size=16, shuffle=True), Net2(), dev)
    want2 = sngan_oracle.logits_pass(sd, xs, 32, dtype=torch.float64)
    assert _logit_close(got2, want2)[0] <= 2e-3


def test_recorder_stylegan2_drop_last(golden_dir, dev):
    """StyleGAN2 recording pass through LogitRecorder: whole batches only, the ragged tail keeps 0.0 like the
    reference's drop_last=True loader (stylegan2/train_ffhq.py:596-602)."""
    from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset
    from oracle import stylegan2 as sg2_oracle
    g = _load(golden_dir, "stylegan2_d32")
    params = sg2_oracle.init_params(32, int(g["param_seed"]))
    x = torch.from_numpy(g["x_u8"][:13])                        # 13 samples, batch 8 -> one batch + 5 dropped
    rec = LogitRecorder(ResidentDataset(x.to(dev)), dev, precision="fp32", batch=8)
    snap = rec.record(params).cpu().numpy()
    assert _logit_close(snap[:8], g["logits"][:8])[0] <= 1e-5
    assert np.all(snap[8:] == 0.0)


# ---------------------------------------------------------------------------------------------------
# BASELINE-size properties (no oracle at this size: invariants instead)
# ---------------------------------------------------------------------------------------------------
def test_full_size_recording_pass_properties(dev):
    """configs[1] at full size (50 000 x 3x32x32, SNGAN-32, tensor-core engine): the pass is deterministic, every
    sample's logit is independent of where it sits in the pass (shards / permutations give bit-identical values:
    this is what makes index sharding across GPUs exact), and a random subset agrees with the oracle."""
    from diagan_b200 import engine, synthetic
    n = 50_000
    x = synthetic.uniform_images_u8(n, 32, seed=1).to(dev)
    sd = synthetic.sngan_state_dict(32, seed=1)
    eng = engine.DiscriminatorEngine(dev).load_sngan(sd, 32, "fp16", True)
    full = eng.forward(x)
    assert torch.equal(full, eng.forward(x))                                   # deterministic
    lo, hi = 12_345, 31_111                                                    # an arbitrary shard
    assert torch.equal(eng.forward(x[lo:hi].contiguous()), full[lo:hi])
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).to(dev)
    assert torch.equal(eng.forward(x[perm].contiguous()), full[perm])          # permutation equivariance
    eng.set_chunk(1000)
    assert torch.equal(eng.forward(x), full)                                   # chunking invariance
    sub = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:192]
    want, l1 = sngan_oracle.logits_pass(sd, x[sub.to(dev)].cpu(), 32, dtype=torch.float64, with_head_l1=True)
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        ytf = sngan_oracle.forward({k: v.to(dev) for k, v in sd.items()}, sngan_oracle.normalise_u8(x[sub.to(dev)].cpu()).to(dev),
                                   32, True).view(-1).cpu().numpy()
    e, etf = _logit_close(full[sub.to(dev)].cpu().numpy(), want)[0], _logit_close(ytf, want)[0]
    eb = float((np.abs(full[sub.to(dev)].cpu().numpy() - want) / l1).max())
    print(f"full-size pass, 192-sample subset vs float64 oracle: A {e:.2e} B {eb:.2e} (torch-eager TF32: A {etf:.2e})")
    assert eb <= 2e-4 and e <= 1e-3



def test_full_size_properties(dev):
    """50k samples x 50 snapshots (configs[1] score stage): window path == Welford path to 1e-12,
    clip bounds hold, idempotent clip, top-k sorted and consistent with the score vector."""
    from diagan_b200 import engine
    n, T = 50_000, 50
    gen = torch.Generator(device="cuda").manual_seed(0)
    snaps = (1.0 + 1.5 * torch.randn(T, n, generator=gen, device=dev)).float()
    mom = engine.window_moments(snaps)
    st = engine.RunningStats(n, dev)
    for t in range(T):
        st.update(snaps[t])
    torch.testing.assert_close(st.mean, mom["mean"], rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(st.ldrv(), mom["var"], rtol=1e-10, atol=1e-14)
    ref_mean = snaps.double().mean(0)
    torch.testing.assert_close(mom["mean"], ref_mean, rtol=1e-12, atol=1e-14)
    t03 = engine.conf_from_key("ldr_conf_0.3_ratio_50")
    s = engine.scores_from_moments(mom["mean"], mom["var"], [t03, 5.0])
    for row in s:
        assert row.min().item() >= engine.FLOOR and row.max().item() <= row.min().item() * engine.RATIO * (1 + 1e-15)
    top = engine.top_indices(s[0], 100, True)
    vals = s[0][top]
    assert torch.all(vals[1:] >= vals[:-1])
    assert vals[0].item() >= torch.kthvalue(s[0], n - 99).values.item()
    bot = engine.top_indices(s[0], 100, False)
    assert s[0][bot].max().item() <= torch.kthvalue(s[0], 100).values.item()
