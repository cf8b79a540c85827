"""The CPU oracle against the golden vectors produced by the reference itself
(tests/golden/*.npz, made by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dcgan as dcgan_oracle
from oracle import drs as drs_oracle
from oracle import scores as so
from oracle import stylegan2 as sg2_oracle

CASES = ["scores_cifar_window", "scores_ffhq_window", "scores_ties"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("faithful", [False, True])
def test_scores_oracle_bit_exact(golden_dir, name, faithful):
    g = _load(golden_dir, name)
    logits = {int(s): g["logits"][i] for i, s in enumerate(g["steps"])}
    out = so.calculate_scores(logits, int(g["start"]), int(g["end"]), faithful=faithful)
    keys = [str(k) for k in g["keys"]]
    assert list(out.keys()) == keys
    for i, k in enumerate(keys):
        assert np.array_equal(out[k], g["scores"][i]), k


@pytest.mark.parametrize("name", CASES)
def test_moments_loop_matches_numpy_reduction(golden_dir, name):
    """The explicit row-by-row order (what the CUDA kernel does) == NumPy's axis-0 reduction."""
    g = _load(golden_dir, name)
    steps = g["steps"]
    sel = (steps >= g["start"]) & (steps < g["end"])
    arr = g["logits"][sel]
    mean, var = so.moments(arr)
    keys = [str(k) for k in g["keys"]]
    assert np.array_equal(mean, g["scores"][keys.index("ldrm")])
    assert np.array_equal(var, g["scores"][keys.index("ldrv")])
    for t in (so.conf_values()[2], so.conf_values()[49]):
        s = so.score_from_moments(mean, var, t)
        assert np.array_equal(s, g["scores"][keys.index(so.conf_key(t))])


@pytest.mark.parametrize("name", CASES)
def test_welford_close(golden_dir, name):
    g = _load(golden_dir, name)
    steps = g["steps"]
    sel = (steps >= g["start"]) & (steps < g["end"])
    arr = g["logits"][sel]
    T = arr.shape[0]
    mean, m2, last, sad = so.welford(arr)
    keys = [str(k) for k in g["keys"]]
    np.testing.assert_allclose(mean, g["scores"][keys.index("ldrm")], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(m2 / (T - 1), g["scores"][keys.index("ldrv")], rtol=1e-10, atol=1e-18)
    assert np.array_equal(last, g["scores"][keys.index("ldr")])
    np.testing.assert_allclose(sad / (T - 1), g["scores"][keys.index("ldrd")], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", CASES)
def test_stream_and_top_indices(golden_dir, name):
    g = _load(golden_dir, name)
    keys = [str(k) for k in g["keys"]]
    w = g["scores"][keys.index("ldr_conf_0.3_ratio_50")]
    stream = so.resample_stream(so.floor_weights(w), int(g["stream_seed"]))
    assert np.array_equal(stream, g["stream"])
    assert np.array_equal(so.top_indices(w, 20, True), g["argsort_stable"][-20:])
    assert np.array_equal(so.top_indices(w, 20, False), g["argsort_stable"][:20])


def test_conf_key_grammar():
    t = so.conf_values()
    assert len(t) == 99
    assert so.conf_key(t[2]) == "ldr_conf_0.3_ratio_50" and t[2] == 0.30000000000000004
    assert so.conf_from_key("ldr_conf_5.0_ratio_50") == 5.0


@pytest.mark.parametrize("name", ["drs_b256", "drs_b128"])
def test_drs_oracle_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    o = drs_oracle.DRSOracle(percentile=int(g["percentile"]))
    o.burn_in(list(g["burn_ldr"]))
    assert np.float32(o.maximum) == g["max_after_burn"]
    for b in range(g["ldr"].shape[0]):
        p, acc = o.accept(g["ldr"][b], g["psi"][b])
        assert np.array_equal(acc, g["accept"][b])
        assert np.float32(o.maximum) == g["max_after"][b]


def test_percentile_restatement():
    rng = np.random.RandomState(0)
    for n in (2, 3, 50, 128, 256, 257, 1000):
        for q in (80, 50, 99, 1):
            x = rng.standard_normal(n).astype(np.float32)
            assert drs_oracle.percentile_f32(x, q) == np.percentile(x.reshape(-1, 1), q)


def test_dcgan_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, "dcgan_eval")
    params = dcgan_oracle.init_params(int(g["param_seed"]))
    chk = sum(float(v.double().sum()) for v in params.values())
    assert chk == float(g["param_checksum"])
    torch.set_num_threads(1)
    y = dcgan_oracle.logits_pass(params, torch.from_numpy(g["x_u8"]))
    np.testing.assert_allclose(y, g["logits"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("size", [32, 128])
def test_stylegan2_oracle_matches_reference(golden_dir, size):
    g = _load(golden_dir, f"stylegan2_d{size}")
    params = sg2_oracle.init_params(size, int(g["param_seed"]))
    chk = sum(float(np.sum(v.numpy().astype(np.float64))) for v in params.values())
    assert chk == float(g["param_checksum"])
    torch.set_num_threads(os.cpu_count() or 1)
    y = sg2_oracle.logits_pass(params, torch.from_numpy(g["x_u8"]), size, int(g["batch"]))
    np.testing.assert_allclose(y, g["logits"], rtol=2e-5, atol=1e-6)
    assert sg2_oracle.block_channels(256) == [(128, 256), (256, 512), (512, 512), (512, 512), (512, 512), (512, 512)]


def sg2_256_inputs(g):
    """The d256 fixture stores a seed instead of 1.5 MB of random bytes (oracle/make_golden.py:make_stylegan2_256)."""
    x = np.random.RandomState(int(g["x_seed"])).randint(0, 256, (int(g["n"]), 256, 256, 3)).astype(np.uint8)
    assert int(x.astype(np.int64).sum()) == int(g["x_checksum"])
    return x


def test_stylegan2_256_oracle_matches_reference(golden_dir):
    """BASELINE configs[4] at its own size: the restatement against logits of the reference StyleGANDiscriminator(256)."""
    g = _load(golden_dir, "stylegan2_d256")
    params = sg2_oracle.init_params(256, int(g["param_seed"]))
    assert sum(float(np.sum(v.numpy().astype(np.float64))) for v in params.values()) == float(g["param_checksum"])
    torch.set_num_threads(os.cpu_count() or 1)
    y = sg2_oracle.logits_pass(params, torch.from_numpy(sg2_256_inputs(g)), 256, int(g["batch"]))
    np.testing.assert_allclose(y, g["logits"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("h,w,size", [(28, 28, 32), (218, 178, 64), (45, 37, 32), (37, 45, 32), (100, 300, 64), (32, 32, 32),
                                      (64, 48, 64), (20, 20, 64)])
def test_resize_oracle_bit_exact_vs_pillow(h, w, size):
    """oracle/resize.py (restatement of Pillow's 8-bit bilinear resample + torchvision geometry) against the real
    libraries, which ARE installed in this image: every dataset shape of the reference plus odd ones."""
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    from oracle import resize as R
    rng = np.random.RandomState(h + 7 * w)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    want = np.asarray(T.Compose([T.Resize(size), T.CenterCrop(size)])(Image.fromarray(img, mode="RGB")))
    assert np.array_equal(R.resize_center_crop_u8(img, size), want)


def test_resize_oracle_golden(golden_dir):
    """Committed fixture generated from Pillow/torchvision by oracle/make_golden.py (travels to boxes without them)."""
    from oracle import resize as R
    g = np.load(os.path.join(golden_dir, "resize_pil.npz"))
    for tag in ("mnist28", "celeba"):
        size = int(g[f"{tag}_size"])
        assert np.array_equal(R.resize_center_crop_u8(g[f"{tag}_in"], size), g[f"{tag}_out"])
