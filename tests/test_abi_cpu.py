"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/sdg.h
declares, the ctypes binding covers exactly that set, and the host logic (key grammar, architecture
detection, shard arithmetic, loud failure without a GPU) behaves.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sdg.h")).read()
    return sorted(set(re.findall(r"SDG_API\s+[\w\s\*]+?\b(sdg_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from diagan_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libsdg.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdg.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == names, "ctypes binding and header disagree"
    assert _lib.load().sdg_abi_version() == _lib.ABI_VERSION


def test_constants_match_header():
    from diagan_b200 import _lib
    src = open(os.path.join(ROOT, "include", "sdg.h")).read()
    val = lambda name: int(re.search(rf"#define\s+{name}\s+\(?(-?\d+)\)?", src).group(1))
    assert val("SDG_ARCH_SNGAN32") == _lib.ARCH_SNGAN32 and val("SDG_ARCH_SNGAN64") == _lib.ARCH_SNGAN64
    assert val("SDG_ARCH_DCGAN32") == _lib.ARCH_DCGAN32
    assert val("SDG_PREC_FP32") == _lib.PREC_FP32 and val("SDG_PREC_BF16") == _lib.PREC_BF16
    assert val("SDG_LAYOUT_U8_NHWC") == _lib.LAYOUT_U8_NHWC and val("SDG_LAYOUT_F32_NCHW") == _lib.LAYOUT_F32_NCHW
    assert val("SDG_ABI_VERSION") == _lib.ABI_VERSION


def test_error_path_without_device_is_loud():
    """Invalid arguments come back as negative codes with a message; nothing falls back to the CPU."""
    from diagan_b200 import _lib
    lib = _lib.load()
    rc = lib.sdg_stats_update(None, None, None, None, None, 4, 0, None)
    assert rc == -1 and b"null pointer" in lib.sdg_last_error()
    rc = lib.sdg_topk_indices(None, 10, 3, 1, None, None, 0, None)
    assert rc == -1
    if not torch.cuda.is_available():
        from diagan_b200 import engine
        with pytest.raises(_lib.SdgError):
            engine.DiscriminatorEngine()
        with pytest.raises(_lib.SdgError):
            engine.window_moments(torch.zeros(3, 4, dtype=torch.float64))


def test_key_grammar_and_arch_detection():
    from diagan_b200 import engine
    from oracle import dcgan, sngan
    t = engine.conf_values()
    assert len(t) == 99 and engine.conf_key(t[2]) == "ldr_conf_0.3_ratio_50"
    assert engine.conf_from_key("ldr_conf_0.3_ratio_50") == 0.30000000000000004
    with pytest.raises(KeyError):
        engine.conf_from_key("ldr_conf_10.0_ratio_50")
    assert engine.detect_arch(sngan.init_params(32)) == "sngan32"
    assert engine.detect_arch(sngan.init_params(64)) == "sngan64"
    assert engine.detect_arch(dcgan.init_params()) == "dcgan32"
    # InfoMax-GAN / SSGAN discriminators (predefined_models.py:36-52): the SNGAN stack under other names / with extra heads
    for arch in (32, 64):
        p = sngan.init_params(arch)
        for variant in ("ssgan", "infomax"):
            sd = sngan.as_variant_state_dict(p, arch, variant)
            kind = engine.detect_arch(sd)
            assert kind == f"{variant}{arch}"
            canon = engine.canonical_sngan_state_dict(sd, kind)
            for k in engine.sngan_layer_keys(arch):
                assert canon[f"{k}.weight"] is p[f"{k}.weight"] and canon[f"{k}.sn_u"] is p[f"{k}.sn_u"]
    assert [k + ".weight" in sngan.init_params(32) for k in engine.sngan_layer_keys(32)] == [True] * 11
    assert len(engine.sngan_layer_keys(64)) == 16
    # layer order of the ABI == forward order of the oracle
    assert engine.sngan_layer_keys(32) == [k for k, _, _, _ in sngan.layer_list(32)]
    assert engine.sngan_layer_keys(64) == [k for k, _, _, _ in sngan.layer_list(64)]


def test_shard_ranges_cover_dataset():
    from diagan_b200 import distributed as D
    for n in (0, 1, 7, 50000, 162770, 202599):
        for w in (1, 2, 3, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = D.shard_range(n, r, w)
                assert 0 <= lo <= hi <= n and hi - lo <= D.shard_size(n, w)
                cover += list(range(lo, hi)) if n < 100 else [(lo, hi)]
            if n < 100:
                assert cover == list(range(n))
            else:
                assert cover[0][0] == 0 and cover[-1][1] == n
                assert all(cover[i][1] == cover[i + 1][0] for i in range(w - 1))


def test_recording_trigger_matches_reference_rule(tmp_path):
    """trainer.py:328: step % logit_save_steps == 0 and save_logit_after <= step <= stop_save_logit_after."""
    from diagan_b200.trainer.trainer import LogTrainer
    tr = LogTrainer.__new__(LogTrainer)
    tr.save_logits, tr.logit_save_steps, tr.save_logit_after, tr.stop_save_logit_after = True, 100, 35000, 40000
    hits = [s for s in range(34000, 41001) if tr.should_record(s)]
    assert hits == list(range(35000, 40001, 100)) and len(hits) == 51      # train_mimicry_phase1.py:88-92
    tr.save_logits = False
    assert not tr.should_record(35000)


def test_save_logit_pickle_schema(tmp_path):
    """logits_<name>.pkl == {step: float64[N]} loadable the way train_mimicry_phase2.py:87-92 does."""
    import pickle
    from diagan_b200.trainer.trainer import LogTrainer
    tr = LogTrainer.__new__(LogTrainer)
    tr.output_path = tmp_path
    res = {"netD_eval": {35000: np.arange(5, dtype=np.float64), 35100: torch.arange(5, dtype=torch.float32)}}
    tr._save_logit(res)
    got = pickle.load(open(tmp_path / "logits_netD_eval.pkl", "rb"))
    assert list(got.keys()) == [35000, 35100]
    assert all(v.dtype == np.float64 and v.shape == (5,) for v in got.values())


def test_save_logit_is_atomic_and_resume_merges_history(tmp_path):
    """the pickle keeps the reference's schema, is replaced atomically, and a restarted run extends
    the history instead of overwriting it (the reference loses it: trainer.py:222 starts empty, :138-140 overwrites)."""
    import pickle
    from collections import defaultdict
    from diagan_b200.trainer import trainer as T
    tr = T.LogTrainer.__new__(T.LogTrainer)
    tr.output_path, tr.save_f32_sidecar = tmp_path, True
    first = {"netD_eval": {35000: np.arange(6, dtype=np.float64), 35100: np.arange(6, dtype=np.float64) + 1}}
    tr._save_logit(first)
    assert not list(tmp_path.glob("*.tmp"))
    # "restart": a fresh trainer records two more steps, one of them again (35100 is overwritten by the new row)
    tr2 = T.LogTrainer.__new__(T.LogTrainer)
    tr2.output_path, tr2.save_f32_sidecar = tmp_path, True
    tr2.logit_results = defaultdict(dict)
    tr2.logit_results["netD_eval"][35100] = np.full(6, 7.0)
    tr2.logit_results["netD_eval"][35200] = np.full(6, 9.0)
    assert tr2._restore_logits() == 2
    tr2._save_logit(tr2.logit_results)
    got = pickle.load(open(tmp_path / "logits_netD_eval.pkl", "rb"))
    assert list(got.keys()) == [35000, 35100, 35200]
    assert np.array_equal(got[35000], first["netD_eval"][35000]) and np.all(got[35100] == 7.0) and np.all(got[35200] == 9.0)
    assert type(got) is dict and all(v.dtype == np.float64 for v in got.values())
    # compact side-file: same steps in the same order, float32 rows
    side = T.load_logits(tmp_path / "logits_netD_eval_f32.npz")
    assert list(side.keys()) == [35000, 35100, 35200] and all(v.dtype == np.float32 for v in side.values())
    assert all(np.array_equal(side[k].astype(np.float64), got[k]) for k in got)
    assert list(T.load_logits(tmp_path / "logits_netD_eval.pkl").keys()) == [35000, 35100, 35200]
