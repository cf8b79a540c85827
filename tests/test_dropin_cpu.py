"""The drop-in claim, on the reference's REAL classes: ``diagan_b200.patch.install()`` must bind onto
``diagan.trainer.trainer.LogTrainer`` / ``diagan.trainer.evaluate.DRS`` / ``diagan.models.drs.DRS`` /
``diagan.utils.plot.calculate_scores`` as they lie under /root/reference (imported in place, nothing copied; packages that
are absent here -- torch_mimicry, tensorflow, lmdb ... -- are stubbed by ``oracle.ref_loader.stub_training_stack``), keep
their call signatures, and keep the scripts the engine does not serve running through the reference's own method.

CPU only: the engine itself cannot run here, so what is exercised is the binding, the signature contract and the
fallback routing; the grafted pass on a GPU is ``tests/test_gpu_parity.py::test_grafted_get_logit_*``.
Skipped where /root/reference does not exist (the GPU box)."""
import inspect
import warnings

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, Dataset

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


class _Items(Dataset):
    """The reference's WeightedDataset item contract (predefined.py:22-24): (data, target, weight, index)."""

    def __init__(self, x):
        self.x = x

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], 0, 1.0, i


@pytest.fixture()
def grafted():
    ref_loader.stub_training_stack()
    import diagan.models.drs as ref_drs
    import diagan.trainer.evaluate as ref_eval
    import diagan.trainer.trainer as ref_trainer
    import diagan.utils.plot as ref_plot
    from diagan_b200 import patch
    originals = {
        "get_logit": ref_trainer.LogTrainer._get_logit, "save_logit": ref_trainer.LogTrainer._save_logit,
        "scores": ref_plot.calculate_scores, "drs": ref_drs.DRS, "eval_drs": ref_eval.DRS,
    }
    done = patch.install(verbose=False)
    yield done, originals, (ref_trainer, ref_eval, ref_drs, ref_plot)
    patch.uninstall()
    assert ref_trainer.LogTrainer._get_logit is originals["get_logit"]
    assert ref_plot.calculate_scores is originals["scores"]
    assert not hasattr(ref_trainer.LogTrainer, "_get_logit_ref")


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]


def test_install_patches_all_four_targets_with_the_same_signatures(grafted):
    done, orig, (ref_trainer, ref_eval, ref_drs, ref_plot) = grafted
    assert done == {"diagan.utils.plot": "patched", "diagan.trainer.trainer": "patched", "diagan.models.drs": "patched",
                    "diagan.trainer.evaluate": "patched"}
    from diagan_b200.models import drs as b_drs
    from diagan_b200.trainer import evaluate as b_eval
    from diagan_b200.trainer import trainer as b_trainer
    from diagan_b200.utils import plot as b_plot
    LT = ref_trainer.LogTrainer
    assert LT._get_logit is b_trainer._get_logit and LT._save_logit is b_trainer._save_logit
    assert LT._get_logit_ref is orig["get_logit"]
    assert ref_plot.calculate_scores is b_plot.calculate_scores
    assert ref_drs.DRS is b_drs.DRS and ref_eval.DRS is b_eval.DRS
    # call signatures: identical for the functions, a superset (same leading parameters and defaults) for the classes
    assert _params(LT._get_logit) == _params(orig["get_logit"])               # (self, netD, eval_mode=False)
    assert _params(LT._save_logit) == _params(orig["save_logit"])             # (self, logits_dict)
    assert _params(ref_plot.calculate_scores) == _params(orig["scores"])      # (logits, start_epoch=50, end_epoch=75, ...)
    for ours, theirs in ((b_drs.DRS, orig["drs"]), (b_eval.DRS, orig["eval_drs"])):
        po, pt = _params(ours.__init__), _params(theirs.__init__)
        assert po[:len(pt)] == pt, (po, pt)
        for name in ("get_fake_samples_and_ldr", "init_drs", "sub_rejection_sampler", "generate_images"):
            a, b = _params(getattr(ours, name)), _params(getattr(theirs, name))
            assert a[:len(b)] == b, (name, a, b)
    assert hasattr(b_eval.DRS, "visualize_images")
    # the grafted methods are bound to instances of the REAL class
    t = LT.__new__(LT)
    assert t._get_logit.__func__ is b_trainer._get_logit


def _trainer_instance(LT, x, batch=16):
    t = LT.__new__(LT)                       # the real constructor builds optimisers / loggers the pass does not need
    t.dataloader = DataLoader(_Items(x), batch_size=batch, shuffle=True)
    t.device = torch.device("cpu")
    return t


@pytest.mark.parametrize("case", ["nc1_eval", "nc3_train"])
def test_unsupported_discriminators_run_through_the_reference_method(grafted, case):
    """train_mimicry_color_mnist_phase1.py (save_eval_logits=False) and the MNIST/FMNIST scripts (nc = 1) must not
    abort at the first recording step once install() is in place."""
    _, orig, (ref_trainer, _, _, _) = grafted
    D = ref_loader.reference_dcgan_discriminator()
    torch.manual_seed(3)
    nc = 1 if case == "nc1_eval" else 3
    netD = D(nc=nc)
    netD.device = torch.device("cpu")
    x = torch.randn(40, nc, 32, 32)
    t = _trainer_instance(ref_trainer.LogTrainer, x)
    eval_mode = case == "nc1_eval"
    with pytest.warns(RuntimeWarning, match="reference's own LogTrainer._get_logit"):
        torch.manual_seed(11)
        got = t._get_logit(netD=netD, eval_mode=eval_mode)
    assert netD.training                                                         # trainer.py:155
    assert got.dtype == np.float64 and got.shape == (40,)
    if eval_mode:                                                                # deterministic: compare with a direct call
        netD.eval()
        with torch.no_grad():
            want = netD(x).view(-1).double().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    else:                                                                        # stochastic (Dropout): same RNG, same result
        torch.manual_seed(11)
        want = orig["get_logit"](t, netD=netD, eval_mode=False)
        np.testing.assert_array_equal(got, want)


def test_without_the_reference_method_unsupported_cases_raise():
    """Standalone ``diagan_b200.trainer.trainer.LogTrainer`` has no reference method to hand over to: it raises, and does so
    BEFORE touching the GPU (so the message is about the model, not about a missing device)."""
    from diagan_b200 import _lib
    from diagan_b200.trainer import trainer as b_trainer
    D = ref_loader.reference_dcgan_discriminator()
    t = b_trainer.LogTrainer.__new__(b_trainer.LogTrainer)
    t.recorder = None
    with pytest.raises(_lib.SdgError, match="nc \\* num_pack = 1"):
        t._get_logit(D(nc=1), eval_mode=True)
    with pytest.raises(_lib.SdgError, match="train-mode logits of the DCGAN"):
        t._get_logit(D(nc=3), eval_mode=False)

    class Other(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(4, 1)
    with pytest.raises(_lib.SdgError, match="unsupported discriminator"):
        t._get_logit(Other(), eval_mode=True)


def test_supported_discriminator_is_not_handed_to_the_reference(grafted):
    """An eval-mode nc = 3 DCGAN is the engine's job: without a GPU the grafted method must fail loudly (no CUDA), not
    quietly fall back to the reference's PyTorch pass."""
    _, _, (ref_trainer, _, _, _) = grafted
    from diagan_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    D = ref_loader.reference_dcgan_discriminator()
    netD = D(nc=3)
    t = _trainer_instance(ref_trainer.LogTrainer, torch.randn(8, 3, 32, 32))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        with pytest.raises(_lib.SdgError, match="no CUDA device|no CPU"):
            t._get_logit(netD=netD, eval_mode=True)
