"""Graft the B200 diagnosis path onto an importable reference ``diagan`` package.

``install()`` is the one call a maintainer adds at the top of ``train_mimicry_phase1.py`` /
``train_mimicry_phase2.py`` / ``eval_gan_drs.py``; the scripts then run unchanged:

    diagan.utils.plot.calculate_scores          <- diagan_b200.utils.plot.calculate_scores
    diagan.trainer.trainer.LogTrainer._get_logit / ._save_logit
                                                <- diagan_b200.trainer.trainer._get_logit / _save_logit
    diagan.models.drs.DRS                       <- diagan_b200.models.drs.DRS
    diagan.trainer.evaluate.DRS                 <- diagan_b200.trainer.evaluate.DRS

Modules of the reference that cannot be imported in the current environment (e.g. ``diagan.trainer``
needs torch_mimicry and tensorflow) are skipped and reported.
"""
from __future__ import annotations

import importlib

_originals = {}      # (module name, attribute path) -> the reference's own object, for uninstall()


def _swap(mod, owner, attr, new):
    key = (mod.__name__, owner.__name__ if owner is not mod else "", attr)
    if key not in _originals:
        _originals[key] = (owner, attr, getattr(owner, attr, None), hasattr(owner, attr))
    setattr(owner, attr, new)


def uninstall() -> int:
    """Put the reference's own objects back (tests; A/B runs of the same script).  Returns how many were restored."""
    n = 0
    for owner, attr, old, existed in _originals.values():
        if existed:
            setattr(owner, attr, old)
        elif hasattr(owner, attr):
            delattr(owner, attr)
        n += 1
    _originals.clear()
    return n


def install(verbose: bool = True) -> dict:
    from .models import drs as b_drs
    from .trainer import evaluate as b_eval
    from .trainer import trainer as b_trainer
    from .utils import plot as b_plot

    done = {}

    def _try(modname, fn):
        try:
            mod = importlib.import_module(modname)
        except Exception as e:      # the reference module's own dependencies are missing
            done[modname] = f"skipped ({type(e).__name__}: {e})"
            return
        fn(mod)
        done[modname] = "patched"

    _try("diagan.utils.plot", lambda m: _swap(m, m, "calculate_scores", b_plot.calculate_scores))

    def _trainer(m):
        # the reference's own pass stays reachable: _get_logit delegates to it (with a warning) for discriminators or
        # modes the engine does not reproduce (train-mode DCGAN of train_mimicry_color_mnist_phase1.py, nc = 1 of the
        # MNIST/FMNIST scripts, architectures detect_arch rejects), so those scripts keep running unchanged
        if not hasattr(m.LogTrainer, "_get_logit_ref"):
            _swap(m, m.LogTrainer, "_get_logit_ref", m.LogTrainer._get_logit)
        _swap(m, m.LogTrainer, "_get_logit", b_trainer._get_logit)
        _swap(m, m.LogTrainer, "_save_logit", b_trainer._save_logit)
        _swap(m, m.LogTrainer, "_restore_logits", b_trainer._restore_logits)   # opt-in resume (call after construction)
    _try("diagan.trainer.trainer", _trainer)
    _try("diagan.models.drs", lambda m: _swap(m, m, "DRS", b_drs.DRS))
    _try("diagan.trainer.evaluate", lambda m: _swap(m, m, "DRS", b_eval.DRS))
    if verbose:
        for k, v in done.items():
            print(f"diagan_b200.patch: {k}: {v}")
    return done
