"""diagan_b200 -- B200-native per-sample diagnosis path of Self-Diagnosing GAN.

Module layout mirrors the slice of the reference's ``diagan`` package that lies on the path:

    diagan.utils.plot.calculate_scores        -> diagan_b200.utils.plot.calculate_scores
    diagan.trainer.trainer.LogTrainer          -> diagan_b200.trainer.trainer.LogTrainer / LogitRecorder
      ._get_logit / ._save_logit
    diagan.models.drs.DRS                      -> diagan_b200.models.drs.DRS
    diagan.trainer.evaluate.DRS                -> diagan_b200.trainer.evaluate.DRS
    stylegan2/train_ffhq.py get_logit,
      concat_all_gather, save_logit            -> diagan_b200.distributed

All arithmetic runs in ``libsdg.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/sdg.h``).
There is no CPU fallback: importing works anywhere, but the first call without the built library or
without a B200 raises ``SdgError``.  ``diagan_b200.patch.install()`` swaps these entry points into an
importable reference ``diagan`` package.
"""
from ._lib import SdgError  # noqa: F401

__version__ = "0.1.0"
