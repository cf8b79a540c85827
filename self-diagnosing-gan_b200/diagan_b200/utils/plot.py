"""Drop-in for ``diagan.utils.plot.calculate_scores`` (diagan-pkg/diagan/utils/plot.py:220-249).

Same signature, same dict of 103 float64 arrays (``ldr, ldrd, ldrv, ldrm`` + 99 ``ldr_conf_<t>_ratio_50``),
same window rule (``start_epoch <= step < end_epoch`` in dict insertion order), same one-line print.
The arithmetic runs on the GPU: one two-pass moments kernel in NumPy's evaluation order (bit-exact
mean / ddof-1 variance), then floor + per-key global min + clip kernels for all 99 multipliers at once.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from ..engine import conf_key, conf_values


def _as_device_window(logits: dict, start_epoch, end_epoch, device):
    rows = [v for k, v in logits.items() if k >= start_epoch and k < end_epoch]
    if len(rows) == 0:
        raise ValueError(f"calculate_scores: no snapshot with {start_epoch} <= step < {end_epoch}")
    if torch.is_tensor(rows[0]):
        arr = torch.stack([r.to(device) for r in rows])
        if arr.dtype not in (torch.float32, torch.float64):
            arr = arr.double()
        return arr.contiguous()
    host = np.ascontiguousarray(np.array(rows, dtype=np.float64))
    return torch.from_numpy(host).to(device)


def calculate_scores_device(logits: dict, start_epoch=50, end_epoch=75, device=None, keys=None, eps=0.0,
                            min_reduce=None) -> dict:
    """Same as :func:`calculate_scores` but returns CUDA float64 tensors (no D2H).  ``keys`` restricts the
    ldr_conf entries that are materialised; ``min_reduce`` hooks the sharded global-min exchange."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    arr = _as_device_window(logits, start_epoch, end_epoch, device)
    print(f'calculate_scores -- start_epoch: {start_epoch} end_epoch: {end_epoch} logits_arr: {tuple(arr.shape)}')
    mom = engine.window_moments(arr)
    out = {"ldr": mom["ldr"], "ldrd": mom["ldrd"], "ldrv": mom["var"], "ldrm": mom["mean"]}
    confs = conf_values() if keys is None else np.array([engine.conf_from_key(k) for k in keys])
    if len(confs):
        sc = engine.scores_from_moments(mom["mean"], mom["var"], confs, engine.FLOOR, engine.RATIO, eps,
                                        min_reduce=min_reduce)
        for j, t in enumerate(confs):
            out[conf_key(t)] = sc[j]
    return out


def calculate_scores(logits, start_epoch=50, end_epoch=75, clip_val=1.5, conf=1):
    """logits: ``{step: float64[N]}`` as pickled by ``LogTrainer._save_logit``.  ``clip_val`` and ``conf``
    are accepted and ignored, exactly like the reference (dead arguments, plot.py:220-237)."""
    dev = calculate_scores_device(logits, start_epoch, end_epoch)
    order = ["ldr", "ldrd", "ldrv", "ldrm"] + [conf_key(t) for t in conf_values()]
    stacked = torch.stack([dev[k] for k in order]).cpu().numpy()      # one D2H copy for all 103 rows
    return {k: stacked[i] for i, k in enumerate(order)}
