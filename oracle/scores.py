"""NumPy restatement of the reference's LDR scoring stage (oracle; test infrastructure only).

Follows, function by function:

* ``calculate_scores``      diagan-pkg/diagan/utils/plot.py:220-249
* weight floor + sampler    train_mimicry_phase2.py:21-34  (eps 1e-6, WeightedRandomSampler)
* top-index consumers       eval_gan_drs_with_index.py:97-99, eval_gan_with_index.py:93-95

Pinned by ``tests/golden/scores_*.npz`` (outputs of the reference's own
``calculate_scores`` imported from /root/reference by ``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

FLOOR = 1e-2          # plot.py:230  clip_min lower bound
RATIO = 50            # plot.py:248  clip_max_ratio(..., ratio=50)


def conf_values() -> np.ndarray:
    """The 99 confidence multipliers the reference iterates (plot.py:247).

    They are ``np.arange(0.1, 10.0, 0.1)[i]`` -- NOT the decimal printed in the key:
    key ``ldr_conf_0.3_ratio_50`` uses t = 0.30000000000000004.
    """
    return np.arange(0.1, 10.0, 0.1)


def conf_key(t: float) -> str:
    return f"ldr_conf_{t:.1f}_ratio_50"


def conf_from_key(key: str) -> float:
    """Map a score key back to the exact float64 multiplier the reference used."""
    for t in conf_values():
        if conf_key(t) == key:
            return float(t)
    raise KeyError(key)


def window(logits: dict, start_epoch: int, end_epoch: int) -> np.ndarray:
    """Snapshot selection by dict key in insertion order (plot.py:239)."""
    return np.array([v for k, v in logits.items() if k >= start_epoch and k < end_epoch])


def moments(arr: np.ndarray):
    """mean over axis 0 and ddof=1 std/var, the way NumPy evaluates plot.py:245-248.

    For a C-contiguous [T, N] float64 array NumPy reduces axis 0 by adding rows in
    order (no pairwise blocking on the strided axis): mean = (((x0+x1)+x2)+...)/T and
    var = sum_t |x_t - mean|^2 / (T-1), again row by row.  The CUDA snapshot kernel
    follows exactly this order so the two agree bit for bit.
    """
    T = arr.shape[0]
    acc = np.zeros(arr.shape[1:], dtype=np.float64)
    for t in range(T):
        acc = acc + arr[t]
    mean = acc / T
    sq = np.zeros_like(acc)
    for t in range(T):
        d = arr[t] - mean
        sq = sq + d * d
    var = sq / (T - 1)
    return mean, var


def score_from_moments(mean, var, t, floor=FLOOR, ratio=RATIO, global_min=None):
    """clip_max_ratio(clip_min(mean + t*std)) (plot.py:226-231,248).

    ``global_min`` lets a sharded caller supply min over all shards of the
    floored score (the only cross-sample coupling in the whole stage).
    """
    s = np.clip(mean + t * np.sqrt(var), a_min=floor, a_max=None)
    m = s.min() if global_min is None else global_min
    return np.clip(s, None, m * ratio)


def calculate_scores(logits: dict, start_epoch=50, end_epoch=75, faithful=False) -> dict:
    """Restatement of plot.py:220-249 -> dict of 103 float64 [N] arrays.

    ``faithful=True`` recomputes mean/std for each of the 99 keys exactly as the
    reference does (that is what its CPU cost is made of; used for the timed CPU
    baseline).  ``faithful=False`` computes them once -- same values bit for bit.
    """
    arr = window(logits, start_epoch, end_epoch)
    out = {}
    out["ldr"] = arr[-1]
    out["ldrd"] = np.abs(arr[1:] - arr[:-1]).mean(0)
    out["ldrv"] = np.var(arr, axis=0, ddof=1)
    out["ldrm"] = arr.mean(0)
    if faithful:
        for t in conf_values():
            s = np.clip(arr.mean(0) + t * np.std(arr, 0, ddof=1), a_min=FLOOR, a_max=None)
            out[conf_key(t)] = np.clip(s, None, s.min() * RATIO)
    else:
        mean = arr.mean(0)
        std = np.std(arr, 0, ddof=1)
        for t in conf_values():
            s = np.clip(mean + t * std, a_min=FLOOR, a_max=None)
            out[conf_key(t)] = np.clip(s, None, s.min() * RATIO)
    return out


def welford(arr: np.ndarray):
    """Streaming restatement (new design): one update per snapshot.

    Returns mean, M2, last, sum|delta| after consuming arr[0..T-1] in order, using
    the update the CUDA ``sdg_stats_update`` kernel performs.  Not bit-identical to
    ``moments`` (different rounding order); tests bound it at 1e-12 relative.
    """
    mean = np.zeros(arr.shape[1:], np.float64)
    m2 = np.zeros_like(mean)
    last = np.zeros_like(mean)
    sad = np.zeros_like(mean)
    for t in range(arr.shape[0]):
        x = arr[t].astype(np.float64)
        if t > 0:
            sad = sad + np.abs(x - last)
        d = x - mean
        mean = mean + d / (t + 1)
        m2 = m2 + d * (x - mean)
        last = x
    return mean, m2, last, sad


def floor_weights(w: np.ndarray, eps=1e-6) -> np.ndarray:
    """train_mimicry_phase2.py:23  [eps if i < eps else i for i in weights]."""
    return np.where(w < eps, eps, w)


def resample_stream(weights: np.ndarray, seed: int, num_samples=None) -> np.ndarray:
    """Index stream of WeightedRandomSampler(w, N, replacement=True) for one epoch.

    torch.utils.data.WeightedRandomSampler.__iter__ draws
    ``torch.multinomial(weights.double(), num_samples, True, generator=None)`` on the
    CPU default generator (train_mimicry_phase2.py:24).
    """
    import torch
    n = len(weights) if num_samples is None else num_samples
    torch.manual_seed(seed)
    w = torch.as_tensor(np.asarray(weights), dtype=torch.double)
    return torch.multinomial(w, n, True).numpy()


def top_indices(score: np.ndarray, k: int, largest=True) -> np.ndarray:
    """``np.argsort(w)[-k:]`` / ``[:k]`` (eval_gan_drs_with_index.py:97-99) with the
    tie-break made explicit: stable sort, i.e. ties ordered by ascending sample index."""
    order = np.argsort(score, kind="stable")
    return order[-k:] if largest else order[:k]
