"""NumPy restatement of Discriminator Rejection Sampling acceptance (oracle; test infrastructure only).

Follows ``diagan-pkg/diagan/models/drs.py:7-69`` and its twin
``diagan-pkg/diagan/trainer/evaluate.py:23-83`` (fixed 80th percentile, ``batch_size`` arg).
All arithmetic is float32 NumPy exactly as in the reference: ``ldr`` is the
``.cpu().numpy()`` of a float32 ``[n, 1]`` discriminator output (drs.py:28), the running
maximum becomes a ``np.float32`` after the first batch (drs.py:34-36), ``eps`` is a weak
Python float, and ``np.percentile`` of a float32 array returns float32.

Pinned by ``tests/golden/drs_*.npz`` (outputs of the reference ``diagan.models.drs.DRS``
imported from /root/reference by ``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np


def percentile_f32(values: np.ndarray, q: float) -> np.float32:
    """Linear-interpolation percentile of a float32 vector, restating what ``np.percentile(F, q)``
    evaluates for a float32 array in the NumPy this repo runs on (2.x;
    numpy/lib/_function_base_impl.py ``percentile`` / ``_quantile`` / ``_lerp``): EVERYTHING is
    float32 -- quantile = q / float32(100), virtual index = (n-1) * quantile, weight t = index -
    floor(index), and ``a + (b-a)*t`` for t < 0.5 or ``b - (b-a)*(1-t)`` otherwise.  The CUDA kernel
    uses this same formula with contraction disabled."""
    f32 = np.float32
    s = np.sort(np.asarray(values, dtype=f32).ravel())
    n = s.shape[0]
    quant = f32(q) / f32(100)
    vidx = f32(n - 1) * quant
    lo = int(np.floor(vidx))
    hi = min(lo + 1, n - 1)
    t = f32(vidx - f32(lo))
    a, b = s[lo], s[hi]
    d = f32(b - a)
    if t >= f32(0.5):
        return f32(b - f32(d * f32(f32(1) - t)))
    return f32(a + f32(d * t))


class DRSOracle:
    """State + per-batch acceptance of drs.py:10-57 without the G/D modules."""

    def __init__(self, percentile=80, gamma=None):
        self.maximum = -100000            # drs.py:15
        self.percentile = percentile
        self.gamma = gamma

    def burn_in(self, ldr_batches):
        """drs.py:31-36: running max over the 50 burn-in batches."""
        for ldr in ldr_batches:
            m = ldr.max()
            if self.maximum < m:
                self.maximum = m

    def accept(self, ldr: np.ndarray, psi: np.ndarray, eps=1e-6):
        """drs.py:38-57.  ``ldr`` float32 [n,1] (or [n]); ``psi`` = the n uniforms the
        reference draws with ``np.random.rand(len(sigF))``.  Returns (p float32 [n],
        accept bool [n])."""
        ldr = np.asarray(ldr, dtype=np.float32).reshape(-1, 1)
        m = ldr.max()
        if m > self.maximum:
            self.maximum = m
        lm = ldr - self.maximum
        F = lm - np.log(1 - np.exp(lm - eps))
        gamma = np.percentile(F, self.percentile) if self.gamma is None else self.gamma
        F = F - gamma
        p = 1 / (1 + np.exp(-F))
        p = p.reshape(-1)
        return p, p > np.asarray(psi).reshape(-1)
