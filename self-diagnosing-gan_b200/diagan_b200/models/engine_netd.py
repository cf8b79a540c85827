"""``netD`` stand-in backed by the CUDA engine, for callers that invoke the discriminator as a module.

``DRS.get_fake_samples_and_ldr`` (diagan-pkg/diagan/models/drs.py:21-29, trainer/evaluate.py:36-43) and
``LogTrainer._get_logit`` (trainer.py:150) call ``netD(x)`` on float32 NCHW batches in [-1, 1] and expect a ``[B, 1]``
tensor.  ``EngineNetD`` keeps that contract while the forward runs in ``libsdg.so`` (tensor-core engine by default):
construct it from the torch-mimicry / reference module (or its ``state_dict``) once, call ``refresh()`` after the
weights change.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine


class EngineNetD(nn.Module):
    def __init__(self, netD_or_state_dict, device=None, precision=None, inplace_relu=True):
        super().__init__()
        self._src = netD_or_state_dict
        self._precision = precision
        self._inplace = inplace_relu
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.engine = engine.DiscriminatorEngine(self.device)
        self.refresh()

    def _state_dict(self):
        s = self._src
        return s.state_dict() if hasattr(s, "state_dict") else s

    def refresh(self):
        """Re-pack the weights (sigma recomputed once, eval semantics) after the source module changed."""
        self.engine.load(self._state_dict(), self._precision, self._inplace)
        return self

    def forward(self, x):
        x = x.to(device=self.device)
        if x.dtype != torch.uint8:
            x = x.to(torch.float32)
        return self.engine.forward(x.contiguous()).view(-1, 1)
