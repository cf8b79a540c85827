"""Drop-in for ``diagan.models.drs.DRS`` (diagan-pkg/diagan/models/drs.py:10-69).

Same constructor, attributes (``maximum, batch_size, percentile, gamma, device``) and methods.  The
acceptance arithmetic (running max, F, percentile gamma, sigmoid, compare with psi) and the compaction
of accepted samples run on the GPU (``sdg_drs_accept``); accepted images are gathered on the device by
index instead of the reference's per-image ``.cpu().numpy()`` loop (drs.py:56-57).  psi still comes
from the global NumPy stream (``np.random.rand``), so a seeded run accepts the same candidates.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, ptr, stream_ptr


class DRS(nn.Module):
    def __init__(self, netG, netD, device, gamma=None, percentile=80, batch_size=256):
        super().__init__()
        self.netG = netG
        self.netD = netD
        if torch.device(device).type != "cuda":
            raise _lib.SdgError("diagan_b200 DRS needs a CUDA device (no CPU fallback)")
        from ..engine import _resolve_device
        self.device = _resolve_device(device)
        self.batch_size = batch_size
        self.percentile = percentile
        self.gamma = gamma
        self._lib = _lib.load()
        self._max = torch.full((1,), -100000.0, dtype=torch.float32, device=self.device)   # drs.py:15
        self._count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.init_drs()

    # reference attribute: a Python/NumPy scalar
    @property
    def maximum(self):
        return np.float32(self._max.item())

    @maximum.setter
    def maximum(self, v):
        self._max.fill_(float(v))

    def _ldr_device(self, num_data):
        with torch.no_grad():
            imgs = self.netG.generate_images(num_data, device=self.device)
            out = self.netD(imgs)
            if type(out) is tuple:
                out = out[0]
            ldr = out.detach().reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
        return imgs, ldr

    def get_fake_samples_and_ldr(self, num_data):
        """Reference contract (drs.py:21-29): (imgs on device, ldr np.float32 [n,1])."""
        imgs, ldr = self._ldr_device(num_data)
        return imgs, ldr.view(-1, 1).cpu().numpy()

    def init_drs(self):
        for _ in range(50):                                         # drs.py:31-36
            _, ldr = self._ldr_device(self.batch_size)
            with torch.cuda.device(self.device):
                check(self._lib.sdg_drs_update_max(ptr(ldr), ldr.numel(), ptr(self._max), stream_ptr(self.device)),
                      "sdg_drs_update_max")

    def accept(self, ldr, psi=None, eps=1e-6, gamma=None):
        """Device acceptance pass -> (p float32 [n], accept uint8 [n], idx int32 [n], count int32 [1])."""
        if not torch.is_tensor(ldr):
            ldr = torch.from_numpy(np.ascontiguousarray(ldr, dtype=np.float32))
        ldr = ldr.reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
        n = ldr.numel()
        if psi is None:
            psi = np.random.rand(n)                                  # drs.py:54: the global NumPy stream
        psi_d = torch.from_numpy(np.ascontiguousarray(psi, dtype=np.float64)).to(self.device, non_blocking=True)
        p = torch.empty(n, dtype=torch.float32, device=self.device)
        acc = torch.empty(n, dtype=torch.uint8, device=self.device)
        idx = torch.empty(n, dtype=torch.int32, device=self.device)
        g = self.gamma if gamma is None else gamma
        with torch.cuda.device(self.device):
            check(self._lib.sdg_drs_accept(ptr(ldr), n, ptr(self._max), float(eps), float(self.percentile),
                                           0 if g is None else 1, 0.0 if g is None else float(g), ptr(psi_d), ptr(p),
                                           ptr(acc), ptr(idx), ptr(self._count), stream_ptr(self.device)),
                  "sdg_drs_accept")
        return p, acc, idx, self._count

    def sub_rejection_sampler(self, fake_samples, ldr, eps=1e-6, gamma=None):
        """Reference contract (drs.py:38-57): CPU float32 tensor of the accepted samples, in order."""
        _, _, idx, count = self.accept(ldr, eps=eps, gamma=gamma)
        k = int(count.item())
        sel = idx[:k].long()
        return fake_samples.to(self.device).index_select(0, sel).float().cpu()

    # batches of candidates scored per host round trip in generate_images (bounds the memory of the pending images)
    max_batches_in_flight = 32

    def generate_images(self, num_images, device=None):
        """drs.py:59-69 -- ``while num < num_images: imgs, ldr = get_fake...(batch); accepted = sub_rejection_sampler(...)``
        -- with the accept / compact step batched: as long as ``k = (num_images - num) // batch_size`` is at least 2, not even
        a 100 % acceptance rate could finish the loop within the next ``k`` batches, so the reference would run all of them;
        they are generated with ``k`` generator calls (same generator RNG stream), scored by ONE discriminator call,
        judged by ``k`` acceptance launches that update the running maximum in order on the device (psi = one draw of
        ``k * batch_size`` from the global NumPy stream == ``k`` draws of ``batch_size``), and read back with ONE host sync.
        The accepted images, their order and both RNG streams after the call are those of the sequential loop."""
        chunks, num = [], 0
        B = self.batch_size
        while num < num_images:
            k = max(1, min(self.max_batches_in_flight, (num_images - num) // B))
            with torch.no_grad():
                imgs = [self.netG.generate_images(B, device=self.device) for _ in range(k)]
                x = torch.cat(imgs, dim=0) if k > 1 else imgs[0]
                out = self.netD(x)
                if type(out) is tuple:
                    out = out[0]
                ldr = out.detach().reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
            psi = torch.from_numpy(np.random.rand(k * B)).to(self.device, non_blocking=True)     # drs.py:54
            idx = torch.empty(k, B, dtype=torch.int32, device=self.device)
            cnt = torch.zeros(k, dtype=torch.int32, device=self.device)
            g = self.gamma
            with torch.cuda.device(self.device):
                for j in range(k):
                    check(self._lib.sdg_drs_accept(ptr(ldr[j * B:(j + 1) * B]), B, ptr(self._max), 1e-6, float(self.percentile),
                                                   0 if g is None else 1, 0.0 if g is None else float(g),
                                                   ptr(psi[j * B:(j + 1) * B]), None, None, ptr(idx[j]), ptr(cnt[j:j + 1]),
                                                   stream_ptr(self.device)), "sdg_drs_accept")
            counts = cnt.cpu().tolist()                                   # the one host sync of this round
            for j in range(k):
                if counts[j]:
                    chunks.append(imgs[j].index_select(0, idx[j, :counts[j]].long()))
                num += counts[j]
        out = torch.cat(chunks, dim=0)[:num_images]
        return out.cpu().float() if device is None else out.to(device)
