"""The recording side of ``diagan.trainer.trainer.LogTrainer`` (diagan-pkg/diagan/trainer/trainer.py).

What is mirrored (same names, argument meaning, return types):

* ``_get_logit(netD, eval_mode=False) -> np.float64 [len(dataset)]`` indexed by dataset index, leaves
  ``netD`` in train mode on exit                                              (trainer.py:142-156)
* ``_save_logit(logits_dict)`` -> ``output_path / 'logits_<name>.pkl'``, ``{step: float64[N]}``
                                                                              (trainer.py:138-140)
* the recording trigger and the ``logit_results`` bookkeeping                 (trainer.py:222,328-351)

What is NOT here: the GAN optimisation loop of ``LogTrainer.train()`` (trainer.py:208-327), which is
outside the diagnosis path.  ``diagan_b200.patch.install()`` grafts ``_get_logit`` / ``_save_logit`` onto
the reference's own ``LogTrainer`` so its ``train()`` runs unchanged.

The B200-native difference: the whole training set is resident in HBM (uint8 NHWC, 154 MB for
CIFAR-10), the discriminator weights are re-packed once per pass (sigma is constant in eval mode) and
the pass is a handful of launches per chunk of samples instead of 782 DataLoader batches with a
device->host sync each.
"""
from __future__ import annotations

import os
import pickle
import warnings
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from .. import _lib, engine


class ResidentDataset:
    """The training set as the recording pass wants it: one device tensor in dataset-index order,
    either raw uint8 NHWC (normalised on load, transform.py:3-11) or float32 NCHW in [-1,1]."""

    def __init__(self, data: torch.Tensor):
        if not data.is_cuda:
            raise _lib.SdgError("ResidentDataset needs a CUDA tensor")
        if data.dtype == torch.uint8:
            assert data.dim() == 4 and data.shape[3] == 3, "uint8 data must be [N,H,W,3]"
        elif data.dtype == torch.float32:
            assert data.dim() == 4 and data.shape[1] == 3, "float32 data must be [N,3,H,W]"
        else:
            raise _lib.SdgError(f"unsupported dataset dtype {data.dtype}")
        self.data = data.contiguous()

    def __len__(self):
        return self.data.shape[0]

    @classmethod
    def from_numpy_u8(cls, arr: np.ndarray, device):
        """e.g. ``torchvision.datasets.CIFAR10(...).data`` ([N,32,32,3] uint8)."""
        return cls(torch.from_numpy(np.ascontiguousarray(arr)).to(device))

    @classmethod
    def from_raw_images(cls, images_u8: np.ndarray, dataset_name: str, device):
        """Raw uint8 images [N,H,W,3] at their stored size (CelebA 218x178, Colour-MNIST 28x28 ...) -> the reference's
        Resize + CenterCrop (datasets/transform.py) done once on the GPU, bit-exact with the PIL pipeline."""
        from ..datasets.transform import get_transform
        return cls(get_transform(dataset_name).from_host(images_u8, device))

    @classmethod
    def from_dataset(cls, dataset, device, batch_size=1024):
        """Materialise any ``WeightedDataset``-style dataset (``(data, target, weight, index)`` items,
        predefined.py:17-27, or ``(img, idx)``, stylegan2/dataset.py:63) once, as float32 NCHW."""
        n = len(dataset)
        first = dataset[0]
        shape = tuple(first[0].shape)
        out = torch.empty((n,) + shape, dtype=torch.float32, device=device)
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=0)
        for item in loader:
            data, idx = item[0], item[-1]
            out[idx.to(device)] = data.to(device=device, dtype=torch.float32)
        return cls(out)


class LogitRecorder:
    """Full-dataset discriminator logit recording + running statistics for one rank's shard.

    ``shard = (lo, hi)`` restricts the pass to dataset indices [lo, hi) (multi-GPU); the
    snapshot vector is always full length N so that ``diagan_b200.distributed`` can all-gather in place.
    """

    def __init__(self, dataset: ResidentDataset = None, device=None, precision="fp16", inplace_relu=True,
                 shard=None, keep_snapshots=True, stats_window=None, batch=4, range_fallback="bf16"):
        self.device = engine._resolve_device(device)
        self.dataset = dataset
        self.precision = precision
        self.inplace_relu = inplace_relu
        self.engine = engine.DiscriminatorEngine(self.device)
        self.n = len(dataset) if dataset is not None else None
        self.shard = shard
        self.keep_snapshots = keep_snapshots
        self.stats_window = stats_window          # (start, end): only steps in [start, end) feed the stats
        self.snapshots = {}                       # step -> float32 [N] device tensor
        self.stats = None
        # StyleGAN2 only: the loader batch size of the reference pass (stylegan2/train_ffhq.py:596-602); minibatch-stddev
        # groups live inside consecutive batches of this size and the ragged tail is dropped (drop_last=True)
        self.batch = batch
        self.shard_multiple = 1                   # set by distributed.get_logit_resident (StyleGAN2: whole loader batches)
        # fp16 range guard: a pass during which a value left the fp16 range is re-run with this precision ("bf16" has the
        # fp32 range at 8x the rounding error; "fp32" is the exact CUDA-core engine, ~50x slower)
        self.range_fallback = range_fallback
        self.range_events = 0                     # passes that had to be re-run

    def _guarded(self, netD, run, range_check):
        """Run one pass (``run(precision)`` loads the weights and launches the forward).  ``range_check``:
        "sync"     -- read the fp16 range flag after the pass (one 4-byte D2H) and, if a value left the fp16 range,
                      warn and re-run the pass with ``range_fallback``;
        "deferred" -- do not touch the host; the flag stays set until :meth:`check_range` is called;
        "off"      -- ignore the flag."""
        out = run(self.precision)
        if range_check != "sync" or self.precision != "fp16":
            return out
        flag = self.engine.range_status(reset=True)
        if flag:
            self.range_events += 1
            what = " and ".join(n for b, n in ((_lib.RANGE_ACT, "an activation"), (_lib.RANGE_WEIGHT, "a packed weight")) if flag & b)
            warnings.warn(f"diagan_b200: {what} left the fp16 range (|v| > 65504) during this recording pass; re-running "
                          f"it with precision={self.range_fallback!r}", RuntimeWarning, stacklevel=3)
            out = run(self.range_fallback)
        return out

    def check_range(self):
        """Deferred form of the range guard: raises if any fp16 pass since the last check left the fp16 range."""
        if self.precision == "fp16" and self.engine.range_status(reset=True):
            raise _lib.SdgError("a value left the fp16 range (|v| > 65504) during a recording pass since the last check: "
                                "its logits and the statistics folded from them are invalid; re-run with "
                                "range_check='sync' or precision='bf16'")

    def _range(self):
        return (0, self.n) if self.shard is None else self.shard

    def load_weights(self, netD, precision=None):
        sd = netD.state_dict() if hasattr(netD, "state_dict") else netD
        kind = engine.detect_arch(sd)
        precision = precision or self.precision
        if kind == "stylegan2":
            self.engine.load_stylegan2(sd, precision, batch=self.batch)
        else:
            self.engine.load(sd, precision, self.inplace_relu)

    def record(self, netD, step=None, out: torch.Tensor = None, range_check="sync") -> torch.Tensor:
        """One recording pass over the resident dataset (this rank's shard) -> float32 [N] on the device.
        ``range_check``: see :meth:`_guarded` ("sync" costs one 4-byte device->host read per pass)."""
        if self.dataset is None:
            raise _lib.SdgError("LogitRecorder.record needs a ResidentDataset")
        snap = torch.zeros(self.n, dtype=torch.float32, device=self.device) if out is None else out

        def run(precision):
            self.load_weights(netD, precision)
            lo, hi = self._range()
            if self.engine.arch == "stylegan2":
                hi = lo + (hi - lo) // self.batch * self.batch      # drop_last: the tail keeps its 0.0 like the reference
            if hi > lo:
                self.engine.forward(self.dataset.data[lo:hi], out=snap[lo:hi])
            return snap

        self._guarded(netD, run, range_check)
        if step is not None:
            self.observe(step, snap)
        return snap

    def observe(self, step, snap: torch.Tensor):
        """Book-keep one finished snapshot: keep it (for the pickle) and fold it into the running stats."""
        if self.keep_snapshots:
            self.snapshots[step] = snap
        w = self.stats_window
        if w is None or (w[0] <= step < w[1]):
            lo, hi = self._range()
            if self.stats is None:
                self.stats = engine.RunningStats(hi - lo, self.device)
            self.stats.update(snap[lo:hi])

    def record_from_host(self, netD, host_u8: torch.Tensor, step=None, chunk: int = 12544, first_chunk: int = 2048,
                         range_check="sync") -> torch.Tensor:
        """Recording pass over a dataset that lives in (pinned) HOST memory: uint8 NHWC chunks are
        copied on a side stream into two staging buffers while the previous chunk is in the engine, so
        the H2D traffic (3 KB/sample for CIFAR shape) overlaps the forward.  The first chunk is small and its copy is
        issued before the weights are packed, so almost nothing of the transfer is exposed."""
        n = host_u8.shape[0]
        self.n = n
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stage = None
        shape = (min(chunk, n),) + tuple(host_u8.shape[1:])
        if self._stage is None or self._stage[0].shape != shape:
            self._stage = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self._ev_copied = [torch.cuda.Event() for _ in range(2)]
            self._ev_used = [torch.cuda.Event() for _ in range(2)]
            for e in self._ev_used:
                e.record(main)
        bounds = [0]
        if 0 < first_chunk < min(chunk, n):
            bounds.append(first_chunk)
        while bounds[-1] < n:
            bounds.append(min(n, bounds[-1] + chunk))

        def issue_copy(i):
            b, s0, s1 = i & 1, bounds[i], bounds[i + 1]
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._ev_used[b])          # staging buffer free again
                self._stage[b][:s1 - s0].copy_(host_u8[s0:s1], non_blocking=True)
                self._ev_copied[b].record(self._copy_stream)

        snap = torch.zeros(n, dtype=torch.float32, device=self.device)

        def run(precision):
            issue_copy(0)                                                   # overlaps sigma + weight packing
            self.load_weights(netD, precision)
            for i in range(len(bounds) - 1):
                b, s0, s1 = i & 1, bounds[i], bounds[i + 1]
                if i + 1 < len(bounds) - 1:
                    issue_copy(i + 1)                                       # next chunk streams in during this forward
                main.wait_event(self._ev_copied[b])
                self.engine.forward(self._stage[b][:s1 - s0], out=snap[s0:s1])
                self._ev_used[b].record(main)
            return snap

        self._guarded(netD, run, range_check)
        if step is not None:
            self.observe(step, snap)
        return snap

    def record_from_loader(self, netD, dataloader, n=None, range_check="sync", group: int = 16) -> torch.Tensor:
        """Generic feeding path with the reference's item contract: batches ``(data, target, weight,
        index)`` (predefined.py:22-24) or ``(img, idx)`` come from a DataLoader, the forward still runs in
        the CUDA engine, logits are scattered by dataset index (trainer.py:148-154).  Up to ``group`` loader batches are
        concatenated per engine call (samples are independent in every discriminator this path serves except StyleGAN2,
        which goes through ``diagan_b200.distributed.get_logit``), so a batch-64 loader costs one launch sequence per 1024
        samples instead of per 64."""
        n = len(dataloader.dataset) if n is None else n
        snap = torch.zeros(n, dtype=torch.float32, device=self.device)

        def run(precision):
            self.load_weights(netD, precision)
            per_call = 1 if self.engine.arch == "stylegan2" else max(1, int(group))
            xs, ids = [], []

            def flush():
                if xs:
                    x = torch.cat(xs) if len(xs) > 1 else xs[0]
                    i = torch.cat(ids) if len(ids) > 1 else ids[0]
                    snap[i] = self.engine.forward(x)
                    xs.clear(); ids.clear()

            for item in dataloader:
                data, idx = item[0], item[-1]
                xs.append(data.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous())
                ids.append(idx.to(self.device, non_blocking=True))
                if len(xs) >= per_call:
                    flush()
            flush()
            return snap

        return self._guarded(netD, run, range_check)


# ---- reference-named methods (usable standalone or grafted onto the reference class) -------------

def _unsupported_reason(sd, eval_mode):
    """None if the CUDA engine reproduces ``netD(x)`` for this state_dict / mode, else why it does not."""
    try:
        kind = engine.detect_arch(sd)
    except _lib.SdgError as e:
        return str(e)
    if kind == "dcgan32":
        nc = int(sd["conv.0.weight"].shape[1])
        if nc != 3:
            return (f"the DCGAN discriminator was built with nc * num_pack = {nc} input channels (mnist.py:156-163); the "
                    "engine implements nc = 3, num_pack = 1")
        if not eval_mode:
            return ("train-mode logits of the DCGAN discriminator (Dropout + batch-statistics BatchNorm, "
                    "save_eval_logits=False) are stochastic and batch dependent; only eval_mode=True has a per-sample "
                    "definition")
    elif kind != "stylegan2" and not eval_mode:
        return ("train-mode logits of a spectral-norm discriminator advance sn_u / sn_sigma once per batch "
                "(trainer.py:145-146 leaves netD in train mode); the engine computes sigma once per pass with eval "
                "semantics")
    return None


def _get_logit(self, netD, eval_mode=False):
    """``LogTrainer._get_logit`` (trainer.py:142-156): float64 [N] by dataset index.

    Uses ``self.recorder`` (a :class:`LogitRecorder` with a resident dataset) when present, otherwise
    feeds the CUDA engine from ``self.dataloader``.

    Discriminators / modes the engine does not reproduce (``_unsupported_reason``: train-mode DCGAN, nc != 3, an
    architecture ``detect_arch`` rejects, train-mode spectral norm) are handed to the reference's own method when
    ``diagan_b200.patch.install()`` kept it as ``_get_logit_ref`` -- with a warning, never silently; without it they
    raise, except train-mode SNGAN, which warns and returns eval-semantics logits (identical up to the sn_u drift).
    A missing ``libsdg.so`` or a non-B200 device always raises: the fallback is for unsupported MODELS only."""
    sd = netD.state_dict()
    why = _unsupported_reason(sd, eval_mode)
    if why is not None:
        ref = getattr(type(self), "_get_logit_ref", None)
        if ref is not None:
            warnings.warn(f"diagan_b200: {why}; this pass runs through the reference's own LogTrainer._get_logit",
                          RuntimeWarning, stacklevel=2)
            return ref(self, netD=netD, eval_mode=eval_mode)
        if "spectral-norm" not in why:
            raise _lib.SdgError(why)
        warnings.warn(f"diagan_b200: {why}; returning eval-semantics logits", RuntimeWarning, stacklevel=2)
    rec = getattr(self, "recorder", None)
    if rec is None:
        rec = LogitRecorder(None, getattr(self, "device", None))
        self.recorder = rec
    if rec.dataset is not None:
        snap = rec.record(netD, range_check="sync")
    else:
        snap = rec.record_from_loader(netD, self.dataloader)
    if hasattr(netD, "train"):
        netD.train()                                        # trainer.py:155
    return snap.double().cpu().numpy()                      # fp32 values widened, like trainer.py:144,154


def _host_f64(v):
    return v.double().cpu().numpy() if torch.is_tensor(v) else np.asarray(v, dtype=np.float64)


def _atomic_write(path: Path, write_fn):
    """Write to ``<path>.tmp`` then rename: a job killed mid-save never leaves a truncated pickle behind (the
    reference overwrites in place, trainer.py:138-140)."""
    tmp = Path(str(path) + ".tmp")
    with open(tmp, "wb") as f:
        write_fn(f)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def _save_logit(self, logits_dict):
    """``LogTrainer._save_logit`` (trainer.py:138-140): one pickle per name, ``{step: float64[N]}`` -- byte-for-byte what
    ``pickle.dump`` of the reference's dict of float64 arrays produces, so train_mimicry_phase2.py:87-92 loads it unchanged.
    With ``self.save_f32_sidecar`` a compact ``logits_<name>_f32.npz`` (steps int64 [T], logits float32 [T,N]: the values
    ARE fp32, trainer.py:154 only widens them) is written beside it; :func:`load_logits` reads either."""
    for name, logits in logits_dict.items():
        host = {k: _host_f64(v) for k, v in logits.items()}
        _atomic_write(Path(self.output_path) / f'logits_{name}.pkl', lambda f: pickle.dump(host, f))
        if getattr(self, "save_f32_sidecar", False) and len(host):
            steps = np.array(list(host.keys()), dtype=np.int64)
            arr = np.stack([host[k].astype(np.float32) for k in host])
            _atomic_write(Path(self.output_path) / f'logits_{name}_f32.npz', lambda f: np.savez(f, steps=steps, logits=arr))


def load_logits(path) -> dict:
    """``{step: array [N]}`` from ``logits_<name>.pkl`` (float64, the reference's format) or the ``_f32.npz`` side-file
    (float32 rows; insertion order = recording order either way, which is what calculate_scores' window filter walks)."""
    path = Path(path)
    if path.suffix == ".npz":
        z = np.load(path)
        return {int(k): z["logits"][i] for i, k in enumerate(z["steps"])}
    with open(path, "rb") as f:
        return pickle.load(f)


def _restore_logits(self):
    """Resume: merge the snapshots already on disk into ``logit_results`` so that the next ``_save_logit`` extends the
    pickle instead of replacing it.  (The reference starts every run with an empty ``logit_results``, trainer.py:222, and
    overwrites the file: a phase-1 job restarted inside the recording window silently loses the earlier snapshots.)
    Steps recorded again after the restart replace their old rows; order stays ascending in step."""
    out = Path(self.output_path)
    restored = 0
    for f in sorted(out.glob("logits_*.pkl")):
        name = f.stem[len("logits_"):]
        try:
            old = load_logits(f)
        except Exception as e:                              # unreadable pickle: keep going, the new run rewrites it
            print(f"WARNING: could not restore {f}: {e}")
            continue
        merged = dict(old)
        merged.update(self.logit_results.get(name, {}))
        self.logit_results[name] = dict(sorted(merged.items()))
        restored += len(old)
    return restored


class LogTrainer:
    """Recording-side stand-in for the reference ``LogTrainer``: same constructor keywords for everything
    that concerns logit recording (trainer.py:16-49) and the same trigger (trainer.py:328-351), driven by
    ``on_step(global_step)`` from whatever loop trains the GAN."""

    _get_logit = _get_logit
    _save_logit = _save_logit
    _restore_logits = _restore_logits

    def __init__(self, output_path, netD, dataloader=None, netD_drs=None, device=None, save_steps=5000,
                 logit_save_steps=500, save_logits=True, save_logit_after=0, stop_save_logit_after=100000,
                 save_eval_logits=True, recorder: LogitRecorder = None, resume_logits=True, save_f32_sidecar=False,
                 **unused):
        self.output_path = Path(output_path)
        self.netD, self.netD_drs = netD, netD_drs
        self.train_drs = netD_drs is not None               # trainer.py:86
        self.dataloader = dataloader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.save_steps = save_steps
        self.logit_save_steps = logit_save_steps
        self.save_logits = save_logits
        self.save_logit_after = save_logit_after
        self.stop_save_logit_after = stop_save_logit_after
        self.save_eval_logits = save_eval_logits
        self.recorder = recorder
        self.logit_results = defaultdict(dict)              # trainer.py:222
        self.save_f32_sidecar = save_f32_sidecar
        if resume_logits and self.save_logits and self.output_path.is_dir():
            self._restore_logits()

    def should_record(self, global_step) -> bool:
        return bool(self.save_logits and global_step % self.logit_save_steps == 0
                    and self.save_logit_after <= global_step <= self.stop_save_logit_after)

    def on_step(self, global_step):
        """Body of trainer.py:328-346 for one finished optimisation step."""
        if self.should_record(global_step):
            netD, name = (self.netD_drs, 'netD_drs') if self.train_drs else (self.netD, 'netD')
            mode = 'eval' if self.save_eval_logits else 'train'
            self.logit_results[f'{name}_{mode}'][global_step] = self._get_logit(netD=netD, eval_mode=mode == 'eval')
        if global_step % self.save_steps == 0 and self.save_logits and global_step >= self.save_logit_after:
            self._save_logit(self.logit_results)
