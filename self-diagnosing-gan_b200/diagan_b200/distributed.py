"""Multi-GPU side of the diagnosis path: shard by sample index, exchange only finished vectors.

Mirrors the module-level functions of ``stylegan2/train_ffhq.py`` that the reference uses for its
distributed logit pass -- ``get_logit`` (:128-143), ``concat_all_gather`` (:150-161), ``save_logit``
(:145-147) -- and the helpers of ``stylegan2/distributed.py`` it relies on (``get_rank``,
``get_world_size``, :7-16,34-41).

Design: rank r owns the contiguous index range ``[r*ceil(N/W), min(N,(r+1)*ceil(N/W)))``,
keeps that slice of the dataset resident, and scores it with replicated weights.  The reference
all-gathers ``(idx, logit)`` twice per batch of 4; here each pass does ONE all-gather of the finished
shard, and the score stage adds ONE MIN all-reduce of the per-key clip bounds (``score.min()`` in
``clip_max_ratio``, plot.py:226-228, is the only cross-sample coupling).  Both are latency-bound NCCL
calls over NVLink (tens of KB); there is no data-path collective inside the discriminator forward.

The collectives work on whatever device the tensors live on, so the index arithmetic is testable with
the ``gloo`` backend on CPU (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

import pickle
from pathlib import Path

import torch
import torch.distributed as dist


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_size(n: int, world: int, multiple: int = 1) -> int:
    """ceil(n / world), rounded up to a multiple of ``multiple`` (StyleGAN2: the loader batch size, so that the
    minibatch-stddev groups of consecutive batches never straddle two ranks)."""
    s = (n + world - 1) // world
    return (s + multiple - 1) // multiple * multiple


def shard_range(n: int, rank: int = None, world: int = None, multiple: int = 1):
    """Contiguous dataset-index range owned by ``rank``; trailing ranks may be short or empty."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    s = shard_size(n, world, multiple)
    lo = min(n, rank * s)
    return lo, min(n, lo + s)


def all_gather_shards(local: torch.Tensor, n: int, multiple: int = 1) -> torch.Tensor:
    """local: this rank's finished slice ``[hi-lo, ...]`` -> the full ``[n, ...]`` vector on every rank.
    One collective; ragged tails are handled by padding each shard to the common shard size."""
    world = get_world_size()
    if world == 1:
        return local[:n]
    s = shard_size(n, world, multiple)
    tail = local.shape[1:]
    padded = local
    if local.shape[0] != s:
        padded = local.new_zeros((s,) + tuple(tail))
        padded[:local.shape[0]] = local
    out = local.new_empty((world * s,) + tuple(tail))
    dist.all_gather_into_tensor(out, padded.contiguous())
    return out[:n]


def all_reduce_min_(t: torch.Tensor) -> torch.Tensor:
    """In-place MIN all-reduce (the clip bound of clip_max_ratio across shards)."""
    if get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t


def all_reduce_max_(t: torch.Tensor) -> torch.Tensor:
    """In-place MAX all-reduce (the DRS running maximum when burn-in batches are spread over ranks)."""
    if get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


@torch.no_grad()
def concat_all_gather(tensor: torch.Tensor) -> torch.Tensor:
    """Same contract as train_ffhq.py:150-161: equal-shaped per-rank tensors concatenated on dim 0."""
    world = get_world_size()
    if world == 1:
        return tensor
    out = tensor.new_empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]))
    dist.all_gather_into_tensor(out, tensor.contiguous())
    return out


def engine_is_stylegan2(netD) -> bool:
    sd = netD.state_dict() if hasattr(netD, "state_dict") else netD
    return "final_conv.0.weight" in sd and "convs.0.0.weight" in sd


def get_logit_resident(recorder, netD, step=None) -> torch.Tensor:
    """Distributed recording pass over a resident dataset: every rank scores its contiguous index shard and ends up with
    the full vector (train_ffhq.py:128-143 semantics).  ``recorder`` is a :class:`diagan_b200.trainer.trainer.LogitRecorder`;
    returns float32 [N] on the device."""
    n = recorder.n
    # StyleGAN2: shard boundaries on whole loader batches (the recorder drops the ragged tail of the LAST shard only)
    mult = recorder.batch if engine_is_stylegan2(netD) else 1
    lo, hi = shard_range(n, multiple=mult)
    recorder.shard = (lo, hi)
    recorder.shard_multiple = mult                 # sharded_score must all-gather with the same shard size
    snap = recorder.record(netD)
    full = all_gather_shards(snap[lo:hi], n, multiple=mult)
    if step is not None:
        recorder.observe(step, full)
    return full


def gather_indexed(idx: torch.Tensor, values: torch.Tensor, n: int):
    """This rank's ``(dataset index, value)`` pairs -> ``np.float64 [n]`` holding every rank's values at their indices
    (zeros elsewhere, train_ffhq.py:130).  One all-gather per tensor for the whole pass; ranks may hold different counts:
    shorter ones are padded with index -1, which is dropped after the exchange."""
    import numpy as np
    if get_world_size() > 1:
        cnt = torch.tensor([idx.numel()], dtype=torch.int64, device=idx.device)
        m = int(concat_all_gather(cnt).max().item())
        pad = m - idx.numel()
        if pad:
            idx = torch.cat([idx, idx.new_full((pad,), -1)])
            values = torch.cat([values, values.new_zeros(pad)])
        idx, values = concat_all_gather(idx), concat_all_gather(values)
        keep = idx >= 0
        idx, values = idx[keep], values[keep]
    out = np.zeros(n)
    out[idx.cpu().numpy()] = values.double().cpu().numpy()
    return out


_loader_recorders = {}


def get_logit(dataloader, netD, device=None, step=None):
    """Same call as ``stylegan2/train_ffhq.py:128-143`` -- ``get_logit(dataloader=loader, netD=discriminator, device=device)``
    -- returning ``np.float64 [len(dataloader.dataset)]`` indexed by dataset index on every rank, ``netD`` left in train mode.

    Each rank walks ITS loader (the reference's DistributedSampler shard, items ``(img, idx)``), the forward runs in the CUDA
    engine with every loader batch treated as one minibatch-stddev batch exactly like ``netD(real_data)`` does, and the
    ``(idx, logit)`` pairs are exchanged with ONE pair of all-gathers at the end of the pass (the reference issues two
    blocking all-gathers per batch of 4).  A :class:`LogitRecorder` as first argument selects the resident-dataset pass."""
    from .trainer.trainer import LogitRecorder
    if isinstance(dataloader, LogitRecorder):
        return get_logit_resident(dataloader, netD, step)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rec = _loader_recorders.get(device)
    if rec is None:
        rec = _loader_recorders[device] = LogitRecorder(None, device)
    n = len(dataloader.dataset)
    sg2 = engine_is_stylegan2(netD)

    def run(precision):
        idx_parts, logit_parts = [], []
        loaded = False
        pending, pending_b = [], 0                         # loader batches of equal size waiting for one engine call

        def flush():
            nonlocal pending
            if not pending:
                return
            if sg2:
                rec.engine.set_batch(pending_b)            # consecutive groups of pending_b samples = the loader's batches
            logit_parts.append(rec.engine.forward(torch.cat(pending) if len(pending) > 1 else pending[0]))
            pending = []

        for item in dataloader:
            data, idx = item[0], item[-1]
            if not loaded:                                 # weights packed once per pass, after the batch size is known
                rec.batch = int(data.shape[0]) if sg2 else rec.batch
                rec.load_weights(netD, precision)
                loaded = True
            x = data.to(device=device, dtype=torch.float32).contiguous()
            if pending and (x.shape[0] != pending_b or len(pending) * pending_b >= 256):
                flush()                                    # a short last batch is its own stddev batch (drop_last=False)
            pending.append(x)
            pending_b = int(x.shape[0])
            idx_parts.append(idx.to(device))
        flush()
        idx_all = torch.cat(idx_parts) if idx_parts else torch.empty(0, dtype=torch.int64, device=device)
        logit_all = torch.cat(logit_parts) if logit_parts else torch.empty(0, dtype=torch.float32, device=device)
        return idx_all, logit_all

    idx_all, logit_all = rec._guarded(netD, run, "sync")   # fp16 range guard: an overflowing pass is re-run in bf16
    out = gather_indexed(idx_all, logit_all, n)
    if hasattr(netD, "train"):
        netD.train()                                       # train_ffhq.py:142
    return out


def sharded_score(recorder, conf: float, eps: float = 0.0) -> torch.Tensor:
    """ldr_conf score of the whole dataset from per-rank running statistics: local floor+min, MIN
    all-reduce of the bound, local clip, one all-gather of the float64 shard -> float64 [N]."""
    local = recorder.stats.score(conf, eps=eps, min_reduce=all_reduce_min_)
    return all_gather_shards(local, recorder.n, multiple=getattr(recorder, "shard_multiple", 1))


def sharded_score_fused(stats, conf: float, n: int, eps: float = 0.0, multiple: int = 1, floor=None, ratio=None) -> torch.Tensor:
    """The same vector as :func:`sharded_score` with ONE collective per call: each rank all-gathers its floor-clipped shard
    with its local minimum appended, and the clip against the global minimum runs on the gathered buffer
    (``sdg_score_clip_gathered``).  ``stats`` is this rank's :class:`diagan_b200.engine.RunningStats`."""
    import ctypes as C

    import numpy as np

    from . import _lib, engine
    lib = _lib.load()
    floor = engine.FLOOR if floor is None else floor
    ratio = engine.RATIO if ratio is None else ratio
    world = get_world_size()
    s = shard_size(n, world, multiple)
    dev = stats.state.device
    payload = torch.zeros(s + 1, dtype=torch.float64, device=dev)
    cs = np.asarray([conf], dtype=np.float64)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(lib.sdg_score_floor_min(_lib.ptr(stats.mean), _lib.ptr(stats.m2), stats.n,
                                           cs.ctypes.data_as(C.POINTER(C.c_double)), 1, floor, float(stats.count - 1),
                                           _lib.ptr(payload), _lib.ptr(payload[s:]), st), "sdg_score_floor_min")
        gathered = payload
        if world > 1:
            gathered = torch.empty(world * (s + 1), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gathered, payload)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(lib.sdg_score_clip_gathered(_lib.ptr(gathered), world, s, n, ratio, eps, _lib.ptr(out), st),
                   "sdg_score_clip_gathered")
    return out


def save_logit(logits_dict, output_path):
    """train_ffhq.py:145-147 (rank 0 calls it, :321-323)."""
    for name, logits in logits_dict.items():
        host = {k: (v.double().cpu().numpy() if torch.is_tensor(v) else v) for k, v in logits.items()}
        with open(Path(output_path) / f'logits_{name}.pkl', 'wb') as f:
            pickle.dump(host, f)
