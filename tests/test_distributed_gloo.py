"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, the single all-gather of
finished shards, the MIN all-reduce of clip bounds -- diagan_b200.distributed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    for p in (ROOT, os.path.join(ROOT, "self-diagnosing-gan_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diagan_b200 import distributed as D
    from oracle import scores as so

    rng = np.random.RandomState(5)
    mean = rng.normal(0.5, 1.0, n)
    var = rng.uniform(0.0, 2.0, n)
    lo, hi = D.shard_range(n)
    assert (lo, hi) == D.shard_range(n, rank, world)

    # (1) one all-gather of ragged shards reproduces the full vector on every rank
    local = torch.from_numpy(mean[lo:hi].copy())
    full = D.all_gather_shards(local, n)
    assert full.shape[0] == n and np.array_equal(full.numpy(), mean)

    # (2) sharded score = local floor, MIN all-reduce of the bound, local clip, all-gather
    t = so.conf_values()[2]
    s_local = np.clip(mean[lo:hi] + t * np.sqrt(var[lo:hi]), a_min=so.FLOOR, a_max=None)
    m = torch.tensor([s_local.min() if hi > lo else np.inf], dtype=torch.float64)
    D.all_reduce_min_(m)
    s_local = np.clip(s_local, None, m.item() * so.RATIO)
    got = D.all_gather_shards(torch.from_numpy(s_local), n).numpy()
    want = so.score_from_moments(mean, var, t)
    assert np.array_equal(got, want)

    # (2b) batch-aligned shards (StyleGAN2 minibatch-stddev groups stay inside one rank): boundaries are multiples of 8,
    # the ranges still tile [0, n) and the same single all-gather reassembles the vector
    blo, bhi = D.shard_range(n, multiple=8)
    assert (blo % 8 == 0 or blo == n) and (bhi % 8 == 0 or bhi == n)
    spans = [D.shard_range(n, r, world, 8) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert np.array_equal(D.all_gather_shards(torch.from_numpy(mean[blo:bhi].copy()), n, multiple=8).numpy(), mean)

    # (2c) get_logit's exchange: ranks hold different numbers of (index, value) pairs in arbitrary order (a shuffled
    # DistributedSampler shard); one padded all-gather per tensor reassembles the dataset-indexed vector on every rank
    perm = np.random.RandomState(9).permutation(n)
    mine = perm[: (2 * n) // 3] if rank == 0 else perm[(2 * n) // 3:]
    full = D.gather_indexed(torch.from_numpy(mine.copy()), torch.from_numpy(mean[mine].astype(np.float32)), n)
    assert full.dtype == np.float64 and np.array_equal(full, mean.astype(np.float32).astype(np.float64))

    # (3) reference-contract concat_all_gather (train_ffhq.py:150-161)
    idx = torch.arange(rank * 4, rank * 4 + 4)
    assert D.concat_all_gather(idx).tolist() == list(range(world * 4))

    mx = torch.tensor([float(rank)])
    assert D.all_reduce_max_(mx).item() == world - 1
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4])
def test_world2_gloo(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
