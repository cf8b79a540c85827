"""Two-rank NCCL run of the sharded diagnosis path on real GPUs (skipped on boxes with a single GPU; run with
``gpurun --gpus 2 -- python -m pytest tests/test_multigpu_nccl.py -m gpu``).  The host-side index arithmetic is covered on CPU
by tests/test_distributed_gloo.py; this file checks that the sharded pass + score exchange over NCCL reproduces the
single-GPU result bit for bit (samples are independent, so sharding must not change a single logit)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "self-diagnosing-gan_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from diagan_b200 import distributed as D
    from diagan_b200 import engine, synthetic
    from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset

    # ---- SNGAN-32: three sharded passes -> running stats -> sharded score == the same on one GPU ----
    n = 3001                                                        # ragged: shards of 1501 and 1500
    x = synthetic.uniform_images_u8(n, 32, seed=1).to(dev)
    base = synthetic.sngan_state_dict(32, seed=1)
    rec = LogitRecorder(ResidentDataset(x), dev, keep_snapshots=False)
    solo = LogitRecorder(ResidentDataset(x), dev, keep_snapshots=False)
    for step in (0, 100, 200):
        sd = synthetic.perturb_(base, step, 2e-2)
        full = D.get_logit(rec, sd, step=None)
        lo, hi = D.shard_range(n)
        rec.observe(step, full)                                     # stats of this rank's shard only
        ref = solo.record(sd, step=step)
        assert torch.equal(full, ref), "sharded logits differ from the single-GPU pass"
    t = engine.conf_from_key("ldr_conf_0.3_ratio_50")
    got = D.sharded_score(rec, t, eps=1e-6)
    want = solo.stats.score(t, eps=1e-6)
    assert torch.equal(got, want), "sharded score differs"
    assert torch.equal(engine.top_indices(got, 100, True), engine.top_indices(want, 100, True))
    # one-collective form (local minimum rides in the all-gather payload): the same bits
    assert torch.equal(D.sharded_score_fused(rec.stats, t, n, eps=1e-6), want)

    # ---- an EMPTY shard (world > N): rank 1 owns nothing, must neither raise nor hang the collective ----
    x1 = x[:1].contiguous()
    r1 = LogitRecorder(ResidentDataset(x1), dev, keep_snapshots=False)
    full1 = D.get_logit(r1, base)
    assert full1.shape == (1,) and torch.equal(full1, LogitRecorder(ResidentDataset(x1), dev).record(base))

    # ---- StyleGAN2 (size 32, loader batch 4): shards are whole batches, logits identical to one GPU ----
    # n2 / world is NOT a multiple of the batch: shards of 24 and 22, the ragged tail (2) is dropped by rank 1
    n2 = 46
    x2 = synthetic.uniform_images_u8(n2, 32, seed=2).to(dev)
    p2 = synthetic.stylegan2_state_dict(32, seed=3)
    r2 = LogitRecorder(ResidentDataset(x2), dev, batch=4, keep_snapshots=False)
    s2 = LogitRecorder(ResidentDataset(x2), dev, batch=4, keep_snapshots=False)
    lo2, hi2 = D.shard_range(n2, multiple=4)
    assert lo2 % 4 == 0 and (lo2, hi2) == ((0, 24) if rank == 0 else (24, 46))
    for step in (0, 100, 200):
        pp = synthetic.perturb_(p2, step, 2e-2)
        assert torch.equal(D.get_logit(r2, pp, step=step), s2.record(pp, step=step))
    assert r2.shard_multiple == 4
    t3 = engine.conf_from_key("ldr_conf_3.0_ratio_50")
    assert torch.equal(D.sharded_score(r2, t3, eps=1e-6), s2.stats.score(t3, eps=1e-6))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_matches_single_gpu(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_drives_two_devices():
    """one-time initialisation (opt-in shared memory sizes, driver entry point) is tracked per device and every
    wrapper guards the current device, so a single process can run engines on cuda:0 and cuda:1 -- without touching
    torch.cuda.set_device -- and gets the same bits on both."""
    for p in (ROOT, os.path.join(ROOT, "self-diagnosing-gan_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from diagan_b200 import engine, synthetic
    from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset
    n = 515
    x = synthetic.uniform_images_u8(n, 32, seed=8)
    sd = synthetic.sngan_state_dict(32, seed=8)
    outs = []
    for d in (1, 0, 1):                                       # device 1 first: nothing may silently default to device 0
        dev = torch.device("cuda", d)
        rec = LogitRecorder(ResidentDataset(x.to(dev)), dev, keep_snapshots=False)
        for step in (0, 100, 200):
            snap = rec.record(synthetic.perturb_(sd, step, 2e-2), step=step)
        assert snap.device == dev and rec.stats.state.device == dev
        w = rec.stats.score(engine.conf_from_key("ldr_conf_0.3_ratio_50"), eps=1e-6)
        top = engine.top_indices(w, 50, True)
        assert w.device == dev and top.device == dev
        outs.append((snap.cpu(), w.cpu(), top.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)
