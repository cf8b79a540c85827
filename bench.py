"""bench.py -- headline benchmark of the per-sample diagnosis path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sngan32|sngan64|stylegan2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full recording pass of the hot path over the synthetic CIFAR-10-shaped training set with the SNGAN-32
discriminator (BASELINE.json configs[1]):
  weight re-pack (sigma once per pass) -> D forward for every sample -> Welford statistics update ->
  ldr_conf_0.3_ratio_50 score (floor, global MIN, clip) -> sampler weights -> top-100 indices.

Scaling (BASELINE.json metric: "SNGAN-32, 50k at 1/2/4/8 B200"; north_star: "the dataset is sharded by sample index across
the 8 GPUs"): STRONG -- ONE dataset of 50 000 samples is sharded by contiguous index range over the N ranks
(diagan_b200.distributed.shard_range), every rank scores its shard and the finished per-sample score vector is
all-gathered.  ``value`` = 50 000 x K / max-over-ranks time.  For N > 1 the line also carries ``weak``: the same step with
50 000 samples PER GPU (the round-1 headline), so both curves come from one run.

Prints ONE JSON line (rank 0).  ``value`` is timed with the dataset shard resident in HBM; ``e2e`` is the same metric
through ``LogitRecorder.record_from_host`` with the uint8 shard in pinned host memory (H2D inside the timed region) and
the score vector + top indices read back to pinned host memory every step (asynchronous copy, the host waits for it one step
later: the consumer of step i's scores runs beside step i + 1).  At N = 1 the line also carries ``cpu_baseline`` (the
oracle port on the host cores) and ``gpu_eager_baseline`` (the oracle network in PyTorch eager on the same B200: the
reference's pass as users run it today, and a best-case library run).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "self-diagnosing-gan_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

UNIT = "samples/s"
T_WINDOW = 50

# The headline line is configs[1] (default).  --workload selects the same measurement for the other discriminator
# configurations of BASELINE.json (parity-test cases by the contract, measured here for completeness).
#   n_total: the ONE dataset that is sharded (strong scaling); n_weak: samples per GPU of the weak-scaling sub-measurement
#   dom_exec_useful: fraction of the dominant kernel's executed MACs that are not structural zeros
#   traffic: (dram bytes read + written per SAMPLE by the dominant kernel, the committed ncu --set full summary they are from)
WORKLOADS = {
    "sngan32": dict(arch="sngan", size=32, n_total=50_000, n_weak=50_000, key="ldr_conf_0.3_ratio_50", flop=2 * 272_072_832,
                    metric="per-sample D logits + LDR scores per second (SNGAN-32, 50k CIFAR-10-shape samples)",
                    desc="configs[1]: SNGAN-32 recording pass (weights re-packed per pass) + Welford stats + "
                         "ldr_conf_0.3_ratio_50 weights + top-100 over ONE 50k x 3x32x32 uint8 dataset",
                    kernel="b1_fused_kernel: the whole of block 1 in one launch -- c1 (3x3 3->128 from the bytes) -> ReLU -> c2 (3x3 128->128 "
                           "@32x32 + avg-pool, run as the algebraically equal 4x4 stride-2 conv) + image shortcut -> ReLU; 56.0% of the "
                           "reference FLOPs; CTA pairs (tcgen05.mma.cta_group::2, M = 256 pixels x N = 128 channels), relu(c1(x)) built "
                           "and consumed in shared memory as a no-swizzle UMMA operand, only the weights stream",
                    dom_ref_flop=2.0 * 9 * 128 * 128 * 1024 + 2.0 * 27 * 128 * 1024, dom_exec_useful=1.0, cpu_sample=16384,
                    traffic=((39.038208e6 + 765.836544e6) / 12504.0, "profiles/r2i_ncu_full_b1fused_summary.txt (12 504 samples: 39.0 MB "
                             "read + 765.8 MB written to DRAM during the launch = 64.4 KB/sample; algorithmic 3 KiB in + 64 KiB out per "
                             "sample = 68.6 KB -- the tail of the output is still dirty in L2 when the launch ends; the two kernels this "
                             "one replaces moved 585 KB/sample)"),
                    eager=dict(ref_batch=64, ref_n=50_000, best_batch=4096, best_n=50_000)),
    "sngan64": dict(arch="sngan", size=64, n_total=202_599, n_weak=25_325, key="ldr_conf_5.0_ratio_50", flop=2 * 644_809_728,
                    metric="per-sample D logits + LDR scores per second (SNGAN-64, CelebA shape, 202 599 samples)",
                    desc="configs[2]: SNGAN-64 recording pass + Welford stats + ldr_conf_5.0_ratio_50 weights + top-100 over ONE "
                         "202 599 x 3x64x64 uint8 dataset",
                    kernel="b1_fused_kernel<64>: the whole of block 1 in one launch per 32 x 32 image quadrant -- c1 (3x3 3->64 from "
                           "the bytes, halo recomputed) -> ReLU -> c2 (3x3 64->64 @64x64 + avg-pool, run as the algebraically equal 4x4 "
                           "stride-2 conv) + image shortcut -> ReLU; 24.5% of the reference FLOPs; CTA pairs (tcgen05.mma.cta_group::2, "
                           "M = 256 pixels x N = 64 channels, no structural zeros), relu(c1(x)) built and consumed in shared memory; "
                           "paced by the CUDA-core warps that build that tile and the shared-memory operand fetch of N = 64 MMAs, not by the tensor pipe (profiles/r3_b1fused64.md, r4_b1fused64.md)",
                    dom_ref_flop=2.0 * 9 * 64 * 64 * 4096 + 2.0 * 27 * 64 * 4096, dom_exec_useful=1.0, cpu_sample=2048,
                    traffic=((101.1e6 + 1020.7e6) / 8192.0, "profiles/r4_ncu_full_b1fused64_summary.txt (8192 samples: 101.1 MB read + "
                             "1020.7 MB written to DRAM during the launch = 137 KB/sample; algorithmic 12 KiB in + 128 KiB out per "
                             "sample = 143 KB; the two kernels this one replaces moved 1.19 MB/sample)"),
                    eager=dict(ref_batch=64, ref_n=16_384, best_batch=1024, best_n=16_384)),
    "stylegan2": dict(arch="stylegan2", size=256, n_total=2048, n_weak=2048, key="ldr_conf_3.0_ratio_50", flop=None,
                      metric="per-sample D logits + LDR scores per second (StyleGAN2-256 discriminator, FFHQ shape, bounded 2048 "
                             "samples)",
                      desc="configs[4]: StyleGAN2-256 recording pass (loader batch 4) + Welford stats + ldr_conf_3.0_ratio_50 "
                           "weights + top-100 over 2048 x 3x256x256 uint8 (bounded sample of the 70k pass)",
                      kernel="conv_swap_kernel ResBlock 1 conv1 (3x3 128->128 @256x256 + FusedLeakyReLU; 20.8% of the FLOPs)",
                      dom_ref_flop=2.0 * 9 * 128 * 128 * 65536, dom_exec_useful=1.0, cpu_sample=8,
                      traffic=((1.900776e9 + 1.841034e9) / 112.0, "profiles/r2c_ncu_full_sg2_conv1_summary.txt (112 samples: 1.901 GB read + "
                               "1.841 GB written; algorithmic 16 MiB in + 16 MiB out per sample = 33.55 MB)"),
                      eager=dict(ref_batch=4, ref_n=64, best_batch=16, best_n=64)),
}


def _workload(name):
    w = dict(WORKLOADS[name])
    if w["flop"] is None:
        from diagan_b200 import synthetic
        w["flop"] = synthetic.stylegan2_flops(w["size"])
    prof = os.path.join(ROOT, "profiles", "r2_dominant_kernel_traffic.json")       # refreshed per round by tools/ncu_traffic.py
    if os.path.exists(prof):
        d = json.load(open(prof)).get(name)
        if d:
            w["traffic"] = (float(d["dram_bytes_per_sample"]), d["source"])
    return w


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent (NVML) samples of SM clock and throttle reasons during the timed region.  NVML is initialised
    in the constructor (before the timed region starts) so that the thread samples from its first millisecond."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.nv = self.handle = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.handle = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception as e:          # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        nv, h = self.nv, self.handle
        if nv is None:
            return
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        try:
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path, restated in oracle/ (kind "port")
# ---------------------------------------------------------------------------------------------------
def cpu_step_rate(sample_n: int, threads: int, score_n: int, workload: str = "sngan32"):
    """Time the CPU path on a bounded sample: the torch fp32 oracle forward of the workload's discriminator (batch 64 --
    StyleGAN2: batch 4 --, eval, no_grad, tensor-slice batches) over ``sample_n`` samples, plus the reference-faithful
    calculate_scores on a full [50, score_n] window (it recomputes mean/std for each of the 99 keys).
    Returns samples/s of a whole ``score_n``-sample step extrapolated linearly from the sample."""
    from oracle import scores as so
    w = WORKLOADS[workload]
    torch.set_num_threads(threads)
    size = w["size"]
    x = torch.from_numpy(np.random.RandomState(1).randint(0, 256, (sample_n, size, size, 3)).astype(np.uint8))
    if w["arch"] == "sngan":
        from oracle import sngan as sngan_oracle
        params = sngan_oracle.init_params(size, seed=1)
        run = lambda xs: sngan_oracle.logits_pass(params, xs, size)
        run(x[:64])                                                  # warm-up
    else:
        from oracle import stylegan2 as sg2_oracle
        params = sg2_oracle.init_params(size, seed=1)
        run = lambda xs: sg2_oracle.logits_pass(params, xs, size, 4)
        run(x[:4])
    t0 = time.perf_counter()
    run(x)
    t_fwd = time.perf_counter() - t0
    rng = np.random.RandomState(0)
    logits = {s: rng.normal(1.0, 1.5, score_n).astype(np.float32).astype(np.float64) for s in range(T_WINDOW)}
    t0 = time.perf_counter()
    so.calculate_scores(logits, 0, T_WINDOW, faithful=True)
    t_score = time.perf_counter() - t0
    per_step = t_fwd * (score_n / sample_n) + t_score / T_WINDOW     # scoring happens once per window
    return score_n / per_step, {"forward_s_per_sample": t_fwd / sample_n, "score_s_per_window": t_score}


def _cpu_sample_note(w, sample):
    return (f"oracle torch fp32 forward on {sample} of {w['n_total']} samples per step (extrapolated linearly) + faithful "
            f"calculate_scores on the full [50,{w['n_total']}] window / 50")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    sample = w["cpu_sample"]                     # the same bounded sample as the in-line cpu_baseline of the other arm
    vals = []
    for _ in range(args.warmup):
        cpu_step_rate(max(4, sample // 64), threads, score_n=2000, workload=args.workload)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_step_rate(sample, threads, score_n=w["n_total"], workload=args.workload)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "n_total": w["n_total"], "score_key": w["key"], "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": _cpu_sample_note(w, sample)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# the library on trial: the oracle network in PyTorch eager on this GPU
# ---------------------------------------------------------------------------------------------------
def gpu_eager_baseline(w, dev, host_u8):
    """(i) "reference" -- LogTrainer._get_logit as users run it on a GPU today (trainer.py:142-156): float32 NCHW batches of
    64 copied from host memory, netD(x) in eval mode under no_grad with PyTorch's defaults (cuDNN TF32 convolutions, sigma
    recomputed by every forward), logits moved to the CPU per batch.  The DataLoader workers and the PIL transform of the
    reference are NOT included (tensor slices): this is an upper bound of the reference's GPU path.
    (ii) "best_case" -- the same network given every library advantage: W / sigma computed once per pass, channels-last,
    bf16 autocast, large batches straight from the resident uint8 dataset, no host synchronisation until the end."""
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    e = w["eager"]
    size = w["size"]
    if w["arch"] == "sngan":
        from oracle import sngan as O
        params = {k: v.to(dev) for k, v in O.init_params(size, seed=1).items()}
        fwd = lambda p, x: O.forward(p, x, size, True)
        pre = dict(params)
        pre["__wn__"] = O.normalised_weights(params, size)
    else:
        from oracle import stylegan2 as O
        params = {k: v.to(dev) for k, v in O.init_params(size, seed=1).items()}
        fwd = lambda p, x: O.forward(p, x, size)
        pre = params
    from oracle.sngan import normalise_u8
    out = {}
    # (i) reference semantics
    n, B = min(e["ref_n"], host_u8.shape[0]), e["ref_batch"]
    xf = normalise_u8(host_u8[:n]).contiguous().pin_memory()                     # what the DataLoader hands over
    res = np.zeros(n)

    def ref_pass(m):
        with torch.no_grad():
            for s in range(0, m, B):
                x = xf[s:s + B].to(dev)                                           # trainer.py:149
                res[s:s + B] = fwd(params, x).view(-1).detach().cpu().numpy()     # trainer.py:150-154
    ref_pass(min(n, 16 * B))
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    ref_pass(n)
    torch.cuda.synchronize(dev)
    t = time.perf_counter() - t0
    out["reference_semantics"] = {"value": n / t, "unit": UNIT, "samples": n, "batch": B,
                                  "how": "fp32 NCHW batches from pinned host memory, cuDNN TF32, sigma per forward, .cpu() per "
                                         "batch (trainer.py:142-156); no DataLoader / PIL cost included"}
    # (ii) best case
    n2, B2 = min(e["best_n"], host_u8.shape[0]), e["best_batch"]
    xd = host_u8[:n2].to(dev)
    logits = torch.empty(n2, device=dev)

    import contextlib

    def best_pass(autocast):
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
        with torch.no_grad(), ctx:
            for s in range(0, n2, B2):
                x = normalise_u8(xd[s:s + B2]).contiguous(memory_format=torch.channels_last)
                logits[s:s + B2] = fwd(pre, x).view(-1).float()

    best = None
    for autocast, label in ((True, "bf16 autocast"), (False, "fp32 storage, cuDNN TF32")):
        best_pass(autocast)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            best_pass(autocast)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 3
        if best is None or ms < best[0]:
            best = (ms, label)
    out["best_case"] = {"value": n2 / (best[0] / 1e3), "unit": UNIT, "samples": n2, "batch": B2,
                        "how": f"resident uint8 dataset, channels-last, {best[1]} (the faster of bf16 autocast and TF32), "
                               "W/sigma once per pass, no host syncs"}
    return out


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes as C

    import torch.distributed as dist
    from diagan_b200 import distributed as D
    from diagan_b200 import engine, synthetic
    from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    w = _workload(args.workload)
    sg2 = w["arch"] == "stylegan2"
    sd0 = synthetic.sngan_state_dict(w["size"], seed=1) if not sg2 else synthetic.stylegan2_state_dict(w["size"], seed=1)
    base = {k: v.to(dev) for k, v in sd0.items()}
    t_conf = engine.conf_from_key(w["key"])
    host_chunk = 12544 if w["size"] <= 32 else (8192 if w["size"] <= 64 else 256)
    lib = engine._lib.load()
    # the discriminator "trains" between passes: one perturbed weight set per step, generated BEFORE the timed region (the
    # optimiser step is not part of the recording path); every step still re-packs its weights (sigma + W/sigma) inside it
    n_sets = max(args.warmup, 3) + args.steps
    weight_sets = [synthetic.perturb_(base, 35000 + 100 * i, 1e-3, device=dev) for i in range(n_sets)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def leg(n_total, lo, hi, profile):
        """Resident + end-to-end timing of the step for a dataset of n_total samples of which this rank owns [lo, hi)."""
        n_local = hi - lo
        mult = 4 if sg2 else 1
        # synthetic shard: uint8 images of the workload's shape, seeded by the shard's first index; pinned host copy for e2e
        host = synthetic.uniform_images_u8(max(n_local, 1), w["size"], seed=1 + lo, pin=True)[:n_local]
        rec = LogitRecorder(ResidentDataset(host.to(dev)), dev, precision=args.precision, inplace_relu=True,
                            keep_snapshots=False, batch=4)
        snap = torch.zeros(n_local, dtype=torch.float32, device=dev)

        def finish():
            """stats -> score -> weights (+ global min / all-gather when sharded) -> top-100."""
            if world > 1:      # ONE collective: the local minimum rides in the all-gather payload
                full = D.sharded_score_fused(rec.stats, t_conf, n_total, eps=1e-6, multiple=mult)
            else:
                full = rec.stats.score(t_conf, eps=1e-6)
            top = engine.top_indices(full, 100, True)
            return full, top

        def step_resident(i):
            rec.record(weight_sets[i % n_sets], step=i, out=snap, range_check="deferred")
            return finish()

        # end-to-end leg: the step's result (score vector + top indices) is read back into pinned host buffers every step; the
        # copy is asynchronous and the host waits for it ONE STEP LATER (double-buffered), so the consumer of step i's scores
        # runs beside step i + 1 instead of stalling the GPU behind the host's launch work
        host_out = [(torch.empty(n_total, dtype=torch.float64).pin_memory(), torch.empty(100, dtype=torch.int64).pin_memory(),
                     torch.cuda.Event()) for _ in range(2)]

        def step_host(i):
            rec.record_from_host(weight_sets[i % n_sets], host, step=i, chunk=host_chunk, first_chunk=min(2048, host_chunk // 4),
                                 range_check="deferred")
            full, top = finish()
            hf, ht, ev = host_out[i & 1]
            hf.copy_(full, non_blocking=True)                  # D2H of the step's result
            ht.copy_(top, non_blocking=True)
            ev.record()
            host_out[(i + 1) & 1][2].synchronize()             # the previous step's result is on the host now
            return hf, ht

        def timed(step_fn, steps, warmup, prof=False):
            rec.stats = None
            sampler = ClockSampler(local_rank) if (prof and rank == 0) else None      # NVML init happens here, untimed
            for i in range(warmup):
                step_fn(i)
            barrier()
            if prof:
                lib.sdg_ctx_profile(rec.engine._h, 1)
            engine.launch_count(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if sampler:
                sampler.start()
            e0.record()
            for i in range(steps):
                step_fn(warmup + i)
            e1.record()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            launches = engine.launch_count()
            if sampler:
                sampler.stop_flag = True
                sampler.join(timeout=2)
            return ms.item(), launches, (sampler.result() if sampler else None)

        r = {"n_local": n_local, "host_bytes": int(host.numel())}
        r["ms"], r["launches"], r["clocks"] = timed(step_resident, args.steps, args.warmup, prof=profile)
        if profile:
            pm, pl, pf = C.c_double(), C.c_int64(), C.c_double()
            lib.sdg_ctx_profile_read(rec.engine._h, C.byref(pm), C.byref(pl), C.byref(pf))
            lib.sdg_ctx_profile(rec.engine._h, 0)
            r["prof"] = (pm.value, pl.value, pf.value)
        rec.check_range()                                      # fp16 range guard, deferred form: raises if any pass overflowed
        r["ms_e2e"], _, _ = timed(step_host, args.steps, max(3, args.warmup))
        rec.check_range()
        r["host"] = host
        return r

    n_total = w["n_total"]
    lo, hi = D.shard_range(n_total, rank, world, multiple=4 if sg2 else 1)
    strong = leg(n_total, lo, hi, profile=True)
    weak = None
    if world > 1:
        weak = leg(w["n_weak"] * world, rank * w["n_weak"], (rank + 1) * w["n_weak"], profile=False)

    if rank != 0:
        return
    peaks = _peaks()
    ms, ms_e2e = strong["ms"], strong["ms_e2e"]
    value = n_total * args.steps / (ms / 1e3)
    e2e_value = n_total * args.steps / (ms_e2e / 1e3)
    pm, pl, pf = strong["prof"]
    n_local = strong["n_local"]
    useful = w["dom_exec_useful"]
    dom_exec = (pf / 1e12) / (pm / 1e3) if pm > 0 else None                    # EXECUTED FLOPs / time, zeros included
    dom_tflops = dom_exec * useful if dom_exec else None                       # ... structural zeros excluded
    # the same launches in the reference formulation (conv3x3 at full resolution, then avg-pool): 2*9*Cin*Cout per pixel
    ref_flops = w["dom_ref_flop"] * n_local * args.steps                       # every sample of every timed step
    dom_ref_tflops = (ref_flops / 1e12) / (pm / 1e3) if pm > 0 else None
    samples_per_launch = n_local * args.steps / max(1, int(pl))
    line = {
        "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": w["desc"], "n_total": n_total, "samples_per_gpu": n_local, "score_key": w["key"],
                   "l2": f"inputs larger than L2 ({strong['host_bytes'] / 1e6:.0f} MB dataset shard per GPU, > 1 GB of "
                         "activations streamed per sweep of the pass); no explicit flush",
                   "parallelism": f"ONE dataset, contiguous sample-index shards x{world}; per step ONE collective: all-gather of the "
                                  "float64 score shard with the shard's minimum appended (clip bound = min over ranks)",
                   "range_guard": "fp16 range flag checked after the timed loops (deferred): no pass overflowed"},
        "clocks": strong["clocks"],
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": strong["host_bytes"] * world,
                "d2h_bytes_per_step": int((n_total * 8 + 100 * 8) * world), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(strong["launches"]),
        "whole_path_tflops": value * w["flop"] / 1e12,
        "roofline": {
            "bound": "tensor",
            "kernel": w["kernel"],
            "achieved": dom_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": (dom_tflops / peaks["bf16_sustained"]) if dom_tflops else None,
            "flops_counted": "USEFUL FLOPs executed by the tensor pipe: 2*M*N*K of the GEMM run (16 taps per pooled pixel in the "
                             "4x4 stride-2 form; SNGAN-32: plus the 27 real MACs per pixel and channel of c1, which the same launch "
                             "computes), structural zeros of the super-pixel packing and K padding excluded",
            "executed_tflops_incl_structural_zeros": dom_exec,
            "reference_formulation_tflops": dom_ref_tflops,
            "reference_formulation_note": "same launches counted as the reference computes them (conv3x3 at full resolution then "
                                          "avg_pool2d: 36/16 of the useful MACs); may exceed the peak because the fused form skips work",
            "peak_source": f"{peaks['source']} (sustained: kernel timed inside a long step)",
            "launches_timed": int(pl), "ms_per_launch": (pm / pl) if pl else None,
            "samples_per_launch": samples_per_launch,
            # DRAM traffic per launch = ncu dram__bytes_read.sum + dram__bytes_write.sum per sample of THIS kernel (from the
            # committed ncu --set full summary named in traffic_source) x the samples one launch of this run processed
            "traffic": (w["traffic"][0] * samples_per_launch) if w["traffic"] else None,
            "traffic_source": w["traffic"][1] if w["traffic"] else "no ncu --set full capture of this workload's dominant kernel "
                                                                   "is committed",
        },
    }
    if weak is not None:
        nw = w["n_weak"] * world
        line["weak"] = {"value": nw * args.steps / (weak["ms"] / 1e3), "unit": UNIT, "samples_per_gpu": w["n_weak"],
                        "ms_per_step": weak["ms"] / args.steps,
                        "e2e": {"value": nw * args.steps / (weak["ms_e2e"] / 1e3), "ms_per_step": weak["ms_e2e"] / args.steps}}
    if world == 1:
        threads = os.cpu_count() or 1
        v, detail = cpu_step_rate(w["cpu_sample"], threads, score_n=n_total, workload=args.workload)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": _cpu_sample_note(w, w["cpu_sample"]), "detail": detail}
        if not args.no_eager:
            try:
                g = gpu_eager_baseline(w, dev, strong["host"])
                g["ratio_vs_reference_semantics"] = value / g["reference_semantics"]["value"]
                g["ratio_vs_best_case"] = value / g["best_case"]["value"]
                line["gpu_eager_baseline"] = g
            except Exception as ex:          # the baseline must never take the bench line down with it
                line["gpu_eager_baseline"] = {"error": f"{type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sngan32", choices=sorted(WORKLOADS),
                    help="sngan32 = the headline (BASELINE configs[1]); sngan64 / stylegan2 = the same measurement for configs[2] / [4]")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="tensor-core operand type (fp32 accumulate either way).  fp16: <= 1e-3 of the logit scale on every tested "
                         "network, with the measured exceedances on near-zero logits listed in DESIGN.md 4.2; bf16: 3e-3..9e-3")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU baseline (N = 1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
