"""bench.py -- headline benchmark of the per-sample diagnosis path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full recording pass of the hot path over this rank's shard of the synthetic
CIFAR-10-shaped training set with the SNGAN-32 discriminator (BASELINE.json configs[1]):
  weight re-pack (sigma once per pass) -> D forward for every sample -> Welford statistics update ->
  ldr_conf_0.3_ratio_50 score (floor, global MIN, clip) -> sampler weights -> top-100 indices.
Weak scaling: every rank holds 50 000 samples; ``value`` = samples of ALL ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  ``value`` is timed with the dataset resident in HBM; ``e2e`` is the
same metric through ``LogitRecorder.record_from_host`` with the uint8 dataset in pinned host memory
(H2D inside the timed region) and the score vector read back to the host every step.
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

METRIC = "per-sample D logits + LDR scores per second (SNGAN-32, 50k CIFAR-10-shape samples per GPU)"
UNIT = "samples/s"
N_PER_GPU = 50_000
SCORE_KEY = "ldr_conf_0.3_ratio_50"
FLOP_PER_SAMPLE = 2 * 272_072_832        # reference formulation, SURVEY 8(a) appendix
T_WINDOW = 50

# The headline line is configs[1] (default).  --workload selects the same measurement for the other discriminator
# configurations of BASELINE.json (parity-test cases by the contract, measured here for completeness).
WORKLOADS = {
    "sngan32": dict(arch="sngan", size=32, n=N_PER_GPU, key=SCORE_KEY, flop=FLOP_PER_SAMPLE, metric=METRIC,
                    desc="configs[1]: SNGAN-32 recording pass (weights re-packed per pass) + Welford stats + "
                         "ldr_conf_0.3_ratio_50 weights + top-100, 50k x 3x32x32 uint8 per GPU",
                    kernel="conv_swap_kernel block1.c2 (3x3 128->128 @32x32 + avg-pool + shortcut; 55.5% of the reference FLOPs), "
                           "run as the algebraically equal 4x4 stride-2 conv with role-swapped operands (M = 128 channels, "
                           "N = 256 pixels per tcgen05.mma)",
                    dom_ref_flop=2.0 * 9 * 128 * 128 * 1024, cpu_sample=32768, ref_sample=2048),
    "sngan64": dict(arch="sngan", size=64, n=25325, key="ldr_conf_5.0_ratio_50", flop=2 * 644_809_728,
                    metric="per-sample D logits + LDR scores per second (SNGAN-64, CelebA shape, 202 599 / 8 samples per GPU)",
                    desc="configs[2]: SNGAN-64 recording pass + Welford stats + ldr_conf_5.0_ratio_50 weights + top-100, "
                         "25 325 x 3x64x64 uint8 per GPU (the 8-way shard of 202 599)",
                    kernel="conv_swap_kernel block1.c2 (3x3 64->64 @64x64 + avg-pool + shortcut; 23.4% of the reference FLOPs) as the "
                           "4x4 stride-2 conv in super-pixel form (two output pixels = one 128-channel GEMM pixel, 4x6 taps; the "
                           "executed FLOPs include its structural zeros)",
                    dom_ref_flop=2.0 * 9 * 64 * 64 * 4096, cpu_sample=4096, ref_sample=512),
    "stylegan2": dict(arch="stylegan2", size=256, n=2048, key="ldr_conf_3.0_ratio_50", flop=None,
                      metric="per-sample D logits + LDR scores per second (StyleGAN2-256 discriminator, FFHQ shape, bounded 2048 "
                             "samples per GPU)",
                      desc="configs[4]: StyleGAN2-256 recording pass (loader batch 4) + Welford stats + ldr_conf_3.0_ratio_50 "
                           "weights + top-100, 2048 x 3x256x256 uint8 per GPU (bounded sample of the 70k pass)",
                      kernel="conv_swap_kernel ResBlock 1 conv1 (3x3 128->128 @256x256 + FusedLeakyReLU; 20.8% of the FLOPs)",
                      dom_ref_flop=2.0 * 9 * 128 * 128 * 65536, cpu_sample=16, ref_sample=8),
}


def _workload(name):
    w = dict(WORKLOADS[name])
    if w["flop"] is None:
        from diagan_b200 import synthetic
        w["flop"] = synthetic.stylegan2_flops(w["size"])
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
                time.sleep(0.01)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path, restated in oracle/ (kind "port")
# ---------------------------------------------------------------------------------------------------
def cpu_step_rate(sample_n: int, threads: int, score_n: int = N_PER_GPU, workload: str = "sngan32"):
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    sample = w["ref_sample"]
    vals = []
    for _ in range(args.warmup):
        cpu_step_rate(max(4, sample // 8), threads, score_n=2000, workload=args.workload)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_step_rate(sample, threads, score_n=w["n"], workload=args.workload)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"oracle torch fp32 forward on {sample} of {w['n']} samples per step (extrapolated "
                                   f"linearly) + faithful calculate_scores on the full [50,{w['n']}] window / 50"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from diagan_b200 import distributed as D
    from diagan_b200 import engine, synthetic
    from diagan_b200.trainer.trainer import LogitRecorder, ResidentDataset

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    w = _workload(args.workload)
    n_local, n_total = w["n"], w["n"] * world
    lo = rank * n_local
    # synthetic shard: uint8 images of the workload's shape, seeded per rank; pinned host copy for the e2e leg
    host = synthetic.uniform_images_u8(n_local, w["size"], seed=1 + rank, pin=True)
    ds = ResidentDataset(host.to(dev))
    sd0 = synthetic.sngan_state_dict(w["size"], seed=1) if w["arch"] == "sngan" else synthetic.stylegan2_state_dict(w["size"], seed=1)
    base = {k: v.to(dev) for k, v in sd0.items()}
    rec = LogitRecorder(ds, dev, precision=args.precision, inplace_relu=True, keep_snapshots=False, batch=4)
    t_conf = engine.conf_from_key(w["key"])
    host_chunk = 12544 if w["size"] <= 32 else (8192 if w["size"] <= 64 else 256)
    snap = torch.zeros(n_local, dtype=torch.float32, device=dev)
    lib = engine._lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def finish(step_idx):
        """stats -> score -> weights (+ global min / all-gather when sharded) -> top-100."""
        local = rec.stats.score(t_conf, eps=1e-6, min_reduce=D.all_reduce_min_ if world > 1 else None)
        full = D.all_gather_shards(local, n_total) if world > 1 else local
        top = engine.top_indices(full, 100, True)
        return full, top

    # the discriminator "trains" between passes: one perturbed weight set per step, generated BEFORE the timed region (the
    # optimiser step is not part of the recording path); every step still re-packs its weights (sigma + W/sigma) inside it
    n_sets = max(args.warmup, 3) + args.steps
    weight_sets = [synthetic.perturb_(base, 35000 + 100 * i, 1e-3, device=dev) for i in range(n_sets)]

    def step_resident(i):
        rec.record(weight_sets[i % n_sets], step=i, out=snap)
        return finish(i)

    def step_host(i):
        rec.record_from_host(weight_sets[i % n_sets], host, step=i, chunk=host_chunk, first_chunk=min(2048, host_chunk // 4))
        full, top = finish(i)
        return full.cpu(), top.cpu()                       # D2H of the step's result

    def timed(step_fn, steps, warmup, profile=False):
        rec.stats = None
        sampler = ClockSampler(local_rank) if (profile and rank == 0) else None      # NVML init happens here, untimed
        for i in range(warmup):
            step_fn(i)
        barrier()
        if profile:
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

    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, profile=True)
    import ctypes as C
    pm, pl, pf = C.c_double(), C.c_int64(), C.c_double()
    lib.sdg_ctx_profile_read(rec.engine._h, C.byref(pm), C.byref(pl), C.byref(pf))
    lib.sdg_ctx_profile(rec.engine._h, 0)
    ms_e2e, _, _ = timed(step_host, args.steps, max(3, args.warmup))

    if rank != 0:
        return
    peaks = _peaks()
    value = n_total * args.steps / (ms / 1e3)
    e2e_value = n_total * args.steps / (ms_e2e / 1e3)
    dom_tflops = (pf.value / 1e12) / (pm.value / 1e3) if pm.value > 0 else None          # EXECUTED FLOPs / time
    # the same launches in the reference formulation (conv3x3 at full resolution, then avg-pool): 2*9*Cin*Cout per pixel
    ref_flops = w["dom_ref_flop"] * n_local * args.steps                 # every sample of every timed step
    dom_ref_tflops = (ref_flops / 1e12) / (pm.value / 1e3) if pm.value > 0 else None
    line = {
        "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": w["desc"],
                   "samples_per_gpu": n_local, "score_key": w["key"],
                   "l2": f"inputs larger than L2 ({host.numel() / 1e6:.0f} MB dataset, >1 GB activations per sweep); no explicit flush",
                   "parallelism": f"sample-index shards x{world}, MIN all-reduce + one all-gather per step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host.numel()),
                "d2h_bytes_per_step": int(n_total * 8 + 100 * 8), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "whole_path_tflops": value / world * w["flop"] / 1e12,
        "roofline": {
            "bound": "tensor",
            "kernel": w["kernel"],
            "achieved": dom_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": (dom_tflops / peaks["bf16_sustained"]) if dom_tflops else None,
            "flops_counted": "EXECUTED by the tensor pipe (2*M*N*K of the GEMM run" + (": 16 taps per pooled pixel)" if w["arch"] == "sngan" else ")"),
            "reference_formulation_tflops": dom_ref_tflops,
            "reference_formulation_note": "same launches counted as the reference computes them (conv3x3 at 32x32 then "
                                          "avg_pool2d: 36/16 of the executed MACs); > peak because the fused form skips work",
            "peak_source": f"{peaks['source']} (sustained: kernel timed inside a long step)",
            "launches_timed": int(pl.value), "ms_per_launch": (pm.value / pl.value) if pl.value else None,
            # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed
            # `ncu --set full` capture (profiles/r1f_ncu_full_swap_summary.txt, launch 0: 4096 samples, 1.087 GB read +
            # 0.247 GB written = 325.7 KB/sample; algorithmic minimum 256 KiB in + 64 KiB out per sample = 327.7 KB),
            # scaled to the samples one launch of THIS run processed
            "traffic": ((1.087083e9 + 247.195904e6) / 4096.0 * (n_local * args.steps / max(1, int(pl.value)))
                        if args.workload == "sngan32" else None),
            "traffic_note": "ncu --set full, 325.7 KB/sample measured on a 4096-sample launch (profiles/"
                            "r1f_ncu_full_swap_summary.txt, launch 0: tensor pipe 67.5 % active) x samples per launch here; "
                            "algorithmic bytes: 327.7 KB/sample (input once, output once)",
        },
    }
    if world == 1:
        threads = os.cpu_count() or 1
        v, detail = cpu_step_rate(w["cpu_sample"], threads, score_n=n_local, workload=args.workload)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"oracle torch fp32 forward on {w['cpu_sample']} of {n_local} samples (extrapolated linearly) "
                                          f"+ faithful calculate_scores on the full [50,{n_local}] window / 50",
                                "detail": detail}
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
                    help="tensor-core operand type (fp32 accumulate either way); fp16 meets the 1e-3 parity bar")
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
