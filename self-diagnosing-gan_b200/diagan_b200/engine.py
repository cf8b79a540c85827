"""Device-side engines of the diagnosis path: thin Python objects over the C ABI (``include/sdg.h``).

``DiscriminatorEngine``  -- the recording-pass forward (replaces ``netD(x)`` inside
                            ``LogTrainer._get_logit``, diagan-pkg/diagan/trainer/trainer.py:145-154)
``RunningStats``         -- per-sample Welford mean / M2 / last / sum|delta| over recording passes
``window_scores`` etc.   -- ``calculate_scores`` arithmetic (diagan-pkg/diagan/utils/plot.py:220-249)
``top_indices``          -- ``np.argsort(w)[-k:]`` / ``[:k]`` (eval_gan_drs_with_index.py:97-99)

Everything here takes and returns torch CUDA tensors; torch only provides memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, ptr_array, stream_ptr

FLOOR = 1e-2      # plot.py:230
RATIO = 50.0      # plot.py:248


def conf_values() -> np.ndarray:
    """np.arange(0.1, 10.0, 0.1): the exact multipliers behind the 99 ldr_conf keys (plot.py:247)."""
    return np.arange(0.1, 10.0, 0.1)


def conf_key(t: float) -> str:
    return f"ldr_conf_{t:.1f}_ratio_50"


def conf_from_key(key: str) -> float:
    for t in conf_values():
        if conf_key(t) == key:
            return float(t)
    raise KeyError(f"{key!r} is not one of the reference's score keys")


def _resolve_device(device=None) -> torch.device:
    """A CUDA device WITH an index (``cuda`` alone means the current device, not device 0)."""
    d = torch.device("cuda") if device is None else torch.device(device)
    if d.type != "cuda":
        raise _lib.SdgError(f"device {d}: diagan_b200 is sm_100a only and has no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if d.index is None else d


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.SdgError(f"{name} must be a CUDA tensor (diagan_b200 has no CPU path)")
    if not t.is_contiguous():
        raise _lib.SdgError(f"{name} must be contiguous")


# mimicry state_dict layer order (include/sdg.h, sdg_sngan_load)
def sngan_layer_keys(arch: int):
    if arch == 32:
        blocks = [("block1", True), ("block2", True), ("block3", False), ("block4", False)]
        head = "l5"
    elif arch == 64:
        blocks = [("block1", True), ("block2", True), ("block3", True), ("block4", True), ("block5", True)]
        head = "l6"
    else:
        raise ValueError(arch)
    keys = []
    for name, sc in blocks:
        keys += [f"{name}.c1", f"{name}.c2"] + ([f"{name}.c_sc"] if sc else [])
    return keys + [head]


# torch-mimicry InfoMaxGANDiscriminator32/64 (predefined_models.py:36-52,74-90: model='infomax_gan') hold the SAME residual
# stack as SNGANDiscriminator32/64 under other attribute names: local_feat_blocks = Sequential(block1 .. block(n-1)),
# global_feat_blocks = Sequential(last block), linear = the SNLinear head; forward returns (output, local_feat, global_feat)
# and the diagnosis path keeps [0] (trainer.py:151-152).  SSGANDiscriminator32/64 ARE the SNGAN stacks plus a rotation head
# l_y whose output is dropped the same way.  [torch-mimicry 0.1.16 sources as recalled: parity unpinned like SNGAN itself.]
def _infomax_key_map(arch: int) -> dict:
    n_blocks = 4 if arch == 32 else 5
    m = {f"local_feat_blocks.{i}": f"block{i + 1}" for i in range(n_blocks - 1)}
    m["global_feat_blocks.0"] = f"block{n_blocks}"
    m["linear"] = "l5" if arch == 32 else "l6"
    return m


def canonical_sngan_state_dict(state_dict, kind: str) -> dict:
    """state_dict of an InfoMax-GAN / SSGAN discriminator -> the SNGAN key names the engine loads (tensors shared, extra
    heads -- l_y, the nrkhs critics -- dropped)."""
    if kind.startswith("infomax"):
        out = {}
        for old, new in _infomax_key_map(int(kind[7:])).items():
            for k, v in state_dict.items():
                if k.startswith(old + "."):
                    out[new + k[len(old):]] = v
        return out
    return state_dict


def detect_arch(state_dict) -> str:
    """'sngan32' | 'sngan64' | 'ssgan32' | 'ssgan64' | 'infomax32' | 'infomax64' | 'dcgan32' | 'stylegan2' from parameter
    names (mimicry / reference state_dict layouts)."""
    keys = set(state_dict.keys())
    if "local_feat_blocks.0.c1.weight" in keys and "global_feat_blocks.0.c1.weight" in keys and "linear.weight" in keys:
        return "infomax64" if "local_feat_blocks.3.c1.weight" in keys else "infomax32"
    if "l_y.weight" in keys and "block1.c_sc.weight" in keys:
        return "ssgan64" if "l6.weight" in keys else "ssgan32"
    if "conv.0.weight" in keys and "out_d.weight" in keys:
        return "dcgan32"
    if "final_conv.0.weight" in keys and "convs.0.0.weight" in keys:
        return "stylegan2"
    if "l6.weight" in keys and "block5.c1.weight" in keys:
        return "sngan64"
    if "l5.weight" in keys and "block4.c1.weight" in keys and "block1.c_sc.weight" in keys:
        return "sngan32"
    raise _lib.SdgError("unsupported discriminator: expected torch-mimicry SNGAN / SSGAN / InfoMax-GAN Discriminator32/64, the "
                        "reference's MNIST_DCGAN_Discriminator or its StyleGANDiscriminator state_dict")


def stylegan2_tensor_keys(state_dict):
    """(size, ordered key list) of the reference StyleGANDiscriminator state_dict (include/sdg.h, sdg_stylegan2_load)."""
    nblk = 0
    while f"convs.{nblk + 1}.conv1.0.weight" in state_dict:
        nblk += 1
    keys = ["convs.0.0.weight", "convs.0.1.bias"]
    for i in range(1, nblk + 1):
        keys += [f"convs.{i}.conv1.0.weight", f"convs.{i}.conv1.1.bias", f"convs.{i}.conv2.1.weight",
                 f"convs.{i}.conv2.2.bias", f"convs.{i}.skip.1.weight"]
    keys += ["final_conv.0.weight", "final_conv.1.bias", "final_linear.0.weight", "final_linear.0.bias",
             "final_linear.1.weight", "final_linear.1.bias"]
    return 4 << nblk, keys


class DiscriminatorEngine:
    """One per rank.  ``load_*`` packs the weights (sigma computed once, eval semantics); ``forward``
    runs the whole recording pass for a device-resident batch of any size."""

    def __init__(self, device=None, chunk: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SdgError("no CUDA device: diagan_b200 is sm_100a only and has no CPU fallback")
        self.device = _resolve_device(device)
        h = C.c_void_p()
        check(self.lib.sdg_ctx_create(self.device.index, C.byref(h)), "sdg_ctx_create")
        self._h = h
        # fp16 range guard (include/sdg.h, sdg_ctx_set_range_flag): a sticky device flag the fp16 kernels OR into
        self._range = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self.lib.sdg_ctx_set_range_flag(self._h, ptr(self._range)), "sdg_ctx_set_range_flag")
        self.arch = None
        self.size = None
        self.n_layers = 0
        if chunk:
            self.set_chunk(chunk)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdg_ctx_set_range_flag(self._h, None)
            self.lib.sdg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def range_status(self, reset: bool = True) -> int:
        """Host read (one 4-byte D2H, synchronises the stream) of the fp16 range flag: 0, or a mask of
        ``_lib.RANGE_ACT`` / ``_lib.RANGE_WEIGHT`` if a value left the fp16 range since the last reset."""
        v = int(self._range.item())
        if v and reset:
            self._range.zero_()
        return v

    def set_chunk(self, samples: int):
        check(self.lib.sdg_ctx_set_chunk(self._h, int(samples)), "sdg_ctx_set_chunk")

    # ---- weights -------------------------------------------------------------------------------
    def _dev(self, t):
        return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def load_sngan(self, state_dict, arch: int, precision: str = "fp16", inplace_relu: bool = True):
        keys = sngan_layer_keys(arch)
        W = [self._dev(state_dict[f"{k}.weight"]) for k in keys]
        b = [self._dev(state_dict[f"{k}.bias"]) for k in keys]
        u = [self._dev(state_dict[f"{k}.sn_u"]).view(-1) for k in keys]
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}[precision]
        with torch.cuda.device(self.device):
            check(self.lib.sdg_sngan_load(self._h, arch, len(keys), ptr_array(W), ptr_array(b), ptr_array(u), prec,
                                          1 if inplace_relu else 0, stream_ptr(self.device)), "sdg_sngan_load")
        self._keep = (W, b, u)          # the pack kernels read them asynchronously on the stream
        self.arch, self.size, self.n_layers = f"sngan{arch}", arch, len(keys)
        return self

    def load_dcgan(self, state_dict, precision: str = "fp16"):
        """MNIST_DCGAN_Discriminator, eval mode.  fp16 / bf16: convs 2..6 on tcgen05 (BatchNorm folded, LeakyReLU in the
        epilogue); fp32: the exact CUDA-core engine (the 1e-5 parity mode)."""
        conv_idx = [0, 3, 7, 11, 15, 19]
        bn_idx = [4, 8, 12, 16, 20]
        W = [self._dev(state_dict[f"conv.{i}.weight"]) for i in conv_idx]
        if W[0].shape[1] != 3:
            raise _lib.SdgError(f"DCGAN discriminator with nc * num_pack = {W[0].shape[1]} input channels: only nc = 3, "
                                "num_pack = 1 is implemented")
        g = [self._dev(state_dict[f"conv.{i}.weight"]) for i in bn_idx]
        be = [self._dev(state_dict[f"conv.{i}.bias"]) for i in bn_idx]
        mu = [self._dev(state_dict[f"conv.{i}.running_mean"]) for i in bn_idx]
        var = [self._dev(state_dict[f"conv.{i}.running_var"]) for i in bn_idx]
        fw, fb = self._dev(state_dict["out_d.weight"]), self._dev(state_dict["out_d.bias"])
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}[precision]
        with torch.cuda.device(self.device):
            check(self.lib.sdg_dcgan_load(self._h, ptr_array(W), ptr_array(g), ptr_array(be), ptr_array(mu),
                                          ptr_array(var), ptr(fw), ptr(fb), prec, stream_ptr(self.device)),
                  "sdg_dcgan_load")
        self._keep = (W, g, be, mu, var, fw, fb)
        self.arch, self.size, self.n_layers = "dcgan32", 32, 7
        return self

    def load_stylegan2(self, state_dict, precision: str = "fp16", batch: int = 4):
        """The reference's StyleGANDiscriminator (any power-of-two size).  ``batch`` = the loader batch size the
        reference would use: minibatch-stddev groups are formed inside consecutive batches of that size."""
        size, keys = stylegan2_tensor_keys(state_dict)
        T = [self._dev(state_dict[k]) for k in keys]
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}[precision]
        with torch.cuda.device(self.device):
            check(self.lib.sdg_stylegan2_load(self._h, size, len(T), ptr_array(T), prec, stream_ptr(self.device)),
                  "sdg_stylegan2_load")
            check(self.lib.sdg_ctx_set_batch(self._h, int(batch)), "sdg_ctx_set_batch")
        self._keep = T
        self.arch, self.size, self.n_layers = "stylegan2", size, len(T)
        return self

    def set_batch(self, batch: int):
        check(self.lib.sdg_ctx_set_batch(self._h, int(batch)), "sdg_ctx_set_batch")

    def load(self, state_dict, precision: str = None, inplace_relu: bool = True):
        kind = detect_arch(state_dict)
        if kind == "dcgan32":
            return self.load_dcgan(state_dict, precision or "fp16")
        if kind == "stylegan2":
            return self.load_stylegan2(state_dict, precision or "fp16")
        arch = int(kind[-2:])                       # sngan / ssgan / infomax + 32 | 64: one residual stack, one engine
        return self.load_sngan(canonical_sngan_state_dict(state_dict, kind), arch, precision or "fp16", inplace_relu)

    def sigmas(self) -> torch.Tensor:
        out = torch.empty(self.n_layers, dtype=torch.float32, device=self.device)
        check(self.lib.sdg_sngan_sigmas(self._h, ptr(out), stream_ptr(self.device)), "sdg_sngan_sigmas")
        return out

    # ---- recording pass ------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """x: uint8 [n,H,W,3] (raw dataset bytes, normalised on load) or float32 [n,3,H,W] in [-1,1];
        returns float32 [n] logits (written into ``out`` if given, e.g. a slice of the snapshot row)."""
        _require_cuda(x, "x")
        n = x.shape[0]
        if x.dtype == torch.uint8:
            layout = _lib.LAYOUT_U8_NHWC
            ok = x.dim() == 4 and x.shape[1] == self.size and x.shape[2] == self.size and x.shape[3] == 3
        elif x.dtype == torch.float32:
            layout = _lib.LAYOUT_F32_NCHW
            ok = x.dim() == 4 and x.shape[1] == 3 and x.shape[2] == self.size and x.shape[3] == self.size
        else:
            raise _lib.SdgError(f"x dtype {x.dtype}: expected uint8 NHWC or float32 NCHW")
        if not ok:
            raise _lib.SdgError(f"x shape {tuple(x.shape)} does not match a {self.arch} input")
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=self.device)
        else:
            _require_cuda(out, "out")
            assert out.dtype == torch.float32 and out.numel() == n
        with torch.cuda.device(self.device):
            check(self.lib.sdg_d_forward(self._h, ptr(x), layout, n, ptr(out), stream_ptr(self.device)), "sdg_d_forward")
        return out


class RunningStats:
    """Per-sample streaming statistics of the recorded logits (the design's replacement for keeping
    every snapshot; diagan-pkg/diagan/trainer/trainer.py:337-338 vs plot.py:243-246)."""

    def __init__(self, n: int, device):
        self.lib = _lib.load()
        self.n, self.device, self.count = int(n), _resolve_device(device), 0
        self.state = torch.zeros(4, self.n, dtype=torch.float64, device=self.device)   # mean, m2, last, sad

    def update(self, snapshot: torch.Tensor):
        _require_cuda(snapshot, "snapshot")
        assert snapshot.dtype == torch.float32 and snapshot.numel() == self.n
        m, q, la, sa = self.state[0], self.state[1], self.state[2], self.state[3]
        with torch.cuda.device(self.device):
            check(self.lib.sdg_stats_update(ptr(snapshot), ptr(m), ptr(q), ptr(la), ptr(sa), self.n, self.count,
                                            stream_ptr(self.device)), "sdg_stats_update")
        self.count += 1

    @property
    def mean(self): return self.state[0]

    @property
    def m2(self): return self.state[1]

    def ldr(self): return self.state[2]

    def ldrd(self): return self.state[3] / (self.count - 1)

    def ldrv(self): return self.state[1] / (self.count - 1)

    def score(self, conf: float, floor=FLOOR, ratio=RATIO, eps=0.0, min_reduce=None) -> torch.Tensor:
        """ldr_conf score from the running moments; ``min_reduce`` (callable on the device [1] min tensor)
        is where a sharded caller all-reduces the clip bound."""
        return scores_from_moments(self.mean, self.m2, [conf], floor, ratio, eps, m2_over=self.count - 1,
                                   min_reduce=min_reduce)[0]


def window_moments(snaps: torch.Tensor, want=("mean", "var", "ldrd", "ldr")):
    """snaps [T,N] float32 or float64 (rows = selected window, in order) -> dict of float64 [N]."""
    lib = _lib.load()
    if not snaps.is_cuda:
        raise _lib.SdgError("snaps must be a CUDA tensor (diagan_b200 has no CPU path)")
    if snaps.dim() != 2 or (snaps.shape[1] > 1 and snaps.stride(1) != 1):
        raise _lib.SdgError("snaps must be [T, N] with unit stride along N (rows may be strided)")
    T, n = snaps.shape
    out = {k: torch.empty(n, dtype=torch.float64, device=snaps.device) for k in want}
    fn = {torch.float32: lib.sdg_window_moments_f32, torch.float64: lib.sdg_window_moments_f64}[snaps.dtype]
    with torch.cuda.device(snaps.device):
        check(fn(ptr(snaps), T, n, snaps.stride(0), ptr(out.get("mean")), ptr(out.get("var")), ptr(out.get("ldrd")),
                 ptr(out.get("ldr")), stream_ptr(snaps.device)), "sdg_window_moments")
    return out


def scores_from_moments(mean, var, confs, floor=FLOOR, ratio=RATIO, eps=0.0, m2_over=0.0, min_reduce=None):
    """-> float64 [len(confs), N]: clip_max_ratio(clip_min(mean + t*std)) for each multiplier t."""
    lib = _lib.load()
    n = mean.numel()
    k = len(confs)
    score = torch.empty(k, n, dtype=torch.float64, device=mean.device)
    out_chunks = []
    for s in range(0, k, 128):                      # the ABI takes at most 128 multipliers per call
        cs = np.ascontiguousarray(confs[s:s + 128], dtype=np.float64)
        kk = len(cs)
        sc = score[s:s + kk]
        mins = torch.empty(kk, dtype=torch.float64, device=mean.device)
        st = stream_ptr(mean.device)
        with torch.cuda.device(mean.device):
            check(lib.sdg_score_floor_min(ptr(mean), ptr(var), n, cs.ctypes.data_as(C.POINTER(C.c_double)), kk, floor,
                                          float(m2_over), ptr(sc), ptr(mins), st), "sdg_score_floor_min")
            if min_reduce is not None:
                min_reduce(mins)
            check(lib.sdg_score_clip(ptr(sc), n, kk, ptr(mins), ratio, eps, st), "sdg_score_clip")
        out_chunks.append(mins)
    return score


_topk_ws = {}


def top_indices(score: torch.Tensor, k: int, largest: bool = True) -> torch.Tensor:
    """== np.argsort(score, kind='stable')[-k:] (largest) or [:k]; int64 [k] on the device."""
    lib = _lib.load()
    _require_cuda(score, "score")
    assert score.dtype == torch.float64 and score.dim() == 1
    n = score.numel()
    need = lib.sdg_topk_workspace_bytes(n)
    ws = _topk_ws.get(score.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=score.device)
        _topk_ws[score.device] = ws
    out = torch.empty(k, dtype=torch.int64, device=score.device)
    with torch.cuda.device(score.device):
        check(lib.sdg_topk_indices(ptr(score), n, k, 1 if largest else 0, ptr(out), ptr(ws), ws.numel(),
                                   stream_ptr(score.device)), "sdg_topk_indices")
    return out


def launch_count(reset: bool = False) -> int:
    lib = _lib.load()
    c = int(lib.sdg_launch_count())
    if reset:
        lib.sdg_launch_count_reset()
    return c
