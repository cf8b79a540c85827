"""CPU/torch fp32 restatement of torch-mimicry 0.1.16 SNGAN discriminators (oracle; test infrastructure only).

PARITY UNPINNED.  ``torch-mimicry==0.1.16`` (requirements.txt:72, environment.yml:122) is a
third-party dependency that is neither vendored under /root/reference nor installable here.
This file restates, from the published sources as recalled:

* ``torch_mimicry/modules/spectral_norm.py``   SpectralNorm._power_iteration / sn_weights
* ``torch_mimicry/modules/resblocks.py``       DBlock, DBlockOptimized
* ``torch_mimicry/nets/sngan/sngan_32.py``     SNGANDiscriminator32
* ``torch_mimicry/nets/sngan/sngan_64.py``     SNGANDiscriminator64

Parity is anchored on the reference's own call sites: ``predefined_models.py:14,36-52,74-90``
(construction), ``trainer.py:145-154`` (eval mode, no_grad, ``netD(x)`` -> [B,1]).

Parameters are a flat dict with torch-mimicry ``state_dict`` key names
(``block1.c1.weight``, ``block1.c1.bias``, ``block1.c1.sn_u``, ..., ``l5.weight``), so a phase-1
checkpoint's ``model_state_dict`` can be fed straight in.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# (name, kind, cin, cout, downsample); kind "opt" = DBlockOptimized, "res" = DBlock
ARCH = {
    32: dict(blocks=[("block1", "opt", 3, 128, True), ("block2", "res", 128, 128, True),
                     ("block3", "res", 128, 128, False), ("block4", "res", 128, 128, False)],
             head="l5", ndf=128, size=32),
    64: dict(blocks=[("block1", "opt", 3, 64, True), ("block2", "res", 64, 128, True),
                     ("block3", "res", 128, 256, True), ("block4", "res", 256, 512, True),
                     ("block5", "res", 512, 1024, True)],
             head="l6", ndf=1024, size=64),
}


def has_shortcut_conv(kind, cin, cout, down):
    """resblocks.py: ``learnable_sc = in != out or downsample`` (DBlockOptimized always has c_sc)."""
    return kind == "opt" or cin != cout or down


def layer_list(arch: int):
    """[(key_prefix, cout, cin, ksize)] for every spectral-normalised layer, forward order."""
    out = []
    for name, kind, cin, cout, down in ARCH[arch]["blocks"]:
        hidden = cout if kind == "opt" else cin          # DBlock: hidden = in_channels
        out.append((f"{name}.c1", hidden, cin, 3))
        out.append((f"{name}.c2", cout, hidden, 3))
        if has_shortcut_conv(kind, cin, cout, down):
            out.append((f"{name}.c_sc", cout, cin, 1))
    out.append((ARCH[arch]["head"], 1, ARCH[arch]["ndf"], 0))
    return out


def init_params(arch: int, seed: int = 1) -> dict:
    """Random init with mimicry's scheme (xavier-uniform, gain sqrt2 for c1/c2, 1 for c_sc and the
    linear head; torch default uniform bias; sn_u ~ N(0,1); sn_sigma = 1), drawn from a NumPy
    RandomState so the values are identical on every platform."""
    rng = np.random.RandomState(seed)
    p = {}
    for key, cout, cin, k in layer_list(arch):
        if k == 0:
            shape, fan_in, fan_out, gain = (cout, cin), cin, cout, 1.0
        else:
            shape, fan_in, fan_out = (cout, cin, k, k), cin * k * k, cout * k * k
            gain = 1.0 if key.endswith("c_sc") else math.sqrt(2.0)
        bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
        p[f"{key}.weight"] = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        bb = 1.0 / math.sqrt(fan_in)
        p[f"{key}.bias"] = torch.from_numpy(rng.uniform(-bb, bb, (cout,)).astype(np.float32))
        p[f"{key}.sn_u"] = torch.from_numpy(rng.standard_normal((1, cout)).astype(np.float32))
        p[f"{key}.sn_sigma"] = torch.ones(1)
    return p


def perturb_params(params: dict, step: int, scale: float = 1e-3) -> dict:
    """Deterministic stand-in for one stretch of GAN training between two recording passes
    (SURVEY 8(d) item 2): W += scale * randn with seed = step; biases and sn_u untouched."""
    rng = np.random.RandomState(step)
    out = dict(params)
    for k in sorted(params):
        if k.endswith(".weight"):
            w = params[k]
            out[k] = w + scale * torch.from_numpy(rng.standard_normal(tuple(w.shape)).astype(np.float32))
    return out


def sigma_eval(weight: torch.Tensor, u: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """spectral_norm.py _power_iteration with num_iters=1; in eval mode the buffers are not
    updated, so sigma is the same for every batch of a recording pass."""
    W = weight.reshape(weight.shape[0], -1)
    v = F.normalize(torch.matmul(u, W), eps=eps)
    u2 = F.normalize(torch.matmul(v, W.t()), eps=eps)
    return torch.mm(u2, torch.mm(W, v.t()))            # [1,1]


def sn_weight(params: dict, key: str) -> torch.Tensor:
    w = params[f"{key}.weight"]
    return w / sigma_eval(w, params[f"{key}.sn_u"])


def normalised_weights(params: dict, arch: int) -> dict:
    """{layer key: W / sigma} for every layer: what eval-mode mimicry recomputes (to the same value) in every forward."""
    return {key: sn_weight(params, key) for key, _, _, _ in layer_list(arch)}


def _conv(params, key, x, pad):
    w = params["__wn__"][key] if "__wn__" in params else sn_weight(params, key)
    return F.conv2d(x, w, params[f"{key}.bias"], stride=1, padding=pad)


def forward(params: dict, x: torch.Tensor, arch: int = 32, inplace_relu: bool = True, with_head_l1: bool = False):
    """x float32 NCHW in [-1,1] -> logits [B,1]  (``with_head_l1``: also the L1 mass of the head's dot product,
    sum_c |w_c * pooled_c| + |b|, [B] -- the magnitude of the terms a logit sums, i.e. the scale against which the rounding
    error of that sum is meaningful when the terms cancel to a near-zero logit).

    ``params["__wn__"]`` (optional, from :func:`normalised_weights`) supplies W / sigma computed once instead of per forward:
    only bench.py's best-case eager baseline uses it (eval-mode sigma is the same in every forward).

    ``inplace_relu=True`` reproduces mimicry's ``nn.ReLU(True)`` aliasing in ``DBlock``: the
    residual branch runs first and rectifies ``x`` in place, so the shortcut branch (the 1x1 conv,
    or the identity of blocks 3/4 of SNGAN-32) sees relu(x).  ``False`` gives the textbook block.
    """
    h = x
    for name, kind, cin, cout, down in ARCH[arch]["blocks"]:
        if kind == "opt":
            r = _conv(params, f"{name}.c1", h, 1)
            r = F.relu(r)
            r = _conv(params, f"{name}.c2", r, 1)
            r = F.avg_pool2d(r, 2)
            s = _conv(params, f"{name}.c_sc", F.avg_pool2d(h, 2), 0)
            h = r + s
        else:
            a = F.relu(h)
            r = _conv(params, f"{name}.c1", a, 1)
            r = F.relu(r)
            r = _conv(params, f"{name}.c2", r, 1)
            if down:
                r = F.avg_pool2d(r, 2)
            s = a if inplace_relu else h
            if has_shortcut_conv(kind, cin, cout, down):
                s = _conv(params, f"{name}.c_sc", s, 0)
                if down:
                    s = F.avg_pool2d(s, 2)
            h = r + s
    h = F.relu(h)
    h = torch.sum(h, dim=(2, 3))
    head = ARCH[arch]["head"]
    w, b = (params["__wn__"][head] if "__wn__" in params else sn_weight(params, head)), params[f"{head}.bias"]
    out = F.linear(h, w, b)
    if with_head_l1:
        return out, (h.abs() * w.abs().view(1, -1)).sum(1) + b.abs().view(-1)
    return out


def normalise_u8(x_u8_nhwc: torch.Tensor) -> torch.Tensor:
    """ToTensor + Normalize(0.5, 0.5) of transform.py:3-11 on a uint8 NHWC batch -> float32 NCHW."""
    x = x_u8_nhwc.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    return (x - 0.5) / 0.5


def logits_pass(params, data_u8_nhwc: torch.Tensor, arch=32, batch=64, inplace_relu=True,
                dtype=torch.float32, device=None, with_head_l1=False):
    """The recording pass of trainer.py:142-156 on an in-memory dataset, sequential batches of 64:
    float64 [N] with fp32 values widened, indexed by dataset index.  ``dtype=torch.float64`` evaluates
    the same network in double precision (the exact-arithmetic yardstick the parity tests use to
    separate the GPU's rounding error from the fp32 CPU path's own).  ``device``: where the checker's own arithmetic
    runs (the float64 evaluation of thousands of samples is run on the GPU by the BASELINE-size tests; still torch, still
    the restatement above, never the product's kernels).  ``with_head_l1``: -> (logits, head L1 mass), see forward()."""
    n = data_u8_nhwc.shape[0]
    out, l1 = np.zeros(n), np.zeros(n)
    if dtype != torch.float32 or device is not None:
        params = {k: v.to(device=device, dtype=dtype) for k, v in params.items()}
    with torch.no_grad():
        for s in range(0, n, batch):
            x = normalise_u8(data_u8_nhwc[s:s + batch].to(device) if device is not None else data_u8_nhwc[s:s + batch]).to(dtype)
            y = forward(params, x, arch, inplace_relu, with_head_l1=with_head_l1)
            if with_head_l1:
                out[s:s + batch], l1[s:s + batch] = y[0].view(-1).cpu().numpy(), y[1].view(-1).cpu().numpy()
            else:
                out[s:s + batch] = y.view(-1).cpu().numpy()
    return (out, l1) if with_head_l1 else out


# ---------------------------------------------------------------------------------------------------
# InfoMax-GAN / SSGAN discriminators (predefined_models.py:36-52,74-90 model='infomax_gan' | 'ssgan')
# PARITY UNPINNED like the rest of this file: torch-mimicry 0.1.16 nets/infomax_gan/infomax_gan_{32,64}.py and
# nets/ssgan/ssgan_{32,64}.py as recalled.  Both are the SNGAN residual stack; what differs is the attribute names of the
# state_dict (InfoMax) and the extra outputs of forward(), of which the diagnosis path keeps [0] (trainer.py:151-152).
# ---------------------------------------------------------------------------------------------------
def infomax_key_map(arch: int) -> dict:
    n_blocks = len(ARCH[arch]["blocks"])
    m = {f"local_feat_blocks.{i}": f"block{i + 1}" for i in range(n_blocks - 1)}
    m["global_feat_blocks.0"] = f"block{n_blocks}"
    m["linear"] = ARCH[arch]["head"]
    return m


def as_variant_state_dict(params: dict, arch: int, variant: str, seed: int = 7) -> dict:
    """SNGAN-named parameters -> a state_dict with the key names (and extra heads) of the mimicry ``variant``:
    'ssgan' adds the 4-way rotation head l_y; 'infomax' renames the blocks and adds the nrkhs critic layers."""
    rng = np.random.RandomState(seed)
    ndf = ARCH[arch]["ndf"]
    r = lambda *shape: torch.from_numpy((0.05 * rng.standard_normal(shape)).astype(np.float32))
    if variant == "ssgan":
        out = dict(params)
        out.update({"l_y.weight": r(4, ndf), "l_y.bias": r(4), "l_y.sn_u": r(1, 4), "l_y.sn_sigma": torch.ones(1)})
        return out
    assert variant == "infomax"
    inv = {v: k for k, v in infomax_key_map(arch).items()}
    out = {}
    for k, v in params.items():
        prefix = k.split(".")[0]
        out[inv[prefix] + k[len(prefix):]] = v
    nrkhs = 1024
    for name, shape in (("local_nrkhs_a", (ndf, ndf, 1, 1)), ("local_nrkhs_b", (nrkhs, ndf, 1, 1)),
                        ("local_nrkhs_sc", (nrkhs, ndf, 1, 1)), ("global_nrkhs_a", (ndf, ndf)),
                        ("global_nrkhs_b", (nrkhs, ndf)), ("global_nrkhs_sc", (nrkhs, ndf))):
        out[f"{name}.weight"], out[f"{name}.bias"] = r(*shape), r(shape[0])
        out[f"{name}.sn_u"], out[f"{name}.sn_sigma"] = r(1, shape[0]), torch.ones(1)
    return out


def forward_variant(state_dict: dict, x: torch.Tensor, arch: int, variant: str, inplace_relu: bool = True):
    """forward() of the mimicry InfoMax-GAN / SSGAN discriminator on its own state_dict: a TUPLE whose [0] is the [B,1]
    logit -- (output, output_classes) for SSGAN, (output, local_feat, global_feat) for InfoMax-GAN."""
    if variant == "ssgan":
        out = forward(state_dict, x, arch, inplace_relu)
        return out, None                       # the rotation logits are not on the diagnosis path
    km = infomax_key_map(arch)
    canon = {}
    for k, v in state_dict.items():
        prefix = ".".join(k.split(".")[:2]) if k.startswith(("local_feat_blocks", "global_feat_blocks")) else k.split(".")[0]
        if prefix in km:
            canon[km[prefix] + k[len(prefix):]] = v
    return forward(canon, x, arch, inplace_relu), None, None
