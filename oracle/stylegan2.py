"""CPU/torch fp32 restatement of the reference's StyleGAN2 discriminator (oracle; test infrastructure only).

Follows ``diagan-pkg/diagan/models/stylegan2.py`` (twin of ``stylegan2/model.py:536-660``):
``StyleGANDiscriminator.forward`` :659-677, ``ResBlock`` :598-616, ``ConvLayer`` :553-595, ``EqualConv2d`` :93-128,
``EqualLinear`` :131-166, ``Blur`` :75-90 with ``make_kernel`` :22-30, and the CPU forms of the two native ops:
``fused_leaky_relu`` (op/fused_act.py:104-116) and ``upfirdn2d_native`` (op/upfirdn2d.py:159-200).

Minibatch-stddev couples the samples of a batch: ``out.view(group, -1, 1, C, H, W)`` puts sample ``b = g*M + m`` in
group column ``m`` (M = B/group, group = min(B, 4)), so a logit depends on its batch mates;
``forward`` therefore takes whole batches exactly as the reference does.

Pinned by ``tests/golden/stylegan2_*.npz`` (outputs of the reference module imported from /root/reference).
Parameters: flat dict with the reference's ``state_dict`` key names.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}   # channel_multiplier = 2
SQRT2 = 2 ** 0.5


def block_channels(size: int):
    """[(in, out)] of the ResBlocks for an input of ``size`` x ``size`` (stylegan2.py:640-648)."""
    log_size = int(math.log(size, 2))
    out, cin = [], CHANNELS[size]
    for i in range(log_size, 2, -1):
        cout = CHANNELS[2 ** (i - 1)]
        out.append((cin, cout))
        cin = cout
    return out


def init_params(size: int, seed: int = 1) -> dict:
    """N(0,1) weights like the reference (equalised learning rate scales them at run time), small non-zero biases
    so the bias paths are exercised; drawn from a NumPy RandomState (platform independent)."""
    rng = np.random.RandomState(seed)
    r = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    p = {}
    c0 = CHANNELS[size]
    p["convs.0.0.weight"] = r(c0, 3, 1, 1)
    p["convs.0.1.bias"] = 0.1 * r(c0)
    for i, (cin, cout) in enumerate(block_channels(size), start=1):
        p[f"convs.{i}.conv1.0.weight"] = r(cin, cin, 3, 3)
        p[f"convs.{i}.conv1.1.bias"] = 0.1 * r(cin)
        p[f"convs.{i}.conv2.1.weight"] = r(cout, cin, 3, 3)
        p[f"convs.{i}.conv2.2.bias"] = 0.1 * r(cout)
        p[f"convs.{i}.skip.1.weight"] = r(cout, cin, 1, 1)
    p["final_conv.0.weight"] = r(512, 513, 3, 3)
    p["final_conv.1.bias"] = 0.1 * r(512)
    p["final_linear.0.weight"] = r(512, 8192)
    p["final_linear.0.bias"] = 0.1 * r(512)
    p["final_linear.1.weight"] = r(1, 512)
    p["final_linear.1.bias"] = 0.1 * r(1)
    return p


def fir_kernel() -> torch.Tensor:
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None, :] * k[:, None]
    return k / k.sum()


def blur(x: torch.Tensor, pad0: int, pad1: int) -> torch.Tensor:
    """upfirdn2d(x, k, up=1, down=1, pad=(pad0, pad1)): zero-pad, correlate with the flipped 4x4 FIR."""
    c = x.shape[1]
    w = torch.flip(fir_kernel(), [0, 1]).to(device=x.device, dtype=x.dtype).view(1, 1, 4, 4).repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, [pad0, pad1, pad0, pad1]), w, groups=c)


def flrelu(x, bias):
    return F.leaky_relu(x + bias.view(1, -1, *([1] * (x.ndim - 2))), 0.2) * SQRT2


def eq_conv(x, w, stride=1, padding=0, bias=None):
    scale = 1 / math.sqrt(w.shape[1] * w.shape[2] ** 2)
    return F.conv2d(x, w * scale, bias=bias, stride=stride, padding=padding)


def forward(params: dict, x: torch.Tensor, size: int, with_head_l1: bool = False):
    """x: one reference batch, float NCHW [B,3,size,size] in [-1,1] -> logits [B,1]  (``with_head_l1``: also
    sum_j |w_j * h_j| + |b| of the last EqualLinear, [B]: the magnitude of the terms the logit sums)."""
    p = params
    h = flrelu(eq_conv(x, p["convs.0.0.weight"]), p["convs.0.1.bias"])
    for i, _ in enumerate(block_channels(size), start=1):
        o = flrelu(eq_conv(h, p[f"convs.{i}.conv1.0.weight"], padding=1), p[f"convs.{i}.conv1.1.bias"])
        o = flrelu(eq_conv(blur(o, 2, 2), p[f"convs.{i}.conv2.1.weight"], stride=2), p[f"convs.{i}.conv2.2.bias"])
        s = eq_conv(blur(h, 1, 1), p[f"convs.{i}.skip.1.weight"], stride=2)
        h = (o + s) / math.sqrt(2)
    b, c, hh, ww = h.shape
    group = min(b, 4)
    sd = h.view(group, -1, 1, c, hh, ww)
    sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8)
    sd = sd.mean([2, 3, 4], keepdims=True).squeeze(2)
    sd = sd.repeat(group, 1, hh, ww)
    h = torch.cat([h, sd], 1)
    h = flrelu(eq_conv(h, p["final_conv.0.weight"], padding=1), p["final_conv.1.bias"])
    h = h.reshape(b, -1)
    w0 = p["final_linear.0.weight"]
    h = flrelu(F.linear(h, w0 * (1 / math.sqrt(w0.shape[1]))), p["final_linear.0.bias"])
    w1 = p["final_linear.1.weight"]
    out = F.linear(h, w1 * (1 / math.sqrt(w1.shape[1])), bias=p["final_linear.1.bias"])
    if with_head_l1:
        return out, (h.abs() * (w1.abs() * (1 / math.sqrt(w1.shape[1]))).view(1, -1)).sum(1) + p["final_linear.1.bias"].abs().view(-1)
    return out


def logits_pass(params, data_u8_nhwc: torch.Tensor, size: int, batch: int, dtype=torch.float32, device=None,
                with_head_l1=False):
    """Recording pass over an in-memory dataset in consecutive batches of ``batch`` (stylegan2/train_ffhq.py:128-143
    with a sequential, un-flipped loader; the tail that does not fill a batch is dropped like drop_last=True)."""
    from .sngan import normalise_u8
    n = data_u8_nhwc.shape[0] // batch * batch
    out, l1 = np.zeros(data_u8_nhwc.shape[0]), np.zeros(data_u8_nhwc.shape[0])
    if dtype != torch.float32 or device is not None:
        params = {k: v.to(device=device, dtype=dtype) for k, v in params.items()}
    with torch.no_grad():
        for s in range(0, n, batch):
            xb = data_u8_nhwc[s:s + batch]
            x = normalise_u8(xb.to(device) if device is not None else xb).to(dtype)
            y = forward(params, x, size, with_head_l1=with_head_l1)
            if with_head_l1:
                out[s:s + batch], l1[s:s + batch] = y[0].view(-1).cpu().numpy(), y[1].view(-1).cpu().numpy()
            else:
                out[s:s + batch] = y.view(-1).cpu().numpy()
    return (out, l1) if with_head_l1 else out
