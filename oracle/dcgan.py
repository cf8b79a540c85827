"""CPU/torch fp32 restatement of the reference's MNIST/Colour-MNIST DCGAN discriminator, eval mode
(oracle; test infrastructure only).

Follows ``diagan-pkg/diagan/models/mnist.py:155-223`` (layers ``:161-192``, forward ``:213-223``) with
``num_pack=1`` and ``use_sn=False`` (``get_norm`` identity, mnist.py:13-17): six bias-free 3x3 convs
(strides 2,1,2,1,2,1), LeakyReLU(0.2) after each, BatchNorm2d on convs 2-6 (eval: running-stat affine),
Dropout(0.5) = identity in eval, flatten ``[B, 512*4*4]`` in NCHW order, ``Linear(8192, 1)``.

Train-mode logits (``save_eval_logits=False`` in the Colour-MNIST scripts) are stochastic and batch
dependent; only eval mode has a defined per-sample value and only it is restated.

Pinned by ``tests/golden/dcgan_eval.npz`` (outputs of the reference module imported from
/root/reference through a torch_mimicry shim by ``oracle/make_golden.py``).

Parameters: flat dict with the reference's ``state_dict`` key names (``conv.0.weight``,
``conv.4.running_mean``, ..., ``out_d.weight``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# (conv index in nn.Sequential, bn index or None, cin, cout, stride)   mnist.py:161-191
LAYERS = [(0, None, 3, 16, 2), (3, 4, 16, 32, 1), (7, 8, 32, 64, 2),
          (11, 12, 64, 128, 1), (15, 16, 128, 256, 2), (19, 20, 256, 512, 1)]
BN_EPS = 1e-5
SLOPE = 0.2


def init_params(seed: int = 1, nc: int = 3) -> dict:
    """PyTorch-default-shaped init (kaiming-uniform(a=sqrt5) convs = U(+-1/sqrt(fan_in)); BN gamma 1,
    beta 0 -- the reference's ``weights_init_3channel(self)`` call is a no-op)
    from a NumPy RandomState, with non-trivial BN running stats so the eval affine is exercised."""
    rng = np.random.RandomState(seed)
    p = {}
    for ci, bi, cin, cout, _ in LAYERS:
        cin_eff = nc if ci == 0 else cin
        b = 1.0 / np.sqrt(cin_eff * 9)
        p[f"conv.{ci}.weight"] = torch.from_numpy(rng.uniform(-b, b, (cout, cin_eff, 3, 3)).astype(np.float32))
        if bi is not None:
            p[f"conv.{bi}.weight"] = torch.from_numpy(rng.uniform(0.8, 1.2, cout).astype(np.float32))
            p[f"conv.{bi}.bias"] = torch.from_numpy(rng.uniform(-0.1, 0.1, cout).astype(np.float32))
            p[f"conv.{bi}.running_mean"] = torch.from_numpy(rng.uniform(-0.05, 0.05, cout).astype(np.float32))
            p[f"conv.{bi}.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    b = 1.0 / np.sqrt(8192)
    p["out_d.weight"] = torch.from_numpy(rng.uniform(-b, b, (1, 8192)).astype(np.float32))
    p["out_d.bias"] = torch.from_numpy(rng.uniform(-b, b, (1,)).astype(np.float32))
    return p


def forward(params: dict, x: torch.Tensor, with_head_l1: bool = False):
    """x float32 NCHW [B,3,32,32] (28x28 also accepted by the reference) -> logits [B,1]  (``with_head_l1``: also
    sum_j |w_j * h_j| + |b| of the Linear(8192, 1) head, [B]: the magnitude of the terms the logit sums)."""
    h = x
    for ci, bi, _, _, stride in LAYERS:
        h = F.conv2d(h, params[f"conv.{ci}.weight"], None, stride=stride, padding=1)
        if bi is not None:
            h = F.batch_norm(h, params[f"conv.{bi}.running_mean"], params[f"conv.{bi}.running_var"],
                             params[f"conv.{bi}.weight"], params[f"conv.{bi}.bias"], False, 0.1, BN_EPS)
        h = F.leaky_relu(h, SLOPE)
    h = h.reshape(-1, 4 * 4 * 512)
    out = F.linear(h, params["out_d.weight"], params["out_d.bias"])
    if with_head_l1:
        return out, (h.abs() * params["out_d.weight"].abs().view(1, -1)).sum(1) + params["out_d.bias"].abs().view(-1)
    return out


def logits_pass(params, data_u8_nhwc: torch.Tensor, batch=64, dtype=torch.float32, device=None, with_head_l1=False):
    """trainer.py:142-156 over an in-memory uint8 NHWC 32x32 dataset (already resized).  ``dtype`` / ``device`` /
    ``with_head_l1``: as in oracle/sngan.py:logits_pass (float64 yardstick, evaluated where the test wants)."""
    from .sngan import normalise_u8
    n = data_u8_nhwc.shape[0]
    out, l1 = np.zeros(n), np.zeros(n)
    if dtype != torch.float32 or device is not None:
        params = {k: v.to(device=device, dtype=dtype) for k, v in params.items()}
    with torch.no_grad():
        for s in range(0, n, batch):
            xb = data_u8_nhwc[s:s + batch]
            y = forward(params, normalise_u8(xb.to(device) if device is not None else xb).to(dtype), with_head_l1=with_head_l1)
            if with_head_l1:
                out[s:s + batch], l1[s:s + batch] = y[0].view(-1).cpu().numpy(), y[1].view(-1).cpu().numpy()
            else:
                out[s:s + batch] = y.view(-1).cpu().numpy()
    return (out, l1) if with_head_l1 else out
