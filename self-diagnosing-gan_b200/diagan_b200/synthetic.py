"""Synthetic stand-ins for datasets and checkpoints (no network on the GPU box): random-init
discriminator ``state_dict``s with the reference's key names and shapes, and uniform uint8 images of
CIFAR-10 / CelebA / Colour-MNIST shape.  Used by ``bench.py`` and ``__graft_entry__.smoke``."""
from __future__ import annotations

import math

import numpy as np
import torch

from .engine import sngan_layer_keys

_SNGAN_BLOCKS = {
    32: [("block1", "opt", 3, 128, True), ("block2", "res", 128, 128, True),
         ("block3", "res", 128, 128, False), ("block4", "res", 128, 128, False)],
    64: [("block1", "opt", 3, 64, True), ("block2", "res", 64, 128, True), ("block3", "res", 128, 256, True),
         ("block4", "res", 256, 512, True), ("block5", "res", 512, 1024, True)],
}


def sngan_shapes(arch: int) -> dict:
    """{layer key: weight shape} of torch-mimicry SNGANDiscriminator32/64."""
    shapes = {}
    for name, kind, cin, cout, down in _SNGAN_BLOCKS[arch]:
        hidden = cout if kind == "opt" else cin
        shapes[f"{name}.c1"] = (hidden, cin, 3, 3)
        shapes[f"{name}.c2"] = (cout, hidden, 3, 3)
        if kind == "opt" or cin != cout or down:
            shapes[f"{name}.c_sc"] = (cout, cin, 1, 1)
    head = "l5" if arch == 32 else "l6"
    shapes[head] = (1, 128 if arch == 32 else 1024)
    assert list(shapes.keys()) == sngan_layer_keys(arch)
    return shapes


def sngan_state_dict(arch: int, seed: int = 1) -> dict:
    """Xavier-uniform weights (gain sqrt2 for 3x3, 1 for 1x1 / linear), uniform bias, sn_u ~ N(0,1)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for key, shape in sngan_shapes(arch).items():
        fan_in = int(np.prod(shape[1:]))
        fan_out = shape[0] * (int(np.prod(shape[2:])) if len(shape) == 4 else 1)
        gain = math.sqrt(2.0) if (len(shape) == 4 and shape[2] == 3) else 1.0
        bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
        sd[f"{key}.weight"] = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        bb = 1.0 / math.sqrt(fan_in)
        sd[f"{key}.bias"] = torch.from_numpy(rng.uniform(-bb, bb, (shape[0],)).astype(np.float32))
        sd[f"{key}.sn_u"] = torch.from_numpy(rng.standard_normal((1, shape[0])).astype(np.float32))
        sd[f"{key}.sn_sigma"] = torch.ones(1)
    return sd


_SG2_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


def stylegan2_state_dict(size: int, seed: int = 1) -> dict:
    """Random-init StyleGANDiscriminator(size) state_dict with the reference's key names (diagan/models/stylegan2.py:
    619-657): N(0,1) weights (equalised learning rate scales them at run time), small non-zero biases."""
    rng = np.random.RandomState(seed)
    r = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    sd = {"convs.0.0.weight": r(_SG2_CHANNELS[size], 3, 1, 1), "convs.0.1.bias": 0.1 * r(_SG2_CHANNELS[size])}
    cin, res, i = _SG2_CHANNELS[size], size, 1
    while res > 4:
        cout = _SG2_CHANNELS[res // 2]
        sd[f"convs.{i}.conv1.0.weight"] = r(cin, cin, 3, 3)
        sd[f"convs.{i}.conv1.1.bias"] = 0.1 * r(cin)
        sd[f"convs.{i}.conv2.1.weight"] = r(cout, cin, 3, 3)
        sd[f"convs.{i}.conv2.2.bias"] = 0.1 * r(cout)
        sd[f"convs.{i}.skip.1.weight"] = r(cout, cin, 1, 1)
        cin, res, i = cout, res // 2, i + 1
    sd["final_conv.0.weight"] = r(512, 513, 3, 3)
    sd["final_conv.1.bias"] = 0.1 * r(512)
    sd["final_linear.0.weight"] = r(512, 8192)
    sd["final_linear.0.bias"] = 0.1 * r(512)
    sd["final_linear.1.weight"] = r(1, 512)
    sd["final_linear.1.bias"] = 0.1 * r(1)
    return sd


def stylegan2_flops(size: int) -> float:
    """2 x MAC per sample of StyleGANDiscriminator(size) in the reference formulation (convs + linears; the blur FIRs,
    ~0.6 % more, are not counted: SURVEY 8(a) appendix)."""
    mac = size * size * 3 * _SG2_CHANNELS[size]
    cin, res = _SG2_CHANNELS[size], size
    while res > 4:
        cout = _SG2_CHANNELS[res // 2]
        mac += res * res * cin * cin * 9 + (res // 2) ** 2 * cout * cin * 10
        cin, res = cout, res // 2
    mac += 16 * 512 * 513 * 9 + 8192 * 512 + 512
    return 2.0 * mac


def dcgan_state_dict(seed: int = 1) -> dict:
    """Random-init MNIST_DCGAN_Discriminator state_dict (diagan/models/mnist.py:161-192, key names of its nn.Sequential):
    PyTorch-default conv init U(+-1/sqrt(fan_in)), BatchNorm affine and running statistics away from the identity."""
    rng = np.random.RandomState(seed)
    u = lambda lo, hi, *shape: torch.from_numpy(rng.uniform(lo, hi, shape).astype(np.float32))
    sd = {}
    for ci, bi, cin, cout in ((0, None, 3, 16), (3, 4, 16, 32), (7, 8, 32, 64), (11, 12, 64, 128), (15, 16, 128, 256),
                              (19, 20, 256, 512)):
        b = 1.0 / math.sqrt(cin * 9)
        sd[f"conv.{ci}.weight"] = u(-b, b, cout, cin, 3, 3)
        if bi is not None:
            sd[f"conv.{bi}.weight"], sd[f"conv.{bi}.bias"] = u(0.8, 1.2, cout), u(-0.1, 0.1, cout)
            sd[f"conv.{bi}.running_mean"], sd[f"conv.{bi}.running_var"] = u(-0.05, 0.05, cout), u(0.5, 1.5, cout)
    b = 1.0 / math.sqrt(8192)
    sd["out_d.weight"], sd["out_d.bias"] = u(-b, b, 1, 8192), u(-b, b, 1)
    return sd


def perturb_(state_dict: dict, step: int, scale: float = 1e-3, device=None) -> dict:
    """W += scale * randn(seed = step): a deterministic stand-in for the training between two recording
    passes, so that per-sample logits move and std > 0."""
    gen = torch.Generator(device=device or "cpu").manual_seed(int(step))
    out = dict(state_dict)
    for k in sorted(state_dict):
        if k.endswith(".weight"):
            w = state_dict[k]
            out[k] = w + scale * torch.randn(w.shape, generator=gen, device=w.device, dtype=w.dtype)
    return out


def uniform_images_u8(n: int, size: int, seed: int = 1, device="cpu", pin: bool = False) -> torch.Tensor:
    """uint8 [n,size,size,3] ~ U{0..255} (CIFAR-10 shape for size 32, CelebA for 64)."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (n, size, size, 3), dtype=torch.uint8, generator=gen)
    if pin:
        x = x.pin_memory()
    return x.to(device) if device != "cpu" else x
