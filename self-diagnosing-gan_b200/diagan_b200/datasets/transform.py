"""Device-side form of ``diagan.datasets.transform`` (diagan-pkg/diagan/datasets/transform.py:3-41).

Every reference transform is ``Resize(s)``, ``CenterCrop(s)``, ``ToTensor()``, ``Normalize(0.5, 0.5)`` with s = 32 (cifar10,
color_mnist, mnist_fmnist) or 64 (celeba), applied to one PIL image at a time inside DataLoader workers -- for every item of
every recording pass.  Here the raw uint8 dataset is transformed ONCE on the GPU (``sdg_resize_center_crop_u8``, bit-exact
with Pillow's bilinear resample) and stays resident at the network's input size; ToTensor + Normalize happen inside the first
conv's operand load.  ``get_transform(name)`` keeps the reference's entry point and dataset names.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr

IMG_SIZE = {"cifar10": 32, "celeba": 64, "color_mnist": 32, "mnist_fmnist": 32}


class DeviceTransform:
    """uint8 [n,H,W,C] CUDA tensor -> uint8 [n,size,size,C] CUDA tensor (Resize + CenterCrop of the reference transform)."""

    def __init__(self, img_size: int):
        self.img_size = int(img_size)

    def __call__(self, images: torch.Tensor, batch: int = 16384) -> torch.Tensor:
        lib = _lib.load()
        if not images.is_cuda:
            raise _lib.SdgError("DeviceTransform needs a CUDA tensor (diagan_b200 has no CPU path)")
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] not in (1, 3):
            raise _lib.SdgError(f"DeviceTransform expects uint8 [n,H,W,C] with C in (1, 3), got {images.dtype} {tuple(images.shape)}")
        images = images.contiguous()
        n, H, W, C = images.shape
        s = self.img_size
        out = torch.empty((n, s, s, C), dtype=torch.uint8, device=images.device)
        with torch.cuda.device(images.device):
            for lo in range(0, n, batch):
                hi = min(n, lo + batch)
                check(lib.sdg_resize_center_crop_u8(ptr(images[lo:hi]), hi - lo, H, W, C, s, ptr(out[lo:hi]),
                                                    stream_ptr(images.device)), "sdg_resize_center_crop_u8")
        return out

    def from_host(self, images_u8: np.ndarray, device, batch: int = 8192) -> torch.Tensor:
        """Raw dataset in host memory (e.g. CelebA 218x178: 23 GB for 202 599 images) -> resident transformed dataset,
        streamed through the GPU in batches so the raw images never have to fit in HBM at once."""
        n = images_u8.shape[0]
        s = self.img_size
        out = torch.empty((n, s, s, images_u8.shape[3]), dtype=torch.uint8, device=device)
        for lo in range(0, n, batch):
            hi = min(n, lo + batch)
            chunk = torch.from_numpy(np.ascontiguousarray(images_u8[lo:hi])).to(device, non_blocking=True)
            out[lo:hi] = self(chunk)
        return out


def get_transform(dataset_name: str) -> DeviceTransform:
    """Same names as ``diagan.datasets.transform.get_transform`` (transform.py:35-41)."""
    return DeviceTransform(IMG_SIZE[dataset_name])
