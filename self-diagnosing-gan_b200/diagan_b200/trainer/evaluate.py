"""Drop-in for the ``DRS`` wrapper of ``diagan.trainer.evaluate`` (diagan-pkg/diagan/trainer/evaluate.py:26-95):
fixed 80th percentile, ``batch_size`` constructor argument, ``generate_images`` returns on ``device``
(evaluate.py:70-83), optional per-call ``gamma`` (evaluate.py:52)."""
from __future__ import annotations

import os

from ..models.drs import DRS as _DRS


class DRS(_DRS):
    def __init__(self, netG, netD, device, batch_size=256):
        super().__init__(netG, netD, device, gamma=None, percentile=80, batch_size=batch_size)

    def generate_images(self, num_images, device=None):
        return super().generate_images(num_images, device=self.device if device is None else device)

    def visualize_images(self, log_dir, evaluate_step, num_images=64):
        """evaluate.py:85-95 (needs torchvision)."""
        import torchvision.utils as vutils
        img_dir = os.path.join(log_dir, 'images')
        os.makedirs(img_dir, exist_ok=True)
        fake = self.generate_images(num_images)
        grid = vutils.make_grid(fake, padding=2, normalize=True)
        vutils.save_image(grid, '{}/fake_samples_step_{}_after_drs.png'.format(img_dir, evaluate_step), normalize=True)
