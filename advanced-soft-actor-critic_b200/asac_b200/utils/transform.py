"""Image augmentations plugin files pass to ``m.Transform`` (reference: algorithm/utils/transform.py:5-112).
Tensor inputs ``[batch, C, H, W]`` in [0, 1]; PIL inputs are converted and converted back."""
from __future__ import annotations

import torch

__all__ = ['GaussianNoise', 'SaltAndPepperNoise', 'DepthNoise', 'DepthSaltAndPepperNoise']


class _Augment:
    def __call__(self, img):
        if isinstance(img, torch.Tensor):
            return self.apply(img)
        from torchvision import transforms as T
        return T.ToPILImage()(self.apply(T.ToTensor()(img)))

    def apply(self, img: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


def _uniform(shape, like: torch.Tensor) -> torch.Tensor:
    return torch.rand(shape, dtype=torch.float32, device=like.device)


def _salt_pepper(img, mask, amount, noise_p, signal_p):
    img = torch.where(mask < noise_p / 2., (img + amount).clamp(0., 1.), img)
    return torch.where(mask > noise_p / 2. + signal_p, (img - amount).clamp(0., 1.), img)


class GaussianNoise(_Augment):
    """img + U[0,1) * std + mean, clamped (the reference draws ``torch.rand``, :16)."""

    def __init__(self, mean=0., std=.1):
        self.std, self.mean = std, mean

    def apply(self, img):
        return (img + _uniform(img.shape, img) * self.std + self.mean).clamp(0., 1.)


class SaltAndPepperNoise(_Augment):
    """Per pixel (shared by the channels): with probability (1-p)/2 each, +snr / -snr."""

    def __init__(self, snr=.3, p=.9):
        self.snr, self.p = snr, p

    def apply(self, img):
        batch, c, h, w = img.shape
        mask = _uniform((batch, 1, h, w), img).repeat(1, c, 1, 1)
        return _salt_pepper(img, mask, self.snr, 1 - self.p, self.p)


class DepthNoise(_Augment):
    """One uniform offset in ``p`` (or (-p, p)) for the whole batch."""

    def __init__(self, p):
        self.p = p if isinstance(p, tuple) else (-p, p)

    def apply(self, img):
        lo, hi = self.p
        return (img + (_uniform(1, img) * (hi - lo) + lo)).clip(0., 1.)


class DepthSaltAndPepperNoise(_Augment):
    """Per element: with probability p/2 each, +snr / -snr."""

    def __init__(self, snr=1., p=0.03):
        self.snr, self.p = snr, p

    def apply(self, img):
        return _salt_pepper(img, _uniform(img.shape, img), torch.tensor(self.snr), self.p, 1 - self.p)
