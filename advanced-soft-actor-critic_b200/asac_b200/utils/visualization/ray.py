"""``RayVisual``: debugging aid for ray-cast observations (reference:
algorithm/utils/visualization/ray.py:8-79).  matplotlib is imported on first use; without it the call
is a no-op with one warning."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from .image import _plt


class RayVisual:
    def __init__(self, model_abs_dir: Path | None = None) -> None:
        self.model_abs_dir = model_abs_dir
        self.fig = None
        self.idx = 0
        self._warned = False

    def __call__(self, *rays: np.ndarray | torch.Tensor, max_batch=5, save_name: str | None = None):
        """rays: ``[batch, ray_size, C]`` with ``ray[..., -1]`` the hit fraction (1 = no hit); one polar plot
        per input and batch row."""
        plt = _plt()
        if plt is None:
            if not self._warned:
                logging.getLogger('visualization').warning('matplotlib is not installed: RayVisual is a no-op')
                self._warned = True
            return
        shown = []
        for ray in rays:
            ray = ray[:, -1] if len(ray.shape) > 3 else ray
            ray = ray[:max_batch]
            shown.append(ray.detach().cpu().numpy() if isinstance(ray, torch.Tensor) else ray)
        cols = len(shown)
        if self.fig is None:
            self.fig, self.axes = plt.subplots(nrows=max_batch, ncols=cols, squeeze=False,
                                               subplot_kw={'projection': 'polar'},
                                               figsize=(3 * cols, 3 * max_batch))
        for r in range(min(shown[0].shape[0], max_batch)):
            for c, ray in enumerate(shown):
                ax = self.axes[r][c]
                ax.clear()
                n = ray.shape[1]
                ax.scatter(np.linspace(0, 2 * np.pi, n, endpoint=False), ray[r, :, -1], s=4)
                ax.set_ylim(0, 1)
        if save_name is not None and self.model_abs_dir is not None:
            out = Path(self.model_abs_dir) / 'ray_visual'
            out.mkdir(parents=True, exist_ok=True)
            self.fig.savefig(out / f'{save_name}_{self.idx}.png')
            self.idx += 1
        else:
            self.fig.canvas.draw_idle()
            self.fig.canvas.flush_events()
