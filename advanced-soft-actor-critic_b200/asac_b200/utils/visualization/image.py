"""``ImageVisual``: debugging aid plugin files construct in ``ModelRep._build_model`` and call on image
batches (reference: algorithm/utils/visualization/image.py:8-74).  matplotlib is imported on first use;
without it the call is a no-op with one warning (the image has no matplotlib)."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:  # noqa: BLE001
        return None


class ImageVisual:
    def __init__(self, model_abs_dir: Path | None = None) -> None:
        self.model_abs_dir = model_abs_dir
        self.fig = None
        self.idx = 0
        self._warned = False

    def __call__(self, *images: np.ndarray | torch.Tensor, max_batch=5, range_min=0, range_max=1,
                 save_name: str | None = None):
        """images: ``[batch, C, H, W]`` tensors or ``[batch, H, W, C]`` arrays (a sequence axis shows its last
        step); one column per image, at most ``max_batch`` rows."""
        plt = _plt()
        if plt is None:
            if not self._warned:
                logging.getLogger('visualization').warning('matplotlib is not installed: ImageVisual is a no-op')
                self._warned = True
            return
        shown = []
        for im in images:
            im = im[:, -1] if len(im.shape) > 4 else im
            im = im[:max_batch]
            shown.append(im.detach().cpu().numpy().transpose(0, 2, 3, 1) if isinstance(im, torch.Tensor) else im)
        cols = len(shown)
        if self.fig is None:
            self.fig, self.axes = plt.subplots(nrows=max_batch, ncols=cols, squeeze=False,
                                               figsize=(3 * cols, 3 * max_batch))
            self.ims = [[None] * cols for _ in range(max_batch)]
            for ax in self.axes.flat:
                ax.axis('off')
        for r in range(min(shown[0].shape[0], max_batch)):
            for c, im in enumerate(shown):
                frame = im[r] if im.shape[-1] != 1 else im[r, ..., 0]
                if self.ims[r][c] is None:
                    self.ims[r][c] = self.axes[r][c].imshow(frame, vmin=range_min, vmax=range_max,
                                                            cmap=None if frame.ndim == 3 else 'gray')
                else:
                    self.ims[r][c].set_data(frame)
        if save_name is not None and self.model_abs_dir is not None:
            out = Path(self.model_abs_dir) / 'image_visual'
            out.mkdir(parents=True, exist_ok=True)
            self.fig.savefig(out / f'{save_name}_{self.idx}.png')
            self.idx += 1
        else:
            self.fig.canvas.draw_idle()
            self.fig.canvas.flush_events()
