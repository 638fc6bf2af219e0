"""TEST INFRASTRUCTURE — generates tests/golden/overlap_losses.npz by running the UNMODIFIED reference
branchy_seg_losses.py (DiceLoss, JaccardLoss, TverskyLoss, FocalTverskyLoss) from /root/reference on seeded inputs.

Run in the build container only:  python -m oracle.make_golden_overlap
(kept separate from make_golden.py so the other committed fixtures are not regenerated)."""
import os

import numpy as np
import torch

from oracle import ref_import
from oracle.make_golden import OUT, blocky_labels


def main():
    bsl = ref_import.load("branchy_seg_losses")
    g = torch.Generator().manual_seed(1301)
    E, N, C, H, W = 3, 2, 7, 13, 17
    y = torch.randn(E, N, C, H, W, generator=g) * 2
    tgt = blocky_labels(g, N, C, H, W, void_frac=0.0, cell=3)            # no void: Dice / Tversky one_hot(C)
    tgt_void = blocky_labels(g, N, C, H, W, void_frac=0.1, cell=3)       # void == C: Jaccard drops it
    out = {"y_pred": y.numpy(), "targets": tgt.numpy(), "targets_void": tgt_void.numpy()}
    cases = {
        "dice_mean": (bsl.DiceLoss(n_branches=2), tgt),
        "dice_sum_w": (bsl.DiceLoss(reduction="sum", n_branches=2, weights=[0.5, 1.0, 2.0]), tgt),
        "jaccard_mean": (bsl.JaccardLoss(n_branches=2), tgt),
        "jaccard_void_bg": (bsl.JaccardLoss(n_branches=2, downgrad_bg=0.3), tgt_void),
        "jaccard_nobg_sum": (bsl.JaccardLoss(reduction="sum", n_branches=1, downgrad_bg=0.0), tgt_void),
    }
    for tag, (fn, t) in cases.items():
        yy = y.clone().requires_grad_(True)
        l = fn(yy, t)
        l.backward()
        out[f"{tag}_loss"], out[f"{tag}_grad"] = l.detach().numpy(), yy.grad.numpy()
    out["dice_none"] = bsl.DiceLoss(reduction="none", n_branches=2)(y, tgt).detach().numpy()
    out["jaccard_none"] = bsl.JaccardLoss(reduction="none", n_branches=2)(y, tgt_void).detach().numpy()
    out["tversky_mean"] = bsl.TverskyLoss(alpha=0.3, beta=0.7, n_branches=2)(y, tgt).detach().numpy()
    out["tversky_none"] = bsl.TverskyLoss(reduction="none", n_branches=2)(y, tgt).detach().numpy()
    out["focal_tversky_mean"] = bsl.FocalTverskyLoss(gamma=0.75, n_branches=2)(y, tgt).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "overlap_losses.npz"), **out)
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items() if "grad" not in k and k not in ("y_pred",)})


if __name__ == "__main__":
    main()
