"""TEST INFRASTRUCTURE — generates tests/golden/focal_loss.npz by running the UNMODIFIED reference
branchy_seg_losses.FocalLoss from /root/reference on seeded inputs (values and autograd gradients).

Run in the build container only:  python -m oracle.make_golden_focal"""
import os

import numpy as np
import torch

from oracle import ref_import
from oracle.make_golden import OUT, blocky_labels


def main():
    bsl = ref_import.load("branchy_seg_losses")
    g = torch.Generator().manual_seed(1777)
    E, N, C, H, W = 3, 2, 7, 13, 17
    y = torch.randn(E, N, C, H, W, generator=g) * 2
    tgt = blocky_labels(g, N, C, H, W, void_frac=0.0, cell=3)            # [N,1,H,W], no void: gather(1, targets)
    alpha = torch.rand(C, generator=g) + 0.25
    out = {"y_pred": y.numpy(), "targets": tgt.numpy(), "alpha": alpha.numpy()}
    cases = {
        "g2_mean": (bsl.FocalLoss(n_branches=2), y, tgt),
        "g15_sum_w": (bsl.FocalLoss(gamma=1.5, reduction="sum", n_branches=2, weights=[0.5, 1.0, 2.0]), y, tgt),
        "g0_mean": (bsl.FocalLoss(gamma=0, n_branches=1), y, tgt),
        "g05_mean": (bsl.FocalLoss(gamma=0.5, n_branches=2), y, tgt),
        "alpha_mean": (bsl.FocalLoss(alpha=alpha, n_branches=2), y, tgt),                   # [N,N,H,W] broadcast
        "alpha_sum": (bsl.FocalLoss(alpha=alpha, gamma=1, reduction="sum", n_branches=2), y, tgt),
        "alpha_n1_mean": (bsl.FocalLoss(alpha=alpha, n_branches=2), y[:, :1], tgt[:1]),
    }
    for tag, (fn, yy, t) in cases.items():
        yy = yy.clone().requires_grad_(True)
        l = fn(yy, t)
        l.backward()
        out[f"{tag}_loss"], out[f"{tag}_grad"] = l.detach().numpy(), yy.grad.numpy()
    out["g2_none"] = bsl.FocalLoss(reduction="none", n_branches=2)(y, tgt).detach().numpy()
    yy = y.clone().requires_grad_(True)
    ln = bsl.FocalLoss(alpha=alpha, reduction="none", n_branches=2)(yy, tgt)
    up = torch.rand(ln.shape, generator=g)
    (ln * up).sum().backward()
    out["alpha_none"], out["alpha_none_up"], out["alpha_none_grad"] = ln.detach().numpy(), up.numpy(), yy.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "focal_loss.npz"), **out)
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items() if "grad" not in k and k not in ("y_pred",)})


if __name__ == "__main__":
    main()
