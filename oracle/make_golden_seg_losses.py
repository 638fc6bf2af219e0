"""TEST INFRASTRUCTURE — generates tests/golden/seg_losses.npz by running the UNMODIFIED reference new_seg_losses.py:
its own __main__ demo (inputs extracted with runpy, printed known answers 0.0504 / 0.4033, new_seg_losses.py:170-256) and
DiceLoss / JaccardLoss / TverskyLoss / FocalTverskyLoss on seeded inputs (values and autograd gradients).

Run in the build container only:  python -m oracle.make_golden_seg_losses"""
import contextlib
import io
import os
import runpy

import numpy as np
import torch

from oracle import ref_import
from oracle.make_golden import OUT, blocky_labels


def main():
    nsl = ref_import.load("new_seg_losses")
    out = {}
    with ref_import.reference_on_path():
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            gl = runpy.run_path(os.path.join(ref_import.REFERENCE_DIR, "new_seg_losses.py"), run_name="__main__")
    out["demo_y_pred"], out["demo_y_true"] = gl["y_pred"].numpy(), gl["y_true"].numpy()
    out["demo_stdout"] = np.array(buf.getvalue())
    yp, yt = gl["y_pred"], gl["y_true"]
    out["demo_jaccard_mean"] = nsl.JaccardLoss()(yp, yt).numpy()
    out["demo_jaccard_sum"] = nsl.JaccardLoss(reduction="sum")(yp, yt).numpy()
    out["demo_dice_mean"] = nsl.DiceLoss()(yp, yt).numpy()
    g = torch.Generator().manual_seed(2024)
    N, C, H, W = 3, 6, 11, 15
    y = torch.randn(N, C, H, W, generator=g) * 2
    tv = blocky_labels(g, N, C, H, W, void_frac=0.1, cell=3)          # void == C: Dice / Jaccard drop it
    t = blocky_labels(g, N, C, H, W, void_frac=0.0, cell=3)           # Tversky: one_hot(num_classes=C)
    out.update(y_pred=y.numpy(), targets_void=tv.numpy(), targets=t.numpy())
    cases = {
        "dice_mean": (nsl.DiceLoss(), tv), "dice_index_sum": (nsl.DiceLoss(reduction="sum", index=True), tv),
        "dice_batchwise": (nsl.DiceLoss(reduction="mean_batchwise"), t),
        "jaccard_mean": (nsl.JaccardLoss(), tv), "jaccard_bg": (nsl.JaccardLoss(downgrad_bg=0.25, reduction="sum"), tv),
        "jaccard_nobg": (nsl.JaccardLoss(downgrad_bg=0.0, reduction="sum_batchwise"), tv),
        "jaccard_index": (nsl.JaccardLoss(index=True, reduction="none"), t),
        "tversky_mean": (nsl.TverskyLoss(alpha=0.3, beta=0.7), t),
        "ftversky_sum": (nsl.FocalTverskyLoss(alpha=0.7, beta=0.3, gamma=4 / 3, reduction="sum"), t),
    }
    for tag, (fn, tg) in cases.items():
        yy = y.clone().requires_grad_(True)
        l = fn(yy, tg)
        l.sum().backward()
        out[f"{tag}_loss"], out[f"{tag}_grad"] = l.detach().numpy(), yy.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "seg_losses.npz"), **out)
    print(str(out["demo_stdout"]))
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items() if k.endswith("_loss") or k.startswith("demo_j")})


if __name__ == "__main__":
    main()
