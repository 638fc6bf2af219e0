"""Drop-in for the reference's compute_mIoU.py: dataset-level `mIoU` (:7-36) and per-image
`img_mIoU` (:38-63), built on the confusion-histogram kernel.

`mIoU` keeps BOTH the reference's float32 [3,C] accumulator (per-call adds of the int64 batch sums,
so `compute()` is bit-identical to the reference including its fp32 saturation at 2^24 and its NaN
for absent classes) and an exact int64 confusion matrix (`cm`, `compute_exact()`), which is what
the multi-GPU all-reduce uses."""
import numpy as np
import torch as tch

from . import ops
from .seg_metrics import SegMetric


class mIoU(SegMetric):
    def __init__(self, n_classes, device='cpu'):
        super().__init__()
        self.C = n_classes
        self.accumulator = tch.zeros((3, self.C))
        self.cm = None  # exact int64 [C+1, C], lives on the prediction device

    def forward(self, y_pred, targets):
        if y_pred.device != self.accumulator.device:
            self.accumulator = self.accumulator.to(y_pred.device)
        C = y_pred.shape[1]
        assert C == self.accumulator.shape[1]
        cm = self._confusion(y_pred, targets).sum(dim=0)
        self.add_confusion(cm)

    def add_confusion(self, cm):
        """cm int64 [C+1,C] of one call (all images of the call summed)."""
        if cm.device != self.accumulator.device:
            self.accumulator = self.accumulator.to(cm.device)
        TP, FP, FN = ops.basics_from_cm(cm)
        self.accumulator[0] += TP
        self.accumulator[1] += FP
        self.accumulator[2] += FN
        self.cm = cm.clone() if self.cm is None else self.cm + cm.to(self.cm.device)

    def compute(self):
        # the [3,C] float32 accumulator is brought to the host first so the few-element reduction
        # is the same arithmetic, in the same order, as the reference's CPU run
        self.accumulator = self.accumulator.cpu()
        den = self.accumulator.sum(dim=0)
        cIoU = tch.div(self.accumulator[0], den)
        cIoU[cIoU == float('nan')] = 1.   # never matches (compute_mIoU.py:35): absent class -> NaN
        return (cIoU.sum() / self.C).cpu()

    def compute_exact(self):
        """mIoU from the exact integer confusion matrix, float64, same NaN rule."""
        if self.cm is None:
            return tch.tensor(float('nan'), dtype=tch.float64)
        TP, FP, FN = (t.double().cpu() for t in ops.basics_from_cm(self.cm))
        return (TP / (TP + FP + FN)).sum() / self.C


class img_mIoU(SegMetric):
    """Per-image mIoU over the classes present in the target (void label included when present),
    compute_mIoU.py:43-58: IoU_i = |gt_i & pred_i| / |gt_i | pred_i| = CM[i][i] / (row_i + col_i - CM[i][i])."""

    def __init__(self):
        super().__init__()
        self.accumulator = [0, 0]

    def forward(self, y_pred, target):
        if len(y_pred.shape) == 4:
            C = y_pred.shape[1]
            pred = y_pred  # argmax happens inside the histogram kernel
            tgt = target.reshape(1, -1)
            pred = pred.reshape(1, C, -1) if y_pred.shape[0] == 1 else \
                pred.permute(1, 0, 2, 3).reshape(1, C, -1)
        else:
            pred = y_pred.reshape(1, -1).to(tch.int64)
            tgt = target.reshape(1, -1)
            C = int(max(pred.max().item(), 0)) + 1
        tgt = tgt.to(tch.int64)
        tmax = int(tgt.max().item())
        K = max(C, tmax + 1)           # classes are taken from the target, void label included
        if pred.dim() == 3 and K > C:  # pad logits so the void label has a (never predicted) column
            pad = tch.full((1, K - C, pred.shape[2]), float('-inf'), dtype=pred.dtype, device=pred.device)
            pred = tch.cat([pred, pad], dim=1)
        cm = ops.confusion_hist(pred, tgt, K)[0].double()   # [K+1, K]
        sq = cm[:K, :]
        inter = tch.diagonal(sq)
        union = sq.sum(dim=1) + cm.sum(dim=0) - inter
        present = sq.sum(dim=1) > 0
        iou = (inter[present].float() / union[present].float())
        self.accumulator[0] += (iou.sum() / present.sum()).item()
        self.accumulator[1] += 1

    def compute(self):
        if self.accumulator[1] <= 0:
            return np.nan
        return self.accumulator[0] / self.accumulator[1]
