"""Drop-in for the reference's seg_metrics.py: `SegMetric._compute_basics` (:13-28) on the
confusion-histogram kernel (csrc/confusion_hist.cu) and the derived Recall / Precision / F_beta /
Accuracy (:30-76), which are small host formulas on the [N,C] counts."""
import torch as tch

from . import ops
from .new_seg_losses import SegLoss


class SegMetric(SegLoss):
    def __init__(self, smooth=1e-6, reduction='mean', avg='macro'):
        super().__init__(smooth, reduction)
        self.avg = avg

    def _confusion(self, y_pred, targets):
        """int64 [N, C+1, C]; row C collects void / out-of-range targets."""
        return ops.confusion_hist(y_pred, targets, y_pred.shape[1])

    def _compute_basics(self, y_pred, targets):
        """TP, FP, FN int64 [N,C]; void pixels count as FP for the predicted class, exactly like the
        one-hot formulation at seg_metrics.py:17-27. No host synchronisation (the reference's
        targets.unique().item() at :15 is not needed)."""
        return ops.basics_from_cm(self._confusion(y_pred, targets))


class Recall(SegMetric):
    def _compute_loss(self, y_pred, targets):
        TP, _, FN = self._compute_basics(y_pred, targets)
        if self.avg == 'macro':
            return ((TP + self.smooth) / (TP + FN + self.smooth)).mean(dim=-1)
        if self.avg == 'micro':
            TP = TP.sum(dim=-1)
            FN = FN.sum(dim=-1)
        return (TP + self.smooth) / (TP + FN + self.smooth)


class Precision(SegMetric):
    def _compute_loss(self, y_pred, targets):
        TP, FP, _ = self._compute_basics(y_pred, targets)
        if self.avg == 'macro':
            return ((TP + self.smooth) / (TP + FP + self.smooth)).mean(dim=-1)
        if self.avg == 'micro':
            TP = TP.sum(dim=-1)
            FP = FP.sum(dim=-1)
        return (TP + self.smooth) / (TP + FP + self.smooth)


class F_beta(SegMetric):
    def __init__(self, beta=1, smooth=1e-6, reduction='mean', avg='macro'):
        super().__init__(smooth, reduction, avg)
        self.beta = beta

    def _compute_loss(self, y_pred, targets):
        TP, FP, FN = self._compute_basics(y_pred, targets)
        b2 = self.beta ** 2
        if self.avg == 'macro':
            return (((1 + b2) * TP + self.smooth) / ((1 + b2) * TP + b2 * FN + FP + self.smooth)).mean(dim=-1)
        if self.avg == 'micro':
            TP = TP.sum(dim=-1)
            FP = FP.sum(dim=-1)
            FN = FN.sum(dim=-1)
        return ((1 + b2) * TP + self.smooth) / ((1 + b2) * TP + b2 * FN + FP + self.smooth)


class Accuracy(SegMetric):
    def _compute_loss(self, y_pred, targets):
        # seg_metrics.py:68-76: #(target == argmax) / #pixels per image == trace of the C x C block
        cm = self._confusion(y_pred, targets)
        C = cm.shape[-1]
        hits = tch.diagonal(cm[:, :C, :], dim1=-2, dim2=-1).sum(dim=-1)
        return hits / cm.sum(dim=(-2, -1))
