"""Drop-in for the reference's new_seg_losses.py: the `SegLoss` base class (:8-32), the single-output
`LovaszSoftmax` (:159-168) and the soft-overlap family `DiceLoss` (:34-56), `JaccardLoss` (:58-88), `TverskyLoss`
(:90-111), `FocalTverskyLoss` (:113-121) — one streaming pass of csrc/soft_overlap.cu over the logits gives
S_pt = sum p*t, S_p = sum p, S_t = sum t per image and class, and the reference's formulas run on those [N,C] tensors
(no softmax / one-hot tensors of the image size). `FocalLoss` / `HybridFocalLoss` (:123-157: a batch-mean nll_loss
scalar times a per-pixel modulating factor) are not provided."""
import torch as tch
from torch import nn

from .lovaszsoftmax import lovasz_softmax


class SegLoss(nn.Module):
    """Same reduction dispatch as new_seg_losses.py:17-32."""

    def __init__(self, smooth=1e-6, reduction='mean'):
        super().__init__()
        self.smooth = smooth
        self.reduction = reduction

    def _compute_loss(self, y_pred, targets):
        pass

    def forward(self, y_pred, targets):
        loss = self._compute_loss(y_pred, targets)
        if self.reduction == 'mean':
            return loss.mean()
        if self.reduction == 'mean_batchwise':
            return loss.mean(dim=list(range(1, loss.dim())))
        if self.reduction == 'sum_batchwise':
            return loss.sum(dim=list(range(1, loss.dim())))
        if self.reduction == 'sum':
            return loss.sum()
        return loss


def _overlap(y_pred, targets):
    """(S_pt, S_p, S_t) as [N,C] tensors; labels outside [0,C) (void) match no class, as after the reference's
    `targets[:, :C, :]` slice (:47-48)."""
    from . import ops
    s_pt, s_p, s_t = ops.soft_overlap_sums(y_pred.unsqueeze(0), targets)
    return s_pt[0], s_p[0], s_t


class DiceLoss(SegLoss):
    def __init__(self, smooth=1e-6, reduction='mean', index=False):
        super().__init__(smooth, reduction)
        self.index = index

    def _compute_loss(self, y_pred, targets):
        s_pt, s_p, s_t = _overlap(y_pred, targets)
        num = 2 * s_pt.sum(dim=1) + self.smooth
        den = (s_p + s_t).sum(dim=1) + self.smooth
        return num / den if self.index else 1 - num / den


class JaccardLoss(DiceLoss):
    def __init__(self, smooth=1e-6, reduction='mean', index=False, downgrad_bg=1.):
        super().__init__(smooth, reduction, index)
        self.downgrad_bg = downgrad_bg if 0 <= downgrad_bg <= 1. else 1.

    def _compute_loss(self, y_pred, targets):
        intersection, s_p, s_t = _overlap(y_pred, targets)
        union = s_p + s_t - intersection
        IoU = (intersection + self.smooth) / (union + self.smooth)
        if self.index:
            return IoU
        if self.downgrad_bg:
            scale = tch.ones(IoU.shape[-1], dtype=IoU.dtype, device=IoU.device)
            scale[0] = self.downgrad_bg
            return (1 - IoU) * scale
        return (1 - IoU).sum(dim=-1)


class TverskyLoss(SegLoss):
    """Soft Tversky index per image and class (:96-108): TP = sum p*t, FP = sum p*(1-t), FN = sum (1-p)*t."""

    def __init__(self, smooth=1e-6, alpha=.5, beta=.5, reduction='mean'):
        super().__init__(smooth, reduction)
        self.alpha = alpha
        self.beta = beta

    def _forward_imp(self, y_pred, targets):
        TP, s_p, s_t = _overlap(y_pred, targets)
        if float(s_t.sum()) != float(targets.numel()):        # F.one_hot(num_classes=C), :101
            raise RuntimeError("Class values must be smaller than num_classes.")
        FP, FN = s_p - TP, s_t - TP
        return 1 - (TP + self.smooth) / (TP + self.alpha * FP + self.beta * FN + self.smooth)

    def _compute_loss(self, y_pred, targets):
        return self._forward_imp(y_pred, targets)


class FocalTverskyLoss(TverskyLoss):
    def __init__(self, smooth=1e-6, alpha=.5, beta=.5, gamma=1., reduction='mean'):
        super().__init__(smooth, alpha, beta, reduction)
        self.gamma = gamma

    def _compute_loss(self, y_pred, targets):
        return self._forward_imp(y_pred, targets) ** (1 / self.gamma)


class LovaszSoftmax(nn.Module):
    def __init__(self, classes='present', per_image=False, ignore=None):
        super().__init__()
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore

    def forward(self, y_pred, targets):
        return lovasz_softmax(y_pred, targets, classes=self.classes, per_image=self.per_image,
                              ignore=self.ignore)
