"""Drop-in for the hot-path part of the reference's new_seg_losses.py: the `SegLoss` base class
(new_seg_losses.py:8-32) and the single-output `LovaszSoftmax` (new_seg_losses.py:159-168).
The Dice/Jaccard/Tversky/Focal/Hybrid alternatives are out of scope."""
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


class LovaszSoftmax(nn.Module):
    def __init__(self, classes='present', per_image=False, ignore=None):
        super().__init__()
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore

    def forward(self, y_pred, targets):
        return lovasz_softmax(y_pred, targets, classes=self.classes, per_image=self.per_image,
                              ignore=self.ignore)
