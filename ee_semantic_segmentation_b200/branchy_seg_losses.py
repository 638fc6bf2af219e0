"""Drop-in for the hot-path part of the reference's branchy_seg_losses.py: the multi-exit wrapper
`BrSegLoss` (:9-38) and the branchy `LovaszSoftmax` (:133-159). All exits are handled by one call
into csrc/lovasz.cu instead of a Python loop of per-class torch.sort calls.
Dice/Jaccard/Tversky/FocalTversky/Focal bodies are out of scope."""
import torch as tch
from torch import nn

from . import ops
from .new_seg_losses import SegLoss


class BrSegLoss(SegLoss):
    def __init__(self, smooth=1e-6, reduction='mean', n_branches=0, weights=None):
        super().__init__(smooth, reduction)
        self.n = n_branches + 1
        if weights and len(weights) == n_branches + 1:
            self.weights = tch.tensor(weights, requires_grad=True)
        else:
            self.weights = tch.ones(self.n, requires_grad=True)

    def update_n(self, n):
        self.n = n + 1

    def _compute_loss(self, y_pred, targets):
        pass

    def forward(self, y_pred, targets):
        losses = [self._compute_loss(y_pred[i], targets).unsqueeze(0) for i in range(self.n)]
        losses = tch.cat(losses)
        dim = list(range(1, len(losses.shape)))
        if self.reduction == 'mean':
            losses = losses.mean(dim=dim)
        elif self.reduction == 'sum':
            losses = losses.sum(dim=dim)
        else:
            return losses
        return tch.dot(self.weights.to(device=losses.device), losses)


class LovaszSoftmax(nn.Module):
    """sum_i w_i * lovasz_softmax(y_pred[i], targets); w = linspace(0,1,E+1)[1:] when prev_out.
    Raw logits are passed as 'probas' exactly like the reference (no softmax on this path)."""

    def __init__(self, classes='present', per_image=False, ignore=None, n_branches=0, prev_out=False):
        super().__init__()
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore
        self.n = n_branches + 1
        self.prev_out = prev_out
        if self.prev_out:
            self.weights = tch.linspace(0, 1, self.n + 1, requires_grad=True)[1:]

    def update_n(self, n):
        self.n = n + 1
        if self.prev_out:
            self.weights = tch.linspace(0, 1, self.n + 1, requires_grad=True)[1:]

    def forward(self, y_pred, targets):
        if y_pred.shape[0] < self.n:
            raise IndexError(f'index {self.n - 1} is out of bounds for dimension 0 with size {y_pred.shape[0]}')
        losses = ops.lovasz_multi_exit(y_pred[:self.n], targets, classes=self.classes,
                                       per_image=self.per_image, ignore=self.ignore)
        if self.prev_out:
            if self.weights.device != losses.device:
                self.weights = self.weights.to(losses.device)
            return tch.dot(self.weights, losses).sum()
        return losses.sum()
