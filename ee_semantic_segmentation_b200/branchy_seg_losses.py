"""Drop-in for the reference's branchy_seg_losses.py: the multi-exit wrapper `BrSegLoss` (:9-38), the branchy
`LovaszSoftmax` (:133-159; one call into csrc/lovasz.cu for all exits instead of a Python loop of per-class
torch.sort calls) and the overlap family `DiceLoss` (:40-48), `JaccardLoss` (:50-77), `TverskyLoss` (:79-103),
`FocalTverskyLoss` (:105-113): one streaming pass over the logits of ALL exits (csrc/soft_overlap.cu) produces
the per-exit / image / class sums, and the reference's formulas are applied verbatim on those small tensors —
instead of materialising softmax and int64 one-hot tensors [N,HW,C] per exit. Tversky works on the arg-max map
(no gradient reaches the logits, as in the reference) and comes from the confusion-matrix kernel.
`FocalLoss` (:115-131) is one streaming pass over all exits (csrc/focal.cu) with the gradient fused into it."""
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

    def _compute_all(self, y_pred, targets):
        """Per-exit losses of all exits at once, [E, ...] (subclasses on the eeseg kernels override this)."""
        return tch.cat([self._compute_loss(y_pred[i], targets).unsqueeze(0) for i in range(y_pred.shape[0])])

    def forward(self, y_pred, targets):
        if y_pred.shape[0] < self.n:
            raise IndexError(f'index {self.n - 1} is out of bounds for dimension 0 with size {y_pred.shape[0]}')
        losses = self._compute_all(y_pred[:self.n], targets)
        dim = list(range(1, len(losses.shape)))
        if self.reduction == 'mean':
            losses = losses.mean(dim=dim)
        elif self.reduction == 'sum':
            losses = losses.sum(dim=dim)
        else:
            return losses
        return tch.dot(self.weights.to(device=losses.device), losses)


def _no_void(s_t, n_pixels):
    """F.one_hot(targets, num_classes=C) of the reference raises on labels outside [0,C) (branchy_seg_losses.py:44,92)."""
    # the per-class counts are fp32 (exact per class up to 2^24 pixels); summed in fp64 so the total stays exact for
    # any batch (an fp32 sum of an odd total above 2^24 rounds and raised spuriously). One host read: these losses
    # cannot be captured in a CUDA graph (train_funcs.GraphedTrainStep takes the CE, Lovasz and Jaccard losses).
    if int(s_t.double().sum().round().item()) != int(n_pixels):
        raise RuntimeError("Class values must be smaller than num_classes.")


class DiceLoss(BrSegLoss):
    """1 - (2*sum p*t + smooth) / (sum (p + t) + smooth) per image (branchy_seg_losses.py:40-48)."""

    def _compute_all(self, y_pred, targets):
        s_pt, s_p, s_t = ops.soft_overlap_sums(y_pred, targets)
        _no_void(s_t, targets.numel())
        num = 2 * s_pt.sum(dim=2) + self.smooth                      # [E,N]
        den = (s_p + s_t.unsqueeze(0)).sum(dim=2) + self.smooth
        return 1 - num / den

    def _compute_loss(self, y_pred, targets):
        return self._compute_all(y_pred.unsqueeze(0), targets)[0]


class JaccardLoss(BrSegLoss):
    """1 - (I + smooth)/(U + smooth) per image and class, class 0 scaled by downgrad_bg; void labels (>= C) are
    dropped (branchy_seg_losses.py:50-77)."""

    def __init__(self, smooth=1e-6, reduction='mean', n_branches=0, downgrad_bg=1.):
        super().__init__(smooth, reduction, n_branches)
        self.downgrad_bg = downgrad_bg if 0 <= downgrad_bg <= 1. else 1.

    def _compute_all(self, y_pred, targets):
        s_pt, s_p, s_t = ops.soft_overlap_sums(y_pred, targets)
        intersection = s_pt                                          # [E,N,C]
        total = s_p + s_t.unsqueeze(0)
        union = total - intersection
        IoU = (intersection + self.smooth) / (union + self.smooth)
        if self.downgrad_bg:
            loss = 1 - IoU
            scale = tch.ones(loss.shape[-1], dtype=loss.dtype, device=loss.device)
            scale[0] = self.downgrad_bg
            return loss * scale
        return (1 - IoU).sum(dim=-1)

    def _compute_loss(self, y_pred, targets):
        return self._compute_all(y_pred.unsqueeze(0), targets)[0]


class TverskyLoss(BrSegLoss):
    """1 - (TP + s)/(TP + alpha*FP + beta*FN + s) per image and class on the ARG-MAX map (branchy_seg_losses.py:79-103);
    TP/FP/FN from the confusion-matrix kernel."""

    def __init__(self, smooth=1e-6, alpha=.5, beta=.5, reduction='mean', n_branches=1, weights=None):
        super().__init__(smooth, reduction, n_branches, weights)
        self.alpha = alpha
        self.beta = beta

    def _forward_imp(self, y_pred, targets):
        N, C = y_pred.shape[:2]
        cm = ops.confusion_hist(y_pred, targets, C)                  # [N, C+1, C]
        if int(cm[:, C].sum()) != 0:
            raise RuntimeError("Class values must be smaller than num_classes.")
        TP, FP, FN = (t.to(tch.float32) for t in ops.basics_from_cm(cm))
        tversky_idx = (TP + self.smooth) / (TP + self.alpha * FP + self.beta * FN + self.smooth)
        return 1 - tversky_idx

    def _compute_loss(self, y_pred, targets):
        return self._forward_imp(y_pred, targets)


class FocalTverskyLoss(TverskyLoss):
    def __init__(self, smooth=1e-6, alpha=.5, beta=.5, gamma=1., reduction='mean', n_branches=1, weights=None):
        super().__init__(smooth, alpha, beta, reduction, n_branches, weights)
        self.gamma = gamma

    def _compute_loss(self, y_pred, targets):
        return self._forward_imp(y_pred, targets) ** self.gamma


class FocalLoss(BrSegLoss):
    """-(1 - p_t)^gamma * log p_t per pixel, times alpha[targets] when alpha is given (branchy_seg_losses.py:115-131).
    As in the reference, targets are [N,1,H,W] (`probs.gather(1, targets)`) with labels in [0,C), and the alpha product
    broadcasts the [N,H,W] loss against alpha[targets] of shape [N,1,H,W] to [N,N,H,W] (:128-129): every image's loss is
    weighted by the alpha of EVERY image at that pixel. Reproduced as a per-pixel weight sum_i alpha[t_i] inside the kernel
    for 'mean' / 'sum'; reduction='none' returns the reference's [E,N,(N,)H,W] tensor."""

    def __init__(self, alpha=None, gamma=2, smooth=1e-6, reduction='mean', n_branches=1, weights=None):
        super().__init__(smooth, reduction, n_branches, weights)
        self.alpha = alpha
        self.gamma = gamma

    def _check(self, y_pred, targets):
        N, C = y_pred.shape[1:3]
        if targets.dim() != y_pred.dim() - 1 or targets.shape[0] != N or targets.shape[1] != 1 \
                or tuple(targets.shape[2:]) != tuple(y_pred.shape[3:]):
            raise RuntimeError(f'gather(): targets must be [N,1,H,W] for logits {tuple(y_pred.shape[1:])}, '
                               f'got {tuple(targets.shape)}')
        if targets.is_floating_point():
            targets = targets.to(tch.int64)
        if bool(((targets < 0) | (targets >= C)).any()):
            raise RuntimeError('index out of range in gather(): labels must be in [0, C)')
        return targets

    def _alpha_map(self, targets, device):
        alpha = tch.as_tensor(self.alpha, dtype=tch.float32, device=device)
        return alpha, alpha[targets.to(tch.int64)]                  # [N,1,H,W]

    def _compute_all(self, y_pred, targets):
        targets = self._check(y_pred, targets)
        E, N = y_pred.shape[:2]
        loss = ops.focal_map(y_pred, targets, self.gamma).view(E, N, *y_pred.shape[3:])
        if self.alpha is not None:
            loss = loss.unsqueeze(1) * self._alpha_map(targets, loss.device)[1].unsqueeze(0)    # [E,N,N,H,W]
        return loss

    def _compute_loss(self, y_pred, targets):
        return self._compute_all(y_pred.unsqueeze(0), targets)[0]

    def forward(self, y_pred, targets):
        if self.reduction not in ('mean', 'sum'):
            return super().forward(y_pred, targets)
        if y_pred.shape[0] < self.n:
            raise IndexError(f'index {self.n - 1} is out of bounds for dimension 0 with size {y_pred.shape[0]}')
        y_pred = y_pred[:self.n]
        targets = self._check(y_pred, targets)
        N = y_pred.shape[1]
        count = targets.numel()
        alpha = pixw = None
        if self.alpha is not None:
            alpha, amap = self._alpha_map(targets, y_pred.device)
            if N > 1:
                alpha, pixw = None, amap.sum(dim=0).reshape(-1)      # sum_i alpha[t_i] per pixel position
                count *= N
        w = self.weights.to(device=y_pred.device)
        scale = 1.0 / count if self.reduction == 'mean' else 1.0
        sums = ops.focal_sums(y_pred, targets, self.gamma, alpha=alpha, pixel_weight=pixw, coef=w.detach() * scale)
        return tch.dot(w, sums * scale)


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
