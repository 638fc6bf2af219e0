"""Drop-in for the reference's my_pixelwise_xentropy.py: `_cross_entropy` (:6-17) and
`BrXEntropyLoss` (:19-46). One fused kernel pass (csrc/multi_exit_ce.cu) computes every exit's
mean NLL and, when a gradient is needed, d loss / d logits in the same sweep."""
import torch as tch

from . import ops


class _cross_entropy(tch.nn.Module):
    def __init__(self, reduction='mean', ignore_index=-100):
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("the eeseg CE kernel implements reduction='mean' "
                                      "(the only one the reference's training scripts use)")
        self.reduction = reduction
        self.ignore_index = ignore_index

    def _squeeze(self, targets):
        # my_pixelwise_xentropy.py:12-13. The reference's bare .squeeze() also drops a batch dim of
        # size 1 (and then fails inside CrossEntropyLoss); here only the channel dim is dropped.
        if len(targets.shape) > 3:
            targets = targets.squeeze(1) if targets.shape[1] == 1 else targets.squeeze()
        return targets

    def _compute_loss(self, y_pred, targets, coef=None):
        per_exit, _ = ops.multi_exit_ce(y_pred.unsqueeze(0), self._squeeze(targets),
                                        self.ignore_index, coef)
        return per_exit[0]

    def forward(self, y_pred, targets):
        return self._compute_loss(y_pred, targets)


class BrXEntropyLoss(_cross_entropy):
    def __init__(self, reduction='mean', ignore_index=-100, b_reduction='mean', n_exits=0, weights=None):
        super().__init__(reduction, ignore_index)
        self.b_reduction = b_reduction
        self.n_exits = n_exits
        if weights and len(weights) == n_exits:
            self.weights = tch.tensor(weights, requires_grad=True)
        else:
            self.weights = weights

    def forward(self, y_pred, targets):
        if not self.n_exits:
            return self._compute_loss(y_pred, targets)
        assert self.n_exits <= y_pred.shape[0]
        E = self.n_exits
        w = None
        if self.weights is not None:
            w = self.weights.to(y_pred.device)
        # expected d total / d per-exit loss, so the forward kernel can emit the final gradient
        coef = tch.ones(E, device=y_pred.device) if w is None else w.detach().float().clone()
        if self.b_reduction == 'mean':
            coef = coef / E
        losses, _ = ops.multi_exit_ce(y_pred[:E], self._squeeze(targets), self.ignore_index, coef)
        if w is not None:
            losses = losses * w
        if self.b_reduction == 'sum':
            return losses.sum()
        if self.b_reduction == 'mean':
            return losses.mean()
        return losses
