"""Drop-in for the reference's eval_mIoU.py: `mIoU_evaluator` (:15-40) — per-exit dataset mIoU."""
import torch as tch

from . import ops
from .compute_mIoU import mIoU


def mIoU_evaluator(net, n_exits, n_classes, test_loader, device):
    accumulator = [mIoU(n_classes=n_classes, device=device) for _ in range(n_exits)]
    n_branches = n_exits - 1
    with tch.no_grad():
        for X, y in test_loader:
            X, y = X.to(device, non_blocking=True), y.to(device, non_blocking=True)
            fast = hasattr(net, 'forward_lowres') and getattr(net, 'fast_inference', False) \
                and X.is_cuda and not getattr(net, 'training', False) and n_branches
            if fast:
                # argmax maps straight from the low-res logits: no [E,N,C,H,W] tensor at all
                H, W = X.shape[-2:]
                for i, lo in enumerate(net.forward_lowres(X)):
                    am = ops.exit_gate(lo, (H, W), layout='NHWC', n_classes=n_classes, want_score=False).amax
                    accumulator[i].add_confusion(ops.confusion_hist(am, y, n_classes).sum(dim=0))
                continue
            y_pred = net(X)
            for i in range(n_branches):
                accumulator[i](y_pred[i], y)
            accumulator[-1](y_pred[-1] if n_branches else y_pred, y)

    res = dict()
    for i in range(n_branches):
        res[f'b{i+1}_mIoU'] = accumulator[i].compute().item()
    res['mIoU'] = accumulator[-1].compute().item()
    del accumulator
    return res
