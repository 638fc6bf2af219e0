"""Drop-in for the multiclass path of the reference's lovaszsoftmax.py (Berman's Lovasz-softmax):
`lovasz_softmax` (:154-169), `lovasz_softmax_flat` (:172-200), `flatten_probas` (:203-219),
`lovasz_grad` (:19-31), `mean` (:233-251). The work (error keys, segmented radix sort, Jaccard
gradient, dot, gradient scatter) runs in csrc/lovasz.cu; there is no PyTorch/CPU fallback.
The binary hinge / iou / xloss helpers of the reference are never called on this path (out of scope).
"""
import torch

from . import ops


def lovasz_softmax(probas, labels, classes='present', per_image=False, ignore=None):
    """probas [B,C,H,W] (or [B,H,W] sigmoid output), labels [B,H,W] / [B,1,H,W]. Returns a 0-d
    tensor, differentiable w.r.t. probas. Same semantics as lovaszsoftmax.py:154-169."""
    if probas.dim() == 3:
        probas = probas.unsqueeze(1)
    return ops.lovasz_multi_exit(probas.unsqueeze(0), labels, classes=classes, per_image=per_image,
                                 ignore=ignore)[0]


def lovasz_softmax_flat(probas, labels, classes='present'):
    """probas [P,C], labels [P] (lovaszsoftmax.py:172-200)."""
    if probas.numel() == 0:
        return probas * 0.
    P, C = probas.shape
    y = probas.t().contiguous().view(1, 1, C, P)
    return ops.lovasz_multi_exit(y, labels.view(1, P), classes=classes)[0]


def flatten_probas(probas, labels, ignore=None):
    """Pure layout helper kept for API compatibility (lovaszsoftmax.py:203-219); the kernel path
    never materialises this [P,C] copy."""
    if probas.dim() == 3:
        B, H, W = probas.size()
        probas = probas.view(B, 1, H, W)
    B, C, H, W = probas.size()
    probas = probas.permute(0, 2, 3, 1).contiguous().view(-1, C)
    labels = labels.view(-1)
    if ignore is None:
        return probas, labels
    valid = labels != ignore
    return probas[valid], labels[valid]


def lovasz_grad(gt_sorted):
    """Gradient of the Lovasz extension w.r.t. sorted errors (lovaszsoftmax.py:19-31) for a
    caller-supplied sorted ground-truth vector. Tiny helper on torch ops (the training path uses
    the fused kernel, which derives the same quantity from exact integer counts)."""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1. - intersection / union
    if p > 1:
        jaccard[1:p] = jaccard[1:p] - jaccard[0:-1]
    return jaccard


def mean(l, ignore_nan=False, empty=0):
    """nanmean compatible with generators (lovaszsoftmax.py:233-251)."""
    it = iter(l)
    if ignore_nan:
        it = (x for x in it if not (x != x))
    try:
        n = 1
        acc = next(it)
    except StopIteration:
        if empty == 'raise':
            raise ValueError('Empty mean')
        return empty
    for n, v in enumerate(it, 2):
        acc += v
    if n == 1:
        return acc
    return acc / n
