"""Drop-in for the reference's eval_br_ent.py: the `img_norm_entropy` confidence functor (:19-36)
and the entropy-gated batch evaluator `br_evaluator` (:38-84).

The reference moves every exit's [C,H,W] probabilities to the host and runs scipy/skimage there
(75 ms per image-exit); here the softmax, entropy, pooling and mean run in csrc/exit_gate.cu and only
the per-image scalar crosses PCIe. `br_evaluator` accepts any batch size (the reference needs 1) and,
when `net` is an eeseg `branchyDeepv3`, gates on the low-resolution logits with the fused
up-sample+entropy kernel so full-resolution logits are never materialised."""
import torch as tch
import torch.nn.functional as F  # noqa: F401  (kept for API parity with the reference module)

from . import ops
from .compute_mIoU import mIoU


class img_norm_entropy:
    def __init__(self, n_classes, pool_min=False, s=1):
        self.pool = s != 1
        self.pool_min = pool_min
        self.size = (s, s)
        self.C = n_classes

    def scores(self, x, kind='probs', out_hw=None, layout='NCHW', want_amax=False):
        """Batched form: x [N,C,h,w] CUDA (probabilities or logits) -> (scores f32 [N] on device,
        GateResult)."""
        if layout == 'NCHW' and x.shape[1] != self.C:
            # the reference normalises by log(self.C) whatever the tensor holds (eval_br_ent.py:29); a mismatch is a caller bug
            raise ValueError(f'img_norm_entropy(n_classes={self.C}) got a tensor with {x.shape[1]} channels')
        res = ops.exit_gate(x, out_hw, layout=layout, kind=kind, n_classes=self.C,
                            want_ent=self.pool, want_amax=want_amax, want_score=not self.pool)
        if self.pool:
            res.score = ops.entropy_pool_mean(res.ent, self.size[0], self.pool_min)
        return res.score, res

    def __call__(self, probs):
        """probs: [C,H,W] probabilities (the reference passes a CPU tensor, eval_br_ent.py:59; a CPU
        tensor is copied to the current CUDA device — the arithmetic always runs on the GPU).
        Returns a Python float like np.mean does."""
        assert len(probs.shape) == 3
        if not probs.is_cuda:
            probs = probs.cuda()
        sc, _ = self.scores(probs.unsqueeze(0).float(), kind='probs')
        return float(sc.item())


def _exit_predictions(net, X, n_classes, l, n_branches, skip):
    """Returns (scores [n_br, N] f32 device or None, amax list of E uint8 [N,H,W] or None, y_pred)."""
    H, W = X.shape[-2:]
    if hasattr(net, 'forward_lowres') and getattr(net, 'fast_inference', False) and X.is_cuda \
            and not getattr(net, 'training', False):
        lows = net.forward_lowres(X)
        scores, amaxes = [], []
        for i, lo in enumerate(lows):
            if i < n_branches and i >= skip:
                sc, res = l.scores(lo, kind='logits', out_hw=(H, W), layout='NHWC', want_amax=True)
                scores.append(sc)
            else:
                res = ops.exit_gate(lo, (H, W), layout='NHWC', n_classes=n_classes, want_score=False)
                scores.append(None)
            amaxes.append(res.amax)
        return scores, amaxes
    y_pred = net(X)
    scores, amaxes = [], []
    for i in range(n_branches + 1):
        yi = y_pred[i]
        if i < n_branches and i >= skip:
            sc, res = l.scores(yi, kind='logits', want_amax=True)
            scores.append(sc)
        else:
            res = ops.exit_gate(yi, None, n_classes=n_classes, want_score=False)
            scores.append(None)
        amaxes.append(res.amax)
    return scores, amaxes


def br_evaluator(net, n_exits, n_classes, test_loader, device, tau, metric='ent', size=1, ignore=(), skip=0):
    accumulator = [mIoU(n_classes=n_classes, device=device) for _ in range(n_exits + 1)]
    out_count = [0 for _ in range(n_exits + 1)]

    if metric.lower() == 'max':
        l = img_norm_entropy(n_classes, s=size)
    elif metric.lower() == 'min':
        l = img_norm_entropy(n_classes, s=size, pool_min=True)
    else:
        l = img_norm_entropy(n_classes)

    n_branches = n_exits - 1
    with tch.no_grad():
        for X, y in test_loader:
            X, y = X.to(device, non_blocking=True), y.to(device, non_blocking=True)
            N = X.shape[0]
            scores, amaxes = _exit_predictions(net, X, n_classes, l, n_branches, skip)
            # one confusion matrix per (exit, image); one D2H of the scalars + matrices per batch
            cms = tch.stack([ops.confusion_hist(a, y, n_classes) for a in amaxes])          # [E,N,C+1,C]
            sc = tch.stack([s if s is not None else tch.full((N,), float('inf'), device=X.device)
                            for s in scores[:n_branches]]).cpu() if n_branches else None
            for k in range(N):                      # images in loader order, like the batch-1 reference
                left = False
                for i in range(skip, n_branches):
                    if float(sc[i, k]) < tau:
                        accumulator[i].add_confusion(cms[i, k])
                        accumulator[-1].add_confusion(cms[i, k])
                        out_count[i] += 1
                        left = True
                        break
                if not left:
                    accumulator[-2].add_confusion(cms[-1, k])
                    accumulator[-1].add_confusion(cms[-1, k])
                    out_count[-2] += 1
                out_count[-1] += 1

    res = dict()
    for i in range(n_branches):
        res[f'b{i+1}_mIoU'] = accumulator[i].compute().item()
        res[f'b{i+1}_count'] = out_count[i]
    res['mIoU_out'] = accumulator[-2].compute().item()
    res['count_out'] = out_count[-2]
    res['mIoU_gl'] = accumulator[-1].compute().item()
    res['out_gl'] = out_count[-1]
    res['t'] = tau
    res['pool'] = metric
    res['pool_size'] = size
    del accumulator
    return res
