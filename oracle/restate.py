"""TEST INFRASTRUCTURE — CPU restatement (numpy, fp32) of the reference's early-exit hot path.

This file is the parity ORACLE. It is imported only by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, never by the product package
(``ee_semantic_segmentation_b200``), whose kernels must fail loudly without their CUDA library.

Pinning: every function here is checked (``tests/test_oracle_pinned.py``, ``-m "not gpu"``)
  (1) against the three known-answer fixtures the reference itself holds
      (compute_mIoU.py:65-149 -> 0.9513888955116272; seg_metrics.py:78-173; new_seg_losses.py:170-256),
  (2) against outputs of the UNMODIFIED reference modules imported from /root/reference in the
      build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``), and
  (3) live against the reference when /root/reference is present.
All file:line citations are relative to /root/reference.
"""
import math

import numpy as np


# ----------------------------------------------------------------------------------------------
# A4  F.interpolate(mode='bilinear', align_corners=False)   (from_deepv3_new.py:149,152)
# ----------------------------------------------------------------------------------------------
def _src_index(out_size, in_size):
    """ATen area_pixel_compute_source_index, align_corners=False, no user scale:
    src = max(scale*(dst+0.5)-0.5, 0), i0=floor(src), i1=min(i0+1,in-1), l1=src-i0, l0=1-l1."""
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.maximum(scale * (dst + np.float32(0.5)) - np.float32(0.5), np.float32(0.0))
    i0 = np.floor(src).astype(np.int64)
    i0 = np.minimum(i0, in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def bilinear_upsample(x, out_hw):
    """x: [..., h, w] float32 -> [..., H, W] float32 (ATen upsample_bilinear2d formula)."""
    x = np.asarray(x, dtype=np.float32)
    H, W = out_hw
    h, w = x.shape[-2:]
    y0, y1, ly0, ly1 = _src_index(H, h)
    x0, x1, lx0, lx1 = _src_index(W, w)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    t = top[..., :, x0] * lx0 + top[..., :, x1] * lx1
    b = bot[..., :, x0] * lx0 + bot[..., :, x1] * lx1
    return (ly0[:, None] * t + ly1[:, None] * b).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# softmax / argmax helpers (F.softmax(y,1) then argmax, seg_metrics.py:16, ee_dnn_op_ne.py:80-82)
# ----------------------------------------------------------------------------------------------
def softmax_c(logits, axis):
    x = np.asarray(logits, dtype=np.float32)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# A5  img_norm_entropy.__call__   (eval_br_ent.py:19-36)
# ----------------------------------------------------------------------------------------------
def pixel_norm_entropy(probs, n_classes):
    """scipy.stats.entropy(p, base=C, axis=0): p is re-normalised along axis 0, entr(0)=0,
    computed in the input dtype (float32)."""
    p = np.asarray(probs, dtype=np.float32)
    p = p / p.sum(axis=0, keepdims=True, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, -p * np.log(p, dtype=np.float32), np.float32(0.0)).astype(np.float32)
    ent = t.sum(axis=0, dtype=np.float32)
    return (ent / np.float32(math.log(n_classes))).astype(np.float32)


def block_reduce_2d(img, s, func):
    """skimage.measure.block_reduce(img,(s,s),func): pad the END of each axis with 0 to a
    multiple of s (eval_br_ent.py:34-35) — so np.min sees zeros in ragged edge blocks."""
    H, W = img.shape
    ph, pw = (-H) % s, (-W) % s
    img = np.pad(img, ((0, ph), (0, pw)), mode="constant", constant_values=0)
    blk = img.reshape(img.shape[0] // s, s, img.shape[1] // s, s)
    return func(blk, axis=(1, 3))


def img_norm_entropy(probs, n_classes, pool_min=False, s=1):
    """Returns np.float32 — mean normalised entropy of one image, optionally block max/min pooled."""
    p = np.asarray(probs)
    assert p.ndim == 3  # eval_br_ent.py:28
    ent = pixel_norm_entropy(p, n_classes)
    if s != 1:
        return np.mean(block_reduce_2d(ent, s, np.min if pool_min else np.max))
    return np.mean(ent)


# ----------------------------------------------------------------------------------------------
# A11  SegMetric._compute_basics   (seg_metrics.py:13-28)  as a (C+1) x C confusion matrix
# ----------------------------------------------------------------------------------------------
def argmax_first(logits, axis):
    """argmax(softmax(y)) == argmax(y) up to fp ties; torch/numpy both return the first max."""
    return np.argmax(np.asarray(logits), axis=axis)


def confusion_matrix(pred, targets, n_classes):
    """pred: [N, P] ints in [0,C); targets: [N, P] any ints (>=C or <0 -> void row C).
    Returns int64 [N, C+1, C] with CM[n, t', p], t' = t if 0<=t<C else C."""
    pred = np.asarray(pred).astype(np.int64)
    N = pred.shape[0]
    pred = pred.reshape(N, -1)
    t = np.asarray(targets).astype(np.int64).reshape(N, -1)
    C = n_classes
    t = np.where((t >= 0) & (t < C), t, C)
    cm = np.zeros((N, C + 1, C), dtype=np.int64)
    for n in range(N):
        idx = t[n] * C + pred[n]
        cm[n] = np.bincount(idx, minlength=(C + 1) * C).reshape(C + 1, C)
    return cm


def basics_from_cm(cm):
    """TP_c=CM[c][c]; FP_c=sum_{t'!=c} CM[t'][c] (void rows count as FP, seg_metrics.py:26);
    FN_c=sum_{p!=c} CM[c][p]. cm: [..., C+1, C] -> three [..., C] int64."""
    C = cm.shape[-1]
    tp = np.diagonal(cm[..., :C, :], axis1=-2, axis2=-1)
    fp = cm.sum(axis=-2) - tp
    fn = cm[..., :C, :].sum(axis=-1) - tp
    return tp.astype(np.int64), fp.astype(np.int64), fn.astype(np.int64)


def compute_basics(logits, targets):
    """logits [N,C,H,W] (or [N,C,P]); targets viewable to [N,-1]. -> TP,FP,FN int64 [N,C]."""
    logits = np.asarray(logits)
    N, C = logits.shape[:2]
    pred = argmax_first(logits.reshape(N, C, -1), axis=1)
    return basics_from_cm(confusion_matrix(pred, np.asarray(targets).reshape(N, -1), C))


# ----------------------------------------------------------------------------------------------
# A12  mIoU   (compute_mIoU.py:7-36)
# ----------------------------------------------------------------------------------------------
class MIoU:
    """float32 [3,C] accumulator, per-call adds of int64 sums cast to fp32 (compute_mIoU.py:25-27);
    compute(): TP/(TP+FP+FN) mean over C; the NaN patch at :35 never matches, so an absent class
    gives NaN."""

    def __init__(self, n_classes):
        self.C = n_classes
        self.acc = np.zeros((3, n_classes), dtype=np.float32)

    def __call__(self, logits, targets):
        tp, fp, fn = compute_basics(logits, targets)
        self.add_counts(tp.sum(0), fp.sum(0), fn.sum(0))

    def add_counts(self, tp, fp, fn):
        self.acc[0] += tp.astype(np.float32)
        self.acc[1] += fp.astype(np.float32)
        self.acc[2] += fn.astype(np.float32)

    def compute(self):
        den = self.acc.sum(axis=0, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ciou = self.acc[0] / den
        return np.float32(ciou.sum(dtype=np.float32) / np.float32(self.C))


def img_miou(pred, target):
    """img_mIoU.forward for one image (compute_mIoU.py:43-58): classes = unique(target) (void
    included), IoU_i = |gt_i & pred_i| / |gt_i | pred_i|, mean over those classes."""
    pred = np.asarray(pred)
    if pred.ndim == 4:
        pred = pred.argmax(axis=1).squeeze()
    target = np.asarray(target).squeeze()
    classes = np.unique(target.reshape(-1))
    s = np.float32(0.0)
    for c in classes:
        gt = target == c
        pr = pred == c
        inter = np.float32((gt & pr).sum())
        union = np.float32((gt | pr).sum())
        s = s + inter / union
    return float(s / np.float32(classes.shape[0]))


# ----------------------------------------------------------------------------------------------
# A8  BrXEntropyLoss   (my_pixelwise_xentropy.py:19-46)
# ----------------------------------------------------------------------------------------------
def pixel_ce(logits, targets, ignore_index):
    """CrossEntropyLoss(reduction='mean', ignore_index): logits [N,C,...], targets [N,...].
    Returns (loss, dlogits) in float64-accumulated float32; all-void gives NaN (0/0) like torch."""
    x = np.asarray(logits, dtype=np.float32)
    N, C = x.shape[:2]
    xs = x.reshape(N, C, -1)
    t = np.asarray(targets).astype(np.int64).reshape(N, -1)
    valid = t != ignore_index
    m = xs.max(axis=1, keepdims=True)
    z = xs - m
    lse = np.log(np.exp(z).sum(axis=1, dtype=np.float32))
    tt = np.where(valid, t, 0)
    picked = np.take_along_axis(z, tt[:, None, :], axis=1)[:, 0, :]
    nll = (lse - picked) * valid
    cnt = valid.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        loss = np.float32(nll.sum(dtype=np.float64) / np.float64(cnt))
        sm = np.exp(z - lse[:, None, :]).astype(np.float32)
        onehot = np.zeros_like(sm)
        np.put_along_axis(onehot, tt[:, None, :], 1.0, axis=1)
        d = (sm - onehot) * valid[:, None, :] / np.float32(cnt)
    return loss, d.reshape(x.shape).astype(np.float32)


def br_xentropy(y_pred, targets, ignore_index=-100, b_reduction="mean", n_exits=0, weights=None):
    """y_pred [E,N,C,H,W]; targets [N,1,H,W] or [N,H,W]. Returns (loss, dy_pred, per_exit)."""
    y = np.asarray(y_pred, dtype=np.float32)
    t = np.asarray(targets)
    if t.ndim > 3:
        t = t.squeeze()  # my_pixelwise_xentropy.py:12-13 (collapses N==1 too — reference quirk)
    if not n_exits:
        loss, d = pixel_ce(y, t, ignore_index)
        return loss, d, np.array([loss], dtype=np.float32)
    assert n_exits <= y.shape[0]
    per = np.zeros(n_exits, dtype=np.float32)
    dy = np.zeros_like(y)
    for i in range(n_exits):
        per[i], dy[i] = pixel_ce(y[i], t, ignore_index)
    w = np.ones(n_exits, dtype=np.float32)
    if weights is not None and len(weights) == n_exits:
        w = np.asarray(weights, dtype=np.float32)
    wl = per * w
    if b_reduction == "sum":
        loss, scale = wl.sum(dtype=np.float32), w
    elif b_reduction == "mean":
        loss, scale = wl.mean(dtype=np.float32), w / np.float32(n_exits)
    else:
        return wl, dy[:n_exits] * w[:, None, None, None, None], per
    dy[:n_exits] *= scale[:, None, None, None, None]
    return np.float32(loss), dy, per


# ----------------------------------------------------------------------------------------------
# Overlap family of branchy_seg_losses.py: Dice :40-48, Jaccard :50-77, Tversky :79-103,
# FocalTversky :105-113, Focal :115-131, under BrSegLoss.forward :24-38. Values and (Dice / Jaccard / Focal) gradients
# w.r.t. the logits.
# ----------------------------------------------------------------------------------------------
def _overlap_sums(logits, targets):
    """logits [N,C,H,W] -> p [N,C,HW] (softmax), one-hot t [N,C,HW] (labels outside [0,C) match no class)."""
    y = np.asarray(logits, dtype=np.float32)
    N, C = y.shape[:2]
    p = softmax_c(y.reshape(N, C, -1).astype(np.float64), 1)
    t = np.asarray(targets).reshape(N, -1)
    oh = (t[:, None, :] == np.arange(C)[None, :, None]).astype(np.float64)
    return p, oh, t


def _softmax_backward(p, g):
    """dL/dz for z -> softmax -> p with dL/dp = g (both [N,C,HW])."""
    return p * (g - (p * g).sum(axis=1, keepdims=True))


def dice_loss(logits, targets, smooth=1e-6):
    """DiceLoss._compute_loss (branchy_seg_losses.py:40-48): [N] losses and d sum(loss*w)/dlogits for w = 1."""
    p, oh, t = _overlap_sums(logits, targets)
    C = p.shape[1]
    if ((t < 0) | (t >= C)).any():
        raise RuntimeError("Class values must be smaller than num_classes.")   # F.one_hot, :44
    num = 2 * (p * oh).sum(axis=(1, 2)) + smooth
    den = (p + oh).sum(axis=(1, 2)) + smooth
    loss = 1 - num / den
    g = (-(2 * oh) / den[:, None, None] + (num / den ** 2)[:, None, None])       # d loss_n / d p
    return loss.astype(np.float32), _softmax_backward(p, g).reshape(np.asarray(logits).shape)


def jaccard_loss(logits, targets, smooth=1e-6, downgrad_bg=1.0):
    """JaccardLoss._compute_loss (:55-77): [N,C] (or [N] when downgrad_bg is falsy) and the gradient of its sum."""
    p, oh, _ = _overlap_sums(logits, targets)
    inter = (p * oh).sum(axis=-1)
    total = (p + oh).sum(axis=-1)
    union = total - inter
    iou = (inter + smooth) / (union + smooth)
    scale = np.ones(p.shape[1])
    if downgrad_bg:
        scale[0] = downgrad_bg
        loss = (1 - iou) * scale
    else:
        loss = (1 - iou).sum(axis=-1)
    # d(1-iou)/dp = -[oh*(union+s) - (inter+s)*(1-oh)] / (union+s)^2
    u = (union + smooth)[:, :, None]
    g = -(oh * u - (inter + smooth)[:, :, None] * (1 - oh)) / u ** 2 * scale[None, :, None]
    return loss.astype(np.float32), _softmax_backward(p, g).reshape(np.asarray(logits).shape)


def tversky_loss(logits, targets, smooth=1e-6, alpha=.5, beta=.5, gamma=None):
    """TverskyLoss._forward_imp (:85-100) on the arg-max map; gamma -> FocalTverskyLoss (:110-113). [N,C]."""
    y = np.asarray(logits, dtype=np.float32)
    N, C = y.shape[:2]
    t = np.asarray(targets).reshape(N, -1)
    if ((t < 0) | (t >= C)).any():
        raise RuntimeError("Class values must be smaller than num_classes.")   # F.one_hot, :92
    pred = argmax_first(y.reshape(N, C, -1), 1)
    cls = np.arange(C)[None, :, None]
    P, T = (pred[:, None, :] == cls), (t[:, None, :] == cls)
    TP = (P & T).sum(-1).astype(np.float32)
    FP = (P & ~T).sum(-1).astype(np.float32)
    FN = (~P & T).sum(-1).astype(np.float32)
    loss = 1 - (TP + np.float32(smooth)) / (TP + np.float32(alpha) * FP + np.float32(beta) * FN + np.float32(smooth))
    return loss if gamma is None else loss ** np.float32(gamma)


def seg_overlap_loss(logits, targets, kind, smooth=1e-6, index=False, downgrad_bg=1.0, alpha=.5, beta=.5, gamma=None):
    """Single-output family of new_seg_losses.py: kind 'dice' (:34-56), 'jaccard' (:58-88), 'tversky' (:90-111; gamma ->
    FocalTverskyLoss :113-121, exponent 1/gamma). Labels >= C are void and dropped by dice / jaccard (:47-48); tversky's
    one_hot(num_classes=C) raises on them (:101). Returns (_compute_loss tensor, d sum(loss) / d logits)."""
    p, oh, t = _overlap_sums(logits, targets)
    C = p.shape[1]
    if (t < 0).any() or (kind == "tversky" and (t >= C).any()):
        raise RuntimeError("Class values must be smaller than num_classes.")
    if kind == "dice":
        num = 2 * (p * oh).sum(axis=(1, 2)) + smooth
        den = (p + oh).sum(axis=(1, 2)) + smooth
        val = num / den
        g = (2 * oh) / den[:, None, None] - (num / den ** 2)[:, None, None]          # d val / d p
        loss, g = (val, g) if index else (1 - val, -g)
    elif kind == "jaccard":
        inter = (p * oh).sum(axis=-1)
        union = (p + oh).sum(axis=-1) - inter
        iou = (inter + smooth) / (union + smooth)
        u = (union + smooth)[:, :, None]
        diou = (oh * u - (inter + smooth)[:, :, None] * (1 - oh)) / u ** 2
        if index:
            loss, g = iou, diou
        elif downgrad_bg:
            scale = np.ones(C)
            scale[0] = downgrad_bg
            loss, g = (1 - iou) * scale, -diou * scale[None, :, None]
        else:
            loss, g = (1 - iou).sum(axis=-1), -diou
    else:
        TP = (p * oh).sum(axis=-1)
        FP = (p * (1 - oh)).sum(axis=-1)
        FN = ((1 - p) * oh).sum(axis=-1)
        D = TP + alpha * FP + beta * FN + smooth
        T = (TP + smooth) / D
        dT = (oh * D[:, :, None] - (TP + smooth)[:, :, None] * (oh + alpha * (1 - oh) - beta * oh)) / D[:, :, None] ** 2
        loss, g = 1 - T, -dT
        if gamma is not None:
            g = g * ((1 / gamma) * loss ** (1 / gamma - 1))[:, :, None]
            loss = loss ** (1 / gamma)
    return loss.astype(np.float32), _softmax_backward(p, g).reshape(np.asarray(logits).shape)


def focal_loss(logits, targets, gamma=2.0, alpha=None):
    """FocalLoss._compute_loss (branchy_seg_losses.py:122-131). logits [N,C,H,W], targets [N,1,H,W] in [0,C).
    Returns the reference's loss tensor — [N,H,W], or [N,N,H,W] when alpha is given: `loss * alpha[targets]` broadcasts
    [N,H,W] against [N,1,H,W] (:128-129), out[i,j] = loss[j] * alpha[t_i] — and the gradient of its SUM w.r.t. logits."""
    y = np.asarray(logits, dtype=np.float64)
    N, C = y.shape[:2]
    t = np.asarray(targets)
    if t.ndim != y.ndim or t.shape[1] != 1 or t.shape[0] != N or t.shape[2:] != y.shape[2:]:
        raise RuntimeError("gather(): index shape")                               # probs.gather(1, targets), :126
    t = t.astype(np.int64)
    if ((t < 0) | (t >= C)).any():
        raise RuntimeError("index out of range in gather()")
    lsm = y - y.max(axis=1, keepdims=True)
    lsm = lsm - np.log(np.exp(lsm).sum(axis=1, keepdims=True))                    # log_softmax, :124
    p = np.exp(lsm)
    lp = np.take_along_axis(lsm, t, axis=1)[:, 0]                                 # [N,H,W]
    pt = np.exp(lp)
    loss = -((1 - pt) ** gamma) * lp
    # d loss / d lp = gamma (1-pt)^(gamma-1) pt lp - (1-pt)^gamma ;  d lp / d z_c = [c == t] - p_c
    with np.errstate(invalid="ignore", divide="ignore"):
        dl = np.where(pt < 1, gamma * (1 - pt) ** (gamma - 1) * pt * lp, 0.0) - (1 - pt) ** gamma if gamma else -np.ones_like(lp)
    w = np.ones_like(lp)
    if alpha is not None:
        a_t = np.asarray(alpha, dtype=np.float64)[t]                              # [N,1,H,W]
        loss = loss[None] * a_t                                                   # [N(i),N(j),H,W]
        w = np.broadcast_to(a_t.sum(axis=0), lp.shape)                            # every image j: sum_i alpha[t_i]
    oh = (np.arange(C)[None, :, None, None] == t)
    grad = (dl * w)[:, None] * (oh - p)
    return loss.astype(np.float32), grad


def br_seg_loss(per_exit, reduction="mean", weights=None):
    """BrSegLoss.forward (:24-38) on the stacked per-exit losses [E, ...]."""
    l = np.asarray(per_exit, dtype=np.float32)
    if reduction == "none":
        return l
    dims = tuple(range(1, l.ndim))
    r = l.mean(axis=dims, dtype=np.float32) if reduction == "mean" else l.sum(axis=dims, dtype=np.float32)
    w = np.ones(l.shape[0], np.float32) if weights is None else np.asarray(weights, np.float32)
    return np.float32(np.dot(w, r))


# ----------------------------------------------------------------------------------------------
# A10  lovasz_grad / lovasz_softmax_flat / flatten_probas   (lovaszsoftmax.py:19-31,172-219)
# ----------------------------------------------------------------------------------------------
def lovasz_grad(gt_sorted):
    gt_sorted = np.asarray(gt_sorted, dtype=np.float32)
    p = gt_sorted.shape[0]
    gts = gt_sorted.sum(dtype=np.float32)
    inter = gts - np.cumsum(gt_sorted, dtype=np.float32)
    union = gts + np.cumsum(np.float32(1.0) - gt_sorted, dtype=np.float32)
    jac = (np.float32(1.0) - inter / union).astype(np.float32)
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


def flatten_probas(probas, labels, ignore=None):
    x = np.asarray(probas, dtype=np.float32)
    if x.ndim == 3:
        x = x[:, None]
    B, C = x.shape[:2]
    x = np.moveaxis(x, 1, -1).reshape(-1, C)
    lab = np.asarray(labels).reshape(-1)
    if ignore is None:
        return x, lab
    valid = lab != ignore
    return x[valid], lab[valid]


def lovasz_softmax_flat(probas, labels, classes="present"):
    """Returns (loss, dprobas[P,C]). Stable descending sort (the loss value does not depend on
    the tie order; the subgradient does — tests compare gradients on tie-free inputs)."""
    P, C = probas.shape
    if probas.size == 0:
        return np.float32(0.0), np.zeros_like(probas)
    grads = np.zeros_like(probas, dtype=np.float32)
    losses = []
    cls = list(range(C)) if classes in ("all", "present") else list(classes)
    used = []
    for c in cls:
        fg = (labels == c).astype(np.float32)
        if classes == "present" and fg.sum() == 0:
            continue
        pc = probas[:, c]
        err = np.abs(fg - pc)
        perm = np.argsort(-err, kind="stable")
        g = lovasz_grad(fg[perm])
        losses.append(np.dot(err[perm].astype(np.float64), g.astype(np.float64)))
        gc = np.zeros(P, dtype=np.float32)
        gc[perm] = g
        grads[:, c] = gc * np.sign(pc - fg)
        used.append(c)
    if not losses:
        return np.float32(0.0), grads  # lovaszsoftmax.py:233-244 mean([]) -> empty=0
    n = len(losses)
    return np.float32(sum(losses) / n), grads / np.float32(n)


def lovasz_softmax(probas, labels, classes="present", per_image=False, ignore=None):
    """probas [B,C,H,W], labels [B,(1,)H,W]. Returns (loss, dprobas[B,C,H,W])."""
    x = np.asarray(probas, dtype=np.float32)
    B, C, H, W = x.shape
    lab = np.asarray(labels).reshape(B, H, W)
    dx = np.zeros_like(x)

    def one(xb, lb):
        flat = np.moveaxis(xb, 1, -1).reshape(-1, C)
        l = lb.reshape(-1)
        valid = np.ones_like(l, dtype=bool) if ignore is None else (l != ignore)
        loss, g = lovasz_softmax_flat(flat[valid], l[valid], classes)
        gfull = np.zeros_like(flat)
        gfull[valid] = g
        return loss, np.moveaxis(gfull.reshape(xb.shape[0], H, W, C), -1, 1)

    if per_image:
        tot = np.float32(0.0)
        for b in range(B):
            l, g = one(x[b:b + 1], lab[b:b + 1])
            tot += l
            dx[b:b + 1] = g / np.float32(B)
        return np.float32(tot / np.float32(B)), dx
    loss, dx = one(x, lab)
    return loss, dx


def br_lovasz(y_pred, targets, classes="present", per_image=False, ignore=None, n_branches=0,
              prev_out=False):
    """BSL.LovaszSoftmax.forward (branchy_seg_losses.py:151-159): sum_i w_i * lovasz(y_pred[i]),
    w = linspace(0,1,E+1)[1:] when prev_out else 1. Raw logits are passed as 'probas'."""
    E = n_branches + 1
    y = np.asarray(y_pred, dtype=np.float32)
    w = np.linspace(0, 1, E + 1, dtype=np.float32)[1:] if prev_out else np.ones(E, np.float32)
    tot = np.float32(0.0)
    dy = np.zeros_like(y)
    per = np.zeros(E, np.float32)
    for i in range(E):
        per[i], g = lovasz_softmax(y[i], targets, classes, per_image, ignore)
        tot += w[i] * per[i]
        dy[i] = w[i] * g
    return np.float32(tot), dy, per


# ----------------------------------------------------------------------------------------------
# A6  br_evaluator decision rule   (eval_br_ent.py:51-70)
# ----------------------------------------------------------------------------------------------
def first_confident_exit(entropies, tau, skip=0):
    """entropies: per early exit (len E-1) image scalars. Returns the 0-based exit index taken:
    first i>=skip with t<tau, else E-1 (the final exit)."""
    n_br = len(entropies)
    for i in range(skip, n_br):
        if entropies[i] < tau:
            return i
    return n_br
