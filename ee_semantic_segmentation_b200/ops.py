"""Tensor-level wrappers: argument checks, output allocation and autograd glue around the registered torch operators
`torch.ops.eeseg.*` (torch_ops.py), each of which forwards to the `extern "C"` launcher of the same name in
libeeseg_b200.so (include/eeseg.h). torch is used for device memory and streams only; every op here raises on CPU
tensors — there is no fallback path (the operators have a CUDA implementation only)."""
import math

import torch

from . import _lib, torch_ops  # noqa: F401  (torch_ops registers torch.ops.eeseg.*)
from ._lib import BF16, F32, check, lib

_ops = torch_ops.fast


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"eeseg kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the eeseg kernels have no CPU fallback")
    return t


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------------------
# confusion histogram
# --------------------------------------------------------------------------------------------------
def confusion_hist(pred, targets, n_classes, out=None, accumulate=False):
    """pred: logits [N,C,...] (f32/bf16) or class map [N,...] (uint8/int64). targets: anything
    viewable to [N,-1] (cast to int64 like seg_metrics.py:18). Returns int64 [N, C+1, C]."""
    _cuda(pred, "pred")
    C = int(n_classes)
    N = pred.shape[0]
    if N == 0:      # empty batch: nothing to launch
        return out if out is not None else torch.zeros((0, C + 1, C), dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        if pred.dtype in (torch.uint8, torch.int64):
            kind = 1 if pred.dtype == torch.uint8 else 2
            pr = pred.reshape(N, -1).contiguous()
            HW = pr.shape[1]
            dt = 0
        else:
            if pred.shape[1] != C:
                raise ValueError(f"logits have {pred.shape[1]} channels, expected {C}")
            kind = 0
            pr = pred.reshape(N, C, -1).contiguous()
            HW = pr.shape[2]
            dt = _dt(pr)
        tg = _cuda(targets, "targets").reshape(N, -1)
        if tg.dtype != torch.int64:
            tg = tg.to(torch.int64)
        tg = tg.contiguous()
        if tg.shape[1] != HW:
            raise ValueError(f"targets have {tg.shape[1]} pixels per image, predictions {HW}")
        if out is None:
            out = torch.empty((N, C + 1, C), dtype=torch.int64, device=pred.device)
            accumulate = False
        _ops.confusion_hist(pr, kind, tg, C, out, bool(accumulate))
    return out


def basics_from_cm(cm):
    """TP/FP/FN [..., C] int64 from cm [..., C+1, C] (void row counts as FP, seg_metrics.py:25-27)."""
    C = cm.shape[-1]
    tp = torch.diagonal(cm[..., :C, :], dim1=-2, dim2=-1)
    fp = cm.sum(dim=-2) - tp
    fn = cm[..., :C, :].sum(dim=-1) - tp
    return tp, fp, fn


# --------------------------------------------------------------------------------------------------
# exit gate
# --------------------------------------------------------------------------------------------------
class GateResult:
    __slots__ = ("ent", "amax", "mask", "score", "exited_px", "up_logits", "part_sum", "part_cnt")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


def exit_gate(x, out_hw=None, *, layout="NCHW", kind="logits", tau=0.0, n_classes=None,
              want_ent=False, want_amax=True, want_mask=False, want_score=True,
              up_out=None, up_dtype=None, amax_out=None, score_out=None):
    """Fused (bilinear up-sample ->) softmax -> normalised entropy -> argmax -> threshold.

    x: [N,C,h,w] (layout 'NCHW') or [N,h,w,Cp] with Cp >= C (layout 'NHWC', pass n_classes).
    out_hw: (H,W) to interpolate to (align_corners=False); None = same size.
    up_out: optional preallocated [N,C,H,W] tensor (contiguous C,H,W planes) receiving the
            up-sampled logits; or up_dtype to allocate one.
    Returns GateResult; score = per-image mean normalised entropy (img_norm_entropy, s == 1)."""
    _cuda(x, "x")
    if layout == "NCHW":
        N, C, h, w = x.shape
        sn, sc, sy, sx = x.stride()
    elif layout == "NHWC":
        N, h, w, Cp = x.shape
        C = int(n_classes) if n_classes is not None else Cp
        sn, sy, sx, sc = x.stride()
    else:
        raise ValueError(layout)
    H, W = (h, w) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    res = GateResult()
    dev = x.device
    with torch.cuda.device(dev):
        npart = lib().eeseg_exit_gate_num_partials(H, W)
        if up_out is None and up_dtype is not None:
            up_out = torch.empty((N, C, H, W), dtype=up_dtype, device=dev)
        up_sn = 0
        if up_out is not None:
            if tuple(up_out.shape) != (N, C, H, W) or up_out.stride()[1:] != (H * W, W, 1):
                raise ValueError("up_out must be [N,C,H,W] with contiguous (C,H,W) planes")
            up_sn = up_out.stride(0)
            res.up_logits = up_out
        if want_ent:
            res.ent = torch.empty((N, H, W), dtype=torch.float32, device=dev)
        if want_amax:
            res.amax = amax_out if amax_out is not None else torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        if want_mask:
            res.mask = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        if want_score:
            res.part_sum = torch.empty((N, npart), dtype=torch.float64, device=dev)
            res.part_cnt = torch.empty((N, npart), dtype=torch.int32, device=dev)
        _dt(x)
        _ops.exit_gate_pixels(x, 0 if kind == "logits" else 1, sn, sc, sy, sx, N, C, h, w, H, W, float(tau), up_out,
                              up_sn, res.ent, res.amax, res.mask, res.part_sum, res.part_cnt)
        if want_score:
            res.score = score_out if score_out is not None else torch.empty((N,), dtype=torch.float32, device=dev)
            res.exited_px = torch.empty((N,), dtype=torch.int64, device=dev)
            _ops.exit_gate_decide(res.part_sum, res.part_cnt, npart, None, N, H * W, float(tau), True, 0, None,
                                  res.score, res.exited_px, None, None)
    return res


def entropy_pool_mean(ent, s, pool_min=False):
    """ent f32 [N,H,W] -> f32 [N]: mean of the s x s block max (or min) with zero end-padding."""
    _cuda(ent, "ent")
    ent = ent.contiguous()
    N, H, W = ent.shape
    out = torch.empty((N,), dtype=torch.float32, device=ent.device)
    with torch.cuda.device(ent.device):
        check(lib().eeseg_entropy_pool_mean(ent.data_ptr(), N, H, W, int(s), 1 if pool_min else 0,
                                            out.data_ptr(), _stream(ent)), "eeseg_entropy_pool_mean")
    return out


def gate_decide(score, tau, exit_id, exit_idx, less_than=True, want_active=True):
    """Per-image decision + compaction on the device. score f32 [N]; exit_idx int32 [N] in/out
    (-1 = active). Returns (active_list int32 [N], active_count int32 [1]) or (None, None)."""
    _cuda(score, "score")
    N = score.shape[0]
    dev = score.device
    al = ac = None
    with torch.cuda.device(dev):
        if want_active:
            al = torch.empty((N,), dtype=torch.int32, device=dev)
            ac = torch.empty((1,), dtype=torch.int32, device=dev)
        _ops.exit_gate_decide(None, None, 0, score, N, 1, float(tau), bool(less_than), int(exit_id), exit_idx,
                              None, None, al, ac)
    return al, ac


def compact_rows(src, active_list, active_count, out):
    """out[j] = src[active_list[j]] for j < active_count (device-side count; rows past it are left untouched).
    src [n, ...] and out [m, ...] dense rows of equal size; active_list int32 [>= m]."""
    _cuda(src, "src")
    if src.shape[1:] != out.shape[1:] or src.dtype != out.dtype:
        raise ValueError("compact_rows: row shapes / dtypes differ")
    if src.shape[0] and (src.stride(0) != src[0].numel() or not src[0].is_contiguous()) or \
            out.shape[0] and (out.stride(0) != out[0].numel() or not out[0].is_contiguous()):
        raise ValueError("compact_rows: rows must be dense")
    row_bytes = (src[0].numel() if src.shape[0] else 0) * src.element_size()
    with torch.cuda.device(src.device):
        check(lib().eeseg_compact_rows(src.data_ptr(), out.data_ptr(), active_list.data_ptr(), _p(active_count),
                                       src.shape[0], out.shape[0], row_bytes, _stream(src)), "eeseg_compact_rows")
    return out


def upsample_bilinear(x, out_hw, out=None, out_dtype=None, layout="NCHW", n_classes=None):
    """F.interpolate(x, size=out_hw, mode='bilinear', align_corners=False) into planes [N,C,H,W]."""
    r = exit_gate(x, out_hw, layout=layout, n_classes=n_classes, want_amax=False, want_score=False,
                  up_out=out, up_dtype=(out_dtype or (x.dtype if out is None else None)))
    return r.up_logits


class _UpsampleBilinear(torch.autograd.Function):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) for NCHW fp32 x with autograd: forward
    on eeseg_upsample_bilinear, backward on the gather-form adjoint kernel (deterministic)."""

    @staticmethod
    def forward(ctx, x, out_hw):
        ctx.in_hw = tuple(x.shape[-2:])
        ctx.out_hw = (int(out_hw[0]), int(out_hw[1]))
        return upsample_bilinear(x.contiguous(), ctx.out_hw)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N, C = g.shape[:2]
        h, w = ctx.in_hw
        H, W = ctx.out_hw
        dx = torch.empty((N, C, h, w), dtype=torch.float32, device=g.device)
        _dt(g)
        _ops.upsample_bilinear_bwd(g, N * C, h, w, H, W, dx)
        return dx, None


def upsample_bilinear_autograd(x, out_hw):
    """Differentiable bilinear up-sampling of an NCHW fp32 CUDA tensor on the eeseg kernels."""
    _cuda(x, "x")
    if x.dtype != torch.float32:
        x = x.float()
    return _UpsampleBilinear.apply(x, out_hw)


# --------------------------------------------------------------------------------------------------
# multi-exit cross-entropy
# --------------------------------------------------------------------------------------------------
class _MultiExitCE(torch.autograd.Function):
    """per_exit[e] = mean over valid pixels of -log softmax(y[e])[t]. When y needs a gradient the
    forward kernel also writes d per_exit/dy scaled by `coef` (the d total/d per_exit the caller is
    about to apply); backward rescales only if the incoming gradient differs (decided on device)."""

    @staticmethod
    def forward(ctx, y, targets, ignore_index, coef):
        E, N, C = y.shape[:3]
        HW = y[0, 0, 0].numel()
        dev = y.device
        need_grad = y.requires_grad
        with torch.cuda.device(dev):
            per_exit = torch.empty((E,), dtype=torch.float32, device=dev)
            valid = torch.empty((1,), dtype=torch.int64, device=dev)
            ws = torch.empty((lib().eeseg_multi_exit_ce_workspace_bytes(E, N, HW),), dtype=torch.uint8, device=dev)
            dy = torch.empty_like(y) if need_grad else None
            _dt(y)
            _ops.multi_exit_ce_fwd(y, targets, int(ignore_index), coef, per_exit, valid, dy, ws)
        ctx.dy = dy
        ctx.save_for_backward(coef)
        ctx.mark_non_differentiable(valid)
        return per_exit, valid

    @staticmethod
    def backward(ctx, g, _gv):
        dy = ctx.dy
        (coef,) = ctx.saved_tensors
        if dy is None:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("multi_exit_ce: the fused gradient was consumed by an earlier backward (the forward "
                                   "kernel writes it once); call the loss again instead of retain_graph=True")
            return None, None, None, None
        g = g.contiguous().float()
        _ops.scale_exits(dy, g, coef)
        ctx.dy = None
        return dy, None, None, None


def multi_exit_ce(y, targets, ignore_index=-100, coef=None):
    """y [E,N,C,*spatial] f32/bf16 CUDA; targets [N,*spatial] int64. Returns (per_exit f32 [E],
    valid_count int64 [1]). coef f32 [E]: expected upstream gradient per exit (default ones)."""
    _cuda(y, "y_pred")
    _cuda(targets, "targets")
    if y.dim() < 4:
        raise ValueError("y_pred must be [E,N,C,...]")
    if not y.is_contiguous():
        y = y.contiguous()
    E, N = y.shape[:2]
    tg = targets.reshape(N, -1)
    if tg.dtype != torch.int64:
        tg = tg.to(torch.int64)
    tg = tg.contiguous()
    if tg.shape[1] != y[0, 0, 0].numel():
        raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(y.shape)}")
    if coef is None:
        coef = torch.ones((E,), dtype=torch.float32, device=y.device)
    coef = coef.detach().to(device=y.device, dtype=torch.float32).contiguous()
    return _MultiExitCE.apply(y, tg, ignore_index, coef)


def multi_exit_ce_backward_unfused(y, targets, ignore_index, g, valid):
    """Second-pass gradient (reads logits again): g[e]/valid * (softmax - onehot)."""
    E, N, C = y.shape[:3]
    HW = y[0, 0, 0].numel()
    dy = torch.empty_like(y)
    tg = targets.reshape(N, -1).to(torch.int64).contiguous()
    _dt(y)
    _ops.multi_exit_ce_bwd(y, tg, int(ignore_index), g.float().contiguous(), valid, dy)
    return dy


# --------------------------------------------------------------------------------------------------
# soft-overlap sums (Dice / Jaccard family)
# --------------------------------------------------------------------------------------------------
class _SoftOverlap(torch.autograd.Function):
    """(S_pt, S_p, S_t) of eeseg_soft_overlap_fwd for y [E,N,C,HW...]; differentiable in y through S_pt and S_p."""

    @staticmethod
    def forward(ctx, y, targets):
        E, N, C = y.shape[:3]
        HW = y[0, 0, 0].numel()
        dev = y.device
        with torch.cuda.device(dev):
            sums = torch.empty((E, N, 3, C), dtype=torch.float32, device=dev)
            ws = torch.empty((lib().eeseg_soft_overlap_workspace_bytes(E, N, C, HW),), dtype=torch.uint8, device=dev)
            check(lib().eeseg_soft_overlap_fwd(y.data_ptr(), _dt(y), y.stride(0), targets.data_ptr(), E, N, C, HW,
                                               sums.data_ptr(), ws.data_ptr(), _stream(y)), "eeseg_soft_overlap_fwd")
        ctx.save_for_backward(y, targets)
        s_t = sums[0, :, 2]
        ctx.mark_non_differentiable(s_t)
        return sums[:, :, 0], sums[:, :, 1], s_t

    @staticmethod
    def backward(ctx, g_pt, g_p, _g_t):
        y, targets = ctx.saved_tensors
        E, N, C = y.shape[:3]
        HW = y[0, 0, 0].numel()
        zeros = None
        if g_pt is None or g_p is None:
            zeros = torch.zeros((E, N, C), dtype=torch.float32, device=y.device)
        a = (zeros if g_pt is None else g_pt).contiguous().float()
        b = (zeros if g_p is None else g_p).contiguous().float()
        dy = torch.empty_like(y)
        with torch.cuda.device(y.device):
            check(lib().eeseg_soft_overlap_bwd(y.data_ptr(), _dt(y), y.stride(0), targets.data_ptr(), E, N, C, HW,
                                               a.data_ptr(), b.data_ptr(), dy.data_ptr(), _stream(y)), "eeseg_soft_overlap_bwd")
        return dy, None


def soft_overlap_sums(y, targets):
    """y [E,N,C,*spatial] f32/bf16 CUDA logits, targets [N,(1,)*spatial] integer labels. Returns
    S_pt [E,N,C] = sum_px softmax(y)[c]*[t==c], S_p [E,N,C] = sum_px softmax(y)[c], S_t [N,C] = sum_px [t==c]
    (fp32; labels outside [0,C) match no class). Differentiable with respect to y."""
    _cuda(y, "y_pred")
    _cuda(targets, "targets")
    if y.dim() < 4:
        raise ValueError("y_pred must be [E,N,C,...]")
    if not y.is_contiguous():
        y = y.contiguous()
    N = y.shape[1]
    tg = targets.reshape(N, -1)
    if tg.dtype != torch.int64:
        tg = tg.to(torch.int64)
    tg = tg.contiguous()
    if tg.shape[1] != y[0, 0, 0].numel():
        raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(y.shape)}")
    return _SoftOverlap.apply(y, tg)


# --------------------------------------------------------------------------------------------------
# focal loss
# --------------------------------------------------------------------------------------------------
def _focal_call(y, tg, gamma, alpha, pixw, coef, per_exit, loss_map, dy):
    E, N, C = y.shape[:3]
    HW = y[0, 0, 0].numel()
    if pixw is None:
        pw, pes, pis = None, 0, 0
    else:
        pw = pixw
        pes = pw.stride(0) if pw.dim() == 3 else 0
        pis = pw.stride(-2) if pw.dim() >= 2 else 0
    with torch.cuda.device(y.device):
        ws = torch.empty((lib().eeseg_focal_workspace_bytes(E, N, HW),), dtype=torch.uint8, device=y.device)
        check(lib().eeseg_focal_fwd(y.data_ptr(), _dt(y), y.stride(0), tg.data_ptr(), E, N, C, HW, float(gamma),
                                    _p(alpha), _p(pw), pes, pis, _p(coef), per_exit.data_ptr(), _p(loss_map), _p(dy),
                                    ws.data_ptr(), _stream(y)), "eeseg_focal_fwd")


class _FocalSum(torch.autograd.Function):
    """per_exit[e] = sum over (n, px) of the weighted focal loss; the same pass writes d per_exit/dy scaled by `coef`
    (the upstream gradient the caller is about to apply; backward rescales on the device only if it differs)."""

    @staticmethod
    def forward(ctx, y, tg, gamma, alpha, pixw, coef):
        E = y.shape[0]
        per_exit = torch.empty((E,), dtype=torch.float32, device=y.device)
        dy = torch.empty_like(y) if y.requires_grad else None
        _focal_call(y, tg, gamma, alpha, pixw, coef, per_exit, None, dy)
        ctx.dy = dy
        ctx.save_for_backward(coef)
        return per_exit

    @staticmethod
    def backward(ctx, g):
        dy = ctx.dy
        (coef,) = ctx.saved_tensors
        if dy is None:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("focal_sums: the fused gradient was consumed by an earlier backward; call the loss again "
                                   "instead of retain_graph=True")
            return None, None, None, None, None, None
        g = g.contiguous().float()
        _ops.scale_exits(dy, g, coef)
        ctx.dy = None
        return dy, None, None, None, None, None


class _FocalMap(torch.autograd.Function):
    """loss[e,n,px] (reduction='none'); backward is a second pass with the incoming gradient map as pixel weight."""

    @staticmethod
    def forward(ctx, y, tg, gamma):
        E, N = y.shape[:2]
        loss_map = torch.empty((E, N, tg.shape[1]), dtype=torch.float32, device=y.device)
        per_exit = torch.empty((E,), dtype=torch.float32, device=y.device)
        _focal_call(y, tg, gamma, None, None, None, per_exit, loss_map, None)
        ctx.save_for_backward(y, tg)
        ctx.gamma = gamma
        return loss_map

    @staticmethod
    def backward(ctx, g):
        y, tg = ctx.saved_tensors
        dy = torch.empty_like(y)
        per_exit = torch.empty((y.shape[0],), dtype=torch.float32, device=y.device)
        _focal_call(y, tg, ctx.gamma, None, g.contiguous().float(), None, per_exit, None, dy)
        return dy, None, None


def _focal_args(y, targets):
    _cuda(y, "y_pred")
    _cuda(targets, "targets")
    if y.dim() < 4:
        raise ValueError("y_pred must be [E,N,C,...]")
    if not y.is_contiguous():
        y = y.contiguous()
    tg = targets.reshape(y.shape[1], -1)
    if tg.dtype != torch.int64:
        tg = tg.to(torch.int64)
    tg = tg.contiguous()
    if tg.shape[1] != y[0, 0, 0].numel():
        raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(y.shape)}")
    return y, tg


def focal_sums(y, targets, gamma=2.0, alpha=None, pixel_weight=None, coef=None):
    """y [E,N,C,*spatial] f32/bf16 CUDA logits, targets [N,(1,)*spatial] labels in [0,C). Returns f32 [E]:
    sum over (n, px) of -alpha[t] * w * (1 - p_t)^gamma * log p_t. alpha f32 [C] and pixel_weight f32 [HW] / [N,HW] /
    [E,N,HW] are optional; coef f32 [E] is the expected upstream gradient per exit (default ones)."""
    y, tg = _focal_args(y, targets)
    dev = y.device
    f = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
    coef = torch.ones((y.shape[0],), dtype=torch.float32, device=dev) if coef is None else f(coef)
    return _FocalSum.apply(y, tg, float(gamma), f(alpha), f(pixel_weight), coef)


def focal_map(y, targets, gamma=2.0):
    """Per-pixel focal loss f32 [E,N,HW] (no class weights), differentiable in y."""
    y, tg = _focal_args(y, targets)
    return _FocalMap.apply(y, tg, float(gamma))


# --------------------------------------------------------------------------------------------------
# Lovasz-softmax
# --------------------------------------------------------------------------------------------------
class _Lovasz(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, labels, has_ignore, ignore, classes_mode, per_image):
        E, N, C = y.shape[:3]
        HW = y[0, 0, 0].numel()
        dev = y.device
        need_grad = y.requires_grad
        with torch.cuda.device(dev):
            per_exit = torch.empty((E,), dtype=torch.float32, device=dev)
            nbytes = lib().eeseg_lovasz_workspace_bytes(E, N, C, HW)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            dy = torch.empty_like(y) if need_grad else None
            _dt(y)
            _ops.lovasz_fwd_bwd(y, labels, bool(has_ignore), int(ignore), int(classes_mode), bool(per_image), per_exit,
                                dy, ws)
        ctx.dy = dy
        return per_exit

    @staticmethod
    def backward(ctx, g):
        dy = ctx.dy
        if dy is None:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("lovasz_multi_exit: the fused gradient was consumed by an earlier backward; call the loss "
                                   "again instead of retain_graph=True")
            return (None,) * 6
        ctx.dy = None
        g = g.contiguous().float()
        _ops.scale_exits(dy, g, None)   # dy[e] *= g[e]; a no-op on the device when g == 1
        return dy, None, None, None, None, None


def lovasz_multi_exit(y, labels, classes="present", per_image=False, ignore=None):
    """y [E,N,C,H,W] ('probas' — raw logits on the reference path), labels [N,(1,)H,W].
    Returns per-exit losses f32 [E] (differentiable w.r.t. y)."""
    _cuda(y, "probas")
    _cuda(labels, "labels")
    if classes not in ("present", "all"):
        raise NotImplementedError("eeseg Lovasz kernel supports classes='present' or 'all'")
    if not y.is_contiguous():
        y = y.contiguous()
    N = y.shape[1]
    lab = labels.reshape(N, -1)
    if lab.dtype != torch.int64:
        lab = lab.to(torch.int64)
    lab = lab.contiguous()
    if lab.shape[1] != y[0, 0, 0].numel():
        raise ValueError(f"labels {tuple(labels.shape)} do not match probas {tuple(y.shape)}")
    return _Lovasz.apply(y, lab, ignore is not None, 0 if ignore is None else int(ignore),
                         0 if classes == "present" else 1, bool(per_image))
