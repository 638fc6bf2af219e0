"""`torch.ops.eeseg.*` — the torch custom-op layer over the C ABI (include/eeseg.h, libeeseg_b200.so).

SURVEY.md §8(b) / north_star: the Python host code reaches the kernels through registered torch operators. Each
operator here is a thin shim: its schema is declared with `torch.library`, its ONLY implementation is registered for
the CUDA dispatch key and forwards raw device pointers, sizes and the current stream to the `extern "C"` launcher of
the same name. There is no CPU (or Meta / autograd-fallback) kernel, so a CPU tensor fails in the dispatcher — there
is no fallback by construction. Outputs are passed in pre-allocated (`Tensor(a!)`), exactly as the C ABI wants them;
allocation, argument checks and autograd glue live in ops.py / head_train.py, which call these operators.

Registered (the launchers §8(b) names): conv_igemm_fwd / _grouped / _dgrad / _wgrad, exit_gate_pixels, exit_gate_decide,
multi_exit_ce_fwd / _bwd, scale_exits, lovasz_fwd_bwd, confusion_hist, upsample_bilinear_bwd. The remaining launchers
(BatchNorm, soft-overlap, focal, max-pool, stem, compaction, stage commit, pooling) are bound the same way through the
ctypes table in _lib.py by the modules that use them; tests/ exercises both routes.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib

_L = torch.library.Library("eeseg", "DEF")


def _p(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dt(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"eeseg kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _op(schema):
    name = schema.split("(", 1)[0]
    _L.define(schema)

    def deco(fn):
        _L.impl(name, fn, "CUDA")
        return fn
    return deco


# ---- convolutions ----------------------------------------------------------------------------------------------
@_op("conv_igemm_fwd(Tensor x, Tensor wt, Tensor? scale, Tensor shift, int shift_sn, int dilation, int stride, int pad, "
     "bool relu, Tensor? residual, Tensor(a!) out, int ldo) -> ()")
def _conv_igemm_fwd(x, wt, scale, shift, shift_sn, dilation, stride, pad, relu, residual, out, ldo):
    N, h, w, Cin = x.shape
    Cout, R, S, _ = wt.shape
    with torch.cuda.device(x.device):
        check(lib().eeseg_conv_igemm_fwd(
            x.data_ptr(), wt.data_ptr(), _p(scale), shift.data_ptr(), shift_sn, N, h, w, Cin, Cout, R, S,
            dilation, stride, pad, 1 if relu else 0, _p(residual), 0 if residual is None else residual.stride(2),
            out.data_ptr(), _dt(out), ldo, _stream(x)), "eeseg_conv_igemm_fwd")


@_op("conv_igemm_grouped(Tensor x, Tensor[] wts, Tensor[]? scales, Tensor[] shifts, int[] ksizes, int[] dils, int[] ch_offs, "
     "bool relu, Tensor(a!) out, int ldo, int out_channels, Tensor schedule, bool cta_pairs) -> ()")
def _conv_igemm_grouped(x, wts, scales, shifts, ksizes, dils, ch_offs, relu, out, ldo, out_channels, schedule, cta_pairs):
    N, h, w, Cin = x.shape
    n = len(wts)
    Cout = wts[0].shape[0]
    PA = ctypes.c_void_p * n
    IA = ctypes.c_int * n
    with torch.cuda.device(x.device):
        check(lib().eeseg_conv_igemm_grouped(
            x.data_ptr(), n, PA(*[t.data_ptr() for t in wts]),
            PA(*([t.data_ptr() for t in scales] if scales is not None else [None] * n)),
            PA(*[t.data_ptr() for t in shifts]), IA(*ksizes), IA(*dils), IA(*ch_offs), N, h, w, Cin, Cout,
            1 if relu else 0, out.data_ptr(), ldo, out_channels, schedule.data_ptr(), schedule.numel(),
            1 if cta_pairs else 0, _stream(x)), "eeseg_conv_igemm_grouped")


@_op("conv_igemm_dgrad(Tensor dy, Tensor wt, int dilation, Tensor(a!) dx, Tensor(b!) workspace) -> ()")
def _conv_igemm_dgrad(dy, wt, dilation, dx, workspace):
    N, h, w, Cin = dx.shape
    Cout, R, S, _ = wt.shape
    with torch.cuda.device(dy.device):
        check(lib().eeseg_conv_igemm_dgrad(dy.data_ptr(), wt.data_ptr(), N, h, w, Cin, Cout, R, S, dilation,
                                           dx.data_ptr(), _dt(dx), Cin, workspace.data_ptr(), _stream(dy)),
              "eeseg_conv_igemm_dgrad")


@_op("conv_igemm_wgrad(Tensor x, Tensor dy, int dilation, Tensor(a!) dw, Tensor(b!) workspace) -> ()")
def _conv_igemm_wgrad(x, dy, dilation, dw, workspace):
    N, h, w, Cin = x.shape
    Cout, R, S, _ = dw.shape
    with torch.cuda.device(x.device):
        check(lib().eeseg_conv_igemm_wgrad(x.data_ptr(), dy.data_ptr(), Cout, Cout, 0, N, h, w, Cin, Cout, R, S,
                                           dilation, dw.data_ptr(), workspace.data_ptr(), _stream(x)),
              "eeseg_conv_igemm_wgrad")


# ---- exit gate -------------------------------------------------------------------------------------------------
@_op("exit_gate_pixels(Tensor x, int kind, int sn, int sc, int sy, int sx, int N, int C, int h, int w, int H, int W, "
     "float tau, Tensor(a!)? up_out, int up_sn, Tensor(b!)? ent, Tensor(c!)? amax, Tensor(d!)? mask, "
     "Tensor(e!)? part_sum, Tensor(f!)? part_cnt) -> ()")
def _exit_gate_pixels(x, kind, sn, sc, sy, sx, N, C, h, w, H, W, tau, up_out, up_sn, ent, amax, mask, part_sum, part_cnt):
    with torch.cuda.device(x.device):
        check(lib().eeseg_exit_gate_pixels(
            x.data_ptr(), _dt(x), kind, sn, sc, sy, sx, N, C, h, w, H, W, float(tau), _p(up_out),
            _dt(up_out) if up_out is not None else 0, up_sn, _p(ent), _p(amax), _p(mask), _p(part_sum), _p(part_cnt),
            _stream(x)), "eeseg_exit_gate_pixels")


@_op("exit_gate_decide(Tensor? part_sum, Tensor? part_cnt, int npart, Tensor? score_in, int N, int HW, float tau, "
     "bool less_than, int exit_id, Tensor(a!)? exit_idx, Tensor(b!)? score_out, Tensor(c!)? exited_px, "
     "Tensor(d!)? active_list, Tensor(e!)? active_count) -> ()")
def _exit_gate_decide(part_sum, part_cnt, npart, score_in, N, HW, tau, less_than, exit_id, exit_idx, score_out,
                      exited_px, active_list, active_count):
    ref = next(t for t in (part_sum, score_in, exit_idx, score_out) if t is not None)
    with torch.cuda.device(ref.device):
        check(lib().eeseg_exit_gate_decide(_p(part_sum), _p(part_cnt), npart, _p(score_in), N, HW, float(tau),
                                           1 if less_than else 0, exit_id, _p(exit_idx), _p(score_out), _p(exited_px),
                                           _p(active_list), _p(active_count), _stream(ref)), "eeseg_exit_gate_decide")


@_op("upsample_bilinear_bwd(Tensor g, int planes, int h, int w, int H, int W, Tensor(a!) dx) -> ()")
def _upsample_bilinear_bwd(g, planes, h, w, H, W, dx):
    with torch.cuda.device(g.device):
        check(lib().eeseg_upsample_bilinear_bwd(g.data_ptr(), _dt(g), planes, h, w, H, W, dx.data_ptr(), _stream(g)),
              "eeseg_upsample_bilinear_bwd")


# ---- losses ----------------------------------------------------------------------------------------------------
@_op("multi_exit_ce_fwd(Tensor y, Tensor targets, int ignore_index, Tensor coef, Tensor(a!) per_exit, Tensor(b!) valid, "
     "Tensor(c!)? dy, Tensor(d!) workspace) -> ()")
def _multi_exit_ce_fwd(y, targets, ignore_index, coef, per_exit, valid, dy, workspace):
    E, N, C = y.shape[:3]
    HW = y[0, 0, 0].numel()
    with torch.cuda.device(y.device):
        check(lib().eeseg_multi_exit_ce_fwd(
            y.data_ptr(), _dt(y), y.stride(0), targets.data_ptr(), E, N, C, HW, int(ignore_index), coef.data_ptr(),
            per_exit.data_ptr(), valid.data_ptr(), _p(dy), workspace.data_ptr(), _stream(y)), "eeseg_multi_exit_ce_fwd")


@_op("multi_exit_ce_bwd(Tensor y, Tensor targets, int ignore_index, Tensor g, Tensor valid, Tensor(a!) dy) -> ()")
def _multi_exit_ce_bwd(y, targets, ignore_index, g, valid, dy):
    E, N, C = y.shape[:3]
    HW = y[0, 0, 0].numel()
    with torch.cuda.device(y.device):
        check(lib().eeseg_multi_exit_ce_bwd(y.data_ptr(), _dt(y), y.stride(0), targets.data_ptr(), E, N, C, HW,
                                            int(ignore_index), g.data_ptr(), valid.data_ptr(), dy.data_ptr(), _stream(y)),
              "eeseg_multi_exit_ce_bwd")


@_op("scale_exits(Tensor(a!) dy, Tensor g, Tensor? coef) -> ()")
def _scale_exits(dy, g, coef):
    with torch.cuda.device(dy.device):
        check(lib().eeseg_scale_exits(dy.data_ptr(), _dt(dy), dy.stride(0), dy.shape[0], dy[0].numel(), g.data_ptr(),
                                      _p(coef), _stream(dy)), "eeseg_scale_exits")


@_op("lovasz_fwd_bwd(Tensor y, Tensor labels, bool has_ignore, int ignore, int classes_mode, bool per_image, "
     "Tensor(a!) per_exit, Tensor(b!)? dy, Tensor(c!) workspace) -> ()")
def _lovasz_fwd_bwd(y, labels, has_ignore, ignore, classes_mode, per_image, per_exit, dy, workspace):
    E, N, C = y.shape[:3]
    HW = y[0, 0, 0].numel()
    with torch.cuda.device(y.device):
        check(lib().eeseg_lovasz_fwd_bwd(
            y.data_ptr(), _dt(y), y.stride(0), labels.data_ptr(), E, N, C, HW, int(has_ignore), int(ignore),
            int(classes_mode), int(per_image), per_exit.data_ptr(), _p(dy), workspace.data_ptr(), workspace.numel(),
            _stream(y)), "eeseg_lovasz_fwd_bwd")


# ---- metric ----------------------------------------------------------------------------------------------------
@_op("confusion_hist(Tensor pred, int pred_kind, Tensor targets, int n_classes, Tensor(a!) out, bool accumulate) -> ()")
def _confusion_hist(pred, pred_kind, targets, n_classes, out, accumulate):
    N = pred.shape[0]
    HW = targets.shape[1]
    with torch.cuda.device(pred.device):
        check(lib().eeseg_confusion_hist(pred.data_ptr(), pred_kind, _dt(pred) if pred_kind == 0 else 0,
                                         targets.data_ptr(), N, int(n_classes), HW, out.data_ptr(),
                                         1 if accumulate else 0, _stream(pred)), "eeseg_confusion_hist")


ops = torch.ops.eeseg


class _Handles:
    """`torch.ops.eeseg.<name>.default` resolved once: calling the OpOverload directly skips the per-call packet lookup and
    overload resolution of `torch.ops.eeseg.<name>(...)` (the eager, un-graphed host path makes ~100 such calls per step)."""

    def __getattr__(self, name):
        h = getattr(torch.ops.eeseg, name).default
        setattr(self, name, h)
        return h


fast = _Handles()
