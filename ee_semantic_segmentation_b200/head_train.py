"""Training-mode forward/backward of a DeepLabHead / my_branch exit head with its convolutions on the
eeseg tcgen05 kernels (forward `eeseg_conv_igemm_fwd`, input gradient `eeseg_conv_igemm_dgrad`, weight
gradient `eeseg_conv_igemm_wgrad`) — what autograd through cuDNN computes for
`branches[i](X)` / `classifier(X)` (from_deepv3_new.py:147,151) inside `train_epoch`
(train_funcs.py:22-27).

Activations of the head are bf16 NHWC; BatchNorm (batch statistics) + ReLU run as fused eeseg nodes
(bn_train.BnActFn) on the modules' own parameters and running-statistic buffers. The pooled ASPP branch (global average
pool, 1x1 convolution on one pixel per image, BatchNorm over the batch, ReLU, broadcast back over the map), Dropout(0.5)
after the projection (a counter-based generator with its state in device memory: graph replays draw fresh masks; the
stream of torch's generator is not reproduced — no kernel-for-kernel port could) and the final Cout = num_classes 1x1
convolution (+ bias; forward on a 32-column tile, gradients on 64-channel padded tensors) run on eeseg kernels as well:
the modules are parameter / buffer containers only, state-dict layout and running statistics are the reference's. Weight
gradients are returned in fp32 in the parameter's own [Cout,Cin,R,S] layout. torch is left with memory plumbing
(layout / precision copies, zero fills, the channel concatenation of the five ASPP branches).
"""
import torch
from torch import nn
from torchvision.models.segmentation.deeplabv3 import ASPP

from . import _lib, torch_ops
from ._lib import check, lib

_UNIT = {}


def _unit_scale_shift(dev, n):
    key = (dev, n)
    if key not in _UNIT:
        _UNIT[key] = (torch.ones(n, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.float32, device=dev))
    return _UNIT[key]


class TrainWeightCache:
    """bf16 copies of every supported convolution weight of a model in the two layouts a training step reads —
    [Cout,R,S,Cin] (forward, weight gradient) and the rotated transpose [Cin,R,S,Cout] (input gradient) — refreshed by ONE
    launch per step (eeseg_weight_prep_multi) instead of three small launches per convolution per step. Only consulted
    inside `with cache.active():`, which the model's training forward opens right after `refresh()`; anywhere else
    ConvIgemmFn converts per call, so a stale copy can never be read."""

    def __init__(self, convs):
        import numpy as np
        self.weights = [c.weight for c in convs]
        dev = self.weights[0].device
        total = sum(w.numel() for w in self.weights)
        self.krsc = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.rot = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.entries, rows, o, self.max_rs = {}, [], 0, 1
        assert lib().eeseg_weight_prep_tile_bytes() == 48
        for w in self.weights:
            Cout, Cin, R, S = w.shape
            n = w.numel()
            wt = self.krsc[o:o + n].view(Cout, R, S, Cin)
            wT = self.rot[o:o + n].view(Cin, R, S, Cout)
            self.entries[id(w)] = (wt, wT)
            self.max_rs = max(self.max_rs, R * S)
            for co0 in range(0, Cout, 32):
                for ci0 in range(0, Cin, 32):
                    rows.append((w.data_ptr(), wt.data_ptr(), wT.data_ptr(), Cout, Cin, R * S, co0, ci0, 0))
            o += n
        table = np.zeros(len(rows), dtype=np.dtype([('s', '<u8'), ('k', '<u8'), ('r', '<u8'), ('co', '<i4'), ('ci', '<i4'),
                                                    ('rs', '<i4'), ('co0', '<i4'), ('ci0', '<i4'), ('pad', '<i4')]))
        for i, r in enumerate(rows):
            table[i] = r
        self.table = torch.from_numpy(table.view(np.uint8).copy()).to(dev)
        self.n_tiles = len(rows)
        self.ptrs = [w.data_ptr() for w in self.weights]

    def valid(self):
        return all(w.data_ptr() == p for w, p in zip(self.weights, self.ptrs))

    def refresh(self):
        dev = self.table.device
        with torch.cuda.device(dev):
            check(lib().eeseg_weight_prep_multi(self.table.data_ptr(), self.n_tiles, self.max_rs,
                                                torch.cuda.current_stream(dev).cuda_stream), "eeseg_weight_prep_multi")

    def active(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            global _CUR_WCACHE
            prev, _CUR_WCACHE = _CUR_WCACHE, self
            try:
                yield
            finally:
                _CUR_WCACHE = prev
        return ctx()


_CUR_WCACHE = None


def trainable_convs(model):
    """The convolutions of a model that ConvIgemmFn runs (everything _conv_ok admits with 32-channel multiples)."""
    return [m for m in model.modules() if isinstance(m, nn.Conv2d) and _conv_ok(m) and m.weight.is_cuda
            and m.weight.dtype == torch.float32]


class ConvIgemmFn(torch.autograd.Function):
    """y = conv2d(x, weight; stride 1 or 2, padding dilation*(R//2), no bias) on NHWC bf16 activations.
    x [N,h,w,Cin] bf16 contiguous (Cin % 64 == 0); weight: the nn.Conv2d parameter [Cout,Cin,R,S]
    (Cout % 64 == 0). Returns [N,ho,wo,Cout] bf16, ho = (h-1)//stride + 1.

    Stride 2 (layer2.0 of the ResNet): the forward kernel strides its TMA boxes; the backward inserts zeros
    between the rows / columns of dY (a [N,h,w,Cout] tensor, 3 of 4 pixels zero) and runs the stride-1 input- and
    weight-gradient kernels on it — the transposed convolution and the strided correlation are exactly those."""

    @staticmethod
    def forward(ctx, x, weight, dilation, stride=1):
        from .head_plan import conv_igemm
        x = x.contiguous()
        N, h, w, Cin = x.shape
        Cout, _, R, S = weight.shape
        ent = _CUR_WCACHE.entries.get(id(weight)) if _CUR_WCACHE is not None else None
        if ent is not None:                                                            # prepared once per step for all layers
            wt, ctx.wT = ent
        else:
            ctx.wT = None
            wt = torch.empty((Cout, R, S, Cin), dtype=torch.bfloat16, device=x.device)    # [Cout,R,S,Cin]
            wt.copy_(weight.detach().permute(0, 2, 3, 1))                                  # layout + precision in one pass
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = torch.empty((N, ho, wo, Cout), dtype=torch.bfloat16, device=x.device)
        one, zero = _unit_scale_shift(x.device, Cout)
        conv_igemm(x, wt, None, zero, dilation, False, out, _lib.BF16, Cout, stride=stride)
        ctx.save_for_backward(x, wt)
        ctx.dilation = dilation
        ctx.stride = stride
        ctx.weight = weight if isinstance(weight, nn.Parameter) else None   # for the direct-to-.grad weight gradient
        return out

    @staticmethod
    def backward(ctx, dy):
        x, wt = ctx.saved_tensors
        N, h, w, Cin = x.shape
        Cout, R, S, _ = wt.shape
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        if ctx.stride != 1:
            up = torch.zeros((N, h, w, Cout), dtype=torch.bfloat16, device=x.device)
            up[:, ::ctx.stride, ::ctx.stride] = dy
            dy = up
        dy = dy.contiguous()
        dx = dw = None
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream(x.device).cuda_stream
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                if ctx.wT is not None:       # the rotated transpose is in the per-step cache: dgrad = the forward kernel on it
                    from .head_plan import conv_igemm
                    one, zero = _unit_scale_shift(x.device, Cin)
                    conv_igemm(dy, ctx.wT, None, zero, ctx.dilation, False, dx, _lib.BF16, Cin)
                else:
                    ws = torch.empty((lib().eeseg_conv_igemm_dgrad_workspace_bytes(Cin, Cout, R, S),), dtype=torch.uint8,
                                     device=x.device)
                    torch_ops.fast.conv_igemm_dgrad(dy, wt, ctx.dilation, dx, ws)
            if ctx.needs_input_grad[1]:
                from .parallel import direct_grad
                g = direct_grad(ctx.weight) if ctx.weight is not None else None
                if g is not None and tuple(g.shape) == (Cout, Cin, R, S):
                    # straight into the parameter's .grad (its own [Cout,Cin,R,S] layout): no AccumulateGrad launch
                    wws = torch.empty((lib().eeseg_conv_igemm_wgrad_to_param_workspace_bytes(N, h, w, Cin, Cout, R, S),),
                                      dtype=torch.uint8, device=x.device)
                    check(lib().eeseg_conv_igemm_wgrad_to_param(x.data_ptr(), dy.data_ptr(), Cout, Cout, 0, N, h, w, Cin, Cout,
                                                                R, S, ctx.dilation, g.data_ptr(), 1, wws.data_ptr(), st),
                          "eeseg_conv_igemm_wgrad_to_param")
                    ctx.weight._eeseg_flat.written(ctx.weight)
                else:
                    dwk = torch.empty((Cout, R, S, Cin), dtype=torch.float32, device=x.device)
                    wws = torch.empty((lib().eeseg_conv_igemm_wgrad_workspace_bytes(N, h, w, Cin, Cout, R, S),), dtype=torch.uint8,
                                      device=x.device)
                    torch_ops.fast.conv_igemm_wgrad(x, dy, ctx.dilation, dwk, wws)
                    dw = dwk.permute(0, 3, 1, 2)                                         # the parameter's layout
        return dx, dw, None, None


def _conv_ok(conv):
    """Shapes the three kernels take: stride 1 or 2, padding dilation*(k//2), no groups/bias, channel multiples."""
    k, d = conv.kernel_size[0], conv.dilation[0]
    return (isinstance(conv, nn.Conv2d) and conv.bias is None and conv.groups == 1 and conv.stride in ((1, 1), (2, 2))
            and conv.kernel_size[0] == conv.kernel_size[1] and (k & 1) and conv.dilation[0] == conv.dilation[1]
            and conv.padding == (d * (k // 2), d * (k // 2)) and conv.padding_mode == 'zeros'
            and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0)


def head_supported(head):
    """DeepLabHead / my_branch without bottleneck: Sequential(ASPP, Conv3x3, BN, ReLU, Conv1x1)."""
    if not (isinstance(head, nn.Sequential) and len(head) == 5 and isinstance(head[0], ASPP)):
        return False
    aspp = head[0]
    convs = [m[0] for m in list(aspp.convs)[:-1]] + [aspp.project[0], head[1]]
    return all(_conv_ok(c) for c in convs) and isinstance(head[4], nn.Conv2d)


def _nhwc(t):
    """[N,C,h,w] (any memory format) -> contiguous [N,h,w,C] bf16."""
    return t.to(dtype=torch.bfloat16, memory_format=torch.channels_last).permute(0, 2, 3, 1)


def _conv(xh, conv):
    """NHWC bf16 in -> NCHW-shaped (channels_last) bf16 out, as the following BatchNorm2d expects."""
    return ConvIgemmFn.apply(xh, conv.weight, conv.dilation[0], conv.stride[0]).permute(0, 3, 1, 2)


class GlobalAvgPoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d(1) of an NHWC bf16 tensor -> fp32 [N,C]; backward broadcasts dy / HW over the map."""

    @staticmethod
    def forward(ctx, xh):
        from .head_plan import global_avgpool_nhwc
        ctx.shape = tuple(xh.shape)
        return global_avgpool_nhwc(xh.contiguous())

    @staticmethod
    def backward(ctx, dy):
        N, h, w, C = ctx.shape
        dx = torch.empty(ctx.shape, dtype=torch.bfloat16, device=dy.device)
        dy = dy.contiguous().float()
        with torch.cuda.device(dy.device):
            check(lib().eeseg_broadcast_rows_nhwc(dy.data_ptr(), N, h * w, C, 1.0 / (h * w), dx.data_ptr(),
                                                  torch.cuda.current_stream(dy.device).cuda_stream), "eeseg_broadcast_rows_nhwc")
        return dx


class BroadcastFn(torch.autograd.Function):
    """v fp32 [N,C] -> bf16 [N,h,w,C] (ASPPPooling's up-sampling of a 1x1 map is this constant, deeplabv3.py:83);
    backward: per-image channel sums of the incoming gradient (global average pool x HW)."""

    @staticmethod
    def forward(ctx, v, h, w):
        N, C = v.shape
        ctx.hw = (h, w)
        out = torch.empty((N, h, w, C), dtype=torch.bfloat16, device=v.device)
        v = v.contiguous().float()
        with torch.cuda.device(v.device):
            check(lib().eeseg_broadcast_rows_nhwc(v.data_ptr(), N, h * w, C, 1.0, out.data_ptr(),
                                                  torch.cuda.current_stream(v.device).cuda_stream), "eeseg_broadcast_rows_nhwc")
        return out

    @staticmethod
    def backward(ctx, g):
        from .head_plan import global_avgpool_nhwc
        h, w = ctx.hw
        g = g.to(torch.bfloat16).contiguous()
        return global_avgpool_nhwc(g) * float(h * w), None, None


class DenseFn(torch.autograd.Function):
    """y[n] = W x[n] for x fp32 [N,K], W the 1x1 conv parameter [O,K,1,1] (no bias): the pooled branch's convolution."""

    @staticmethod
    def forward(ctx, x, weight):
        from .head_plan import dense_bn_act
        W = weight.detach().reshape(weight.shape[0], -1).float().contiguous()
        x = x.contiguous().float()
        ctx.save_for_backward(x, W)
        ctx.wshape = tuple(weight.shape)
        N, O = x.shape[0], W.shape[0]
        y = torch.empty((N, O), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().eeseg_dense_bn_act(x.data_ptr(), W.data_ptr(), None, None, N, x.shape[1], O, 0, y.data_ptr(),
                                           torch.cuda.current_stream(x.device).cuda_stream), "eeseg_dense_bn_act")
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        N, K = x.shape
        O = W.shape[0]
        dy = dy.contiguous().float()
        dW = torch.empty_like(W) if ctx.needs_input_grad[1] else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(x.device):
            check(lib().eeseg_dense_bwd(dy.data_ptr(), x.data_ptr(), W.data_ptr(), N, K, O,
                                        None if dW is None else dW.data_ptr(), None if dx is None else dx.data_ptr(),
                                        torch.cuda.current_stream(x.device).cuda_stream), "eeseg_dense_bwd")
        return dx, None if dW is None else dW.view(ctx.wshape)


_RNG_STATE = {}


def dropout_state(dev, seed=None):
    """DEVICE uint64[2] {seed, offset} of the eeseg dropout generator for `dev` (created from torch's seed on first use;
    `seed=` re-seeds). The offset advances on the device with every dropout launch, graph replays included."""
    key = str(dev)
    if key not in _RNG_STATE or seed is not None:
        s = int(torch.initial_seed() if seed is None else seed) & 0x7fffffffffffffff
        _RNG_STATE[key] = torch.tensor([s, 0], dtype=torch.int64, device=dev)
    return _RNG_STATE[key]


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) in training mode on a bf16 tensor (numel % 8 == 0): eeseg_dropout_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        n = x.numel()
        y = torch.empty_like(x)
        mask = torch.empty((n // 8,), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().eeseg_dropout_fwd(x.data_ptr(), n, float(p), dropout_state(x.device).data_ptr(), y.data_ptr(),
                                          mask.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream), "eeseg_dropout_fwd")
        ctx.save_for_backward(mask)
        ctx.p = float(p)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dy = dy.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            check(lib().eeseg_dropout_bwd(dy.data_ptr(), mask.data_ptr(), dy.numel(), ctx.p, dx.data_ptr(),
                                          torch.cuda.current_stream(dy.device).cuda_stream), "eeseg_dropout_bwd")
        return dx, None


class FinalConvFn(torch.autograd.Function):
    """The head's last layer, Conv2d(256, num_classes, 1) with bias: y NHWC bf16 [N,h,w,Cin] -> logits fp32 NHWC
    [N,h,w,Cp] (Cp = classes padded to the 16-column MMA multiple; columns >= num_classes are zero). Backward: the
    incoming gradient (fp32, same padded layout) is rounded to bf16 on 64 padded channels for the input-gradient and
    weight-gradient kernels; the bias gradient is its channel sum (global-average-pool kernel x HW, then the batch)."""

    @staticmethod
    def forward(ctx, y, weight, bias):
        from .head_plan import conv_igemm
        y = y.contiguous()
        N, h, w, Cin = y.shape
        C = weight.shape[0]
        Cp = (C + 15) // 16 * 16
        dev = y.device
        wt = torch.zeros((64, 1, 1, Cin), dtype=torch.bfloat16, device=dev)       # 64 rows: shared with the backward
        wt[:C] = weight.detach().permute(0, 2, 3, 1)
        sh = torch.zeros((Cp,), dtype=torch.float32, device=dev)
        if bias is not None:
            sh[:C] = bias.detach()
        one, _ = _unit_scale_shift(dev, Cp)
        out = torch.empty((N, h, w, Cp), dtype=torch.float32, device=dev)
        conv_igemm(y, wt[:Cp], one, sh, 1, False, out, _lib.F32, Cp)
        ctx.save_for_backward(y, wt)
        ctx.C, ctx.has_bias = C, bias is not None
        ctx.wshape = tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        from .head_plan import global_avgpool_nhwc
        y, wt = ctx.saved_tensors
        N, h, w, Cin = y.shape
        C = ctx.C
        dev = y.device
        gp = torch.zeros((N, h, w, 64), dtype=torch.bfloat16, device=dev)
        gp[..., :C] = g[..., :C]
        dy = dw = db = None
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[0]:
                ws = torch.empty((lib().eeseg_conv_igemm_dgrad_workspace_bytes(Cin, 64, 1, 1),), dtype=torch.uint8, device=dev)
                dy = torch.empty_like(y)
                torch_ops.fast.conv_igemm_dgrad(gp, wt, 1, dy, ws)
            if ctx.needs_input_grad[1]:
                dwk = torch.empty((64, 1, 1, Cin), dtype=torch.float32, device=dev)
                wws = torch.empty((lib().eeseg_conv_igemm_wgrad_workspace_bytes(N, h, w, Cin, 64, 1, 1),), dtype=torch.uint8,
                                  device=dev)
                torch_ops.fast.conv_igemm_wgrad(y, gp, 1, dwk, wws)
                dw = dwk[:C].permute(0, 3, 1, 2).reshape(ctx.wshape)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                per_img = (global_avgpool_nhwc(gp) * float(h * w)).contiguous()       # [N,64] channel sums per image
                ones = torch.ones((N, 1), dtype=torch.float32, device=dev)
                dbk = torch.empty((64, 1), dtype=torch.float32, device=dev)
                check(lib().eeseg_dense_bwd(per_img.data_ptr(), ones.data_ptr(), ones.data_ptr(), N, 1, 64, dbk.data_ptr(), None,
                                            torch.cuda.current_stream(dev).cuda_stream), "eeseg_dense_bwd")
                db = dbk[:C, 0]
        return dy, dw, db


class BnRowsFn(torch.autograd.Function):
    """BatchNorm2d (training: statistics over the N rows) + ReLU on fp32 [N,C] row vectors, running statistics updated in
    place (eeseg_bn_rows_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous().float()
        N, C = x.shape
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty((C,), dtype=torch.float32, device=dev)
        invstd = torch.empty((C,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().eeseg_bn_rows_fwd(x.data_ptr(), N, C, weight.data_ptr(), bias.data_ptr(),
                                          None if running_mean is None else running_mean.data_ptr(),
                                          None if running_var is None else running_var.data_ptr(), float(momentum), float(eps),
                                          1 if relu else 0, y.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "eeseg_bn_rows_fwd")
        ctx.save_for_backward(x, y, weight, mean, invstd)
        ctx.relu = bool(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        N, C = x.shape
        dev = x.device
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dg = torch.empty((C,), dtype=torch.float32, device=dev)
        db = torch.empty((C,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().eeseg_bn_rows_bwd(dy.data_ptr(), x.data_ptr(), y.data_ptr(), N, C, weight.data_ptr(), mean.data_ptr(),
                                          invstd.data_ptr(), 1 if ctx.relu else 0, dx.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "eeseg_bn_rows_bwd")
        return dx, dg, db, None, None, None, None, None


def pooled_branch_train(pool_seq, xh):
    """ASPPPooling (deeplabv3.py:70-83) with autograd on the eeseg kernels: xh NHWC bf16 [N,h,w,Cin] -> bf16 NCHW-shaped
    (channels_last) [N,mid,h,w]. BatchNorm takes its batch statistics over the N pooled vectors, as the module does."""
    from .bn_train import bn_act
    N, h, w, _ = xh.shape
    conv, bn = pool_seq[1], pool_seq[2]
    v = DenseFn.apply(GlobalAvgPoolFn.apply(xh), conv.weight)                       # [N, mid] fp32
    # fp32 all the way: with a handful of samples the normalised values sit near +-1 and the input gradient is a small
    # difference of large terms
    v = BnRowsFn.apply(v, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, True)
    with torch.no_grad():
        bn.num_batches_tracked += 1
    return BroadcastFn.apply(v, h, w).permute(0, 3, 1, 2)


def head_forward_train(head, x, lowres_nhwc=False):
    """head(x) with autograd, x [N,Cin,h,w] (fp32 or bf16, any memory format) -> logits [N,C,h,w] fp32 (a view of the
    kernel's NHWC [N,h,w,Cp] output; lowres_nhwc=True returns that padded NHWC tensor itself)."""
    aspp = head[0]
    xh = _nhwc(x)
    outs = []
    from .bn_train import bn_act
    for m in list(aspp.convs)[:-1]:                      # 1x1 and the atrous 3x3 branches: conv, BN + ReLU
        outs.append(bn_act(_conv(xh, m[0]), m[1], True))
    outs.append(pooled_branch_train(aspp.convs[-1], xh))
    cat = torch.cat(outs, dim=1)
    y = bn_act(_conv(_nhwc(cat), aspp.project[0]), aspp.project[1], True)   # projection, BN + ReLU
    for m in list(aspp.project)[3:]:                     # Dropout(0.5)
        if isinstance(m, nn.Dropout) and m.training and m.p > 0 and y.numel() % 8 == 0:
            y = DropoutFn.apply(y.permute(0, 2, 3, 1), m.p).permute(0, 3, 1, 2)      # on the NHWC buffer underneath
        else:
            y = m(y)
    y = bn_act(_conv(_nhwc(y), head[1]), head[2], True)  # 3x3, BN + ReLU
    last = head[4]
    out = FinalConvFn.apply(_nhwc(y), last.weight, last.bias)               # final 1x1 (+bias) to num_classes
    if lowres_nhwc:
        return out
    return out[..., :last.out_channels].permute(0, 3, 1, 2)
