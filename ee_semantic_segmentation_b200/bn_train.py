"""Training-mode BatchNorm2d (+ residual add) (+ ReLU) on the eeseg kernels behind autograd — what the
nn.BatchNorm2d / nn.ReLU modules and the `out += identity` of torchvision's DeepLabHead, ASPP and ResNet
Bottleneck compute under net.train() inside train_epoch (train_funcs.py:12-33). The nn.BatchNorm2d module stays
the parameter / buffer container: gamma, beta, running_mean, running_var and num_batches_tracked are the
module's own tensors and are updated exactly as the module would (momentum, unbiased running variance)."""
import torch
from torch import nn

from ._lib import check, lib


class BnActFn(torch.autograd.Function):
    """y = act(batch_norm(x; batch statistics) (+ residual)) for x [N,h,w,C] bf16 contiguous (C % 64 == 0)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous()
        C = x.shape[-1]
        P = x.numel() // C
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty((C,), dtype=torch.float32, device=dev)
        invstd = torch.empty((C,), dtype=torch.float32, device=dev)
        ws = torch.empty((lib().eeseg_bn_train_workspace_bytes(C),), dtype=torch.uint8, device=dev)
        res = None if residual is None else residual.contiguous()
        with torch.cuda.device(dev):
            check(lib().eeseg_bn_train_fwd(
                x.data_ptr(), P, C, weight.data_ptr(), bias.data_ptr(),
                None if running_mean is None else running_mean.data_ptr(),
                None if running_var is None else running_var.data_ptr(), float(momentum), float(eps), 1 if relu else 0,
                None if res is None else res.data_ptr(), y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream), "eeseg_bn_train_fwd")
        ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
        ctx.params = (weight, bias) if isinstance(weight, nn.Parameter) and isinstance(bias, nn.Parameter) else None
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        C = x.shape[-1]
        P = x.numel() // C
        dev = x.device
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (ctx.has_res and ctx.needs_input_grad[3]) else None
        ws = torch.empty((lib().eeseg_bn_train_workspace_bytes(C),), dtype=torch.uint8, device=dev)
        from .parallel import direct_grad
        gw = gb = None
        if ctx.params is not None and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            gw, gb = direct_grad(ctx.params[0]), direct_grad(ctx.params[1])
        if gw is not None and gb is not None:
            # dgamma / dbeta added straight into the parameters' .grad: two AccumulateGrad launches less per BatchNorm
            with torch.cuda.device(dev):
                check(lib().eeseg_bn_train_bwd_acc(
                    dy.data_ptr(), x.data_ptr(), None if y is None else y.data_ptr(), P, C, weight.data_ptr(), mean.data_ptr(),
                    invstd.data_ptr(), 1 if ctx.relu else 0, dx.data_ptr(), None if dres is None else dres.data_ptr(),
                    gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                    "eeseg_bn_train_bwd_acc")
            for p in ctx.params:
                p._eeseg_flat.written(p)
            return dx, None, None, dres, None, None, None, None, None
        dgamma = torch.empty((C,), dtype=torch.float32, device=dev)
        dbeta = torch.empty((C,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().eeseg_bn_train_bwd(
                dy.data_ptr(), x.data_ptr(), None if y is None else y.data_ptr(), P, C, weight.data_ptr(), mean.data_ptr(),
                invstd.data_ptr(), 1 if ctx.relu else 0, dx.data_ptr(), None if dres is None else dres.data_ptr(),
                dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "eeseg_bn_train_bwd")
        return dx, dgamma.to(weight.dtype), dbeta.to(weight.dtype), dres, None, None, None, None, None


def bn_supported(bn, x):
    return (isinstance(bn, nn.BatchNorm2d) and bn.training and bn.affine and bn.track_running_stats
            and bn.momentum is not None and bn.weight.dtype == torch.float32 and x.is_cuda
            and x.dtype == torch.bfloat16 and x.shape[1] % 64 == 0)


def bn_act(x, bn, relu, residual=None):
    """x, residual: [N,C,h,w] bf16 channels_last. Returns act(bn(x) (+ residual)) in the same format:
    one autograd node on the eeseg kernels when supported, else the PyTorch modules."""
    if bn_supported(bn, x) and (residual is None or (residual.dtype == torch.bfloat16 and residual.shape == x.shape)):
        y = BnActFn.apply(x.permute(0, 2, 3, 1), bn.weight, bn.bias,
                          None if residual is None else residual.permute(0, 2, 3, 1),
                          bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu)
        with torch.no_grad():
            bn.num_batches_tracked += 1
        return y.permute(0, 3, 1, 2)
    y = bn(x)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y
