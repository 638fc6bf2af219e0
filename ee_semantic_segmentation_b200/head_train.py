"""Training-mode forward/backward of a DeepLabHead / my_branch exit head with its convolutions on the
eeseg tcgen05 kernels (forward `eeseg_conv_igemm_fwd`, input gradient `eeseg_conv_igemm_dgrad`, weight
gradient `eeseg_conv_igemm_wgrad`) — what autograd through cuDNN computes for
`branches[i](X)` / `classifier(X)` (from_deepv3_new.py:147,151) inside `train_epoch`
(train_funcs.py:22-27).

Activations of the head are bf16 NHWC; BatchNorm (batch statistics) + ReLU run as fused eeseg nodes
(bn_train.BnActFn) on the modules' own parameters and running-statistic buffers; Dropout, the pooled ASPP branch
and the final Cout = num_classes 1x1 convolution stay on the PyTorch modules (0.03 % of the head's FLOPs), so
state-dict layout, running statistics and RNG consumption are those of the reference modules. Weight gradients are returned in fp32 in the
parameter's own [Cout,Cin,R,S] layout.
"""
import torch
from torch import nn
from torchvision.models.segmentation.deeplabv3 import ASPP

from . import _lib
from ._lib import check, lib

_UNIT = {}


def _unit_scale_shift(dev, n):
    key = (dev, n)
    if key not in _UNIT:
        _UNIT[key] = (torch.ones(n, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.float32, device=dev))
    return _UNIT[key]


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
        wt = torch.empty((Cout, R, S, Cin), dtype=torch.bfloat16, device=x.device)    # [Cout,R,S,Cin]
        wt.copy_(weight.detach().permute(0, 2, 3, 1))                                  # layout + precision in one pass
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = torch.empty((N, ho, wo, Cout), dtype=torch.bfloat16, device=x.device)
        one, zero = _unit_scale_shift(x.device, Cout)
        conv_igemm(x, wt, one, zero, dilation, False, out, _lib.BF16, Cout, stride=stride)
        ctx.save_for_backward(x, wt)
        ctx.dilation = dilation
        ctx.stride = stride
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
                ws = torch.empty((lib().eeseg_conv_igemm_dgrad_workspace_bytes(Cin, Cout, R, S),), dtype=torch.uint8,
                                 device=x.device)
                dx = torch.empty_like(x)
                torch.ops.eeseg.conv_igemm_dgrad(dy, wt, ctx.dilation, dx, ws)
            if ctx.needs_input_grad[1]:
                dwk = torch.empty((Cout, R, S, Cin), dtype=torch.float32, device=x.device)
                wws = torch.empty((lib().eeseg_conv_igemm_wgrad_workspace_bytes(N, h, w, Cin, Cout, R, S),), dtype=torch.uint8,
                                  device=x.device)
                torch.ops.eeseg.conv_igemm_wgrad(x, dy, ctx.dilation, dwk, wws)
                dw = dwk.permute(0, 3, 1, 2)                                             # the parameter's layout
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


def head_forward_train(head, x):
    """head(x) with autograd, x [N,Cin,h,w] (fp32 or bf16, any memory format) -> logits [N,C,h,w] fp32."""
    aspp = head[0]
    xh = _nhwc(x)
    outs = []
    from .bn_train import bn_act
    for m in list(aspp.convs)[:-1]:                      # 1x1 and the atrous 3x3 branches: conv, BN + ReLU
        outs.append(bn_act(_conv(xh, m[0]), m[1], True))
    # ASPPPooling on its PyTorch modules (global average pool, 1x1 conv, BN, ReLU). Its "bilinear" up-sampling
    # of a 1x1 map (deeplabv3.py:83) is a broadcast: expand() instead of F.interpolate, whose backward is a plain
    # sum instead of ATen's atomic scatter onto one pixel (0.42 ms per head at 65x65)
    pooled = x
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=x.dtype != torch.float32):
        for mod in aspp.convs[-1]:
            pooled = mod(pooled)
    pooled = pooled.to(torch.bfloat16).expand(-1, -1, x.shape[-2], x.shape[-1]).contiguous(memory_format=torch.channels_last)
    outs.append(pooled)
    cat = torch.cat(outs, dim=1)
    y = bn_act(_conv(_nhwc(cat), aspp.project[0]), aspp.project[1], True)   # projection, BN + ReLU
    for m in list(aspp.project)[3:]:                     # Dropout(0.5)
        y = m(y)
    y = bn_act(_conv(_nhwc(y), head[1]), head[2], True)  # 3x3, BN + ReLU
    return head[4](y.float())                            # final 1x1 (+bias) to num_classes
