"""Training-mode forward/backward of a backbone section (`base_model[i]`, from_deepv3_new.py:146,151 inside
`train_epoch`, train_funcs.py:22-27): every convolution of the torchvision ResNet `Bottleneck`s (Cin, Cout
multiples of 64; stride 1, or 2 in layer2.0) runs forward,
input-gradient and weight-gradient on the eeseg tcgen05 kernels (head_train.ConvIgemmFn); BatchNorm (batch
statistics, running-stat updates) + residual add + ReLU run as one fused eeseg node per BatchNorm
(bn_train.BnActFn); activations are bf16 channels_last end to end. The 7x7/stride-2 stem runs on the kernels too
(stem_forward_train): space-to-depth of the image, the 4x1 implicit GEMM of the inference stem with the weights
gathered from the 7x7 parameter by a differentiable index map, its weight gradient on the wgrad kernel, BatchNorm +
ReLU as for every other layer, and the max-pool forward (recording the winning tap) / gather-form backward kernels.
Parameters, buffers and state-dict layout are the reference's.
Master weights and their gradients stay fp32 (mixed precision); the reference trains in fp32 with TF32
allowed (train_funcs.py:117-118) — parity is within the bf16 bound of north_star and is tested as such.
"""
import torch
from torch import nn
from torchvision.models.resnet import Bottleneck

from . import _lib
from ._lib import check, lib
from .backbone_plan import StemPlan
from .bn_train import bn_act
from .head_train import ConvIgemmFn, _conv_ok, _unit_scale_shift

_STEM_MAP = {}


def _stem_index_map(dev):
    """Column of the [Cout, 3*7*7] weight matrix behind every slot of the space-to-depth'ed [Cout, 4, 1, 64] kernel
    (StemPlan: input row 2(y+t-2)+a = 2y-3+r with r = 2t+a-1, columns likewise with u, b; slot channel
    u*12 + (a*2+b)*3 + c), and a mask for the slots that have no tap (r or q outside 0..6, channels 48..63)."""
    if dev not in _STEM_MAP:
        idx = torch.zeros((4, 64), dtype=torch.int64)
        mask = torch.zeros((4, 64), dtype=torch.float32)
        for t in range(4):
            for a in range(2):
                r = 2 * t + a - 1
                for u in range(4):
                    for b in range(2):
                        q = 2 * u + b - 1
                        if 0 <= r <= 6 and 0 <= q <= 6:
                            for c in range(3):
                                ch = u * 12 + (a * 2 + b) * 3 + c
                                idx[t, ch] = (c * 7 + r) * 7 + q
                                mask[t, ch] = 1.0
        _STEM_MAP[dev] = (idx.reshape(-1).to(dev), mask.reshape(-1).to(dev))
    return _STEM_MAP[dev]


class StemConvFn(torch.autograd.Function):
    """conv1 (7x7 / stride 2 / pad 3, 3 -> Cout) as the 4x1 stride-1 implicit GEMM over the space-to-depth'ed image.
    s2d [N,H2,W2,64] bf16 (no gradient: it is the input image), w2 [Cout,4,1,64] fp32 -> [N,H2,W2,Cout] bf16;
    backward: dW2 on the weight-gradient kernel (pixels as the contraction dimension)."""

    @staticmethod
    def forward(ctx, s2d, w2):
        from .head_plan import conv_igemm
        N, H2, W2, _ = s2d.shape
        Cout = w2.shape[0]
        wt = w2.detach().to(torch.bfloat16).contiguous()
        out = torch.empty((N, H2, W2, Cout), dtype=torch.bfloat16, device=s2d.device)
        one, zero = _unit_scale_shift(s2d.device, Cout)
        conv_igemm(s2d, wt, None, zero, 1, False, out, _lib.BF16, Cout, pad=2)
        ctx.save_for_backward(s2d)
        ctx.cout = Cout
        return out

    @staticmethod
    def backward(ctx, dy):
        (s2d,) = ctx.saved_tensors
        N, H2, W2, Cin = s2d.shape
        Cout = ctx.cout
        dy = dy.to(torch.bfloat16).contiguous()
        dev = s2d.device
        dw = torch.empty((Cout, 4, 1, Cin), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = torch.empty((lib().eeseg_conv_igemm_wgrad_workspace_bytes(N, H2, W2, Cin, Cout, 4, 1),), dtype=torch.uint8,
                             device=dev)
            check(lib().eeseg_conv_igemm_wgrad(s2d.data_ptr(), dy.data_ptr(), Cout, Cout, 0, N, H2, W2, Cin, Cout, 4, 1, 1,
                                               dw.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                  "eeseg_conv_igemm_wgrad")
        return None, dw


class MaxPool3x3s2Fn(torch.autograd.Function):
    """nn.MaxPool2d(3, 2, 1) on NHWC bf16: forward records the winning tap, backward gathers (no atomics)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, h, w, C = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        out = torch.empty((N, ho, wo, C), dtype=torch.bfloat16, device=x.device)
        idx = torch.empty((N, ho, wo, C), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().eeseg_maxpool3x3s2_nhwc_train(x.data_ptr(), N, h, w, C, out.data_ptr(), idx.data_ptr(),
                                                      torch.cuda.current_stream(x.device).cuda_stream),
                  "eeseg_maxpool3x3s2_nhwc_train")
        ctx.save_for_backward(idx)
        ctx.shape = (N, h, w, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        N, h, w, C = ctx.shape
        dout = dout.to(torch.bfloat16).contiguous()
        dx = torch.empty((N, h, w, C), dtype=torch.bfloat16, device=dout.device)
        with torch.cuda.device(dout.device):
            check(lib().eeseg_maxpool3x3s2_nhwc_bwd(dout.data_ptr(), idx.data_ptr(), N, h, w, C, dx.data_ptr(),
                                                    torch.cuda.current_stream(dout.device).cuda_stream),
                  "eeseg_maxpool3x3s2_nhwc_bwd")
        return dx


def stem_supported(mods, x):
    return (StemPlan.matches(mods) and mods[0].out_channels % 64 == 0 and x.is_cuda and not x.requires_grad
            and mods[0].weight.dtype == torch.float32)


def stem_forward_train(conv, bn, x):
    """maxpool(relu(bn1(conv1(x)))) of the torchvision ResNet stem with autograd, x [N,3,H,W] fp32/bf16 NCHW ->
    bf16 channels_last [N,Cout,H/4,W/4]."""
    N, _, H, W = x.shape
    dev = x.device
    x = x.detach().float().contiguous()
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    s2d = torch.empty((N, H2, W2, 64), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        check(lib().eeseg_stem_space_to_depth(x.data_ptr(), N, H, W, s2d.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
              "eeseg_stem_space_to_depth")
    idx, mask = _stem_index_map(dev)
    cout = conv.out_channels
    w2 = (conv.weight.reshape(cout, -1).index_select(1, idx) * mask).view(cout, 4, 1, 64)
    y = StemConvFn.apply(s2d, w2).permute(0, 3, 1, 2)
    y = bn_act(y, bn, True)
    return MaxPool3x3s2Fn.apply(y.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def _conv(x, conv):
    """x: [N,C,h,w] bf16 channels_last -> conv(x) in the same format."""
    if _conv_ok(conv):
        return ConvIgemmFn.apply(x.permute(0, 2, 3, 1), conv.weight, conv.dilation[0], conv.stride[0]).permute(0, 3, 1, 2)
    return conv(x)       # under autocast: cuDNN bf16


def bottleneck_forward_train(blk, x):
    """torchvision.models.resnet.Bottleneck.forward with the supported convolutions on ConvIgemmFn."""
    identity = x
    out = bn_act(_conv(x, blk.conv1), blk.bn1, True)
    out = bn_act(_conv(out, blk.conv2), blk.bn2, True)
    if blk.downsample is not None:
        identity = bn_act(_conv(x, blk.downsample[0]), blk.downsample[1], False)
    # bn3 + residual add + ReLU: one pass
    return bn_act(_conv(out, blk.conv3), blk.bn3, True, residual=identity)


def section_supported(section, x):
    """True when every unit of the section trains on the eeseg kernels (stem + Bottlenecks with supported convolutions);
    anything else runs on the PyTorch modules inside section_forward_train."""
    mods = list(section)
    if stem_supported(mods, x):
        mods = mods[4:]
    for unit in mods:
        if not (isinstance(unit, Bottleneck) and len(unit.downsample or [0, 0]) == 2):
            return False
        convs = [unit.conv1, unit.conv2, unit.conv3] + ([unit.downsample[0]] if unit.downsample is not None else [])
        if not all(_conv_ok(c) for c in convs):
            return False
    return True


def section_forward_train(section, x):
    """section(x) with autograd. x: fp32/bf16 NCHW (any memory format) -> bf16 channels_last."""
    mods = list(section)
    if stem_supported(mods, x):
        x = stem_forward_train(mods[0], mods[1], x)
        mods = mods[4:]
    x = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        for unit in mods:
            if isinstance(unit, Bottleneck) and len(unit.downsample or [0, 0]) == 2:
                x = bottleneck_forward_train(unit, x)
            else:
                x = unit(x)
    return x
