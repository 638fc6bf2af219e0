"""Inference plan of one backbone section (`base_model[i]`, from_deepv3_new.py:146,151) on the eeseg
implicit-GEMM kernel: every torchvision ResNet `Bottleneck` becomes three (four with a projection
shortcut) `eeseg_conv_igemm_fwd` launches on bf16 NHWC activations with BatchNorm folded into the
epilogue scale/shift, ReLU fused, and the residual add fused into the last 1x1 — instead of
conv + BN + ReLU + add as separate library kernels. Stride-2 convolutions (layer2.0) use TMA element
strides. The 7x7/stride-2 stem (3 input channels) and the max-pool stay on PyTorch (bf16,
channels_last): their K is not a multiple of the 64-channel MMA K-block.
"""
import torch
from torch import nn
from torchvision.models.resnet import Bottleneck

from . import _lib, head_plan
from .head_plan import _fold_bn, _krsc, conv_igemm


class _Conv:
    def __init__(self, conv, bn):
        assert conv.groups == 1 and conv.bias is None
        self.w = _krsc(conv)
        s, b = _fold_bn(bn)
        self.s, self.b = s.contiguous(), b.contiguous()
        self.dil = conv.dilation[0]
        self.stride = conv.stride[0]
        self.cout = conv.out_channels
        assert conv.padding[0] == self.dil * (conv.kernel_size[0] // 2)

    def __call__(self, x, relu, residual=None):
        N, h, w, _ = x.shape
        ho, wo = (h - 1) // self.stride + 1, (w - 1) // self.stride + 1
        out = torch.empty((N, ho, wo, self.cout), dtype=torch.bfloat16, device=x.device)
        conv_igemm(x, self.w, self.s, self.b, self.dil, relu, out, _lib.BF16, self.cout,
                   stride=self.stride, residual=residual)
        return out


class BottleneckPlan:
    def __init__(self, blk):
        self.c1 = _Conv(blk.conv1, blk.bn1)
        self.c2 = _Conv(blk.conv2, blk.bn2)
        self.c3 = _Conv(blk.conv3, blk.bn3)
        self.down = _Conv(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None

    def run(self, x):
        t = self.c1(x, True)
        t = self.c2(t, True)
        idn = x if self.down is None else self.down(x, False)
        return self.c3(t, True, residual=idn)


def supported(section):
    for m in section:
        if isinstance(m, Bottleneck):
            for c in (m.conv1, m.conv2, m.conv3):
                if c.in_channels % 64 or c.out_channels % 16 or c.groups != 1 or c.stride[0] not in (1, 2):
                    return False
    return True


class SectionPlan:
    def __init__(self, section):
        self.ops = [BottleneckPlan(m) if isinstance(m, Bottleneck) else m for m in section]

    def run(self, x):
        """x: [N,C,h,w] tensor (any float dtype / memory format). Returns a bf16 [N,C',h',w'] tensor
        in channels_last memory format (an NHWC buffer viewed as NCHW)."""
        nhwc = None
        head_plan.PROFILE_TAG = "backbone"
        try:
            return self._run(x, nhwc)
        finally:
            head_plan.PROFILE_TAG = "head"

    def _run(self, x, nhwc):
        for op in self.ops:
            if isinstance(op, BottleneckPlan):
                if nhwc is None:
                    nhwc = x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
                nhwc = op.run(nhwc)
            else:
                if nhwc is not None:
                    x, nhwc = nhwc.permute(0, 3, 1, 2), None
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    x = op(x.contiguous(memory_format=torch.channels_last))
        return nhwc.permute(0, 3, 1, 2) if nhwc is not None else x
