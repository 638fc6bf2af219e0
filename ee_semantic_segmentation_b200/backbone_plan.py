"""Inference plan of one backbone section (`base_model[i]`, from_deepv3_new.py:146,151) on the eeseg
implicit-GEMM kernel: every torchvision ResNet `Bottleneck` becomes three (four with a projection
shortcut) `eeseg_conv_igemm_fwd` launches on bf16 NHWC activations with BatchNorm folded into the
epilogue scale/shift, ReLU fused, and the residual add fused into the last 1x1 — instead of
conv + BN + ReLU + add as separate library kernels (the BatchNorm scale goes into the bf16 weights, the shift stays in the
epilogue: head_plan.FOLD_SCALE). Stride-2 convolutions (layer2.0) use TMA element
strides. The 7x7/stride-2 stem (3 input channels) is a space-to-depth 4x1 implicit GEMM followed by the NHWC
max-pool kernel (StemPlan). A section with any other unit is not `supported` (the model then warns once, or raises
under `strict_kernels`).
"""
import torch
from torch import nn
from torchvision.models.resnet import Bottleneck

from . import _lib, head_plan
from .head_plan import _fold_bn, _krsc, conv_igemm


FOLD_SCALE = head_plan.FOLD_SCALE


class _Conv:
    def __init__(self, conv, bn, fold_scale=False):
        assert conv.groups == 1 and conv.bias is None
        s, b = _fold_bn(bn)
        fold_scale = fold_scale or FOLD_SCALE >= 2
        if fold_scale:
            # convs that take a residual: the kernel adds the residual inside the accumulator (identity
            # K blocks on the tensor core), so the BN scale goes into the weights and the epilogue scale is 1
            self.w = (conv.weight.detach().float() * s.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
            s = torch.ones_like(s) if FOLD_SCALE == 0 else None
        else:
            self.w = _krsc(conv)
        self.s, self.b = (None if s is None else s.contiguous()), b.contiguous()
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
        self.c3 = _Conv(blk.conv3, blk.bn3, fold_scale=True)
        self.down = _Conv(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None

    def run(self, x):
        t = self.c1(x, True)
        t = self.c2(t, True)
        idn = x if self.down is None else self.down(x, False)
        return self.c3(t, True, residual=idn)


class StemPlan:
    """conv1 7x7/s2/p3 (3->64) + bn1 + relu + maxpool 3x3/s2/p1 of torchvision's ResNet on the eeseg
    kernels: the fp32 NCHW image is space-to-depth'ed (2x2 -> 12 channels, zero-padded to the 64-channel
    MMA K block together with the 4 horizontal taps: channel u*12 + (a*2+b)*3 + c) so the stride-2 7x7 conv
    is a stride-1 4x1 implicit GEMM (4 K blocks) with the first tap at row offset -2 (input row
    2y-3+r = 2(y+t-2)+a with r = 2t+a-1); BatchNorm and ReLU are the conv epilogue; the
    max-pool is one streaming kernel. Replaces a cuDNN conv + 3 elementwise kernels on a 4x larger
    intermediate."""

    def __init__(self, conv, bn):
        dev = conv.weight.device
        W = conv.weight.detach().float()                         # [64, 3, 7, 7]
        cout = W.shape[0]
        w2 = torch.zeros((cout, 4, 1, 64), dtype=torch.float32, device=dev)
        for t in range(4):
            for a in range(2):
                r = 2 * t + a - 1
                if not 0 <= r <= 6:
                    continue
                for u in range(4):
                    for b in range(2):
                        q = 2 * u + b - 1
                        if not 0 <= q <= 6:
                            continue
                        ch = u * 12 + (a * 2 + b) * 3
                        w2[:, t, 0, ch:ch + 3] = W[:, :, r, q]
        self.w = w2.to(torch.bfloat16).contiguous()
        s, b = _fold_bn(bn)
        self.s, self.b = s.contiguous(), b.contiguous()
        self.cout = cout

    @staticmethod
    def matches(mods):
        if len(mods) < 4:
            return False
        c, b, r, m = mods[:4]
        return (isinstance(c, nn.Conv2d) and c.in_channels == 3 and c.kernel_size == (7, 7) and c.stride == (2, 2)
                and c.padding == (3, 3) and c.bias is None and c.out_channels % 16 == 0
                and isinstance(b, nn.BatchNorm2d) and isinstance(r, nn.ReLU) and isinstance(m, nn.MaxPool2d)
                and m.kernel_size in (3, (3, 3)) and m.stride in (2, (2, 2)) and m.padding in (1, (1, 1))
                and m.dilation in (1, (1, 1)) and not m.ceil_mode)

    def run(self, x, norm=None):
        """x NCHW [N,3,H,W] fp32, bf16 or uint8 -> bf16 NHWC [N, H/4-ish, W/4-ish, 64]. The kernel's first step rounds
        the image to bf16, so a bf16 upload gives bit-identical results at half the bytes; uint8 pixels are normalised
        inside the kernel, (u/255 - mean[c]) / std[c] with norm = (mean, std) (3 floats each; None: mean 0, std 1) —
        what the reference's loader does on the CPU (get_seg_datasets.py:62-70)."""
        import ctypes
        N, _, H, W = x.shape
        dev = x.device
        if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            x = x.float()
        x = x.contiguous()
        kind = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.uint8: _lib.U8}[x.dtype]
        mean = std = None
        if norm is not None and x.dtype == torch.uint8:
            mean = (ctypes.c_float * 3)(*[float(v) for v in norm[0]])
            std = (ctypes.c_float * 3)(*[float(v) for v in norm[1]])
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            s2d = torch.empty((N, H2, W2, 64), dtype=torch.bfloat16, device=dev)
            _lib.check(_lib.lib().eeseg_stem_space_to_depth_any(x.data_ptr(), kind, mean, std, N, H, W, s2d.data_ptr(),
                                                                stream), "eeseg_stem_space_to_depth_any")
            y = torch.empty((N, H2, W2, self.cout), dtype=torch.bfloat16, device=dev)
            conv_igemm(s2d, self.w, self.s, self.b, 1, True, y, _lib.BF16, self.cout, pad=2)
            ho, wo = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
            out = torch.empty((N, ho, wo, self.cout), dtype=torch.bfloat16, device=dev)
            _lib.check(_lib.lib().eeseg_maxpool3x3s2_nhwc(y.data_ptr(), N, H2, W2, self.cout, out.data_ptr(), stream),
                       "eeseg_maxpool3x3s2_nhwc")
        return out


def supported(section):
    """True when EVERY unit of the section runs on the eeseg kernels: an optional ResNet stem (conv 7x7/s2 + BN + ReLU +
    max-pool, StemPlan.matches) followed by Bottlenecks whose convolutions the implicit-GEMM kernel takes."""
    mods = list(section)
    if StemPlan.matches(mods):
        mods = mods[4:]
    for m in mods:
        if not isinstance(m, Bottleneck):
            return False
        for c in (m.conv1, m.conv2, m.conv3):
            if c.in_channels % 64 or c.out_channels % 16 or c.groups != 1 or c.stride[0] not in (1, 2):
                return False
    return True


class SectionPlan:
    def __init__(self, section):
        mods = list(section)
        self.ops = []
        if StemPlan.matches(mods):
            self.ops.append(StemPlan(mods[0], mods[1]))
            mods = mods[4:]
        self.ops += [BottleneckPlan(m) if isinstance(m, Bottleneck) else m for m in mods]

    def run(self, x, norm=None):
        """x: [N,C,h,w] tensor (any float dtype / memory format; uint8 images for a section that starts with the stem).
        Returns a bf16 [N,C',h',w'] tensor in channels_last memory format (an NHWC buffer viewed as NCHW)."""
        nhwc = None
        head_plan.PROFILE_TAG = "backbone"
        try:
            return self._run(x, nhwc, norm)
        finally:
            head_plan.PROFILE_TAG = "head"

    def _run(self, x, nhwc, norm=None):
        for op in self.ops:
            if isinstance(op, StemPlan):
                nhwc = op.run(x, norm)
            elif isinstance(op, BottleneckPlan):
                if nhwc is None:
                    nhwc = x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
                nhwc = op.run(nhwc)
            else:       # supported() admits stems and Bottlenecks only
                raise RuntimeError(f'SectionPlan: {type(op).__name__} has no eeseg kernel plan')
        return nhwc.permute(0, 3, 1, 2) if nhwc is not None else x
