"""Training-mode forward/backward of a backbone section (`base_model[i]`, from_deepv3_new.py:146,151 inside
`train_epoch`, train_funcs.py:22-27): every convolution of the torchvision ResNet `Bottleneck`s (Cin, Cout
multiples of 64; stride 1, or 2 in layer2.0) runs forward,
input-gradient and weight-gradient on the eeseg tcgen05 kernels (head_train.ConvIgemmFn); BatchNorm (batch
statistics, running-stat updates) + residual add + ReLU run as one fused eeseg node per BatchNorm
(bn_train.BnActFn); activations are bf16 channels_last end to end. The 7x7 stem (conv, BN, ReLU, max-pool: 3 input
channels) stays on the PyTorch modules under bf16
autocast: parameters, buffers and state-dict layout are the reference's.
Master weights and their gradients stay fp32 (mixed precision); the reference trains in fp32 with TF32
allowed (train_funcs.py:117-118) — parity is within the bf16 bound of north_star and is tested as such.
"""
import torch
from torch import nn
from torchvision.models.resnet import Bottleneck

from .bn_train import bn_act
from .head_train import ConvIgemmFn, _conv_ok


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


def section_forward_train(section, x):
    """section(x) with autograd. x: fp32/bf16 NCHW (any memory format) -> bf16 channels_last."""
    x = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        for unit in section:
            if isinstance(unit, Bottleneck) and len(unit.downsample or [0, 0]) == 2:
                x = bottleneck_forward_train(unit, x)
            else:
                x = unit(x)
    return x
