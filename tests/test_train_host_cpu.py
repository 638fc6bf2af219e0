"""Host-side logic of the training path and of bench.py that needs no GPU: which torchvision modules are routed
to the eeseg kernels (shape rules of the tensor-core tiles), the no-CPU-fallback behaviour, workload switching."""
import pytest
import torch
from torch import nn


def test_conv_routing_rules_match_the_kernel_constraints():
    from ee_semantic_segmentation_b200.head_train import _conv_ok, head_supported
    from torchvision.models.segmentation.deeplabv3 import DeepLabHead
    ok = nn.Conv2d(256, 256, 3, padding=2, dilation=2, bias=False)
    assert _conv_ok(ok)
    assert not _conv_ok(nn.Conv2d(256, 256, 3, padding=1, dilation=2, bias=False))      # not 'same'
    assert _conv_ok(nn.Conv2d(256, 256, 3, stride=2, padding=1, bias=False))            # stride 2 (layer2.0)
    assert not _conv_ok(nn.Conv2d(256, 256, 3, stride=4, padding=1, bias=False))
    assert _conv_ok(nn.Conv2d(256, 64, 1, bias=False))                                  # 64 output channels (layer1)
    assert not _conv_ok(nn.Conv2d(256, 96, 1, bias=False))                              # Cout % 64
    assert not _conv_ok(nn.Conv2d(3, 128, 7, padding=3, bias=False))                    # Cin % 64 (stem)
    assert not _conv_ok(nn.Conv2d(256, 128, 1, bias=True))                              # bias
    assert not _conv_ok(nn.Conv2d(256, 256, 3, padding=1, groups=2, bias=False))
    assert head_supported(DeepLabHead(2048, 21)) and head_supported(DeepLabHead(1024, 19))
    assert not head_supported(DeepLabHead(100, 21))                                     # Cin % 64
    assert not head_supported(nn.Sequential(nn.Conv2d(64, 21, 1)))


def test_resnet50_backbone_routing_counts():
    """All 52 Bottleneck convolutions of the DeepLab ResNet-50 backbone (stride 8) take the eeseg tiles; the stem's
    7x7 convolution (3 input channels) does not."""
    import torchvision
    from torchvision.models.resnet import Bottleneck
    from ee_semantic_segmentation_b200.head_train import _conv_ok
    bb = torchvision.models.resnet50(weights=None, replace_stride_with_dilation=[False, True, True])
    convs = [m for blk in bb.modules() if isinstance(blk, Bottleneck) for m in blk.modules() if isinstance(m, nn.Conv2d)]
    assert len(convs) == 52
    bad = [c for c in convs if not _conv_ok(c)]
    assert len(bad) == 0
    assert not _conv_ok(bb.conv1)


def test_training_kernels_refuse_cpu_tensors():
    from ee_semantic_segmentation_b200.bn_train import bn_act, bn_supported
    from ee_semantic_segmentation_b200 import ops
    bn = nn.BatchNorm2d(64).train()
    x = torch.randn(2, 64, 4, 4)
    assert not bn_supported(bn, x)                       # CPU tensor: the PyTorch module handles it
    y = bn_act(x, bn, True)
    assert torch.allclose(y, torch.relu(nn.BatchNorm2d(64).train()(x)), atol=1e-6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.upsample_bilinear_autograd(torch.randn(1, 2, 3, 3), (6, 6))


def test_bench_workloads():
    import bench
    try:
        bench.set_workload("cityscapes")
        assert bench.img_hw() == (1024, 2048) and bench.N_CLASSES == 19
        X, y = bench.synth_batch(0, 1, img=(32, 64), n_classes=19)
        assert X.shape == (1, 3, 32, 64) and y.shape == (1, 1, 32, 64) and int(y.max()) <= 19
        assert bench.workload_config(8)["global_batch"] == 16
    finally:
        bench.set_workload("voc513")
    assert bench.img_hw() == (513, 513) and bench.METRIC == "early_exit_images_per_sec_513"
    h = (513 - 1) // 8 + 1
    assert bench.head_flops(h, h, 4, [1024, 2048, 2048]) == 1327685632000      # 3 heads, N = 4
