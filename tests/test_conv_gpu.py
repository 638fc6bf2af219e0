"""Parity of the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu) and of the whole exit head
against PyTorch fp32 on the same bf16-rounded operands. Tolerance: bf16 storage of the result,
1e-2 relative (north_star) on the normalised error."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def run_conv(N, h, w, Cin, Cout, R, dil, relu, out_f32=False, per_image_shift=False, seed=0):
    from ee_semantic_segmentation_b200 import _lib
    from ee_semantic_segmentation_b200.head_plan import conv_igemm
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, h, w, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cout, R, R, Cin, generator=g) / np.sqrt(R * R * Cin)).to(torch.bfloat16)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(N if per_image_shift else 1, Cout, generator=g)
    xd, wd = x.to(dev()), wt.to(dev())
    out = torch.full((N, h, w, Cout + 16), -7.0, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev())
    conv_igemm(xd, wd, scale.to(dev()), shift.to(dev()).contiguous(), dil, relu, out[..., 8:],
               _lib.F32 if out_f32 else _lib.BF16, Cout + 16, shift_sn=Cout if per_image_shift else 0)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), padding=dil * (R // 2), dilation=dil)
    ref = ref * scale.view(1, -1, 1, 1) + shift.view(-1, Cout, 1, 1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    got = out[..., 8:8 + Cout].float().cpu()
    # untouched guard channels on both sides (ldo > Cout, channel offset)
    assert torch.all(out[..., :8].float() == -7.0) and torch.all(out[..., 8 + Cout:].float() == -7.0)
    err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
    return err


@pytest.mark.parametrize("cfg", [
    dict(N=1, h=16, w=8, Cin=64, Cout=32, R=1, dil=1, relu=False, out_f32=True),    # one exact tile, one k-block
    dict(N=1, h=16, w=8, Cin=256, Cout=256, R=1, dil=1, relu=True),
    dict(N=2, h=65, w=65, Cin=128, Cout=256, R=3, dil=1, relu=True),                # ragged 11x11 tiles
    dict(N=1, h=65, w=65, Cin=128, Cout=256, R=3, dil=12, relu=True),
    dict(N=1, h=65, w=65, Cin=64, Cout=64, R=3, dil=36, relu=False),                # most taps all-padding
    dict(N=2, h=33, w=47, Cin=192, Cout=32, R=1, dil=1, relu=False, out_f32=True, per_image_shift=True),
    dict(N=1, h=24, w=40, Cin=1024, Cout=512, R=1, dil=1, relu=True),               # two N tiles, deep K
])
def test_conv_igemm_vs_torch(cfg):
    err = run_conv(**cfg)
    assert err < 1e-2, err


def test_conv_igemm_full_head_shapes():
    """The ASPP shapes at 513x513 (65x65 maps): Cin=2048, d=24, and the K=1280-like projection."""
    assert run_conv(N=1, h=65, w=65, Cin=2048, Cout=256, R=3, dil=24, relu=True, seed=1) < 1e-2
    assert run_conv(N=2, h=65, w=65, Cin=1024, Cout=256, R=1, dil=1, relu=True, per_image_shift=True, seed=2) < 1e-2


def test_avgpool():
    from ee_semantic_segmentation_b200 import _lib
    x = torch.randn(3, 65 * 65, 192).to(torch.bfloat16).to(dev())
    out = torch.empty(3, 192, device=dev())
    _lib.check(_lib.lib().eeseg_global_avgpool_nhwc(x.data_ptr(), 3, 65 * 65, 192, out.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "avgpool")
    np.testing.assert_allclose(out.cpu().numpy(), x.float().mean(1).cpu().numpy(), atol=1e-5)


@pytest.mark.parametrize("cin", [1024, 2048])
def test_head_plan_vs_torchvision(cin):
    """Whole DeepLabHead (eval, BN folded) on the eeseg kernels vs the torchvision module in fp32."""
    from torchvision.models.segmentation.deeplabv3 import DeepLabHead
    from ee_semantic_segmentation_b200.head_plan import HeadPlan
    torch.manual_seed(cin)
    head = DeepLabHead(cin, 21).eval()
    for m in head.modules():            # non-trivial BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    x = torch.randn(2, cin, 33, 33)
    with torch.no_grad():
        ref = head(x)
    head_d = head.to(dev())
    low = HeadPlan(head_d).run(x.to(dev()))
    got = low[..., :21].permute(0, 3, 1, 2).float().cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, err      # five bf16 layers deep; logits within 1e-2 relative of fp32 typical
