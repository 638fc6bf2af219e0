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
    # 88 x 57 = 5016 spatial tiles = 34 per CTA: past the 32-entry tile table of the prologue (items 32+ are decoded inline)
    dict(N=1, h=700, w=900, Cin=64, Cout=64, R=3, dil=1, relu=True),
])
def test_conv_igemm_vs_torch(cfg):
    err = run_conv(**cfg)
    assert err < 1e-2, err


@pytest.mark.parametrize("cfg", [
    dict(N=1, h=13, w=15, Cin=512, Cout=512, R=3, dil=4),      # 72 K blocks: CTA pairs; 4 spatial tiles, 2-4 channel tiles
    dict(N=3, h=13, w=15, Cin=512, Cout=512, R=3, dil=4),
    dict(N=3, h=9, w=9, Cin=2048, Cout=512, R=1, dil=1),       # 1x1, 32 K blocks: pairs; 3 spatial tiles (odd: last repeated)
    dict(N=4, h=65, w=65, Cin=2048, Cout=512, R=1, dil=1),     # layer4 conv1 of the headline shape
    dict(N=2, h=33, w=33, Cin=512, Cout=512, R=3, dil=4),
])
def test_cta_pair_launches_are_deterministic_and_correct(cfg):
    """Launches the launcher runs as CTA pairs (cta_group::2; >= 64 K blocks per tile, or a 1x1 with >= 32) against the fp32
    convolution, and twenty repeats against the first result bit for bit (two SMs share the barriers of one tile)."""
    from ee_semantic_segmentation_b200 import _lib
    from ee_semantic_segmentation_b200.head_plan import conv_igemm
    N, h, w, Cin, Cout, R, dil = (cfg[k] for k in ("N", "h", "w", "Cin", "Cout", "R", "dil"))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, h, w, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cout, R, R, Cin, generator=g) / np.sqrt(R * R * Cin)).to(torch.bfloat16)
    scale, shift = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    xd, wd, sd, bd = x.to(dev()), wt.to(dev()), scale.to(dev()), shift.to(dev())
    outs = []
    for _ in range(21):
        out = torch.empty(N, h, w, Cout, dtype=torch.bfloat16, device=dev())
        conv_igemm(xd, wd, sd, bd, dil, True, out, _lib.BF16, Cout)
        outs.append(out)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), padding=dil * (R // 2), dilation=dil)
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu().permute(0, 2, 3, 1)
    err = (outs[0].float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err


def test_conv_igemm_full_head_shapes():
    """The ASPP shapes at 513x513 (65x65 maps): Cin=2048, d=24, and the K=1280-like projection."""
    assert run_conv(N=1, h=65, w=65, Cin=2048, Cout=256, R=3, dil=24, relu=True, seed=1) < 1e-2
    assert run_conv(N=2, h=65, w=65, Cin=1024, Cout=256, R=1, dil=1, relu=True, per_image_shift=True, seed=2) < 1e-2


def test_avgpool():
    from ee_semantic_segmentation_b200.head_plan import global_avgpool_nhwc
    x = torch.randn(3, 65, 65, 192).to(torch.bfloat16).to(dev())
    out = global_avgpool_nhwc(x)
    np.testing.assert_allclose(out.cpu().numpy(), x.float().mean((1, 2)).cpu().numpy(), atol=1e-5)
    x = torch.randn(1, 3, 5, 64).to(torch.bfloat16).to(dev())          # fewer pixels than splits
    np.testing.assert_allclose(global_avgpool_nhwc(x).cpu().numpy(), x.float().mean((1, 2)).cpu().numpy(), atol=1e-5)


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


@pytest.mark.parametrize("cfg", [
    dict(N=2, hin=129, win=129, Cin=128, Cout=128, R=3, stride=2),     # layer2.0.conv2
    dict(N=1, hin=129, win=129, Cin=256, Cout=512, R=1, stride=2),     # layer2.0.downsample
    dict(N=1, hin=34, win=50, Cin=64, Cout=64, R=3, stride=2),         # even sizes
    dict(N=2, hin=65, win=65, Cin=256, Cout=1024, R=1, stride=1),      # conv3 + residual
])
def test_conv_stride_and_residual(cfg):
    from ee_semantic_segmentation_b200 import _lib
    from ee_semantic_segmentation_b200.head_plan import conv_igemm
    N, hin, win, Cin, Cout, R, stride = (cfg[k] for k in ("N", "hin", "win", "Cin", "Cout", "R", "stride"))
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, hin, win, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cout, R, R, Cin, generator=g) / np.sqrt(R * R * Cin)).to(torch.bfloat16)
    scale, shift = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    ho, wo = (hin - 1) // stride + 1, (win - 1) // stride + 1
    res = torch.randn(N, ho, wo, Cout, generator=g).to(torch.bfloat16)
    out = torch.empty(N, ho, wo, Cout, dtype=torch.bfloat16, device=dev())
    # the residual is added inside the accumulator, before the epilogue scale: scale must be 1
    scale = torch.ones(Cout)
    conv_igemm(x.to(dev()), wt.to(dev()), scale.to(dev()), shift.to(dev()), 1, True, out, _lib.BF16, Cout,
               stride=stride, residual=res.to(dev()))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), padding=R // 2, stride=stride)
    ref = (ref + shift.view(1, -1, 1, 1) + res.float().permute(0, 3, 1, 2)).relu()
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err


def test_section_plan_vs_torchvision():
    """ResNet-50 sections (stem + bottlenecks incl. the stride-2 and dilated blocks) on the eeseg conv
    kernel vs the PyTorch modules in fp32."""
    import torchvision
    from ee_semantic_segmentation_b200.backbone_plan import SectionPlan
    from oracle.model_port import backbone_units
    torch.manual_seed(1)
    base = torchvision.models.segmentation.deeplabv3_resnet50(weights=None, weights_backbone=None, num_classes=21)
    for m in base.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    units = backbone_units(base.backbone)
    sec_a = torch.nn.Sequential(*units[:9]).eval()        # stem + layer1 + layer2.0/2.1 (stride 2 inside)
    sec_b = torch.nn.Sequential(*units[9:]).eval()        # rest of layer2, layer3 (d=2), layer4 (d=4)
    x = torch.randn(2, 3, 97, 113)
    with torch.no_grad():
        ra = sec_a(x)
        rb = sec_b(ra)
    ga = SectionPlan(sec_a.to(dev())).run(x.to(dev()))
    assert ga.shape == ra.shape and ga.dtype == torch.bfloat16
    assert (ga.float().cpu() - ra).abs().max().item() < 2e-2 * ra.abs().max().item()
    gb = SectionPlan(sec_b.to(dev())).run(ga)
    assert gb.shape == rb.shape
    from conftest import assert_bf16_model_close
    assert_bf16_model_close(gb.float().cpu(), rb)      # a whole section (up to 16 Bottlenecks) in bf16: DESIGN §2


def test_stem_plan_vs_torch():
    """conv1 7x7/s2 + bn + relu + maxpool via space-to-depth + 4x4 implicit GEMM + NHWC max-pool."""
    import torchvision
    from ee_semantic_segmentation_b200.backbone_plan import StemPlan
    torch.manual_seed(3)
    r = torchvision.models.resnet50(weights=None)
    r.bn1.running_mean.normal_(0, 0.1); r.bn1.running_var.uniform_(0.5, 1.5)
    r.bn1.weight.data.uniform_(0.5, 1.5); r.bn1.bias.data.normal_(0, 0.1)
    stem = torch.nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool).eval()
    assert StemPlan.matches(list(stem))
    for shape in [(2, 3, 513, 513), (1, 3, 97, 130)]:
        x = torch.randn(*shape)
        with torch.no_grad():
            ref = stem(x)
        stem_d = stem.to(dev())
        got = StemPlan(stem_d[0], stem_d[1]).run(x.to(dev())).permute(0, 3, 1, 2).float().cpu()
        stem.cpu()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-2, err


@pytest.mark.parametrize("N,pairs", [(2, False), (2, True), (4, True), (3, False)])
def test_grouped_conv_matches_separate_launches(N, pairs):
    """The grouped ASPP launch (one persistent kernel over a cost-sorted work list) writes exactly
    what the four separate launches write — as single CTAs and as CTA pairs (cta_group::2: the same tile position of two
    images per cluster, pair work list)."""
    from ee_semantic_segmentation_b200 import _lib
    from ee_semantic_segmentation_b200.head_plan import conv_igemm, conv_igemm_grouped, group_schedule
    g = torch.Generator().manual_seed(8)
    h, w, cin, mid = 65, 65, 256, 256
    x = torch.randn(N, h, w, cin, generator=g).to(torch.bfloat16).to(dev())
    ks, dl = [1, 3, 3, 3], [1, 12, 24, 36]
    wts = [(torch.randn(mid, k, k, cin, generator=g) / np.sqrt(k * k * cin)).to(torch.bfloat16).to(dev()) for k in ks]
    scs = [(torch.rand(mid, generator=g) + 0.5).to(dev()) for _ in ks]
    shs = [torch.randn(mid, generator=g).to(dev()) for _ in ks]
    ref = torch.empty(N, h, w, 4 * mid, dtype=torch.bfloat16, device=dev())
    for k in range(4):
        conv_igemm(x, wts[k], scs[k], shs[k], dl[k], True, ref[..., k * mid:], _lib.BF16, 4 * mid)
    out = torch.zeros_like(ref)
    n_cl = _lib.lib().eeseg_conv_pair_clusters() if pairs else None
    assert not pairs or n_cl >= 32, n_cl
    sched = group_schedule(N, h, w, cin, mid, ks, dl, pairs=pairs, n_clusters=n_cl).to(dev())
    assert sorted(sched.tolist()) == sorted((gi << 24) | t for gi in range(4) for t in range(N * 36))
    conv_igemm_grouped(x, wts, scs, shs, ks, dl, [k * mid for k in range(4)], True, out, 4 * mid, 4 * mid, sched,
                       cta_pairs=pairs)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


# ------------------------------------------------------------------------------------ training: gradients
def run_conv_grads(N, h, w, Cin, Cout, R, dil, seed=0, co_off=0, extra=0):
    """eeseg_conv_igemm_wgrad / _dgrad against autograd through F.conv2d in fp32 on the same bf16-rounded
    operands. dY may be a channel window [co_off, co_off+Cout) of a wider NHWC buffer (extra channels)."""
    import ctypes
    from ee_semantic_segmentation_b200 import _lib
    from ee_semantic_segmentation_b200._lib import check, lib
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, h, w, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cout, R, R, Cin, generator=g) / np.sqrt(R * R * Cin)).to(torch.bfloat16)
    dy_full = torch.randn(N, h, w, Cout + extra, generator=g).to(torch.bfloat16)
    dy = dy_full[..., co_off:co_off + Cout]
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, wr, padding=dil * (R // 2), dilation=dil).backward(dy.float().permute(0, 3, 1, 2))
    ref_dw = wr.grad.permute(0, 2, 3, 1)          # [Cout][R][S][Cin]
    ref_dx = xr.grad.permute(0, 2, 3, 1)          # NHWC
    xd, wd, dyd = x.to(dev()), wt.to(dev()), dy_full.to(dev())
    st = torch.cuda.current_stream().cuda_stream
    dw = torch.full((Cout, R, R, Cin), 7.0, dtype=torch.float32, device=dev())
    wws = torch.empty(lib().eeseg_conv_igemm_wgrad_workspace_bytes(N, h, w, Cin, Cout, R, R), dtype=torch.uint8, device=dev())
    check(lib().eeseg_conv_igemm_wgrad(xd.data_ptr(), dyd.data_ptr(), Cout + extra, Cout + extra, co_off, N, h, w, Cin, Cout,
                                       R, R, dil, dw.data_ptr(), wws.data_ptr(), st), "wgrad")
    dw2 = torch.empty_like(dw)      # split-K partials are summed in a fixed order: bit-reproducible
    check(lib().eeseg_conv_igemm_wgrad(xd.data_ptr(), dyd.data_ptr(), Cout + extra, Cout + extra, co_off, N, h, w, Cin, Cout,
                                       R, R, dil, dw2.data_ptr(), wws.data_ptr(), st), "wgrad")
    assert torch.equal(dw, dw2)
    e_w = (dw.cpu() - ref_dw).abs().max().item() / (ref_dw.abs().max().item() + 1e-9)
    e_x = None
    if extra == 0:
        ws = torch.empty(lib().eeseg_conv_igemm_dgrad_workspace_bytes(Cin, Cout, R, R), dtype=torch.uint8, device=dev())
        dx = torch.empty(N, h, w, Cin, dtype=torch.float32, device=dev())
        check(lib().eeseg_conv_igemm_dgrad(dyd.data_ptr(), wd.data_ptr(), N, h, w, Cin, Cout, R, R, dil, dx.data_ptr(),
                                           _lib.F32, Cin, ws.data_ptr(), st), "dgrad")
        e_x = (dx.cpu() - ref_dx).abs().max().item() / (ref_dx.abs().max().item() + 1e-9)
    torch.cuda.synchronize()
    return e_w, e_x


@pytest.mark.parametrize("cfg", [
    dict(N=1, h=16, w=8, Cin=64, Cout=128, R=1, dil=1),             # one exact 128-pixel tile, one MMA block
    dict(N=1, h=16, w=8, Cin=256, Cout=128, R=3, dil=1),            # taps, zero fill
    dict(N=2, h=65, w=65, Cin=128, Cout=256, R=3, dil=1),           # ragged 11x11 tiles (7 zero rows per box), K split
    dict(N=1, h=65, w=65, Cin=192, Cout=128, R=3, dil=12),          # 64-wide ci tiles
    dict(N=2, h=65, w=65, Cin=256, Cout=256, R=3, dil=36),          # most (tap, tile) pairs all padding
    dict(N=2, h=33, w=47, Cin=512, Cout=256, R=1, dil=1),           # 1x1, deep K split
    dict(N=1, h=24, w=40, Cin=320, Cout=256, R=1, dil=1, co_off=64, extra=128),   # dY is a window of a wider buffer
    dict(N=2, h=33, w=33, Cin=64, Cout=64, R=3, dil=1),             # layer1: 64 output channels = half a 128-row tile
    dict(N=2, h=33, w=33, Cin=256, Cout=64, R=1, dil=1),
    dict(N=1, h=17, w=17, Cin=128, Cout=192, R=1, dil=1, co_off=64, extra=64),    # ragged last co block inside a wider dY
])
def test_conv_wgrad_dgrad_vs_autograd(cfg):
    e_w, e_x = run_conv_grads(**cfg)
    assert e_w < 5e-3, ("wgrad", e_w)      # fp32 accumulation of bf16 products: far inside the 1e-2 bf16 bound
    assert e_x is None or e_x < 5e-3, ("dgrad", e_x)


def test_conv_grads_aspp_shapes():
    """The ASPP shapes of the 513x513 training crop (65x65 maps): Cin = 2048, d = 24; the projection K = 1280."""
    e_w, e_x = run_conv_grads(N=2, h=65, w=65, Cin=2048, Cout=256, R=3, dil=24, seed=3)
    assert e_w < 5e-3 and e_x < 5e-3, (e_w, e_x)
    e_w, e_x = run_conv_grads(N=2, h=65, w=65, Cin=1280, Cout=256, R=1, dil=1, seed=4)
    assert e_w < 5e-3 and e_x < 5e-3, (e_w, e_x)


@pytest.mark.parametrize("cfg", [dict(Cin=128, Cout=128, R=3, h=33, w=33), dict(Cin=256, Cout=512, R=1, h=33, w=33),
                                 dict(Cin=64, Cout=128, R=3, h=24, w=30)])
def test_conv_autograd_stride2_vs_torch(cfg):
    """ConvIgemmFn with stride 2 (layer2.0 of the ResNet): forward on strided TMA boxes, backward through the
    zero-inserted dY on the stride-1 gradient kernels, against autograd through F.conv2d (odd and even sizes)."""
    from ee_semantic_segmentation_b200.head_train import ConvIgemmFn
    Cin, Cout, R, h, w = cfg["Cin"], cfg["Cout"], cfg["R"], cfg["h"], cfg["w"]
    g = torch.Generator().manual_seed(Cin + Cout + R)
    x = torch.randn(2, h, w, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cout, Cin, R, R, generator=g) / np.sqrt(R * R * Cin)).to(torch.bfloat16).float()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    go = torch.randn(2, ho, wo, Cout, generator=g).to(torch.bfloat16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=R // 2)
    assert yr.shape[-2:] == (ho, wo)
    yr.backward(go.float().permute(0, 3, 1, 2))
    xd = x.to(dev()).requires_grad_(True)
    wd = wt.to(dev()).requires_grad_(True)
    yd = ConvIgemmFn.apply(xd, wd, 1, 2)
    yd.backward(go.to(dev()))
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(yd.detach().float().cpu(), yr.detach().permute(0, 2, 3, 1)) < 1e-2
    assert rel(xd.grad.float().cpu(), xr.grad.permute(0, 2, 3, 1)) < 1e-2
    assert rel(wd.grad.cpu(), wr.grad) < 5e-3


@pytest.mark.parametrize("shape", [(2, 33, 37, 64), (1, 64, 64, 128), (2, 1, 5, 8)])
def test_maxpool_train_fwd_bwd_vs_torch(shape):
    """MaxPool2d(3, 2, 1) on NHWC bf16: values, the recorded winning tap (ties: first maximum in window order, as
    ATen; the inputs are quantised so ties are frequent) and the gather-form backward against autograd."""
    from ee_semantic_segmentation_b200.backbone_train import MaxPool3x3s2Fn
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 2).round().clamp_(min=0).to(torch.bfloat16)     # post-ReLU-like, many ties
    N, h, w, C = shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    go = torch.randn(N, ho, wo, C, generator=g).to(torch.bfloat16)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    yr.backward(go.float().permute(0, 3, 1, 2))
    xd = x.to(dev()).requires_grad_(True)
    yd = MaxPool3x3s2Fn.apply(xd)
    yd.backward(go.to(dev()))
    assert torch.equal(yd.detach().float().cpu(), yr.detach().permute(0, 2, 3, 1))
    ref = xr.grad.permute(0, 2, 3, 1)
    got = xd.grad.float().cpu()
    # up to 4 bf16 gradients are summed in fp32 and rounded once
    assert (got - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert torch.equal(got != 0, ref.to(torch.bfloat16).float() != 0) or (got - ref).abs().max().item() < 1e-2


def test_stem_train_vs_torch_modules():
    """Training stem (conv1 7x7/s2 as the space-to-depth 4x1 implicit GEMM, BatchNorm batch statistics + ReLU, max-pool)
    on the eeseg kernels against autograd through the fp32 PyTorch modules: output, conv / BN parameter gradients
    and the running statistics."""
    import copy
    import torchvision
    from ee_semantic_segmentation_b200 import backbone_train
    torch.manual_seed(5)
    r = torchvision.models.resnet50(weights=None)
    r.bn1.weight.data.uniform_(0.5, 1.5); r.bn1.bias.data.normal_(0, 0.2)
    stem_ref = torch.nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool).train()
    stem_dev = copy.deepcopy(stem_ref).to(dev()).train()
    for shape in [(2, 3, 129, 161), (1, 3, 64, 96)]:
        x = torch.randn(*shape)
        stem_ref.zero_grad(); stem_dev.zero_grad()
        yr = stem_ref(x)
        go = torch.randn_like(yr)
        yr.backward(go)
        assert backbone_train.stem_supported(list(stem_dev), x.to(dev()))
        yd = backbone_train.section_forward_train(stem_dev, x.to(dev()))
        assert yd.shape == yr.shape and yd.dtype == torch.bfloat16
        yd.backward(go.to(dev()).to(torch.bfloat16))
        rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
        # the same step on the PyTorch modules under bf16 autocast (cuDNN): the bf16 error level of this layer — the
        # conv1 gradient behind a batch-statistics BatchNorm is a difference of large sums
        stem_amp = copy.deepcopy(stem_ref).to(dev()).train()
        stem_amp.zero_grad()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            ya = stem_amp(x.to(dev()))
        ya.backward(go.to(dev()).to(ya.dtype))
        assert rel(yd.detach(), yr.detach()) < 1e-2
        for got, amp, ref in [(stem_dev[0].weight.grad, stem_amp[0].weight.grad, stem_ref[0].weight.grad),
                              (stem_dev[1].weight.grad, stem_amp[1].weight.grad, stem_ref[1].weight.grad),
                              (stem_dev[1].bias.grad, stem_amp[1].bias.grad, stem_ref[1].bias.grad)]:
            assert rel(got, ref) < max(2e-2, 1.5 * rel(amp, ref)), (rel(got, ref), rel(amp, ref))
        assert rel(stem_dev[1].running_mean, stem_ref[1].running_mean) < 1e-2
        assert rel(stem_dev[1].running_var, stem_ref[1].running_var) < 1e-2
        stem_ref[1].load_state_dict(stem_dev[1].state_dict())       # keep both on the same running statistics
    assert int(stem_dev[1].num_batches_tracked) == 2
