"""The kernels that complete the training step on eeseg code (csrc/train_misc.cu) against plain PyTorch references:
dropout, the multi-tensor SGD update, the pooled ASPP branch, the final 1x1 classifier, and the bucketed flat
gradient buffer on one rank."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_dropout_mask_scale_backward_and_graph_replay():
    from ee_semantic_segmentation_b200.head_train import DropoutFn, dropout_state
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(4, 65, 65, 256, generator=g) + 3).to(dev()).to(torch.bfloat16).requires_grad_(True)
    dropout_state(dev(), seed=7)
    y = DropoutFn.apply(x, 0.5)
    kept = y != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.5) < 2e-3, frac                        # 4.3 M Bernoulli(0.5) draws
    assert torch.equal(y[kept].float(), (x.detach()[kept].float() * 2).to(torch.bfloat16).float())
    # no structure along channels / pixels: per-channel and per-pixel keep rates are all close to 0.5
    assert (kept.float().mean(dim=(0, 1, 2)) - 0.5).abs().max().item() < 0.02
    assert (kept.float().mean(dim=3) - 0.5).abs().max().item() < 0.2
    gy = torch.randn_like(y)
    y.backward(gy)
    assert torch.equal(x.grad[kept].float(), (gy[kept].float() * 2).to(torch.bfloat16).float())
    assert torch.count_nonzero(x.grad[~kept]) == 0
    # a second call draws a different mask; re-seeding reproduces the first
    y2 = DropoutFn.apply(x.detach(), 0.5)
    assert not torch.equal(y2 != 0, kept)
    dropout_state(dev(), seed=7)
    y3 = DropoutFn.apply(x.detach(), 0.5)
    assert torch.equal(y3, y.detach())
    # p = 0.25 -> keep 0.75, scale 4/3
    y4 = DropoutFn.apply(x.detach(), 0.25)
    assert abs((y4 != 0).float().mean().item() - 0.75) < 2e-3
    # captured in a CUDA graph: the offset lives on the device, so every replay draws a fresh mask
    xs = x.detach().clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        DropoutFn.apply(xs, 0.5)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = DropoutFn.apply(xs, 0.5)
    gr.replay()
    m1 = (out != 0).clone()
    gr.replay()
    m2 = (out != 0).clone()
    assert not torch.equal(m1, m2) and abs(m2.float().mean().item() - 0.5) < 2e-3


def test_sgd_matches_torch_sgd_and_shares_its_state_dict():
    from ee_semantic_segmentation_b200.train_funcs import SGD
    torch.manual_seed(3)
    shapes = [(64, 3, 7, 7), (64,), (21, 256, 1, 1), (21,), (256, 2048, 1, 1), (300001,)]
    pa = [nn.Parameter(torch.randn(s, device=dev())) for s in shapes]
    pb = [nn.Parameter(p.detach().clone()) for p in pa]
    groups = lambda ps: [{'params': ps[:2], 'lr': 1e-2}, {'params': ps[2:], 'lr': 3e-2}]
    oa = SGD(groups(pa), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    ob = torch.optim.SGD(groups(pb), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    for step in range(4):
        oa.zero_grad()
        ob.zero_grad()
        for a, b in zip(pa, pb):
            gr = torch.randn_like(a)
            a.grad.add_(gr)                              # the flat views are accumulated into, never replaced
            b.grad = gr.clone()
        if step == 2:                                    # a scheduler changes the learning rates
            for o in (oa, ob):
                o.param_groups[0]['lr'] = 5e-3
                o.param_groups[1]['lr'] = 1e-3
        oa.step()
        ob.step()
        for a, b in zip(pa, pb):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), step
    # state dict: torch's layout both ways
    sd = oa.state_dict()
    ob2 = torch.optim.SGD(groups(pb), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    ob2.load_state_dict(copy.deepcopy(sd))
    assert torch.allclose(ob2.state[pb[0]]['momentum_buffer'], oa.state[pa[0]]['momentum_buffer'])
    oa2 = SGD(groups(pa), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    oa2.load_state_dict(copy.deepcopy(ob.state_dict()))
    assert torch.allclose(oa2.state[pa[-1]]['momentum_buffer'], ob.state[pb[-1]]['momentum_buffer'])
    assert oa2.state[pa[-1]]['momentum_buffer'].data_ptr() >= oa2._mom.data_ptr()      # still a view of the flat buffer
    assert oa2.param_groups[1]['lr'] == pytest.approx(1e-3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SGD([nn.Parameter(torch.zeros(3))], lr=0.1)


def test_pooled_branch_and_final_conv_match_torch_modules():
    from torchvision.models.segmentation.deeplabv3 import ASPPPooling
    from ee_semantic_segmentation_b200.head_train import FinalConvFn, pooled_branch_train
    torch.manual_seed(5)
    N, Cin, h, w = 4, 512, 33, 29
    pool_a = ASPPPooling(Cin, 256).to(dev()).train()
    pool_b = copy.deepcopy(pool_a)
    x = torch.randn(N, Cin, h, w, device=dev())
    xa = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xb = xa.detach().float().requires_grad_(True)
    ya = pooled_branch_train(pool_a, xa.permute(0, 2, 3, 1))
    yb = pool_b(xb)
    assert ya.shape == yb.shape == (N, 256, h, w)
    assert (ya.float() - yb).abs().max().item() < 3e-2 * yb.abs().max().item()
    g = torch.randn_like(yb)
    ya.backward(g.to(torch.bfloat16))
    yb.backward(g)
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()
    assert cos(pool_a[1].weight.grad, pool_b[1].weight.grad) > 0.999
    assert cos(pool_a[2].weight.grad, pool_b[2].weight.grad) > 0.999 and cos(pool_a[2].bias.grad, pool_b[2].bias.grad) > 0.999
    assert cos(xa.grad.float(), xb.grad) > 0.999
    assert torch.allclose(pool_a[2].running_var, pool_b[2].running_var, rtol=2e-2, atol=1e-4)
    assert int(pool_a[2].num_batches_tracked) == 1

    # final classifier
    conv_a = nn.Conv2d(256, 21, 1).to(dev())
    conv_b = copy.deepcopy(conv_a)
    y = torch.randn(N, 256, h, w, device=dev())
    ya_in = y.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    yb_in = ya_in.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    oa = FinalConvFn.apply(ya_in, conv_a.weight, conv_a.bias)                    # [N,h,w,32] fp32
    ob = conv_b(yb_in)
    assert oa.shape == (N, h, w, 32) and torch.count_nonzero(oa[..., 21:]) == 0
    assert (oa[..., :21].permute(0, 3, 1, 2) - ob).abs().max().item() < 1e-2 * ob.abs().max().item()
    g = torch.randn_like(ob)
    gp = torch.zeros_like(oa)
    gp[..., :21] = g.permute(0, 2, 3, 1)
    oa.backward(gp)
    ob.backward(g)
    assert cos(conv_a.weight.grad, conv_b.weight.grad) > 0.9995 and cos(conv_a.bias.grad, conv_b.bias.grad) > 0.9995
    assert abs(conv_a.weight.grad.norm().item() / conv_b.weight.grad.norm().item() - 1) < 1e-2
    assert cos(ya_in.grad.float().permute(0, 3, 1, 2), yb_in.grad) > 0.9995


def test_graphed_train_step_with_eeseg_sgd_follows_the_scheduler():
    """GraphedTrainStep + the eeseg SGD: two replays at different learning rates WITHOUT re-capturing equal two eager
    steps of the same model with torch.optim.SGD at those rates (Dropout off: the two runs draw different masks)."""
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    from ee_semantic_segmentation_b200.train_funcs import SGD, GraphedTrainStep, make_optimizer
    torch.manual_seed(0)
    net_a = branchyDeepv3(None, "deeplabv3_resnet50", 1, 65, sections=[18, 2], pretrained=False).to(dev()).train()
    for mod in net_a.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    net_a.strict_kernels = True
    net_b = copy.deepcopy(net_a)
    g = torch.Generator().manual_seed(9)
    X = torch.randn(2, 3, 65, 65, generator=g).to(dev())
    y = torch.randint(0, 22, (2, 1, 65, 65), generator=g).to(dev())
    loss = BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=2)
    oa = make_optimizer(net_a, lr=1e-2, base_lr=1e-3)
    assert isinstance(oa, SGD)
    groups = [{'params': [p for p in grp['params']], 'lr': grp['lr']} for grp in make_optimizer(net_b, lr=1e-2, base_lr=1e-3).param_groups]
    for p in net_b.parameters():
        p.grad = None
    ob = torch.optim.SGD(groups, lr=1e-2, momentum=0.9, weight_decay=5e-4)
    step = GraphedTrainStep(net_a, loss, oa, X, y)
    for k, scale in enumerate((1.0, 0.25)):
        for o in (oa, ob):
            for grp, base in zip(o.param_groups, (1e-3, 1e-2, 1.1e-2)):
                grp['lr'] = base * scale
        la = step(X, y)
        ob.zero_grad(set_to_none=True)
        lb = loss(net_b(X), y)
        lb.backward()
        ob.step()
        assert float(la) == pytest.approx(float(lb.detach()), rel=2e-3), k
    pa = torch.cat([p.detach().flatten() for p in net_a.classifier.parameters()])
    pb = torch.cat([p.detach().flatten() for p in net_b.classifier.parameters()])
    assert torch.allclose(pa, pb, rtol=2e-2, atol=2e-5)
    step.release()


def test_direct_parameter_gradients_equal_autograd_accumulation():
    """Weight gradients of the convolutions and dgamma / dbeta of the BatchNorms written straight into the flat-buffer views
    of the eeseg SGD (no AccumulateGrad launch) against the same backward through autograd's accumulation (torch SGD, plain
    .grad tensors): every parameter gradient of a head + a Bottleneck section, and the gradient-hook bucket countdown."""
    from ee_semantic_segmentation_b200 import parallel
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    from ee_semantic_segmentation_b200.train_funcs import SGD
    torch.manual_seed(0)
    net_a = branchyDeepv3(None, "deeplabv3_resnet50", 1, 65, sections=[18, 2], pretrained=False).to(dev()).train()
    for mod in net_a.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    net_b = copy.deepcopy(net_a)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(2, 3, 65, 65, generator=g).to(dev())
    y = torch.randint(0, 22, (2, 1, 65, 65), generator=g).to(dev())
    loss = BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=2)
    opt = SGD(list(net_a.parameters()), lr=0.0, momentum=0.9, weight_decay=0.0, buckets=3)
    fg = opt.flat_grads
    fg._world = lambda: 2                      # pretend two ranks so that the countdown reduces buckets ...
    reduced = []
    fg._reduce_range = lambda a, b: reduced.append((a, b))     # ... without a process group
    opt.zero_grad()
    fg.begin()
    loss(net_a(X), y).backward()
    assert all(v == -1 for v in fg._left), (fg._left, fg._need)   # hooks AND direct writes counted every parameter
    assert sorted(reduced) == sorted(fg.buckets)
    fg._armed = False
    loss(net_b(X), y).backward()                                # plain autograd accumulation into fresh .grad tensors
    n_direct = 0
    for (name, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        assert pa.grad.data_ptr() >= fg.flat.data_ptr()
        assert torch.allclose(pa.grad, pb.grad, rtol=2e-3, atol=1e-6), name
        n_direct += parallel.direct_grad(pa) is not None
    assert n_direct == len(list(net_a.parameters()))


def test_weight_prep_multi_matches_permute_and_rotation():
    """eeseg_weight_prep_multi over a mixed set of conv weights (1x1, 3x3, 64..2048 channels): krsc = w.permute(0,2,3,1) in
    bf16, rot = the 180-degree rotated transpose [Cin,R,S,Cout] (what eeseg_conv_weight_rot180_t produces per layer)."""
    from ee_semantic_segmentation_b200.head_train import TrainWeightCache
    torch.manual_seed(2)
    convs = [nn.Conv2d(64, 64, 3, padding=1, bias=False), nn.Conv2d(256, 64, 1, bias=False),
             nn.Conv2d(128, 512, 1, bias=False), nn.Conv2d(2048, 256, 3, padding=12, dilation=12, bias=False),
             nn.Conv2d(512, 512, 3, padding=2, dilation=2, bias=False)]
    convs = [c.to(dev()) for c in convs]
    cache = TrainWeightCache(convs)
    cache.refresh()
    for c in convs:
        wt, wT = cache.entries[id(c.weight)]
        ref = c.weight.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
        assert torch.equal(wt, ref)
        ref_rot = c.weight.detach().flip(2, 3).permute(1, 2, 3, 0).to(torch.bfloat16)
        assert torch.equal(wT, ref_rot)
    assert cache.valid()
