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


def test_train_driver_tracker_checkpoint_and_early_stopping(tmp_path):
    """train_funcs.train (reference :60-269) on a toy model: epochs 1..num_epochs-1, branchy tracker keys
    val_mIoU_<key>, the followed value = mean over the exits, best-validation checkpoint dict, patience stop, start_from."""
    import torch
    from ee_semantic_segmentation_b200.train_funcs import train

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.n_branches = 2

        def forward(self, X):
            return X * self.w

    net = Toy()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    data = [(torch.ones(2, 1), torch.ones(2, 1))] * 3
    seen = []
    vals = iter([0.2, 0.5, 0.4, 0.4, 0.4, 0.9])           # mean mIoU per epoch (maximised)

    def fake_miou(net_, n_exits, n_classes, loader, device):
        assert n_exits == 3 and n_classes == 21
        v = next(vals)
        seen.append(v)
        return {"b1_mIoU": v - 0.1, "b2_mIoU": v, "mIoU": v + 0.1}
    ck = tmp_path / "best.pth"
    tr = train(net, data, lambda a, b: ((a - b) ** 2).mean(), 20, opt, val_iter=data, metrics=[("mIoU", fake_miou)],
               patience=2, saveat=str(ck), device="cpu", minimize=False, n_branches=2, nout_channels=21, ret_lr=True)
    assert set(tr) == {"val_mIoU_b1_mIoU", "val_mIoU_b2_mIoU", "val_mIoU_mIoU", "lr"}
    # best at epoch 2 (0.5); epochs 3, 4 do not improve -> counter reaches the patience, epoch 5 stops the loop
    assert seen == [0.2, 0.5, 0.4, 0.4, 0.4] and len(tr["lr"]) == 5
    sd = torch.load(str(ck), weights_only=False)
    assert set(sd) == {"model_state_dict", "opt_state_dict", "epoch"} and sd["epoch"] == 2
    w_best = sd["model_state_dict"]["w"].clone()
    # no patience: runs epochs 1 .. num_epochs-1 and keeps the best
    net2, seen2 = Toy(), []
    opt2 = torch.optim.SGD(net2.parameters(), lr=0.1)
    vals2 = iter([3.0, 1.0, 2.0])

    def fake2(net_, n_exits, n_classes, loader, device):
        seen2.append(1)
        return {"mIoU": next(vals2)}
    tr2 = train(net2, data, lambda a, b: ((a - b) ** 2).mean(), 4, opt2, val_iter=data, metrics=[("mIoU", fake2)],
                saveat=str(tmp_path / "b2.pth"), start_from=str(ck), nout_channels=21)
    assert len(seen2) == 3 and tr2["val_mIoU"] == [3.0, 1.0, 2.0]
    assert torch.load(str(tmp_path / "b2.pth"), weights_only=False)["epoch"] == 2         # minimised: 1.0 at epoch 2
    assert not torch.equal(net2.w.detach(), w_best)                                        # started from it, then trained on


def test_train_and_load_state_dict_drop_plans_and_graphs():
    """Everything derived from the parameters — folded kernel plans, captured CUDA graphs — is dropped when the model
    enters training mode or loads a state dict (weights_epoch): a training step replayed as a CUDA graph does not
    advance the parameters' version counters, so version-keyed plans alone would go stale. Also: pickles and deep
    copies carry no plans / graphs, and reference-style pickles get the run-time defaults."""
    import copy
    import io
    import torch
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    net = branchyDeepv3(None, "deeplabv3_resnet50", 1, 65, sections=[18, 2], pretrained=False).eval()
    e0 = net.weights_epoch
    net._plans[0] = ("key", object())
    net._section_plans[0] = ("key", object())
    net._lowres_graphs["shape"] = object()
    net.eval()
    assert net._plans and net.weights_epoch == e0                     # eval() keeps them
    net.train()
    assert net._plans == {} and net._section_plans == {} and net._lowres_graphs == {} and net.weights_epoch == e0 + 1
    net.eval()
    net._plans[0] = ("key", object())
    sd = copy.deepcopy(net.state_dict())
    net.load_state_dict(sd)
    assert net._plans == {} and net.weights_epoch > e0 + 1
    net._plans[0] = ("key", torch.zeros(1))
    n2 = copy.deepcopy(net)
    assert n2._plans == {} and n2.weights_epoch == net.weights_epoch and net._plans
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    n3 = torch.load(buf, weights_only=False)
    assert n3._plans == {} and n3.fast_inference and n3.graph_inference
    st = n3.__getstate__()
    for k in ("fast_inference", "fast_backbone", "graph_inference", "weights_epoch", "_plans"):
        st.pop(k)
    n4 = branchyDeepv3.__new__(branchyDeepv3)
    n4.__setstate__(st)                                               # a pickle without the run-time attributes
    assert n4.fast_inference and n4.fast_backbone and n4.graph_inference and n4.weights_epoch == 0 and n4._plans == {}


def test_grouped_conv_work_list_tile_width_rule():
    """head_plan.group_schedule: the grouped ASPP work list (4 problems on a 65x65 map, Cout 256). With 4 images the
    256-column tiles already cover the 148 SMs; with one image the list switches to 128-column tiles (twice the items),
    every (problem, tile) exactly once, sorted by cost inside each round of n_ctas items (snake order)."""
    import numpy as np
    from ee_semantic_segmentation_b200.head_plan import group_schedule
    ks, ds = [1, 3, 3, 3], [1, 12, 24, 36]
    s4 = group_schedule(4, 65, 65, 2048, 256, ks, ds).numpy()
    s1 = group_schedule(1, 65, 65, 2048, 256, ks, ds).numpy()
    s2 = group_schedule(2, 65, 65, 2048, 256, ks, ds).numpy()
    assert len(s4) == 4 * 4 * 36 and len(s2) == 4 * 2 * 36           # 256-column tiles: one channel tile per position
    assert len(s1) == 4 * 1 * 36 * 2                                  # 128-column tiles
    for s, n_items in ((s4, 4 * 36), (s1, 2 * 36), (s2, 2 * 36)):
        assert len(set(s.tolist())) == len(s)
        for g in range(4):
            tiles = np.sort(s[(s >> 24) == g] & 0xffffff)
            np.testing.assert_array_equal(tiles, np.arange(n_items))
    # the first round holds the most expensive items: none of the cheap 1x1 tiles (problem 0)
    assert not np.any((s4[:148] >> 24) == 0)
    # pair list (CTA pairs): the same tiles, entries 2i / 2i+1 = one tile position in images 2k / 2k+1 of one problem
    p4 = group_schedule(4, 65, 65, 2048, 256, ks, ds, pairs=True, n_clusters=74).numpy()
    assert sorted(p4.tolist()) == sorted(s4.tolist())
    a = p4.reshape(-1, 2)
    assert np.all((a[:, 0] >> 24) == (a[:, 1] >> 24)) and np.all((a[:, 1] & 0xffffff) - (a[:, 0] & 0xffffff) == 36)
    assert np.all(((a[:, 0] & 0xffffff) // 36) % 2 == 0)
    assert not np.any((a[:74, 0] >> 24) == 0)
    with pytest.raises(ValueError):
        group_schedule(3, 65, 65, 2048, 256, ks, ds, pairs=True, n_clusters=74)


def test_old_pickles_and_weights_token():
    """(1) a whole-module pickle written by an older torch lacks the newer hook dictionaries: __setstate__ goes through
    nn.Module.__setstate__, which restores them (the reference's checkpoint format is tch.save(net), eval_br_ent.py:146).
    (2) weights_token() — what every captured CUDA graph is tagged with and compares before each replay — moves on a
    SUB-module load_state_dict, an in-place edit in eval mode and an explicit _bump_epoch (graph-replayed training),
    and stays put otherwise."""
    import copy
    import torch
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    net = branchyDeepv3(None, "deeplabv3_resnet50", 1, 65, sections=[18, 2], pretrained=False).eval()
    st = net.__getstate__()
    for k in ("_forward_pre_hooks_with_kwargs", "_forward_hooks_with_kwargs", "_forward_hooks_always_called",
              "_state_dict_pre_hooks", "_load_state_dict_post_hooks", "_backward_pre_hooks"):
        st.pop(k, None)
    old = branchyDeepv3.__new__(branchyDeepv3)
    old.__setstate__(st)
    for k in ("_forward_pre_hooks_with_kwargs", "_forward_hooks_always_called", "_backward_pre_hooks"):
        assert hasattr(old, k), k
    old.state_dict()                                   # walks _state_dict_pre_hooks: raised AttributeError before
    with torch.no_grad():
        try:
            old(torch.zeros(1, 3, 33, 33))              # __call__ reads the hook dicts before reaching forward()
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)

    t0 = net.weights_token()
    assert net.weights_token() == t0
    net.branches.load_state_dict(copy.deepcopy(net.branches.state_dict()))      # never reaches branchyDeepv3's hooks
    t1 = net.weights_token()
    assert t1 != t0 and t1[0] == t0[0]
    with torch.no_grad():
        net.classifier[-1].bias.add_(1.0)                                        # in-place edit in eval mode
    t2 = net.weights_token()
    assert t2 != t1
    net._bump_epoch()
    assert net.weights_token()[0] == t2[0] + 1
    net.float()                                                                  # _apply: storages may have moved
    assert net.weights_token()[0] == t2[0] + 2


def test_fastdiv_formula_is_exact_below_2_31():
    """The multiply-high division the conv kernel uses for its tile coordinates (csrc/conv_igemm.cu: make_fastdiv / fd_div):
    mul = ceil(2^(31+l) / d), l = ceil(log2 d); q = umulhi(n, mul) >> (l - 1). Restated here and checked against integer
    division for every divisor up to 4096, a spread of larger ones, and edge / random n below 2^31 (the kernel's n are work-item
    and tile indices, < 2^24)."""
    import numpy as np
    rng = np.random.default_rng(0)
    ds = list(range(2, 4097)) + [4097, 5016, 16384, 16385, 65535, 65536, 1 << 20, (1 << 24) - 1, (1 << 24) + 1]
    n = np.concatenate([np.arange(0, 5000), rng.integers(0, 1 << 31, 20000), np.array([(1 << 31) - 1, (1 << 31) - 2, (1 << 24) - 1])]
                       ).astype(np.uint64)
    for d in ds:
        l = int(np.ceil(np.log2(d)))
        if (1 << l) < d:
            l += 1
        p = 31 + l
        mul = ((1 << p) + d - 1) // d
        assert mul < (1 << 32)
        q = ((n * np.uint64(mul)) >> np.uint64(32)) >> np.uint64(p - 32)
        extra = np.array([d - 1, d, d + 1, 2 * d - 1, 2 * d, 7 * d - 1, 7 * d], dtype=np.uint64)
        qe = ((extra * np.uint64(mul)) >> np.uint64(32)) >> np.uint64(p - 32)
        assert np.array_equal(q, n // np.uint64(d)), d
        assert np.array_equal(qe, extra // np.uint64(d)), d
