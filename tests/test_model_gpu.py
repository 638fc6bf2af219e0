"""End-to-end parity on the GPU: the product model / evaluators / engine against the CPU oracle port
(oracle/model_port.py, itself pinned to the reference by tests/golden/model.npz) and against the
reference's br_evaluator fixture (tests/golden/br_eval.npz)."""
import numpy as np
import pytest
import torch

from oracle import model_port
from oracle import restate as R

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def nets():
    sections = [16, 3, 1]
    port = model_port.build_port(sections, seed=0, branch_seed=7).eval()
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=sections, pretrained=False)
    net.load_state_dict(port.state_dict())
    return port, net.to(dev()).eval()


def test_forward_logits_vs_oracle_port(nets):
    port, net = nets
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 129, 129, generator=g)
    with torch.no_grad():
        ref = port(x)                       # fp32 CPU, reference data flow
        got = net(x.to(dev()))
    assert got.shape == ref.shape == (3, 2, 21, 129, 129) and got.dtype == torch.float32
    from conftest import assert_bf16_model_close
    for e in range(3):
        assert_bf16_model_close(got[e], ref[e], e)      # relative L2 < 1.25e-2 and max-abs < 2.5e-2 of the range
    # fp32 PyTorch-module path of the same model object (training path) agrees to fp32 accuracy
    net.fast_inference = False
    with torch.no_grad():
        got32 = net(x.to(dev()))
    net.fast_inference = True
    assert (got32.cpu() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


def test_golden_forward_slice(golden):
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    d = golden("model")
    sections = [int(s) for s in d["n2_sections"]]
    port = model_port.build_port(sections, seed=0, branch_seed=102)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=sections, pretrained=False)
    net.load_state_dict(port.state_dict())
    net = net.to(dev()).eval()
    with torch.no_grad():
        y = net(torch.tensor(d["x"]).to(dev()))
    ref = d["n2_out_slice"]
    got = y[..., ::8, ::8].cpu().numpy()
    from conftest import assert_bf16_model_close
    for e in range(ref.shape[0]):
        assert_bf16_model_close(got[e], ref[e], e)


def test_engine_decisions_and_confusion_vs_oracle(nets):
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    port, net = nets
    g = torch.Generator().manual_seed(22)
    N = 3
    x = torch.randn(N, 3, 97, 113, generator=g)
    y = torch.randint(0, 22, (N, 1, 97, 113), generator=g)
    with torch.no_grad():
        logits = port(x)
    ents = np.array([[R.img_norm_entropy(R.softmax_c(logits[i, k].numpy(), 0), 21) for k in range(N)] for i in range(2)])
    tau = float(np.median(ents))            # some images leave early, some do not
    for skip_compute in (False, True):
        eng = EarlyExitEngine(net, 21, tau, skip_compute=skip_compute)
        out = eng.evaluate(x.to(dev()), y.to(dev()))
        sc = out["scores"].cpu().numpy()
        exits = out["exit"].cpu().numpy()
        for k in range(N):
            # decisions identical except within 1e-4 of the threshold; bf16 logits move the mean
            # entropy by O(1e-3), so the comparison uses the engine's own scores for the rule and the
            # oracle's for the value
            rule = R.first_confident_exit([sc[0, k], sc[1, k]] if not skip_compute else
                                          [s for s in (sc[0, k], sc[1, k])], tau)
            assert exits[k] == rule
            assert abs(sc[0, k] - ents[0, k]) < 5e-3
        # confusion matrix of the exit taken == oracle confusion of the engine's own prediction map
        pred = out["pred"].cpu().numpy().reshape(N, -1)
        cm_ref = R.confusion_matrix(pred, y.numpy().reshape(N, -1), 21)
        np.testing.assert_array_equal(eng.cm[-1].cpu().numpy(), cm_ref.sum(0))
        assert int(eng.counts[-1]) == N and int(eng.counts[:3].sum()) == N
        r = eng.results()
        assert r["out_gl"] == N and set(r) >= {"b1_mIoU", "b1_count", "b2_mIoU", "mIoU_out", "mIoU_gl", "t", "pool", "pool_size"}


def test_engine_graphed_skipping_equals_eager_skipping(nets):
    """skip_compute + use_graph (one CUDA graph per exit stage and active-image count, 4-byte host read per gate)
    against the eager compute-skipping engine: same exits, maps, scores, confusion matrices and counters, over
    several batches (graphs reused) and thresholds that make all / some / no images leave early."""
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    _, net = nets
    g = torch.Generator().manual_seed(31)
    N = 4
    batches = [(torch.randn(N, 3, 97, 113, generator=g).to(dev()), torch.randint(0, 22, (N, 1, 97, 113), generator=g).to(dev()))
               for _ in range(3)]
    sc = EarlyExitEngine(net, 21, 0.5).evaluate(*batches[0])["scores"].cpu()
    taus = [0.0, float(sc[0].median()), float(sc.max()) + 1e-3, float((sc[0].min() + sc[0].sort().values[1]) / 2)]
    for tau in taus:
        eager = EarlyExitEngine(net, 21, tau, skip_compute=True)
        graphed = EarlyExitEngine(net, 21, tau, skip_compute=True, use_graph=True)
        for rep in range(2):                        # second round: every graph already captured
            for X, y in batches:
                a = eager.evaluate(X, y)
                b = graphed.evaluate(X, y)
                assert torch.equal(a["exit"], b["exit"])
                assert torch.equal(a["pred"], b["pred"])
                took = a["exit"].cpu()
                for k in range(N):                   # scores of the gates an image actually reached
                    for i in range(min(int(took[k]) + 1, 2)):
                        assert float(a["scores"][i, k]) == float(b["scores"][i, k])
        assert torch.equal(eager.cm, graphed.cm)
        assert torch.equal(eager.counts, graphed.counts)
        assert torch.equal(eager.exited_px, graphed.exited_px)
        assert int(graphed.counts[-1]) == 2 * 3 * N
    # skip=1 (the first early exit is not allowed to answer: its stage passes every image on) and a single image
    for X, y in [batches[0], (batches[2][0][:1], batches[2][1][:1])]:
        eager = EarlyExitEngine(net, 21, taus[2], skip=1, skip_compute=True)
        graphed = EarlyExitEngine(net, 21, taus[2], skip=1, skip_compute=True, use_graph=True)
        a, b = eager.evaluate(X, y), graphed.evaluate(X, y)
        assert torch.equal(a["exit"], b["exit"]) and torch.equal(a["pred"], b["pred"]) and int(b["exit"].min()) >= 1
        assert torch.equal(eager.cm, graphed.cm) and torch.equal(eager.counts, graphed.counts)
    # inference-only entry
    graphed = EarlyExitEngine(net, 21, taus[1], skip_compute=True, use_graph=True)
    eager = EarlyExitEngine(net, 21, taus[1], skip_compute=True)
    a, b = eager.infer(batches[1][0]), graphed.infer(batches[1][0])
    assert torch.equal(a["exit"], b["exit"]) and torch.equal(a["pred"], b["pred"])


def test_br_evaluator_golden(golden):
    """The reference's br_evaluator result dicts (fake 3-exit net, several tau / pool modes)."""
    from ee_semantic_segmentation_b200.eval_br_ent import br_evaluator
    d = golden("br_eval")
    ys, tg = torch.tensor(d["y"]), torch.tensor(d["targets"])
    n_img, E, _, C, H, W = ys.shape

    class FakeNet:
        def __init__(self): self.k = 0
        def __call__(self, X):
            out = ys[self.k].to(dev()); self.k += 1
            return out
    loader = [(torch.zeros(1, 3, H, W), tg[k]) for k in range(n_img)]
    for key in d["configs"]:
        key = str(key)
        tau, size = float(d[f"{key}/t"]), int(d[f"{key}/pool_size"])
        metric = "ent" if "_ent" in key else ("max" if "_max" in key else "min")
        res = br_evaluator(FakeNet(), E, C, loader, dev(), tau, metric=metric, size=size)
        for k, v in res.items():
            if k == "pool":
                continue
            ref = float(d[f"{key}/{k}"])
            assert (np.isnan(v) and np.isnan(ref)) or v == pytest.approx(ref, abs=1e-6), (key, k, v, ref)


def test_mIoU_evaluator_and_operator(nets):
    from ee_semantic_segmentation_b200.ee_dnn_op_ne import eval_ee_deeplabv3
    from ee_semantic_segmentation_b200.eval_br_ent import img_norm_entropy
    from ee_semantic_segmentation_b200.eval_mIoU import mIoU_evaluator
    port, net = nets
    g = torch.Generator().manual_seed(23)
    x = torch.randn(2, 3, 65, 81, generator=g)
    y = torch.randint(0, 22, (2, 1, 65, 81), generator=g)
    res = mIoU_evaluator(net, 3, 21, [(x, y)], dev())
    assert set(res) == {"b1_mIoU", "b2_mIoU", "mIoU"}
    # exact integer path == oracle on the model's own argmax
    with torch.no_grad():
        logits = net(x.to(dev()))
    mo = R.MIoU(21); mo(logits[-1].cpu().numpy(), y.numpy())
    a, b = res["mIoU"], float(mo.compute())
    assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 5e-3
    op = eval_ee_deeplabv3(net, img_norm_entropy(21), 2.0, device=dev())     # tau=2: leaves at exit 1
    out = op(x[0].to(dev()))
    assert out["n"] == 1 and out["exit"].shape == (65, 81) and out["exit"].dtype == torch.int64
    assert out["exit_flops"] < out["last_flops"] and out["edge_flops"] == out["exit_flops"]
    op = eval_ee_deeplabv3(net, img_norm_entropy(21), -1.0, device=dev())    # never confident
    out = op(x[0].to(dev()))
    assert out["n"] == 3 and torch.equal(out["exit"], out["last"])
    # graph-replayed stages (default) == eagerly launched stages, for every exit pattern / option
    sc = [eval_ee_deeplabv3(net, img_norm_entropy(21), -1.0, use_graph=False)._score(lo, (65, 81))[0]
          for lo in net.forward_lowres(x[:1].to(dev()))[:2]]
    for th in (2.0, -1.0, (sc[0] + sc[1]) / 2):
        for kw in (dict(), dict(ignore=[0]), dict(compute_last=False), dict(less_than=False)):
            a = eval_ee_deeplabv3(net, img_norm_entropy(21), th, use_graph=False, **kw)
            b = eval_ee_deeplabv3(net, img_norm_entropy(21), th, **kw)
            assert b.use_graph and not a.use_graph
            for img in (x[0], x[1], x[0]):           # the third call replays graphs captured by the first
                oa, ob = a(img.to(dev())), b(img.to(dev()))
                assert set(oa) == set(ob), (th, kw)
                for k in oa:
                    assert torch.equal(oa[k], ob[k]) if torch.is_tensor(oa[k]) else oa[k] == ob[k], (th, kw, k)


def test_engine_graph_and_pipelined_match_eager(nets):
    """CUDA-graph replay and the double-buffered streaming API give the eager engine's results."""
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    port, net = nets
    g = torch.Generator().manual_seed(24)
    batches = [(torch.randn(2, 3, 65, 81, generator=g).pin_memory(),
                torch.randint(0, 22, (2, 1, 65, 81), generator=g).pin_memory()) for _ in range(5)]
    tau = 0.97
    eager = EarlyExitEngine(net, 21, tau)
    ref = [eager.evaluate(X.to(dev()), y.to(dev())) for X, y in batches]
    ref = [(r["exit"].cpu(), r["scores"].cpu()) for r in ref]
    graph = EarlyExitEngine(net, 21, tau, use_graph=True)
    outs = list(graph.evaluate_pipelined(iter(batches)))
    assert len(outs) == len(batches)
    for (e1, s1), (e2, s2) in zip(ref, outs):
        assert torch.equal(e1, e2)
        assert torch.allclose(s1, s2, atol=1e-6)
    assert torch.equal(graph.cm.cpu(), eager.cm.cpu()) and torch.equal(graph.counts.cpu(), eager.counts.cpu())


def test_threshold_sweep_equals_per_tau_runs(nets):
    """One-pass sweep == the per-tau evaluator, count for count."""
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine, ThresholdSweep
    port, net = nets
    g = torch.Generator().manual_seed(25)
    batches = [(torch.randn(2, 3, 65, 81, generator=g).to(dev()), torch.randint(0, 22, (2, 1, 65, 81), generator=g).to(dev()))
               for _ in range(3)]
    probe = EarlyExitEngine(net, 21, 0.5)
    sc = torch.cat([probe.infer(X)["scores"] for X, _ in batches], dim=1).cpu()
    taus = [0.0, float(sc[0].median()), float(sc[1].median()), 2.0]
    sweep = ThresholdSweep(net, 21, taus)
    for X, y in batches:
        sweep.update(X, y)
    res = sweep.results()
    for t, tau in enumerate(taus):
        eng = EarlyExitEngine(net, 21, tau)
        for X, y in batches:
            eng.evaluate(X, y)
        assert torch.equal(eng.cm.cpu(), sweep.cm[t].cpu()), tau
        assert torch.equal(eng.counts.cpu(), sweep.counts[t].cpu()), tau
        r = eng.results()
        for k in ("b1_count", "b2_count", "count_out", "out_gl"):
            assert r[k] == res[t][k]
    assert res[0]["count_out"] == 6 and res[-1]["b1_count"] == 6


def test_training_step_matches_torch_losses(nets):
    """One train_epoch step (config 3: multi-exit CE, b_reduction='sum') with the eeseg loss gives the
    same parameter update as the same step with torch.nn.CrossEntropyLoss (the reference's loss)."""
    import copy
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    from ee_semantic_segmentation_b200.train_funcs import make_optimizer, poly_scheduler, train_epoch
    port, net = nets
    g = torch.Generator().manual_seed(26)
    X = torch.randn(2, 3, 65, 65, generator=g)
    y = torch.randint(0, 22, (2, 1, 65, 65), generator=g)
    net_a = copy.deepcopy(net)
    net_b = copy.deepcopy(net)
    opt_a = make_optimizer(net_a, lr=1e-2, base_lr=1e-3)
    assert [grp['lr'] for grp in opt_a.param_groups] == pytest.approx([1e-3, 1e-2, 1.1e-2])
    sched = poly_scheduler(opt_a, 10)
    from ee_semantic_segmentation_b200.head_train import dropout_state
    dropout_state(dev(), seed=123)   # same Dropout(0.5) masks in both runs (ASPP.project is active in train mode)
    l_a = train_epoch(net_a, [(X, y)], BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=3), opt_a, dev())
    sched.step()
    assert opt_a.param_groups[1]['lr'] == pytest.approx(1e-2 * (1 - 1 / 10) ** .9)

    class RefLoss(torch.nn.Module):      # my_pixelwise_xentropy.py:36-46 written with torch ops
        def forward(self, y_pred, t):
            ce = torch.nn.CrossEntropyLoss(ignore_index=21)
            return sum(ce(y_pred[i], t.squeeze(1)) for i in range(3))
    opt_b = make_optimizer(net_b, lr=1e-2, base_lr=1e-3)
    dropout_state(dev(), seed=123)
    l_b = train_epoch(net_b, [(X, y)], RefLoss(), opt_b, dev())
    assert l_a.item() == pytest.approx(l_b.item(), rel=1e-4)
    pa = torch.cat([p.detach().flatten() for p in net_a.branches.parameters()])
    pb = torch.cat([p.detach().flatten() for p in net_b.branches.parameters()])
    assert torch.allclose(pa, pb, rtol=1e-3, atol=1e-6)
    net.eval()


def test_training_heads_on_eeseg_convs_match_torch_autograd(nets):
    """Training forward/backward with the head convolutions on the tcgen05 kernels (fwd, dgrad, wgrad; bf16
    activations) against the same step on the PyTorch modules (fp32 autograd through cuDNN): loss and logits
    within the bf16 bound, gradients of head AND backbone parameters (the latter flow through dgrad) close in
    direction and size. Dropout is switched off so both runs see the same network."""
    import copy
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    port, net = nets
    g = torch.Generator().manual_seed(31)
    X = torch.randn(2, 3, 129, 129, generator=g).to(dev())
    y = torch.randint(0, 22, (2, 1, 129, 129), generator=g).to(dev())
    loss_fn = BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=3)
    runs = {}
    for fast in (True, False):
        m = copy.deepcopy(net).train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.fast_training_heads = fast
        m.fast_training_backbone = False
        out = m(X)
        l = loss_fn(out, y)
        l.backward()
        runs[fast] = (out.detach(), l.item(),
                      torch.cat([p.grad.flatten() for p in m.branches.parameters()] +
                                [p.grad.flatten() for p in m.classifier.parameters()]),
                      torch.cat([p.grad.flatten() for p in m.base_model.parameters()]),
                      {k: v.clone() for k, v in m.state_dict().items() if k.endswith('running_var') and 'branches' in k})
    from conftest import assert_bf16_model_close
    (of, lf, ghf, gbf, rvf), (ot, lt, ght, gbt, rvt) = runs[True], runs[False]
    assert abs(lf - lt) < 1e-2 * abs(lt), (lf, lt)
    for e in range(3):
        assert_bf16_model_close(of[e], ot[e], e)
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a.double(), b.double(), dim=0).item()
    assert cos(ghf, ght) > 0.995 and cos(gbf, gbt) > 0.99, (cos(ghf, ght), cos(gbf, gbt))
    assert abs(ghf.norm().item() / ght.norm().item() - 1) < 3e-2
    assert abs(gbf.norm().item() / gbt.norm().item() - 1) < 5e-2
    for k in rvf:          # BatchNorm running statistics were updated by the same modules
        assert torch.allclose(rvf[k], rvt[k], rtol=5e-2, atol=1e-4), k
    net.eval()


def test_training_backbone_on_eeseg_convs_matches_torch_autograd(nets):
    """Backbone sections in training with the Bottleneck convolutions on the tcgen05 kernels (forward, dgrad,
    wgrad; bf16 activations, fp32 master weights and gradients) against the SAME PyTorch modules under bf16
    autocast (cuDNN): identical rounding points, so outputs and every gradient agree tightly. (Against fp32 a
    random-init train-mode ResNet amplifies bf16 rounding by ~1.15x per block — 1 % after the first block, 50 %
    after twenty, identically for cuDNN-bf16 and for these kernels — so fp32 is only checked on the loss.)"""
    import copy
    from ee_semantic_segmentation_b200 import backbone_train
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    port, net = nets
    g = torch.Generator().manual_seed(33)
    X = torch.randn(4, 3, 129, 129, generator=g).to(dev())
    # block by block (a whole section would only measure the amplification): every Bottleneck variant of the
    # backbone — 64-channel layer1 convs and the stride-2 block stay on cuDNN inside, dilations 1 / 2 / 4,
    # projection shortcuts
    from torchvision.models.resnet import Bottleneck
    blocks = [u for sec in net.base_model for u in sec if isinstance(u, Bottleneck)]
    assert len(blocks) == 16
    for bi in (0, 1, 3, 4, 7, 8, 13, 14):
        blk_a = copy.deepcopy(blocks[bi]).train()
        blk_b = copy.deepcopy(blocks[bi]).train()
        cin = blk_a.conv1.in_channels
        hw = 33 if bi < 4 else 17
        x0 = torch.randn(4, cin, hw, hw, generator=g).to(dev()).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        xa = x0.clone().requires_grad_(True)
        xb = x0.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            ya = backbone_train.bottleneck_forward_train(blk_a, xa)
            yb = blk_b(xb)
        assert ya.dtype == yb.dtype == torch.bfloat16 and ya.shape == yb.shape
        rel = ((ya.float() - yb.float()).norm() / yb.float().norm()).item()
        assert rel < 1e-2, (bi, rel)
        go = torch.randn(yb.shape, generator=g).to(dev()).to(torch.bfloat16)
        ya.backward(go)
        yb.backward(go)
        cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
        assert cos(xa.grad, xb.grad) > 0.999, (bi, cos(xa.grad, xb.grad))
        for (n, pa), (_, pb) in zip(blk_a.named_parameters(), blk_b.named_parameters()):
            assert cos(pa.grad, pb.grad) > 0.999, (bi, n, cos(pa.grad, pb.grad))
            assert abs(pa.grad.norm().item() / (pb.grad.norm().item() + 1e-12) - 1) < 2e-2, (bi, n)
        for (n, ba), (_, bb) in zip(blk_a.named_buffers(), blk_b.named_buffers()):     # BN running statistics
            if ba.dtype.is_floating_point:
                assert torch.allclose(ba, bb, rtol=2e-2, atol=1e-3), (bi, n)
    # whole model: loss next to the fp32 PyTorch-module step, every parameter gets a finite gradient
    y = torch.randint(0, 22, (4, 1, 129, 129), generator=g).to(dev())
    loss_fn = BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=3)
    losses = {}
    for fast in (True, False):
        m = copy.deepcopy(net).train()
        m.fast_training_heads = m.fast_training_backbone = fast
        torch.manual_seed(5)
        l = loss_fn(m(X), y)
        l.backward()
        losses[fast] = l.item()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert abs(losses[True] - losses[False]) < 2e-2 * abs(losses[False]), losses
    net.eval()


def test_graphed_train_step_matches_eager_steps(nets):
    """train_funcs.GraphedTrainStep (one CUDA graph per step) applies the same updates as the eager loop of
    train_epoch: two steps each on identical copies, Dropout off, same batches; the warm-up it runs before the
    capture is rolled back."""
    import copy
    from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
    from ee_semantic_segmentation_b200.train_funcs import GraphedTrainStep, make_optimizer, train_epoch
    port, net = nets
    g = torch.Generator().manual_seed(41)
    batches = [(torch.randn(2, 3, 97, 97, generator=g).to(dev()), torch.randint(0, 22, (2, 1, 97, 97), generator=g).to(dev()))
               for _ in range(2)]
    loss_fn = BrXEntropyLoss(ignore_index=21, b_reduction='sum', n_exits=3)
    nets2 = []
    for _ in range(2):
        m = copy.deepcopy(net).train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        nets2.append(m)
    na, nb = nets2
    opt_a = make_optimizer(na, lr=1e-2, base_lr=1e-3)
    l_eager = [train_epoch(na, [b], loss_fn, opt_a, dev()).item() for b in batches]
    opt_b = make_optimizer(nb, lr=1e-2, base_lr=1e-3)
    step = GraphedTrainStep(nb, loss_fn, opt_b, *batches[0])
    l_graph = [step(*b).item() for b in batches]
    assert l_graph == pytest.approx(l_eager, rel=2e-3)
    pa = torch.cat([p.detach().flatten() for p in na.parameters()])
    pb = torch.cat([p.detach().flatten() for p in nb.parameters()])
    assert ((pa - pb).norm() / pa.norm()).item() < 1e-4
    for (n, ba), (_, bb) in zip(na.named_buffers(), nb.named_buffers()):
        if ba.dtype.is_floating_point:
            assert torch.allclose(ba, bb, rtol=1e-2, atol=1e-3), n
        else:
            assert torch.equal(ba, bb), n          # num_batches_tracked: 2, the warm-up was rolled back
    # the data-parallel form on one rank (gradients accumulated in place into one flat buffer, no collective): same updates
    nc = copy.deepcopy(net).train()
    for mod in nc.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    opt_c = make_optimizer(nc, lr=1e-2, base_lr=1e-3)
    step_c = GraphedTrainStep(nc, loss_fn, opt_c, *batches[0], data_parallel=True)
    l_flat = [step_c(*b).item() for b in batches]
    assert l_flat == pytest.approx(l_graph, rel=2e-3)
    pc = torch.cat([p.detach().flatten() for p in nc.parameters()])
    assert ((pc - pb).norm() / pb.norm()).item() < 1e-4
    assert all(p.grad.data_ptr() >= step_c.flat.data_ptr() for p in nc.parameters())
    step_c.release()
    net.eval()


def test_cityscapes_shaped_full_res_sweep():
    """BASELINE config 5 shape: 19 classes, one 1024x2048 image (128x256 feature maps, 16x8 conv tiles),
    threshold sweep + integer confusion matrices; heads checked against the fp32 PyTorch modules."""
    from ee_semantic_segmentation_b200.engine import ThresholdSweep
    from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
    torch.manual_seed(5)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=[16, 3, 1], num_classes=19, pretrained=False).to(dev()).eval()
    g = torch.Generator().manual_seed(27)
    X = torch.randn(1, 3, 1024, 2048, generator=g).to(dev())
    y = torch.randint(0, 20, (1, 1, 1024, 2048), generator=g).to(dev())
    lows = net.forward_lowres(X)
    assert [tuple(l.shape) for l in lows] == [(1, 128, 256, 32)] * 3
    with torch.no_grad():                       # fp32 reference of the first head on the same features
        feat = net.run_section(0, X)
        ref = net.branches[0](feat.float())
    got = lows[0][..., :19].permute(0, 3, 1, 2)
    from conftest import assert_bf16_model_close
    assert_bf16_model_close(got, ref)
    sweep = ThresholdSweep(net, 19, [0.0, 0.5, 2.0])
    sweep.update(X, y)
    res = sweep.results()
    assert [r["out_gl"] for r in res] == [1, 1, 1]
    assert res[0]["count_out"] == 1 and res[2]["b1_count"] == 1
    assert int(sweep.cm[:, -1].sum()) == 3 * 1024 * 2048       # every pixel counted once per tau


def test_graph_replayed_forward_tracks_weights(nets):
    """forward_lowres replays a CUDA graph per input shape: same logits as the eager launches, re-captured after
    load_state_dict() / train() (weights_epoch), static outputs documented as overwritten by the next call."""
    import copy
    _, net0 = nets
    net = copy.deepcopy(net0)
    assert net._lowres_graphs == {} and net.graph_inference
    g = torch.Generator().manual_seed(41)
    x1 = torch.randn(1, 3, 65, 81, generator=g).to(dev())
    x2 = torch.randn(1, 3, 65, 81, generator=g).to(dev())
    with torch.no_grad():
        ya = net(x1).clone()                     # first sight of the shape: eager
        assert [k for k in net._lowres_graphs if k != 'seen'] == []
        yb = net(x2).clone()                     # second: captured
        assert torch.equal(net(x1), ya)          # replay with new contents
        net.graph_inference = False
        assert torch.equal(net(x1), ya) and torch.equal(net(x2), yb)
        net.graph_inference = True
        assert len([k for k in net._lowres_graphs if k != 'seen']) == 1
        sd = copy.deepcopy(net.state_dict())
        key = next(k for k in sd if k.endswith("classifier.4.weight") or k.endswith("4.weight"))
        sd[key] = sd[key] * 1.5
        e0 = net.weights_epoch
        net.load_state_dict(sd)
        assert net.weights_epoch > e0 and net._lowres_graphs == {}
        yc = net(x1).clone()
        assert torch.equal(net(x1), yc)          # captured again from the new weights
        net.graph_inference = False
        assert torch.equal(net(x1), yc) and not torch.equal(yc, ya)


def test_engine_graphs_follow_weight_updates(nets):
    """A graph-replaying engine kept across a training phase: train() bumps the model's weights_epoch, the engine drops
    the graphs captured from the old folded weights and evaluates the new ones (same results as a fresh eager engine)."""
    import copy
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    _, net0 = nets
    net = copy.deepcopy(net0).eval()
    g = torch.Generator().manual_seed(43)
    X = torch.randn(2, 3, 65, 81, generator=g).to(dev())
    y = torch.randint(0, 22, (2, 1, 65, 81), generator=g).to(dev())
    for kw in (dict(use_graph=True), dict(use_graph=True, skip_compute=True)):
        eng = EarlyExitEngine(net, 21, 0.97, **kw)
        before = {k: v.clone() for k, v in eng.evaluate(X, y).items()}
        net.train()
        with torch.no_grad():
            for p in net.classifier[-1].parameters():
                p.add_(torch.randn_like(p) * 0.5)
            for p in net.branches[0][-1].parameters():
                p.add_(torch.randn_like(p) * 0.5)
        net.eval()
        after = eng.evaluate(X, y)
        ref = EarlyExitEngine(net, 21, 0.97, skip_compute=kw.get("skip_compute", False)).evaluate(X, y)
        assert torch.equal(after["pred"], ref["pred"]) and torch.equal(after["exit"], ref["exit"])
        assert not torch.equal(after["pred"], before["pred"])
        assert len(eng._graphs) <= 2


def test_similarity_operator_keys_and_rule(nets):
    """ee_dnn_op.eval_ee_deeplabv3 (similarity twin, ee_dnn_op.py:40-118): the first exit is only the reference image,
    an image leaves at the first later exit whose map is close enough to the previous one; result keys of the reference
    including the *_flops_2 variants."""
    from ee_semantic_segmentation_b200.ee_dnn_op import eval_ee_deeplabv3 as SimOp
    _, net = nets
    g = torch.Generator().manual_seed(51)
    x = torch.randn(3, 65, 81, generator=g).to(dev())
    diff = lambda a, b: float((a != b).float().mean())
    out = SimOp(net, diff, 2.0, device=dev())(x)                    # always "similar": leaves at exit 2, never at exit 1
    assert out["n"] == 2 and out["exit"].shape == (65, 81) and out["exit"].dtype == torch.int64
    assert set(out) == {"exit", "exit_flops", "exit_flops_2", "edge_flops", "edge_flops_2", "n", "last", "last_flops",
                        "last_flops_2"}
    assert out["exit_flops_2"] < out["exit_flops"] < out["last_flops"] and out["edge_flops_2"] == out["exit_flops_2"]
    out = SimOp(net, diff, -1.0, device=dev())(x)                   # never similar
    assert out["n"] == 3 and torch.equal(out["exit"], out["last"]) and out["exit_flops_2"] == out["last_flops_2"]
    assert "edge_flops_2" in out
    with torch.no_grad():
        ref = net(x.unsqueeze(0))[-1, 0].argmax(0).cpu()
    assert (out["last"] == ref).float().mean().item() > 0.999


def test_engine_input_formats_bf16_and_uint8(nets):
    """The streaming engine's input formats: bf16 images + uint8 labels give bit-identical exits, maps and confusion
    matrices to fp32 images + int64 labels (the stem rounds to bf16 first; labels are widened in the accumulate kernel);
    uint8 images are normalised inside the stem kernel like ToTensor + Normalize (get_seg_datasets.py:62-70)."""
    from ee_semantic_segmentation_b200.engine import EarlyExitEngine
    _, net = nets
    g = torch.Generator().manual_seed(41)
    N, H, W = 3, 97, 113
    X = torch.randn(N, 3, H, W, generator=g)
    y = torch.randint(0, 22, (N, 1, H, W), generator=g)
    y[0, 0, :5] = 255                                              # VOC's void value: >= C is void in either dtype
    y64 = torch.where(y == 255, torch.full_like(y, 21), y)
    sc = EarlyExitEngine(net, 21, 0.5).evaluate(X.to(dev()), y64.to(dev()))["scores"][0].float().cpu()
    tau = float(sc.median())
    for skip_compute in (False, True):
        a = EarlyExitEngine(net, 21, tau, skip_compute=skip_compute, use_graph=True)
        b = EarlyExitEngine(net, 21, tau, skip_compute=skip_compute, use_graph=True, input_dtype=torch.bfloat16,
                            target_dtype=torch.uint8)
        for _ in range(2):
            ra = a.evaluate(X.to(dev()), y64.to(dev()))
            rb = b.evaluate(X.to(torch.bfloat16).to(dev()), y.to(torch.uint8).to(dev()))
        assert torch.equal(ra["exit"], rb["exit"]) and torch.equal(ra["pred"], rb["pred"])
        assert torch.equal(a.cm, b.cm) and torch.equal(a.counts, b.counts)
    # uint8 pixels with the ImageNet normalisation inside the kernel == normalised fp32 images
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    U = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    Xn = (U.float() / 255 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    net.input_norm = (mean, std)
    try:
        lo_u8 = [t.clone() for t in net._lowres_eager(U.to(dev()))]
        lo_f = net._lowres_eager(Xn.to(dev()))
    finally:
        net.input_norm = None
    for p, q in zip(lo_u8, lo_f):
        # the two paths round slightly different fp32 values to bf16 (u * a + b vs (u / 255 - m) / s): one bf16 ulp at the input
        assert (p - q).abs().max().item() < 2e-2 * q.abs().max().item()
