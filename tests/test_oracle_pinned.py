"""Pins the oracle (oracle/restate.py) against (1) the reference's own known-answer demo values and
(2) fixtures produced by the unmodified reference modules (oracle/make_golden.py). CPU only."""
import numpy as np
import pytest

from oracle import restate as R


def test_demo_miou_known_answer(golden):
    d = golden("demo_fixtures")
    # compute_mIoU.py:139-149 prints 0.9513888955116272 for both evaluators
    assert float(d["compute_mIoU_value"]) == pytest.approx(0.9513888955116272, abs=0)
    m = R.MIoU(4)
    m(d["compute_mIoU_y_pred"], d["compute_mIoU_y_true"])
    assert m.compute() == np.float32(0.9513888955116272)
    vals = [R.img_miou(d["compute_mIoU_y_pred"][i:i + 1], d["compute_mIoU_y_true"][i]) for i in range(2)]
    # img_mIoU is fed the whole batch at once in the demo (squeeze keeps N=2): restate that call
    assert R.img_miou(d["compute_mIoU_y_pred"], d["compute_mIoU_y_true"]) == pytest.approx(
        float(d["compute_mIoU_img_value"]), abs=1e-7)
    assert len(vals) == 2


def test_demo_seg_metrics_known_answer(golden):
    d = golden("demo_fixtures")
    tp, fp, fn = R.compute_basics(d["seg_metrics_y_pred"], d["seg_metrics_y_true"])
    np.testing.assert_array_equal(tp, d["seg_metrics_tp"])
    np.testing.assert_array_equal(fp, d["seg_metrics_fp"])
    np.testing.assert_array_equal(fn, d["seg_metrics_fn"])
    # printed digits in seg_metrics.py's demo: macro R/P/F1 0.8500/0.9667/0.8419
    s = 1e-6
    rec = ((tp + s) / (tp + fn + s)).mean(-1).mean()
    pre = ((tp + s) / (tp + fp + s)).mean(-1).mean()
    f1 = ((2 * tp + s) / (2 * tp + fn + fp + s)).mean(-1).mean()
    assert round(float(rec), 4) == 0.8500 and round(float(pre), 4) == 0.9667 and round(float(f1), 4) == 0.8419
    assert float(d["seg_metrics_recall_macro"]) == pytest.approx(rec, abs=1e-6)


def test_metrics_vs_reference(golden):
    d = golden("metrics")
    tp, fp, fn = R.compute_basics(d["logits"], d["targets"])
    np.testing.assert_array_equal(tp, d["tp"])
    np.testing.assert_array_equal(fp, d["fp"])
    np.testing.assert_array_equal(fn, d["fn"])
    m = R.MIoU(21)
    m(d["logits"], d["targets"])
    m(d["logits"][::-1], d["targets"])
    np.testing.assert_array_equal(m.acc, d["acc"])
    assert m.compute() == d["miou"] or (np.isnan(m.compute()) and np.isnan(d["miou"]))
    im = np.mean([R.img_miou(d["logits"][i:i + 1], d["targets"][i:i + 1]) for i in range(2)])
    assert im == pytest.approx(float(d["img_miou"]), abs=1e-6)


def test_entropy_vs_reference(golden):
    d = golden("entropy")
    assert R.img_norm_entropy(d["probs"], 21) == pytest.approx(float(d["ent"]), abs=2e-6)
    for s in (2, 4, 5):
        assert R.img_norm_entropy(d["probs"], 21, s=s) == pytest.approx(float(d[f"max_{s}"]), abs=2e-6)
        assert R.img_norm_entropy(d["probs"], 21, pool_min=True, s=s) == pytest.approx(
            float(d[f"min_{s}"]), abs=2e-6)
    # from logits: softmax then entropy
    assert R.img_norm_entropy(R.softmax_c(d["logits"], 0), 21) == pytest.approx(float(d["ent"]), abs=2e-6)


CE_CASES = {
    "sum": dict(ignore_index=21, b_reduction="sum", n_exits=3),
    "mean": dict(ignore_index=21, b_reduction="mean", n_exits=3),
    "none": dict(ignore_index=21, b_reduction="none", n_exits=3),
    "wsum": dict(ignore_index=21, b_reduction="sum", n_exits=3, weights=[0.25, 0.5, 1.0]),
    "two": dict(ignore_index=21, b_reduction="sum", n_exits=2),
}


@pytest.mark.parametrize("tag", list(CE_CASES))
def test_ce_vs_reference(golden, tag):
    d = golden("ce")
    loss, grad, _ = R.br_xentropy(d["y_pred"], d["targets"], **CE_CASES[tag])
    np.testing.assert_allclose(loss, d[f"{tag}_loss"], rtol=1e-5)
    np.testing.assert_allclose(grad, d[f"{tag}_grad"], rtol=1e-4, atol=1e-9)


def test_ce_single_and_noignore(golden):
    d = golden("ce")
    loss, grad, _ = R.br_xentropy(d["y_pred"][0], d["targets"], ignore_index=21)
    np.testing.assert_allclose(loss, d["single_loss"], rtol=1e-5)
    np.testing.assert_allclose(grad, d["single_grad"], rtol=1e-4, atol=1e-9)
    loss, grad, _ = R.br_xentropy(d["y_pred"], np.minimum(d["targets"], 20), b_reduction="mean", n_exits=3)
    np.testing.assert_allclose(loss, d["noign_loss"], rtol=1e-5)
    np.testing.assert_allclose(grad, d["noign_grad"], rtol=1e-4, atol=1e-9)


LOV_CASES = {
    "present": dict(classes="present", ignore=19, n_branches=2),
    "all": dict(classes="all", ignore=19, n_branches=2),
    "per_image": dict(classes="present", per_image=True, ignore=19, n_branches=2),
    "prev_out": dict(classes="present", ignore=19, n_branches=2, prev_out=True),
    "noignore": dict(classes="present", ignore=None, n_branches=1),
}


@pytest.mark.parametrize("tag", list(LOV_CASES))
def test_lovasz_vs_reference(golden, tag):
    d = golden("lovasz")
    loss, grad, _ = R.br_lovasz(d["y_pred"], d["targets"], **LOV_CASES[tag])
    np.testing.assert_allclose(loss, d[f"{tag}_loss"], rtol=2e-5)
    g_ref = d[f"{tag}_grad"]
    np.testing.assert_allclose(grad[: g_ref.shape[0]], g_ref, rtol=1e-3, atol=1e-7)


def test_lovasz_probas_vs_reference(golden):
    d = golden("lovasz")
    loss, grad = R.lovasz_softmax(d["probas"], d["targets"], classes="present", ignore=19)
    np.testing.assert_allclose(loss, d["probas_loss"], rtol=2e-5)
    np.testing.assert_allclose(grad, d["probas_grad"], rtol=1e-3, atol=1e-7)


def test_upsample_vs_reference(golden):
    d = golden("upsample")
    np.testing.assert_allclose(R.bilinear_upsample(d["a"], (65, 49)), d["a_up"], atol=3e-6)
    np.testing.assert_allclose(R.bilinear_upsample(d["b"], (513, 513))[..., ::7, ::5], d["b_up"], atol=2e-5)


def test_br_evaluator_vs_reference(golden):
    d = golden("br_eval")
    y, tg = d["y"], d["targets"]
    n_img, E = y.shape[:2]
    C = y.shape[3]
    for key in d["configs"]:
        key = str(key)
        tau = float(d[f"{key}/t"])
        size = int(d[f"{key}/pool_size"])
        mode = "ent" if "_ent" in key else ("max" if "_max" in key else "min")
        acc = [R.MIoU(C) for _ in range(E + 1)]
        cnt = [0] * (E + 1)
        for k in range(n_img):
            ents = [R.img_norm_entropy(R.softmax_c(y[k, i, 0], 0), C, pool_min=(mode == "min"),
                                       s=size if mode != "ent" else 1) for i in range(E - 1)]
            ex = R.first_confident_exit(ents, tau)
            slot = ex if ex < E - 1 else E - 1  # accumulator[-2] is the final exit
            acc[slot](y[k, ex], tg[k]); acc[-1](y[k, ex], tg[k])
            cnt[slot] += 1; cnt[-1] += 1
        for i in range(E - 1):
            assert cnt[i] == int(d[f"{key}/b{i+1}_count"]), key
            a, b = float(acc[i].compute()), float(d[f"{key}/b{i+1}_mIoU"])
            assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, abs=1e-6), key
        assert cnt[-2] == int(d[f"{key}/count_out"]) and cnt[-1] == int(d[f"{key}/out_gl"])
        a, b = float(acc[-1].compute()), float(d[f"{key}/mIoU_gl"])
        assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, abs=1e-6), key


def test_overlap_losses_oracle_vs_reference_golden(golden):
    """Dice / Jaccard / Tversky / FocalTversky restatements against the unmodified reference's values and autograd
    gradients (tests/golden/overlap_losses.npz, oracle/make_golden_overlap.py)."""
    G = golden("overlap_losses")
    y, t, tv = G["y_pred"], G["targets"], G["targets_void"]
    E = y.shape[0]

    def stack(fn, tgt, n, **kw):
        ls, gs = zip(*(fn(y[e], tgt, **kw) for e in range(n)))
        return np.stack(ls), np.stack(gs)
    l, g = stack(R.dice_loss, t, E)
    np.testing.assert_allclose(l, G["dice_none"], rtol=1e-5)
    np.testing.assert_allclose(R.br_seg_loss(l), G["dice_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l.shape[1], G["dice_mean_grad"], rtol=1e-4, atol=1e-9)
    w = np.array([0.5, 1.0, 2.0], np.float32)
    np.testing.assert_allclose(R.br_seg_loss(l, "sum", w), G["dice_sum_w_loss"], rtol=1e-5)
    np.testing.assert_allclose(g * w[:, None, None, None, None], G["dice_sum_w_grad"], rtol=1e-4, atol=1e-9)
    l, g = stack(R.jaccard_loss, t, E)
    np.testing.assert_allclose(R.br_seg_loss(l), G["jaccard_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / (l.shape[1] * l.shape[2]), G["jaccard_mean_grad"], rtol=1e-4, atol=1e-9)
    l, g = stack(R.jaccard_loss, tv, E, downgrad_bg=0.3)
    np.testing.assert_allclose(R.br_seg_loss(l), G["jaccard_void_bg_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / (l.shape[1] * l.shape[2]), G["jaccard_void_bg_grad"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(stack(R.jaccard_loss, tv, E)[0], G["jaccard_none"], rtol=1e-5)
    l, g = stack(R.jaccard_loss, tv, 2, downgrad_bg=0.0)
    np.testing.assert_allclose(R.br_seg_loss(l, "sum"), G["jaccard_nobg_sum_loss"], rtol=1e-5)
    np.testing.assert_allclose(g, G["jaccard_nobg_sum_grad"][:2], rtol=1e-4, atol=1e-9)
    assert np.all(G["jaccard_nobg_sum_grad"][2] == 0)
    tl = np.stack([R.tversky_loss(y[e], t) for e in range(E)])
    np.testing.assert_allclose(tl, G["tversky_none"], rtol=1e-6)
    tl = np.stack([R.tversky_loss(y[e], t, alpha=0.3, beta=0.7) for e in range(E)])
    np.testing.assert_allclose(R.br_seg_loss(tl), G["tversky_mean"], rtol=1e-5)
    tl = np.stack([R.tversky_loss(y[e], t, gamma=0.75) for e in range(E)])
    np.testing.assert_allclose(R.br_seg_loss(tl), G["focal_tversky_mean"], rtol=1e-5)
    with pytest.raises(RuntimeError):
        R.dice_loss(y[0], tv)        # void label: F.one_hot(num_classes=C) raises in the reference


def test_focal_loss_oracle_vs_reference_golden(golden):
    """FocalLoss restatement (incl. the reference's [N,N,H,W] alpha broadcast) against the unmodified reference's values
    and autograd gradients (tests/golden/focal_loss.npz, oracle/make_golden_focal.py)."""
    G = golden("focal_loss")
    y, t, alpha = G["y_pred"], G["targets"], G["alpha"]

    def run(n, yy=y, tt=t, **kw):
        ls, gs = zip(*(R.focal_loss(yy[e], tt, **kw) for e in range(n)))
        return np.stack(ls), np.stack(gs)
    l, g = run(3)
    np.testing.assert_allclose(l, G["g2_none"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(R.br_seg_loss(l), G["g2_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l[0].size, G["g2_mean_grad"], rtol=1e-4, atol=1e-8)
    w = np.array([0.5, 1.0, 2.0], np.float32)
    l, g = run(3, gamma=1.5)
    np.testing.assert_allclose(R.br_seg_loss(l, "sum", w), G["g15_sum_w_loss"], rtol=1e-5)
    np.testing.assert_allclose(g * w[:, None, None, None, None], G["g15_sum_w_grad"], rtol=1e-4, atol=1e-6)
    l, g = run(2, gamma=0)
    np.testing.assert_allclose(R.br_seg_loss(l), G["g0_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l[0].size, G["g0_mean_grad"][:2], rtol=1e-4, atol=1e-8)
    assert np.all(G["g0_mean_grad"][2] == 0)
    l, g = run(3, gamma=0.5)
    np.testing.assert_allclose(R.br_seg_loss(l), G["g05_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l[0].size, G["g05_mean_grad"], rtol=1e-4, atol=1e-8)
    l, g = run(3, alpha=alpha)
    assert l.shape == G["alpha_none"].shape == (3, 2, 2, 13, 17)
    np.testing.assert_allclose(l, G["alpha_none"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(R.br_seg_loss(l), G["alpha_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l[0].size, G["alpha_mean_grad"], rtol=1e-4, atol=1e-8)
    l, g = run(3, alpha=alpha, gamma=1)
    np.testing.assert_allclose(R.br_seg_loss(l, "sum"), G["alpha_sum_loss"], rtol=1e-5)
    np.testing.assert_allclose(g, G["alpha_sum_grad"], rtol=1e-4, atol=1e-6)
    l, g = run(3, yy=y[:, :1], tt=t[:1], alpha=alpha)
    np.testing.assert_allclose(R.br_seg_loss(l), G["alpha_n1_mean_loss"], rtol=1e-5)
    np.testing.assert_allclose(g / l[0].size, G["alpha_n1_mean_grad"], rtol=1e-4, atol=1e-8)
    with pytest.raises(RuntimeError):
        R.focal_loss(y[0], t[:, 0])          # [N,H,W] targets: gather() needs [N,1,H,W]
    with pytest.raises(RuntimeError):
        R.focal_loss(y[0], np.full_like(t, 7))


SEG_LOSS_CASES = {
    "dice_mean": ("dice", dict(), "mean", "targets_void"),
    "dice_index_sum": ("dice", dict(index=True), "sum", "targets_void"),
    "dice_batchwise": ("dice", dict(), "mean_batchwise", "targets"),
    "jaccard_mean": ("jaccard", dict(), "mean", "targets_void"),
    "jaccard_bg": ("jaccard", dict(downgrad_bg=0.25), "sum", "targets_void"),
    "jaccard_nobg": ("jaccard", dict(downgrad_bg=0.0), "sum_batchwise", "targets_void"),
    "jaccard_index": ("jaccard", dict(index=True), "none", "targets"),
    "tversky_mean": ("tversky", dict(alpha=0.3, beta=0.7), "mean", "targets"),
    "ftversky_sum": ("tversky", dict(alpha=0.7, beta=0.3, gamma=4 / 3), "sum", "targets"),
}


def _seg_reduce(l, reduction):
    """SegLoss.forward (new_seg_losses.py:17-32); mean over an empty dim list reduces everything, as torch does."""
    if reduction == "mean":
        return l.mean(), 1.0 / l.size
    if reduction == "sum":
        return l.sum(), 1.0
    if reduction in ("mean_batchwise", "sum_batchwise"):
        dims = tuple(range(1, l.ndim))
        if reduction == "mean_batchwise":
            return (l.mean(axis=dims), 1.0 / np.prod(l.shape[1:])) if dims else (l.mean(), 1.0 / l.size)
        return (l.sum(axis=dims), 1.0) if dims else (l.sum(), 1.0)
    return l, 1.0


def test_seg_losses_oracle_vs_reference_golden(golden):
    """new_seg_losses.py Dice / Jaccard / Tversky / FocalTversky restatement against (1) the values the reference's own
    __main__ demo prints (0.0504 / 0.4033, new_seg_losses.py:170-256) and (2) the unmodified reference's values and
    autograd gradients on seeded inputs (tests/golden/seg_losses.npz, oracle/make_golden_seg_losses.py)."""
    G = golden("seg_losses")
    out = str(G["demo_stdout"])
    assert "0.0504" in out and "0.4033" in out
    l, _ = R.seg_overlap_loss(G["demo_y_pred"], G["demo_y_true"], "jaccard")
    assert f"{l.mean():.4f}" == "0.0504" and f"{l.sum():.4f}" == "0.4033"
    np.testing.assert_allclose(l.mean(), G["demo_jaccard_mean"], rtol=1e-5)
    np.testing.assert_allclose(R.seg_overlap_loss(G["demo_y_pred"], G["demo_y_true"], "dice")[0].mean(), G["demo_dice_mean"],
                               rtol=1e-5)
    for tag, (kind, kw, red, tk) in SEG_LOSS_CASES.items():
        l, g = R.seg_overlap_loss(G["y_pred"], G[tk], kind, **kw)
        val, gscale = _seg_reduce(l, red)
        np.testing.assert_allclose(val, G[f"{tag}_loss"], rtol=1e-5, err_msg=tag)
        ref = G[f"{tag}_grad"]
        assert np.abs(g * gscale - ref).max() < 1e-4 * np.abs(ref).max() + 1e-9, tag
    with pytest.raises(RuntimeError):
        R.seg_overlap_loss(G["y_pred"], G["targets_void"], "tversky")
