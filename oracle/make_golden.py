"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference modules
from /root/reference (through oracle/ref_import.py's stub set) on seeded inputs.

Run in the build container only:  python -m oracle.make_golden
The fixtures are small (< 1 MB in total) and committed; the GPU box has no /root/reference.

Fixtures
  demo_fixtures.npz   inputs of the reference's own __main__ demo blocks (extracted with runpy, not
                      copied) + the values those blocks print (compute_mIoU.py:65-149,
                      seg_metrics.py:78-173)
  metrics.npz         _compute_basics / mIoU / img_mIoU on seeded logits with void + out-of-range labels
  entropy.npz         img_norm_entropy (ent / max-pool / min-pool) on seeded probabilities
  ce.npz              BrXEntropyLoss value + autograd gradient (sum / mean / none, weights, n_exits=0)
  lovasz.npz          lovasz_softmax / BSL.LovaszSoftmax value + gradient (present / all / per_image / prev_out)
  upsample.npz        F.interpolate bilinear align_corners=False, 9x7 -> 65x49 and 65x65 -> 129x129 slice
  br_eval.npz         br_evaluator result dict on a fake 3-exit net for several tau / pool modes
  model.npz           branchyDeepv3 (ResNet-50, seed 0, n=1 and n=2) section split + strided slice of
                      net(x) at 65x65
"""
import contextlib
import io
import os
import runpy
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import model_port, ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def blocky_labels(g, N, C, H, W, void_frac=0.05, cell=4):
    """SURVEY.md §8(d) synthetic targets: blocky randint at 1/cell resolution, nearest-upsampled,
    then void (= C) with probability void_frac."""
    low = torch.randint(0, C, (N, 1, (H + cell - 1) // cell, (W + cell - 1) // cell), generator=g)
    lab = F.interpolate(low.float(), size=(H, W), mode="nearest").long()
    void = torch.rand(N, 1, H, W, generator=g) < void_frac
    return torch.where(void, torch.full_like(lab, C), lab)


def main():
    os.makedirs(OUT, exist_ok=True)
    (cm_mod, sm_mod, xe_mod, bsl_mod, lov_mod, ebe_mod, fd_new, fd_old) = ref_import.load(
        "compute_mIoU", "seg_metrics", "my_pixelwise_xentropy", "branchy_seg_losses",
        "lovaszsoftmax", "eval_br_ent", "from_deepv3_new", "from_deepv3")

    # ---- demo fixtures: run the reference's own __main__ blocks and keep their tensors ----------
    demo = {}
    with ref_import.reference_on_path():
        for name in ("compute_mIoU", "seg_metrics"):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                gl = runpy.run_path(os.path.join(ref_import.REFERENCE_DIR, name + ".py"),
                                    run_name="__main__")
            demo[name + "_y_true"] = gl["y_true"].numpy()
            demo[name + "_y_pred"] = gl["y_pred"].numpy()
            demo[name + "_stdout"] = np.array(buf.getvalue())
    yt, yp = torch.tensor(demo["compute_mIoU_y_true"]), torch.tensor(demo["compute_mIoU_y_pred"])
    ev = cm_mod.mIoU(n_classes=4); ev(yp, yt)
    demo["compute_mIoU_value"] = ev.compute().numpy()
    ev2 = cm_mod.img_mIoU(); ev2(yp, yt)
    demo["compute_mIoU_img_value"] = np.float64(ev2.compute())
    yt, yp = torch.tensor(demo["seg_metrics_y_true"]), torch.tensor(demo["seg_metrics_y_pred"])
    demo["seg_metrics_acc"] = sm_mod.Accuracy(reduction=None)(yp, yt).numpy()
    for avg in ("macro", "micro"):
        demo[f"seg_metrics_recall_{avg}"] = sm_mod.Recall(avg=avg)(yp, yt).numpy()
        demo[f"seg_metrics_precision_{avg}"] = sm_mod.Precision(avg=avg)(yp, yt).numpy()
        demo[f"seg_metrics_f1_{avg}"] = sm_mod.F_beta(avg=avg)(yp, yt).numpy()
    tp, fp, fn = sm_mod.SegMetric()._compute_basics(yp, yt)
    demo["seg_metrics_tp"], demo["seg_metrics_fp"], demo["seg_metrics_fn"] = tp.numpy(), fp.numpy(), fn.numpy()
    np.savez_compressed(os.path.join(OUT, "demo_fixtures.npz"), **demo)

    # ---- metrics ---------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    N, C, H, W = 3, 21, 33, 47
    logits = torch.randn(N, C, H, W, generator=g) * 3
    tgt = blocky_labels(g, N, C, H, W)
    tgt[0, 0, 0, :5] = 30  # out-of-range label, like seg_metrics.py:93
    tp, fp, fn = sm_mod.SegMetric()._compute_basics(logits, tgt)
    m = cm_mod.mIoU(C); m(logits, tgt); m(logits.flip(0), tgt)
    im = cm_mod.img_mIoU(); im(logits[:1], tgt[:1]); im(logits[1:2], tgt[1:2])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), logits=logits.numpy(), targets=tgt.numpy(),
                        tp=tp.numpy(), fp=fp.numpy(), fn=fn.numpy(), miou=m.compute().numpy(),
                        acc=m.accumulator.numpy(), img_miou=np.float64(im.compute()))

    # ---- entropy ---------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1235)
    C, H, W = 21, 37, 53
    lg = torch.randn(C, H, W, generator=g) * 4
    lg[:, :5] *= 20  # near one-hot rows: exercises entr(0)=0 underflow
    probs = F.softmax(lg, 0)
    ent = {"logits": lg.numpy(), "probs": probs.numpy()}
    ent["ent"] = np.float32(ebe_mod.img_norm_entropy(C)(probs))
    for s in (2, 4, 5):
        ent[f"max_{s}"] = np.float32(ebe_mod.img_norm_entropy(C, s=s)(probs))
        ent[f"min_{s}"] = np.float32(ebe_mod.img_norm_entropy(C, s=s, pool_min=True)(probs))
    np.savez_compressed(os.path.join(OUT, "entropy.npz"), **ent)

    # ---- multi-exit CE ---------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1236)
    E, N, C, H, W = 3, 2, 21, 19, 23
    y = (torch.randn(E, N, C, H, W, generator=g) * 3)
    tgt = blocky_labels(g, N, C, H, W)
    ce = {"y_pred": y.numpy(), "targets": tgt.numpy()}
    for tag, kw in {
        "sum": dict(ignore_index=21, b_reduction="sum", n_exits=3),
        "mean": dict(ignore_index=21, b_reduction="mean", n_exits=3),
        "none": dict(ignore_index=21, b_reduction="none", n_exits=3),
        "wsum": dict(ignore_index=21, b_reduction="sum", n_exits=3, weights=[0.25, 0.5, 1.0]),
        "two": dict(ignore_index=21, b_reduction="sum", n_exits=2),
        "noign": dict(b_reduction="mean", n_exits=3),
    }.items():
        yy = y.clone().requires_grad_(True)
        t_in = tgt.clamp(max=20) if tag == "noign" else tgt
        out = xe_mod.BrXEntropyLoss(**kw)(yy, t_in)
        out.sum().backward()
        ce[f"{tag}_loss"], ce[f"{tag}_grad"] = out.detach().numpy(), yy.grad.numpy()
    yy = y[0].clone().requires_grad_(True)
    out = xe_mod.BrXEntropyLoss(ignore_index=21)(yy, tgt); out.backward()
    ce["single_loss"], ce["single_grad"] = out.detach().numpy(), yy.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ce.npz"), **ce)

    # ---- Lovasz ----------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1237)
    E, N, C, H, W = 3, 2, 19, 17, 21
    y = torch.randn(E, N, C, H, W, generator=g) * 2  # raw logits, as on the reference path
    tgt = blocky_labels(g, N, C - 3, H, W, cell=3)   # classes 16..18 absent -> 'present' skips them
    tgt = torch.where(tgt == C - 3, torch.full_like(tgt, C), tgt)  # void == 19
    lv = {"y_pred": y.numpy(), "targets": tgt.numpy()}
    for tag, kw in {
        "present": dict(classes="present", ignore=19, n_branches=2),
        "all": dict(classes="all", ignore=19, n_branches=2),
        "per_image": dict(classes="present", per_image=True, ignore=19, n_branches=2),
        "prev_out": dict(classes="present", ignore=19, n_branches=2, prev_out=True),
        "noignore": dict(classes="present", ignore=None, n_branches=1),
    }.items():
        yy = y.clone().requires_grad_(True)
        out = bsl_mod.LovaszSoftmax(**kw)(yy, tgt)
        out.backward()
        lv[f"{tag}_loss"], lv[f"{tag}_grad"] = out.detach().numpy(), yy.grad.numpy()
    pr = F.softmax(y[0], 1).clone().requires_grad_(True)
    out = lov_mod.lovasz_softmax(pr, tgt, classes="present", ignore=19); out.backward()
    lv["probas"], lv["probas_loss"], lv["probas_grad"] = pr.detach().numpy(), out.detach().numpy(), pr.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "lovasz.npz"), **lv)

    # ---- bilinear upsample -----------------------------------------------------------------------
    g = torch.Generator().manual_seed(1238)
    a = torch.randn(2, 5, 9, 7, generator=g)
    b = torch.randn(1, 3, 65, 65, generator=g)
    np.savez_compressed(
        os.path.join(OUT, "upsample.npz"), a=a.numpy(),
        a_up=F.interpolate(a, size=(65, 49), mode="bilinear", align_corners=False).numpy(),
        b=b.numpy(),
        b_up=F.interpolate(b, size=(513, 513), mode="bilinear", align_corners=False)[..., ::7, ::5].numpy())

    # ---- br_evaluator on a fake net --------------------------------------------------------------
    g = torch.Generator().manual_seed(1239)
    n_img, E, C, H, W = 6, 3, 21, 24, 30
    sharp = torch.tensor([[0.3, 1.0, 4.0], [3.0, 0.5, 4.0], [0.2, 0.4, 5.0], [6.0, 6.0, 6.0],
                          [0.1, 5.0, 5.0], [1.5, 2.5, 0.5]])
    ys = torch.randn(n_img, E, 1, C, H, W, generator=g) * sharp[:, :, None, None, None, None]
    tg = torch.stack([blocky_labels(g, 1, C, H, W) for _ in range(n_img)])

    class FakeNet:
        def __init__(self): self.k = 0
        def __call__(self, X):
            out = ys[self.k]; self.k += 1
            return out
    loader = [(torch.zeros(1, 3, H, W), tg[k]) for k in range(n_img)]
    be = {"y": ys.numpy(), "targets": tg.numpy()}
    cfgs = []
    for tau in (0.2, 0.6, 0.8, 0.95):
        for metric, size in (("ent", 1), ("max", 4), ("min", 5)):
            res = ebe_mod.br_evaluator(FakeNet(), E, C, loader, torch.device("cpu"), tau,
                                       metric=metric, size=size)
            key = f"tau{tau}_{metric}{size}"
            cfgs.append(key)
            for k, v in res.items():
                if k not in ("pool",):
                    be[f"{key}/{k}"] = np.float64(v)
    be["configs"] = np.array(cfgs)
    np.savez_compressed(os.path.join(OUT, "br_eval.npz"), **be)

    # ---- model -----------------------------------------------------------------------------------
    import torchvision
    md = {}
    base_path = "/tmp/eeseg_oracle_base_r50.pth"
    torch.manual_seed(0)
    base = torchvision.models.segmentation.deeplabv3_resnet50(
        weights=None, weights_backbone=None, num_classes=21, aux_loss=True)
    torch.save(base, base_path)
    g = torch.Generator().manual_seed(1240)
    x = torch.randn(2, 3, 65, 65, generator=g)
    md["x"] = x.numpy()
    for n in (1, 2):
        net = fd_new.branchyDeepv3(base_path, "deeplabv3_resnet50", n, 513)
        model_port.reinit_branches(net.branches, 100 + n)  # see model_port.reinit_branches
        net.eval()
        with torch.no_grad():
            y = net(x)
        md[f"n{n}_sections"] = np.array([len(s) for s in net.base_model])
        md[f"n{n}_cin"] = np.array([b[0].convs[0][0].in_channels for b in net.branches])
        md[f"n{n}_out_shape"] = np.array(y.shape)
        md[f"n{n}_out_slice"] = y[..., ::8, ::8].numpy()
    np.savez_compressed(os.path.join(OUT, "model.npz"), **md)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
