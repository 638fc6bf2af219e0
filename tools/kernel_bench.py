#!/usr/bin/env python
"""Per-kernel roofline report for the HBM-bound kernels of the hot path:
exit gate, multi-exit CE fwd+bwd, confusion histogram, Lovasz, plus the conv kernel per layer shape.
Each row: algorithmic bytes (or FLOPs) per launch, CUDA-event time (L2 flushed between iterations),
achieved GB/s (TFLOP/s) and the fraction of the measured peak (MEASURED_PEAKS.json), next to the
PyTorch-eager implementation of the same op on the same GPU (what the reference runs on a GPU).

    python tools/kernel_bench.py [--iters 20] [--out profiles/r01_kernels.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import measured_peaks  # noqa: E402
from ee_semantic_segmentation_b200 import _lib, ops  # noqa: E402
from ee_semantic_segmentation_b200.head_plan import conv_igemm  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return sum(ts[: max(1, len(ts) // 2)]) / max(1, len(ts) // 2)   # mean of the faster half (ms)


def blocky(N, C, H, W, dev):
    g = torch.Generator().manual_seed(1234)
    low = torch.randint(0, C, (N, 1, (H + 15) // 16, (W + 15) // 16), generator=g)
    y = F.interpolate(low.float(), size=(H, W), mode="nearest").long()
    void = torch.rand(N, 1, H, W, generator=g) < 0.05
    return torch.where(void, torch.full_like(y, C), y).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = measured_peaks()
    hbm, tf = peaks["hbm_gbs"], peaks["bf16_tflops"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def row(name, ms, nbytes=None, flops=None, eager_ms=None, note=""):
        r = {"kernel": name, "ms": ms, "note": note}
        if nbytes is not None:
            r.update(bytes=nbytes, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / hbm)
        if flops is not None:
            r.update(flops=flops, tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / tf)
        if eager_ms is not None:
            r.update(eager_ms=eager_ms, speedup_vs_eager=eager_ms / ms)
        rows.append(r)
        print(json.dumps(r), flush=True)

    want = lambda k: (not args.only) or args.only in k
    N, C, H, W, h, w = 4, 21, 513, 513, 65, 65

    # ---------------- exit gate ----------------
    if want("gate"):
        for dt, e in ((torch.float32, 4), (torch.bfloat16, 2)):
            low = (torch.randn(N, h, w, 32, device=dev) * 3)
            up = torch.empty(N, C, H, W, dtype=dt, device=dev)
            tag = "f32" if e == 4 else "bf16"
            ms = timeit(lambda: ops.exit_gate(low, (H, W), layout="NHWC", n_classes=C, tau=0.5, want_ent=True,
                                              want_mask=True, up_out=up), args.iters, flush)
            nb = N * (h * w * C * 4 + C * H * W * e + H * W * 6)
            def eager():
                u = F.interpolate(low[..., :C].permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False).to(dt)
                p = F.softmax(u.float(), 1)
                ent = -(p * torch.log(p.clamp_min(1e-30))).sum(1) / 3.0445
                return u, ent, p.argmax(1), ent < 0.5, ent.mean((1, 2))
            row(f"exit_gate fused upsample+softmax+entropy+argmax+mask, materialise {tag} logits", ms, nb,
                eager_ms=timeit(eager, 5, flush), note="N=4 C=21 65x65->513x513")
        ms = timeit(lambda: ops.exit_gate(low, (H, W), layout="NHWC", n_classes=C, tau=0.5), args.iters, flush)
        row("exit_gate non-materialising (argmax u8 + per-image score only)", ms, N * (h * w * C * 4 + H * W * 1),
            note="SFU/latency-bound by design; report px/s: %.2f Gpx/s" % (N * H * W / ms / 1e6))
        for dt, e in ((torch.float32, 4), (torch.bfloat16, 2)):
            full = (torch.randn(N, C, H, W, device=dev) * 3).to(dt)
            ms = timeit(lambda: ops.exit_gate(full, None, tau=0.5, want_ent=True), args.iters, flush)
            def eager():
                p = F.softmax(full.float(), 1)
                ent = -(p * torch.log(p.clamp_min(1e-30))).sum(1) / 3.0445
                return ent, p.argmax(1), ent.mean((1, 2))
            row(f"exit_gate standalone on full-res {'f32' if e == 4 else 'bf16'} logits (A5 input)", ms,
                N * (C * H * W * e + H * W * 5), eager_ms=timeit(eager, 5, flush))

    # ---------------- multi-exit CE ----------------
    if want("ce"):
        E = 3
        tgt = blocky(N, C, H, W, dev)
        for dt, e in ((torch.float32, 4), (torch.bfloat16, 2)):
            y = (torch.randn(E, N, C, H, W, device=dev) * 3).to(dt).requires_grad_(True)
            coef = torch.ones(E, device=dev)
            def fused():
                per, _ = ops.multi_exit_ce(y, tgt, 21, coef)
                per.sum().backward()
                y.grad = None
            ms = timeit(fused, args.iters, flush)
            nb = 2 * E * N * C * H * W * e + N * H * W * 8 * 2
            def eager():
                y.grad = None
                l = sum(F.cross_entropy(y[i].float(), tgt.squeeze(1), ignore_index=21) for i in range(E))
                l.backward()
            row(f"multi_exit_ce fused fwd+bwd {'f32' if e == 4 else 'bf16'} (E=3)", ms, nb, eager_ms=timeit(eager, 5, flush))
            yd = y.detach()
            ms = timeit(lambda: ops.multi_exit_ce(yd, tgt, 21, coef), args.iters, flush)
            row(f"multi_exit_ce forward only {'f32' if e == 4 else 'bf16'}", ms, E * N * C * H * W * e + N * H * W * 8 * 2)
            del y, yd

    # ---------------- confusion histogram ----------------
    if want("hist"):
        tgt = blocky(N, C, H, W, dev)
        for dt, e in ((torch.float32, 4), (torch.bfloat16, 2)):
            lg = (torch.randn(N, C, H, W, device=dev) * 3).to(dt)
            ms = timeit(lambda: ops.confusion_hist(lg, tgt, C), args.iters, flush)
            def eager():
                pred = lg.argmax(1).view(N, -1)
                t = tgt.view(N, -1).clamp(max=C)
                return torch.stack([torch.bincount(t[n] * C + pred[n], minlength=(C + 1) * C) for n in range(N)])
            row(f"confusion_hist from {'f32' if e == 4 else 'bf16'} logits", ms, N * H * W * (C * e + 8), eager_ms=timeit(eager, 5, flush))
        pm = lg.argmax(1).to(torch.uint8)
        ms = timeit(lambda: ops.confusion_hist(pm, tgt, C), args.iters, flush)
        row("confusion_hist from uint8 argmax map", ms, N * H * W * 9)
        Hc, Wc = 1024, 2048
        pmc = torch.randint(0, 19, (N, Hc, Wc), device=dev, dtype=torch.uint8)
        tgc = blocky(N, 19, Hc, Wc, dev)
        ms = timeit(lambda: ops.confusion_hist(pmc, tgc, 19), args.iters, flush)
        row("confusion_hist from uint8 map, Cityscapes 1024x2048 N=4", ms, N * Hc * Wc * 9)

    # ---------------- Lovasz ----------------
    if want("lovasz"):
        E, Nl, Cl, Hl, Wl = 3, 1, 19, 768, 768
        y = torch.randn(E, Nl, Cl, Hl, Wl, device=dev).requires_grad_(True)
        tgt = blocky(Nl, Cl, Hl, Wl, dev)
        def run():
            per = ops.lovasz_multi_exit(y, tgt, ignore=19)
            per.sum().backward()
            y.grad = None
        ms = timeit(run, max(3, args.iters // 4), flush)
        P = Nl * Hl * Wl
        must = E * (2 * P * Cl * 4 + P * 8)
        sort_model = E * 4 * 2 * 8 * P * Cl
        row("lovasz fwd+bwd E=3 N=1 C=19 768x768 (must-touch + 4-pass sort model)", ms, must + sort_model,
            note=f"must-touch {must/1e6:.0f} MB, sort model {sort_model/1e6:.0f} MB")

    # ---------------- conv igemm per layer ----------------
    if want("conv"):
        for (cin, cout, R, dil, name) in [(2048, 256, 1, 1, "ASPP 1x1 Cin=2048"), (2048, 256, 3, 12, "ASPP 3x3 d=12 Cin=2048"),
                                          (2048, 256, 3, 24, "ASPP 3x3 d=24 Cin=2048"), (2048, 256, 3, 36, "ASPP 3x3 d=36 Cin=2048"),
                                          (1024, 256, 3, 12, "ASPP 3x3 d=12 Cin=1024"), (1024, 256, 1, 1, "project 1x1 K=1024"),
                                          (256, 256, 3, 1, "head 3x3 256->256"), (256, 32, 1, 1, "classifier 1x1 256->32")]:
            x = torch.randn(N, h, w, cin, device=dev).to(torch.bfloat16)
            wt = (torch.randn(cout, R, R, cin, device=dev) * 0.02).to(torch.bfloat16)
            sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
            out = torch.empty(N, h, w, cout, dtype=torch.bfloat16, device=dev)
            ms = timeit(lambda: conv_igemm(x, wt, sc, sh, dil, True, out, _lib.BF16, cout), args.iters, flush)
            fl = 2 * N * h * w * cout * cin * R * R
            xc = x.permute(0, 3, 1, 2)
            wc = wt.permute(0, 3, 1, 2)
            ems = timeit(lambda: F.relu(F.conv2d(xc, wc, padding=dil * (R // 2), dilation=dil)), 5, flush)
            row(f"conv_igemm {name} (N=4, 65x65)", ms, flops=fl, eager_ms=ems, note="nominal dense FLOPs; eager = cuDNN bf16 channels_last")

    out = {"peaks": peaks, "rows": rows}
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
