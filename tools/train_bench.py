#!/usr/bin/env python
"""Training-step throughput (BASELINE configs[2]: multi-exit weighted pixelwise CE, synthetic VOC 513x513
crops; configs[3]: Lovasz branchy loss, 19-class 768x768 crops) on one GPU:

  eeseg-graph : eeseg with the whole step (fwd, loss, bwd, SGD) replayed as one CUDA graph (train_funcs.GraphedTrainStep)
  eeseg       : backbone + head convolutions fwd/dgrad/wgrad on the tcgen05 kernels, fused BatchNorm/ReLU and loss kernels
  eeseg-loss  : PyTorch-module heads (cuDNN) + fused multi-exit loss kernel
  torch       : the reference's GPU path — PyTorch modules + torch losses (TF32 allowed, train_funcs.py:117-118)

    python tools/train_bench.py [--config ce|lovasz] [--batch 4] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from ee_semantic_segmentation_b200.branchy_seg_losses import LovaszSoftmax  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402
from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss  # noqa: E402
from ee_semantic_segmentation_b200.train_funcs import GraphedTrainStep, make_optimizer  # noqa: E402


def torch_lovasz(probas, labels, ignore):
    """Lovasz-softmax (classes='present', batch-level) with torch ops, the way the reference runs it on a GPU
    (lovaszsoftmax.py:154-219): per-class sort of the errors, Jaccard-gradient weights, dot product."""
    N, C = probas.shape[:2]
    p = probas.permute(0, 2, 3, 1).reshape(-1, C)
    lab = labels.reshape(-1)
    keep = lab != ignore
    p, lab = p[keep], lab[keep]
    losses = []
    for c in range(C):
        fg = (lab == c).float()
        if fg.sum() == 0:
            continue
        err, perm = torch.sort((fg - p[:, c]).abs(), 0, descending=True)
        fgs = fg[perm]
        inter = fgs.sum() - fgs.cumsum(0)
        union = fgs.sum() + (1 - fgs).cumsum(0)
        jac = 1 - inter / union
        jac[1:] = jac[1:] - jac[:-1]
        losses.append(torch.dot(err, jac))
    return torch.stack(losses).mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ce", choices=["ce", "lovasz"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    C, img = (21, 513) if args.config == "ce" else (19, 768)
    X, y = bench.synth_batch(0, args.batch, img=img, n_classes=C)
    X, y = X.to(dev), y.to(dev)
    rows = []
    ap_modes = ("eeseg-graph", "eeseg", "eeseg-loss", "torch")
    for mode in ap_modes:
        torch.manual_seed(0)
        net = branchyDeepv3(None, "deeplabv3_resnet50", 2, img, sections=bench.SECTIONS, pretrained=False,
                            num_classes=C).to(dev).train()
        net.fast_training_heads = net.fast_training_backbone = mode in ("eeseg", "eeseg-graph")
        opt = make_optimizer(net, lr=1e-3, base_lr=1e-4)
        if mode == "torch":
            if args.config == "ce":
                ce = torch.nn.CrossEntropyLoss(ignore_index=C)
                loss_fn = lambda out, t: sum(ce(out[i], t.squeeze(1)) for i in range(out.shape[0]))
            else:
                loss_fn = lambda out, t: sum(torch_lovasz(out[i], t.squeeze(1), C) for i in range(out.shape[0]))
        else:
            loss_fn = (BrXEntropyLoss(ignore_index=C, b_reduction="sum", n_exits=3) if args.config == "ce"
                       else LovaszSoftmax(ignore=C, n_branches=2))

        def step():
            out = net(X)
            l = loss_fn(out, y)
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
            return l
        if mode == "eeseg-graph":
            gstep = GraphedTrainStep(net, loss_fn, opt, X, y)
            step = lambda: gstep(X, y)
        for _ in range(3):
            l = step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            l = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        r = {"mode": mode, "config": args.config, "batch": args.batch, "img": img, "ms_per_step": ms,
             "images_per_s": args.batch / ms * 1e3, "loss": float(l), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        rows.append(r)
        print(json.dumps(r), flush=True)
        del net, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
