#!/usr/bin/env python
"""Per-shape timing of eeseg_conv_igemm_fwd on the convolutions of one VOC 513x513 step (N=4): CUDA events around
`reps` launches with an L2 flush before each, bit-level checksum of the output (compare runs with different
EESEG_CONV_CLUSTER settings), and a check against torch's fp32 conv for every shape.

    [EESEG_CONV_CLUSTER=1|2|4] python tools/conv_shape_bench.py [--reps 20] [--out file.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ee_semantic_segmentation_b200 import _lib  # noqa: E402
from ee_semantic_segmentation_b200.head_plan import conv_igemm  # noqa: E402

# name, N, h, w, Cin, Cout, k, dil, stride, residual
SHAPES = [
    ("l1.conv1 256->64 1x1 @129", 4, 129, 129, 256, 64, 1, 1, 1, False),
    ("l1.conv2 64->64 3x3 @129", 4, 129, 129, 64, 64, 3, 1, 1, False),
    ("l1.conv3 64->256 1x1+res @129", 4, 129, 129, 64, 256, 1, 1, 1, True),
    ("l2.conv1 512->128 1x1 @65", 4, 65, 65, 512, 128, 1, 1, 1, False),
    ("l2.conv2 128->128 3x3 @65", 4, 65, 65, 128, 128, 3, 1, 1, False),
    ("l2.conv3 128->512 1x1+res @65", 4, 65, 65, 128, 512, 1, 1, 1, True),
    ("l3.conv1 1024->256 1x1 @65", 4, 65, 65, 1024, 256, 1, 1, 1, False),
    ("l3.conv2 256->256 3x3 d2 @65", 4, 65, 65, 256, 256, 3, 2, 1, False),
    ("l3.conv3 256->1024 1x1+res @65", 4, 65, 65, 256, 1024, 1, 1, 1, True),
    ("l4.conv1 2048->512 1x1 @65", 4, 65, 65, 2048, 512, 1, 1, 1, False),
    ("l4.conv2 512->512 3x3 d4 @65", 4, 65, 65, 512, 512, 3, 4, 1, False),
    ("l4.conv3 512->2048 1x1+res @65", 4, 65, 65, 512, 2048, 1, 1, 1, True),
    ("aspp 2048->256 1x1 @65", 4, 65, 65, 2048, 256, 1, 1, 1, False),
    ("aspp 2048->256 3x3 d12 @65", 4, 65, 65, 2048, 256, 3, 12, 1, False),
    ("aspp 2048->256 3x3 d36 @65", 4, 65, 65, 2048, 256, 3, 36, 1, False),
    ("head 256->256 3x3 @65", 4, 65, 65, 256, 256, 3, 1, 1, False),
    ("proj 1024->256 1x1 @65", 4, 65, 65, 1024, 256, 1, 1, 1, False),
    ("l3.conv2 N=1", 1, 65, 65, 256, 256, 3, 2, 1, False),
    ("l3.conv2 N=3", 3, 65, 65, 256, 256, 3, 2, 1, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--only", default="", help="substring of the shape name")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, N, h, w, Cin, Cout, k, dil, stride, res in SHAPES:
        if args.only and args.only not in name:
            continue
        g = torch.Generator(device="cpu").manual_seed(hash(name) & 0xffff)
        x = torch.randn(N, h, w, Cin, generator=g).to(dev).to(torch.bfloat16)
        wt = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev).to(torch.bfloat16)
        scale = torch.ones(Cout, device=dev) if res else (torch.rand(Cout, generator=g) + 0.5).to(dev)
        shift = (torch.randn(Cout, generator=g) * 0.1).to(dev)
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = torch.empty((N, ho, wo, Cout), dtype=torch.bfloat16, device=dev)
        r = torch.randn(N, ho, wo, Cout, generator=g).to(dev).to(torch.bfloat16) if res else None
        run = lambda: conv_igemm(x, wt, scale, shift, dil, True, out, _lib.BF16, Cout, stride=stride, residual=r)
        run()
        torch.cuda.synchronize()
        err = None
        if not args.no_check:
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), stride=stride,
                           padding=dil * (k // 2), dilation=dil)
            ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
            if res:
                ref = ref + r.float().permute(0, 3, 1, 2)
            ref = torch.relu(ref).permute(0, 2, 3, 1)
            err = float((out.float() - ref).abs().max() / ref.abs().max())
            assert err < 1e-2, (name, err)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            run()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
        us = ts[len(ts) // 2]
        fl = 2 * N * ho * wo * Cout * k * k * Cin
        csum = int(out.view(torch.int16).to(torch.int64).sum())
        rows.append({"shape": name, "us": round(us, 2), "tflops": round(fl / us / 1e6, 1), "checksum": csum, "err": err})
        print(f"{name:36s} {us:8.2f} us  {fl / us / 1e6:8.1f} TFLOP/s  err {err}  checksum {csum}", flush=True)
    if args.out:
        json.dump({"cluster": os.environ.get("EESEG_CONV_CLUSTER", "default"), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
