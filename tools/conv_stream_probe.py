#!/usr/bin/env python
"""What bounds the mid-size conv launches? In-kernel wall time (first CTA start -> last CTA end, %globaltimer) of a few
layer shapes with parts of the operand stream switched off and with the grid capped (tuning build only):

    python -m ee_semantic_segmentation_b200.build --tuning
    EESEG_LIB=ee_semantic_segmentation_b200/libeeseg_b200_tuning.so python tools/conv_stream_probe.py [--out f.json]

skip bits: A = activation tiles, B = weight tiles, R = residual tiles, S = output stores. `warm`: launches back to back on
the same tensors (operands L2-resident), `cold`: a 256 MiB fill before each launch.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ee_semantic_segmentation_b200 import _lib  # noqa: E402
from ee_semantic_segmentation_b200.head_plan import conv_igemm  # noqa: E402

L = _lib.lib()
if not hasattr(L, "eeseg_conv_probe"):
    raise SystemExit("needs the tuning build (EESEG_LIB=.../libeeseg_b200_tuning.so)")

SHAPES = [  # name, N, h, w, Cin, Cout, R, dil, residual
    ("l2.c3 128>512+res", 4, 65, 65, 128, 512, 1, 1, True),
    ("l3.c1 1024>256", 4, 65, 65, 1024, 256, 1, 1, False),
    ("l3.c2 3x3 256 d2", 4, 65, 65, 256, 256, 3, 2, False),
    ("l3.c3 256>1024+res", 4, 65, 65, 256, 1024, 1, 1, True),
    ("l4.c1 2048>512", 4, 65, 65, 2048, 512, 1, 1, False),
    ("l4.c2 3x3 512 d4", 4, 65, 65, 512, 512, 3, 4, False),
    ("l4.c3 512>2048+res", 4, 65, 65, 512, 2048, 1, 1, True),
    ("aspp 3x3 d12 2048", 4, 65, 65, 2048, 256, 3, 12, False),
]
MASKS = [("full", 0), ("-A", 1), ("-B", 2), ("-R", 4), ("-S", 8), ("-A-B", 3), ("-A-B-R", 7), ("none", 15),
         ("direct", 16), ("2stages", 32)]   # 16: register -> global epilogue (one more ring stage), 32: ring capped at 2
GRIDS = [148, 111, 74, 37]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tbuf = torch.empty(64, 2, dtype=torch.int64, device=dev)
    rows = []

    def timed(run, cold, reps):
        out = []
        for _ in range(reps):
            if cold:
                flush.fill_(1)
            tbuf[:, 0] = torch.iinfo(torch.int64).max
            tbuf[:, 1] = 0
            torch.cuda.synchronize()
            L.eeseg_conv_timing(tbuf.data_ptr(), 64)
            run()
            torch.cuda.synchronize()
            L.eeseg_conv_timing(None, 0)
            t = tbuf[0].cpu()
            out.append((t[1] - t[0]).item() / 1e3)
        out.sort()
        return out[len(out) // 2]

    for name, N, h, w, cin, cout, R, dil, res in SHAPES:
        x = torch.randn(N, h, w, cin, device=dev).to(torch.bfloat16)
        wt = (torch.randn(cout, R, R, cin, device=dev) * 0.02).to(torch.bfloat16)
        sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        out = torch.empty(N, h, w, cout, dtype=torch.bfloat16, device=dev)
        r = torch.randn(N, h, w, cout, device=dev).to(torch.bfloat16) if res else None
        run = lambda: conv_igemm(x, wt, sc, sh, dil, True, out, _lib.BF16, cout, residual=r)
        for _ in range(3):
            run()
        row = {"shape": name}
        for mname, mask in MASKS:
            if (mask & 4) and mask < 16 and not res and mask != 15 and mname != "-A-B-R":
                continue
            L.eeseg_conv_probe(mask, 0)
            row[f"warm {mname}"] = timed(run, False, args.reps)
            if mname in ("full", "-R", "-A-B-R", "direct"):
                row[f"cold {mname}"] = timed(run, True, args.reps)
        for g in GRIDS[1:]:
            L.eeseg_conv_probe(0, g)
            row[f"warm grid{g}"] = timed(run, False, args.reps)
            L.eeseg_conv_probe(3, g)
            row[f"warm grid{g} -A-B"] = timed(run, False, args.reps)
        L.eeseg_conv_probe(0, 0)
        rows.append(row)
        print(name, " ".join(f"{k}={v:.1f}" for k, v in row.items() if k != "shape"), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
