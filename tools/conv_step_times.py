#!/usr/bin/env python
"""True (in-kernel wall clock) duration of every conv launch of one bench step, grouped by layer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ee_semantic_segmentation_b200 import _lib, head_plan
from ee_semantic_segmentation_b200.engine import EarlyExitEngine
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3

if not hasattr(_lib.lib(), "eeseg_conv_timing"):
    raise SystemExit("this tool needs the tuning build: python -m ee_semantic_segmentation_b200.build --tuning, then run with "
                     "EESEG_LIB=ee_semantic_segmentation_b200/libeeseg_b200_tuning.so (the product library has no EESEG_TUNING hooks)")
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = branchyDeepv3(None, "deeplabv3_resnet50", 2, bench.IMG, sections=bench.SECTIONS, pretrained=False).to(dev).eval()
eng = EarlyExitEngine(net, bench.N_CLASSES, bench.TAU)
X, y = bench.synth_batch(0, bench.PER_GPU_BATCH)
X, y = X.to(dev), y.to(dev)
for _ in range(3):
    eng.evaluate(X, y)
cap = 256
buf = torch.zeros(cap, 2, dtype=torch.int64, device=dev)
buf[:, 0] = torch.iinfo(torch.int64).max
prof = []
head_plan.PROFILE = prof
torch.cuda.synchronize()
_lib.lib().eeseg_conv_timing(buf.data_ptr(), cap)
torch.cuda._sleep(int(4e7))
eng.evaluate(X, y)
torch.cuda.synchronize()
n = _lib.lib().eeseg_conv_timing(None, 0)
head_plan.PROFILE = None
t = buf[:n].cpu()
dur = (t[:, 1] - t[:, 0]).double() / 1e3
tot = dur.sum().item()
print(f"{n} conv launches, sum of kernel durations {tot:.1f} us; span first start -> last end {(t[:,1].max()-t[:,0].min()).item()/1e3:.1f} us")
for i in range(n):
    fl = prof[i][2]
    print(f"{i:3d} {prof[i][3]:9s} {dur[i]:7.1f} us  {fl/1e9:7.2f} GF  {fl/dur[i].item()/1e6:7.0f} TF/s")
