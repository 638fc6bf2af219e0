#!/usr/bin/env python
"""A few eagerly launched steps of bench.py's default workload (4 x 513 x 513, 3 exits, tau 0.5) for profiler captures:
every kernel of a step is an ordinary launch, so `ncu -k regex:conv_igemm -s <warm-up launches> -c <launches per step>`
attributes metrics per launch. Prints the number of eeseg launches per step.

    ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 195 -c 65 -o out python tools/one_step_eager.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ee_semantic_segmentation_b200 import _lib  # noqa: E402
from ee_semantic_segmentation_b200.engine import EarlyExitEngine  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = branchyDeepv3(None, "deeplabv3_resnet50", 2, bench.IMG, sections=bench.SECTIONS, pretrained=False).to(dev).eval()
net.strict_kernels = True
eng = EarlyExitEngine(net, bench.N_CLASSES, bench.TAU)
X, y = bench.synth_batch(0, bench.PER_GPU_BATCH)
X, y = X.to(dev), y.to(dev)
for k in range(steps):
    c0 = _lib.launch_count()
    eng.evaluate(X, y)
    torch.cuda.synchronize()
print("eeseg launches per step:", _lib.launch_count() - c0)
