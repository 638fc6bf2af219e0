#!/usr/bin/env python
"""Probe: throughput of TWO step graphs replayed concurrently on two streams (independent batches, separate engines)
against one graph replayed back to back. Every conv launch is a persistent grid over all SMs, so two replays cannot share
an SM, but the tail of one stream's launch can be filled by the other stream's next launch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ee_semantic_segmentation_b200.engine import EarlyExitEngine  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = branchyDeepv3(None, "deeplabv3_resnet50", 2, bench.IMG, sections=bench.SECTIONS, pretrained=False).to(dev).eval()
net.strict_kernels = True
X, y = bench.synth_batch(0, bench.PER_GPU_BATCH)
X, y = X.to(dev), y.to(dev)
engs = [EarlyExitEngine(net, bench.N_CLASSES, bench.TAU, use_graph=True) for _ in range(2)]
for e in engs:
    for _ in range(3):
        e.evaluate(X, y)
torch.cuda.synchronize()
K = 100


def run(n_streams):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        s = streams[k % n_streams]
        with torch.cuda.stream(s):
            engs[k % n_streams].replay(tuple(X.shape))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return K * bench.PER_GPU_BATCH / dt, dt / K * 1e3


for n in (1, 2, 1, 2):
    ips, ms = run(n)
    print(f"{n} stream(s): {ips:8.1f} img/s  {ms:.3f} ms per batch")
