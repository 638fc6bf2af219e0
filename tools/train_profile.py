#!/usr/bin/env python
"""Kernel-time breakdown of one training step (tools/train_bench.py 'eeseg' mode) with torch.profiler."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3
from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss
from ee_semantic_segmentation_b200.train_funcs import make_optimizer

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
fast = "--torch-heads" not in sys.argv
X, y = bench.synth_batch(0, 4)
from ee_semantic_segmentation_b200 import parallel
if "--no-direct" in sys.argv:
    parallel.DIRECT_GRADS = False
X, y = X.to(dev), y.to(dev)
torch.manual_seed(0)
net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=bench.SECTIONS, pretrained=False).to(dev).train()
net.fast_training_heads = fast
opt = make_optimizer(net, lr=1e-3, base_lr=1e-4)
loss_fn = BrXEntropyLoss(ignore_index=21, b_reduction="sum", n_exits=3)
def step():
    out = net(X); l = loss_fn(out, y); opt.zero_grad(set_to_none=True); l.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
