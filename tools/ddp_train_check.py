#!/usr/bin/env python
"""Data-parallel training check / throughput (BASELINE configs[2] at N GPUs, weak scaling): one process per GPU
under torchrun, torch DDP over NCCL for the gradient all-reduce (overlapped with the backward), the model's
convolution / BatchNorm / loss kernels from eeseg. Every rank trains on its own synthetic shard; after K steps
the parameters must be bit-identical on all ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_train_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ee_semantic_segmentation_b200 import parallel  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402
from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss  # noqa: E402
from ee_semantic_segmentation_b200.train_funcs import GraphedTrainStep, make_optimizer  # noqa: E402


def main():
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    steps = int(os.environ.get("STEPS", "10"))
    torch.manual_seed(0)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=bench.SECTIONS, pretrained=False).to(dev).train()
    mode = os.environ.get("DDP_MODE", "ddp")       # "graph": GraphedTrainStep with the all-reduce captured in the graph
    if world > 1 and mode == "ddp":
        ddp = parallel.wrap_ddp(net, local)
    elif world > 1 and mode == "nobcast":
        from torch.nn.parallel import DistributedDataParallel as DDP
        ddp = DDP(net, device_ids=[local], gradient_as_bucket_view=True, broadcast_buffers=False)
    else:
        ddp = net
    opt = make_optimizer(ddp, lr=1e-3, base_lr=1e-4)
    loss_fn = BrXEntropyLoss(ignore_index=21, b_reduction="sum", n_exits=3)
    X, y = bench.synth_batch(rank, 4)
    X, y = X.to(dev), y.to(dev)

    gstep = None
    if mode == "graph":
        gstep = GraphedTrainStep(net, loss_fn, opt, X, y)

        def step():
            return gstep(X, y)
    else:
        def step():
            l = loss_fn(ddp(X), y)
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
            return l
    for _ in range(3):
        l = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        l = step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().view(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        same = all(torch.equal(g, gathered[0]) for g in gathered)
    else:
        same = True
    if rank == 0:
        print(json.dumps({"n_gpus": world, "per_gpu_batch": 4, "ms_per_step": float(ms), "images_per_s": 4 * world / float(ms) * 1e3,
                          "loss_rank0": float(l), "params_identical_across_ranks": bool(same), "mode": mode}), flush=True)
    assert same or mode == "none", "parameters diverged across ranks"
    if gstep is not None:
        gstep.release()            # the graph references the communicator: destroy it first
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
