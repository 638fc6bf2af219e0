"""Data-parallel plumbing for the early-exit path: one process per GPU, images
sharded across ranks with no data-path collective; the only exchanges are (1) one integer all-reduce
of the stacked confusion matrices / exit counters at the end of an evaluation sweep and (2) the
gradient all-reduce of training (torch DDP over NCCL). The reference has no distributed code at all.
Works with the `gloo` backend on CPU tensors (used by the CPU tests) and `nccl` on the GPU box."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Image indices rank `rank` evaluates: r, r+world, r+2*world, ...."""
    return range(rank, n_items, world)


def all_reduce_counts(*tensors):
    """In-place SUM all-reduce of integer accumulators; exact and independent of the rank order."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            assert not t.dtype.is_floating_point, "only integer accumulators are reduced across ranks"
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def miou_from_cm(cm):
    """cm int64 [..., C+1, C] -> float64 mIoU per leading index (NaN when a class never occurs)."""
    C = cm.shape[-1]
    tp = torch.diagonal(cm[..., :C, :], dim1=-2, dim2=-1).double()
    fp = cm.sum(dim=-2).double() - tp
    fn = cm[..., :C, :].sum(dim=-1).double() - tp
    return (tp / (tp + fp + fn)).sum(dim=-1) / C


def wrap_ddp(net, local_rank):
    """Training: torch DDP (bucketed gradient all-reduce over NCCL, overlapped with backward).
    BatchNorm stays per rank, as in a single-GPU reference run at the per-GPU batch."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    return DDP(net, device_ids=[local_rank], gradient_as_bucket_view=True)


class FlatGradients:
    """Gradients of a set of parameters as views of ONE flat fp32 buffer, so that a data-parallel step needs a single
    all-reduce: autograd accumulates into the views in place, `zero()` clears them with one fill, `all_reduce_mean()`
    averages them over the ranks (NCCL: one AVG collective, capturable in a CUDA graph — train_funcs.GraphedTrainStep;
    gloo on CPU tensors: SUM and a division, used by the CPU tests), and the optimizer reads the views as usual."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.is_contiguous() and p.device == dev for p in self.params), \
            "flat gradients need contiguous fp32 parameters on one device"
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if self.flat.is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
