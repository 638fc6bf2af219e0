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


# backward kernels may add parameter gradients straight into FlatGradients views (direct_grad); EESEG_DIRECT_GRADS=0: A/B
DIRECT_GRADS = os.environ.get("EESEG_DIRECT_GRADS", "1") != "0"


def wrap_ddp(net, local_rank):
    """Training: torch DDP (bucketed gradient all-reduce over NCCL, overlapped with backward).
    BatchNorm stays per rank, as in a single-GPU reference run at the per-GPU batch. DDP's reducer listens to
    autograd's AccumulateGrad nodes, so the direct-to-.grad kernels are switched off for the process."""
    global DIRECT_GRADS
    DIRECT_GRADS = False
    from torch.nn.parallel import DistributedDataParallel as DDP
    return DDP(net, device_ids=[local_rank], gradient_as_bucket_view=True)


class FlatGradients:
    """Gradients of a set of parameters as views of ONE flat fp32 buffer: autograd accumulates into the views in place,
    `zero()` clears them with one fill, and a data-parallel step averages them over the ranks with a few large
    collectives instead of one per tensor (NCCL: AVG all-reduces, capturable in a CUDA graph —
    train_funcs.GraphedTrainStep; gloo on CPU tensors: SUM and a division, used by the CPU tests). The optimizer reads
    the views as usual.

    `buckets` > 1 overlaps the exchange with the backward pass: the buffer is laid out in REVERSE parameter order (the
    last layers' gradients are complete first) and cut into `buckets` contiguous ranges of about equal size; a
    post-accumulate hook on every parameter counts its bucket down, and the bucket's all-reduce is issued on a side
    stream the moment its last gradient has been written, while the rest of the backward keeps the main stream busy
    (`begin()` before the backward, `finish()` before the optimizer). With one bucket, or without hooks, `all_reduce_mean()`
    reduces everything after the backward."""

    def __init__(self, params, buckets=1):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.dtype == torch.float32 and p.is_contiguous() and p.device == dev for p in self.params), \
            "flat gradients need contiguous fp32 parameters on one device"
        total = sum((p.numel() + 3) // 4 * 4 for p in self.params)     # every view starts 16-byte aligned
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        order = list(reversed(self.params))
        self.buckets = []          # [start, end) element ranges, in the order their gradients complete
        self._bucket_of = {}
        n_b = max(1, int(buckets))
        o, start, cur = 0, 0, 0
        for p in order:
            size = (p.numel() + 3) // 4 * 4
            if cur < n_b - 1 and o > start and (o - start) + size / 2 > total / n_b:
                self.buckets.append((start, o))          # close the bucket before a tensor that would overshoot its share
                start, cur = o, cur + 1
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            p._eeseg_flat = self           # kernels may add their gradient straight into p.grad (direct_grad below)
            self._bucket_of[id(p)] = cur
            o += size
        self.buckets.append((start, o))
        self._need = [0] * len(self.buckets)
        for p in order:
            self._need[self._bucket_of[id(p)]] += 1
        self._left = list(self._need)
        self._seen = set()
        self._hooks = []
        self._side = None
        self._armed = False
        if len(self.buckets) > 1 and hasattr(self.params[0], "register_post_accumulate_grad_hook"):
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- exchange -------------------------------------------------------------------------------------------------
    def _world(self):
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _reduce_range(self, a, b):
        t = self.flat[a:b]
        if t.is_cuda:
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(dist.get_world_size())

    def _on_grad(self, p):
        if not self._armed or id(p) in self._seen:      # a parameter counts once per backward, however it was notified
            return
        self._seen.add(id(p))
        b = self._bucket_of[id(p)]
        self._left[b] -= 1
        if self._left[b] == 0 and self._world() > 1:
            a, e = self.buckets[b]
            if self.flat.is_cuda:
                main = torch.cuda.current_stream(self.flat.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.flat.device)
                self._side.wait_stream(main)           # the bucket's gradients are complete on the main stream
                with torch.cuda.stream(self._side):
                    self._reduce_range(a, e)
            else:
                self._reduce_range(a, e)
            self._left[b] = -1                          # reduced

    def written(self, p):
        """A kernel has added p's gradient straight into p.grad (no AccumulateGrad node ran, so no hook fired)."""
        self._on_grad(p)

    def begin(self):
        """Call before backward(): arms the per-bucket countdown (overlapped exchange)."""
        self._left = list(self._need)
        self._seen = set()
        self._armed = bool(self._hooks)

    def finish(self):
        """Call after backward(), before the optimizer: reduces whatever the hooks did not, joins the side stream."""
        if self._world() > 1:
            for b, (a, e) in enumerate(self.buckets):
                if not self._armed or self._left[b] != -1:
                    self._reduce_range(a, e)
            if self._side is not None and self.flat.is_cuda:
                torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
        self._armed = False

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """Everything in one go after the backward (no overlap)."""
        self._armed = False
        self.finish()


def direct_grad(p):
    """p.grad when a backward kernel may ADD its result straight into it (a dense fp32 view of a FlatGradients buffer that
    the owner zeroes before every backward), else None. The caller then returns None for this gradient from its
    autograd.Function and calls `p._eeseg_flat.written(p)` — one read-modify-write launch less per parameter than autograd's
    AccumulateGrad."""
    fg = getattr(p, '_eeseg_flat', None)
    g = p.grad
    if not DIRECT_GRADS or fg is None or g is None or g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda:
        return None
    lo, hi = fg.flat.data_ptr(), fg.flat.data_ptr() + fg.flat.numel() * 4
    if not (lo <= g.data_ptr() < hi):
        return None                      # someone replaced .grad (zero_grad(set_to_none=True), another optimizer)
    return g
