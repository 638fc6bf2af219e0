"""Drop-in for the reference's training driver: `train_epoch` (train_funcs.py:12-33), the epoch loop `train`
(:60-269: validation tracker, best-validation checkpoint dict, early stopping) and the optimiser / schedule
construction of `train_deepv3` (deepv3_funcs.py:74-101,138-156) as small helpers, with one-process-per-GPU data
parallelism.

Under net.train() the model's convolutions (forward, input and weight gradients), BatchNorm, max-pool, up-sampling
and the multi-exit losses run on the eeseg kernels behind autograd (head_train / backbone_train / bn_train / ops);
`GraphedTrainStep` replays the whole step as one CUDA graph. Under `torchrun`, wrap the model with
`parallel.wrap_ddp` — gradients are averaged by NCCL bucket all-reduces overlapped with the backward pass; BatchNorm
stays per rank, as in a single-GPU reference run at the per-GPU batch."""
import os
import re
from collections import defaultdict
from copy import deepcopy

import numpy as np
import torch as tch
from torch import optim


def train_epoch(net, train_iter, loss, updater, device=tch.device('cpu')):
    """Same loop as train_funcs.py:12-33 (returns the last batch loss for logging)."""
    if isinstance(net, tch.nn.Module):
        net.train()
    last = None
    for X, y in train_iter:
        X, y = X.to(device, non_blocking=True), y.to(device, non_blocking=True)
        y_hat = net(X)
        l = loss(y_hat, y)
        if isinstance(updater, tch.optim.Optimizer):
            updater.zero_grad()
            l.mean().backward()
            updater.step()
        else:
            l.sum().backward()
            updater(X.shape[0])
        last = l.detach()
        del X, y
    return last


def train(net, train_iter, loss, num_epochs, updater, val_iter=None, metrics=None, patience=None, saveat=None,
          start_from=None, verbose=False, device='cpu', scheduler=None, use_file=None, up_updater=False, ret_lr=False,
          ae_train=False, transform=None, name=None, minimize=True, start_counting=0, **kwargs):
    """The reference's epoch loop (train_funcs.py:60-269) for the segmentation path: per epoch `train_epoch`, then
    every `(name, f)` of `metrics` on `val_iter` — 'mIoU' is called as `f(net, n_exits, kwargs['nout_channels'], val_iter,
    device)` (eval_mIoU.mIoU_evaluator) and, for a branchy net (`n_branches=` given), its result dict lands in
    `tracker['val_mIoU_<key>']` (b1_mIoU, ..., mIoU); the followed value is the mean of the tracker entries matching
    `val_<metrics[0][0]>`; a better value saves `{"model_state_dict", "opt_state_dict", "epoch", ...}` to `saveat`;
    `patience` epochs without improvement stop the loop (after `start_counting`). Returns the tracker.

    Kept quirks of the reference: the loop runs epochs 1 .. num_epochs-1 (`if epoch >= num_epochs: break` after the
    increment, :139-141); `start_from` restores model (and optionally optimiser) state and the best value; the learning
    rate logged is the LAST param group's for a branchy net. Not carried over: the auto-encoder branch (`ae_train`) and
    the `funcs.eval_branches / eval_results` wrappers for non-mIoU metrics (`funcs` is missing from the reference tree) —
    other metrics are called as `f(net, val_iter, device)` and must return a dict (branchy) or a number."""
    if ae_train:
        raise NotImplementedError("ae_train (auto-encoder pre-training) is outside the segmentation path")
    metrics = metrics or []
    follow = f'val_{metrics[0][0]}' if metrics else None
    tracker = defaultdict(list)
    net.to(device)
    name = name or 'unspecified'

    def say(msg):
        if not verbose:
            return
        if use_file:
            with open(use_file, 'a') as f:
                f.write(msg + '\n')
        else:
            print(msg)
    counter = 0
    best_val = np.inf if minimize else 0.
    saveat = saveat or os.path.join('.', 'model.pth')
    if patience:
        say(f'<< {name} progress update >> Earlystopping will follow {follow} with patience set to {patience}.')
    else:
        patience = None
        say(f'<< {name} progress update >> Earlystopping not set.')
    if start_from:
        save_dict = tch.load(start_from, weights_only=False)
        net.load_state_dict(save_dict['model_state_dict'])
        if up_updater:
            lr_aux = updater.param_groups[0]['lr']
            updater.load_state_dict(save_dict['opt_state_dict'])
            updater.param_groups[0]['lr'] = lr_aux
        if patience and follow in save_dict.keys():
            best_val = save_dict[follow]
    branchy = bool(kwargs.get('n_branches'))
    epoch, last_lr = 0, 0
    num_epochs = num_epochs or np.inf
    tch.backends.cuda.matmul.allow_tf32 = True       # :117-118 (only the PyTorch-module fallbacks are affected)
    tch.backends.cudnn.allow_tf32 = True

    def save(cur_val, branch_val):
        save_dict = {"model_state_dict": deepcopy(net.state_dict()), "opt_state_dict": deepcopy(updater.state_dict()),
                     "epoch": epoch}
        for key, _ in metrics:                         # :211-214 (the pattern is the tracker key, as in the reference)
            for k in list(tracker.keys()):
                if re.search(k, key):
                    save_dict[f'val_{k}'] = tracker[f'val_{k}'][-1]
        tch.save(save_dict, saveat)
        msg = f'<< {name} progress update >> saved @ {epoch} epoch. Best score: {cur_val:.5g}'
        if branchy:
            msg += '\nFor each branch:\n\t' + '\n\t'.join(f'b{i + 1} = {v:.5g}' for i, v in enumerate(branch_val))
        say(msg)

    while True:
        epoch += 1
        if epoch >= num_epochs:
            break
        cur_lr = updater.state_dict()['param_groups'][-1 if branchy else 0]['lr']
        say(f'<< {name} progress update >> starting #{epoch} training epoch; lr = {cur_lr}, no updates since {counter} epochs')
        train_epoch(net, train_iter, loss, updater, device)
        if val_iter:
            with tch.no_grad():
                for met, f in metrics:
                    if met == 'mIoU':
                        cur_res = f(net, net.n_branches + 1 if branchy else 1, kwargs['nout_channels'], val_iter, device)
                    else:
                        cur_res = f(net, val_iter, device)
                    if branchy:
                        for key, value in cur_res.items():
                            tracker[f'val_{met}_{key}'].append(value)
                    elif met == 'mIoU':
                        tracker[f'val_{met}'].append(cur_res['mIoU'])
                    else:
                        tracker[f'val_{met}'].append(cur_res)
        if ret_lr or scheduler:
            tracker['lr'].append(cur_lr)
        branch_val = []
        if follow is None or not val_iter:
            cur_val = np.inf if minimize else 0.
        elif branchy:
            branch_val = [tracker[key][-1] for key in tracker.keys() if re.search(follow, key)]
            if kwargs.get('max2min'):
                # the reference reads an undefined `cur_val` here (:188); the evident intent: weights 1..E, optionally flipped
                weights = np.arange(len(branch_val)) + 1
                w_max = np.max(weights)
                if kwargs['max2min']:
                    weights = np.flip(weights)
                cur_val = np.average(branch_val, weights=weights / w_max)
            else:
                cur_val = np.average(branch_val)
        else:
            cur_val = tracker[follow][-1]
        if scheduler:
            scheduler.step()
        better = best_val > cur_val if minimize else best_val < cur_val
        if patience:
            if counter < patience:
                if better:
                    save(cur_val, branch_val)
                    best_val, counter = cur_val, 0
                elif 'lr' in tracker and last_lr != cur_lr:
                    counter, last_lr = 1, cur_lr
                else:
                    counter += 1
            elif epoch > start_counting:
                break
            else:
                if 'lr' in tracker and last_lr != cur_lr:
                    counter, last_lr = 0, cur_lr
                counter += 1
        elif better:
            save(cur_val, branch_val)
            best_val, counter = cur_val, 0
        else:
            counter += 1
    return tracker


class SGD(optim.Optimizer):
    """torch.optim.SGD's update (momentum, weight decay, one learning rate per parameter group; dampening 0, no
    nesterov — what deepv3_funcs.py:74-101 constructs) on ONE eeseg launch per step (eeseg_sgd_multi): a device table
    of 64 K-element chunks over all parameters, gradients and momentum buffers, the groups' learning rates read from a
    small device array that `step()` refreshes — so a scheduler's `param_groups[i]['lr']` changes reach a CAPTURED step
    without re-capturing it (GraphedTrainStep refreshes the array before every replay).

    Gradients are views of one flat buffer (parallel.FlatGradients; `buckets` > 1 = the overlapped data-parallel
    exchange), momentum buffers views of another: pointers never change, `zero_grad()` is one fill. `state_dict()` /
    `load_state_dict()` use torch.optim.SGD's layout ('momentum_buffer' per parameter), so the reference's checkpoint
    dict ("opt_state_dict", train_funcs.py:208-216) round-trips with torch's optimizer. CUDA parameters only."""

    CHUNK = 65536

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, buckets=1):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        from . import _lib
        from .parallel import FlatGradients
        ps = [p for g in self.param_groups for p in g['params']]
        if not ps or not all(p.is_cuda and p.dtype == tch.float32 for p in ps):
            raise RuntimeError('eeseg SGD needs fp32 CUDA parameters (no CPU fallback)')
        m0, w0 = self.param_groups[0]['momentum'], self.param_groups[0]['weight_decay']
        if any(g['momentum'] != m0 or g['weight_decay'] != w0 for g in self.param_groups):
            raise ValueError('eeseg SGD takes one momentum / weight decay for all groups (per-group learning rates)')
        self._lib = _lib
        self.flat_grads = FlatGradients(ps, buckets=buckets)
        dev = ps[0].device
        self._mom = tch.zeros_like(self.flat_grads.flat)
        recs = []
        for gi, g in enumerate(self.param_groups):
            for p in g['params']:
                if not p.requires_grad:
                    continue
                off = (p.grad.data_ptr() - self.flat_grads.flat.data_ptr()) // 4
                buf = self._mom[off:off + p.numel()].view_as(p)
                self.state[p]['momentum_buffer'] = buf
                for c0 in range(0, p.numel(), self.CHUNK):
                    n = min(self.CHUNK, p.numel() - c0)
                    recs.append((p.data_ptr() + 4 * c0, p.grad.data_ptr() + 4 * c0, buf.data_ptr() + 4 * c0, n, gi))
        assert _lib.lib().eeseg_sgd_chunk_bytes() == 32
        table = np.zeros(len(recs), dtype=np.dtype([('p', '<u8'), ('g', '<u8'), ('b', '<u8'), ('n', '<i4'), ('grp', '<i4')]))
        for i, r in enumerate(recs):
            table[i] = r
        self._table = tch.from_numpy(table.view(np.uint8).copy()).to(dev)
        self._n_chunks = len(recs)
        self._lrs = tch.zeros(len(self.param_groups), dtype=tch.float32, device=dev)
        self._lrs_host = tch.zeros(len(self.param_groups), dtype=tch.float32).pin_memory()
        self._ptrs = [p.data_ptr() for p in ps]
        self.sync_lrs()

    def sync_lrs(self):
        """Host learning rates -> the device array the kernel reads (async copy from pinned memory)."""
        for i, g in enumerate(self.param_groups):
            self._lrs_host[i] = float(g['lr'])
        self._lrs.copy_(self._lrs_host, non_blocking=True)

    def zero_grad(self, set_to_none=False):
        self.flat_grads.zero()          # the views stay: the chunk table holds their addresses

    @tch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if not tch.cuda.is_current_stream_capturing():
            self.sync_lrs()
        g0 = self.param_groups[0]
        dev = self._table.device
        with tch.cuda.device(dev):
            self._lib.check(self._lib.lib().eeseg_sgd_multi(
                self._table.data_ptr(), self._n_chunks, self._lrs.data_ptr(), float(g0['momentum']),
                float(g0['weight_decay']), tch.cuda.current_stream(dev).cuda_stream), 'eeseg_sgd_multi')
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for g in self.param_groups:                     # loaded buffers are fresh tensors: copy them into the flat views
            for p in g['params']:
                if not p.requires_grad:
                    continue
                off = (p.grad.data_ptr() - self.flat_grads.flat.data_ptr()) // 4
                view = self._mom[off:off + p.numel()].view_as(p)
                loaded = self.state[p].get('momentum_buffer')
                if loaded is not None and loaded.data_ptr() != view.data_ptr():
                    view.copy_(loaded)
                elif loaded is None:
                    view.zero_()
                self.state[p]['momentum_buffer'] = view
        self.sync_lrs()


def make_optimizer(net, lr, base_lr=None, weighted_lr=False, buckets=1):
    """SGD(momentum 0.9, weight decay 5e-4) with the reference's parameter groups
    (deepv3_funcs.py:77-101): backbone at base_lr, branches at lr, final classifier at 1.1*lr. CUDA parameters get the
    eeseg single-launch optimizer (`SGD` above), CPU parameters (host-side tests) torch.optim.SGD."""
    net = getattr(net, 'module', net)      # DDP
    on_gpu = all(p.is_cuda for p in net.parameters())
    ctor = (lambda params, **kw: SGD(params, buckets=buckets, **kw)) if on_gpu else optim.SGD
    if base_lr and getattr(net, 'n_branches', 0):
        params = [{'params': list(net.base_model.parameters()), 'lr': base_lr}]
        if weighted_lr:
            w = np.linspace(1, 1.2, num=net.n_branches + 1)
            params += [{'params': list(net.branches[i].parameters()), 'lr': lr * w[i]} for i in range(net.n_branches)]
            params.append({'params': list(net.classifier.parameters()), 'lr': lr * w[-1]})
        else:
            params.append({'params': list(net.branches.parameters()), 'lr': lr})
            params.append({'params': list(net.classifier.parameters()), 'lr': lr * 1.1})
        return ctor(params, lr=lr, momentum=.9, weight_decay=5e-4)
    return ctor(list(net.parameters()), lr=lr, momentum=.9, weight_decay=5e-4)


def poly_scheduler(optimizer, num_epochs, lr=None, min_lr=None):
    """Per-epoch poly schedule (1 - k/num_epochs)^0.9 (deepv3_funcs.py:146-152)."""
    if min_lr and lr:
        w = (min_lr / lr) ** (1 / .9)
        n0 = num_epochs * w / (1 - w)
        return optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda k: (1 - k / (num_epochs + n0)) ** .9)
    return optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda k: (1 - k / num_epochs) ** .9)


class GraphedTrainStep:
    """One training step — forward, loss, backward, optimizer.step() — captured into ONE CUDA graph and replayed
    per batch: the ~650 kernel launches of a step (eeseg conv / BatchNorm / loss kernels, optimizer) cost one
    graph launch instead of Python + driver time each, which is what bounds the eager step once the kernels are
    fast (24 ms wall for 20 ms of GPU work at batch 4, 513x513).

    Same arithmetic as the body of `train_epoch` (train_funcs.py:22-27) for a fixed batch shape. The warm-up
    steps needed before capture run on the example batch and are rolled back (parameters, BatchNorm buffers,
    momentum), so the first replay is the first real update. With the eeseg `SGD` (make_optimizer on CUDA) the
    learning rates are read from device memory at run time, so `scheduler.step()` needs nothing; torch's optimizers
    bake them into the graph: call `recapture()` after `scheduler.step()` (the reference steps it once per epoch)."""

    def __init__(self, net, loss, optimizer, X, y, warmup=3, data_parallel=None):
        """data_parallel (default: torch.distributed initialised with more than one rank): every parameter's .grad is a
        view into ONE flat fp32 buffer, which the step averages over the ranks with a single NCCL all-reduce captured in
        the graph between backward and optimizer.step() — the reference's update at the per-GPU batch on every rank, with
        identical parameters everywhere (ranks must start from identical weights). BatchNorm stays per rank.
        Call `release()` (or drop the object) before `destroy_process_group()`: a live graph keeps the communicator busy."""
        import torch.distributed as dist
        self.net, self.loss_fn, self.opt = net, loss, optimizer
        self.X, self.y = X.clone(), y.clone()
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.dp = (self.world > 1) if data_parallel is None else bool(data_parallel)
        self.flat = self._fg = None
        net.train()
        if getattr(optimizer, 'flat_grads', None) is not None:      # the eeseg SGD owns the flat gradient buffer
            self._fg = optimizer.flat_grads
            self.flat = self._fg.flat
        elif self.dp:
            from .parallel import FlatGradients
            self._fg = FlatGradients([p for g in optimizer.param_groups for p in g['params']])
            self.flat = self._fg.flat
        saved = [t.detach().clone() for t in list(net.parameters()) + list(net.buffers())]
        cur = tch.cuda.current_stream()
        side = tch.cuda.Stream()
        side.wait_stream(cur)
        with tch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        cur.wait_stream(side)
        with tch.no_grad():
            for t, s in zip(list(net.parameters()) + list(net.buffers()), saved):
                t.copy_(s)
            for st in optimizer.state.values():          # momentum buffers must exist before capture: zero them
                for v in st.values():
                    if isinstance(v, tch.Tensor):
                        v.zero_()
        self.recapture()

    def _step(self):
        if self.flat is not None:
            self.flat.zero_()                            # gradients accumulate in place into the flat buffer
        else:
            self.opt.zero_grad(set_to_none=True)
        l = self.loss_fn(self.net(self.X), self.y)
        if self._fg is not None and self.dp:
            self._fg.begin()                             # bucketed exchange: each bucket's NCCL all-reduce starts on a side
        l.mean().backward()                              # stream as soon as its gradients are complete, under the backward
        if self._fg is not None and self.dp:
            self._fg.finish()                            # joins the side stream (none of this on one rank)
        self.opt.step()
        return l

    def recapture(self):
        self.graph = tch.cuda.CUDAGraph()
        if self.flat is None:
            self.opt.zero_grad(set_to_none=True)
        # NCCL's watchdog thread may touch the CUDA API while this thread captures
        kw = {'capture_error_mode': 'thread_local'} if self.world > 1 else {}
        with tch.cuda.graph(self.graph, **kw):
            self.loss = self._step().detach()

    def release(self):
        """Destroys the captured graph (it references the NCCL communicator when data parallel)."""
        tch.cuda.synchronize()
        self.graph.reset()
        self.graph = None

    def __call__(self, X, y):
        self.X.copy_(X, non_blocking=True)
        self.y.copy_(y.view_as(self.y), non_blocking=True)
        if hasattr(self.opt, 'sync_lrs'):
            self.opt.sync_lrs()                          # a scheduler's new learning rates reach the captured update
        self.graph.replay()
        # a replay updates the parameters without touching their version counters: announce it, so that inference plans
        # and graphs derived from the old values (model, engines, operators) are dropped
        bump = getattr(getattr(self.net, 'module', self.net), '_bump_epoch', None)
        if bump is not None:
            bump()
        return self.loss
