"""Drop-in for the step-level part of the reference's training driver: `train_epoch`
(train_funcs.py:12-33) plus the optimiser / schedule construction of `train_deepv3`
(deepv3_funcs.py:74-101,138-156) as small helpers, with one-process-per-GPU data parallelism.

The forward/backward of the convolutions runs on the PyTorch modules (autograd through cuDNN; conv
dgrad/wgrad kernels are future work); the multi-exit loss (forward + fused backward) runs on the
eeseg kernels. Under `torchrun`, wrap the model with `parallel.wrap_ddp` — gradients are averaged
by NCCL bucket all-reduces overlapped with the backward pass; BatchNorm stays per rank, as in a
single-GPU reference run at the per-GPU batch."""
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


def _sgd_impl(net):
    """torch's single-kernel ("fused") SGD update when every parameter lives on a GPU: the same update rule as the
    default multi-tensor implementation in one pass over parameters, gradients and momentum buffers."""
    return {'fused': True} if all(p.is_cuda for p in net.parameters()) else {}


def make_optimizer(net, lr, base_lr=None, weighted_lr=False):
    """SGD(momentum 0.9, weight decay 5e-4) with the reference's parameter groups
    (deepv3_funcs.py:77-101): backbone at base_lr, branches at lr, final classifier at 1.1*lr."""
    net = getattr(net, 'module', net)      # DDP
    if base_lr and getattr(net, 'n_branches', 0):
        params = [{'params': net.base_model.parameters(), 'lr': base_lr}]
        if weighted_lr:
            import numpy as np
            w = np.linspace(1, 1.2, num=net.n_branches + 1)
            params += [{'params': net.branches[i].parameters(), 'lr': lr * w[i]} for i in range(net.n_branches)]
            params.append({'params': net.classifier.parameters(), 'lr': lr * w[-1]})
        else:
            params.append({'params': net.branches.parameters(), 'lr': lr})
            params.append({'params': net.classifier.parameters(), 'lr': lr * 1.1})
        return optim.SGD(params, lr=lr, momentum=.9, weight_decay=5e-4, **_sgd_impl(net))
    return optim.SGD(net.parameters(), lr=lr, momentum=.9, weight_decay=5e-4, **_sgd_impl(net))


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
    momentum), so the first replay is the first real update. Learning-rate changes made by a scheduler are baked
    into the graph: call `recapture()` after `scheduler.step()` (the reference steps it once per epoch)."""

    def __init__(self, net, loss, optimizer, X, y, warmup=3):
        self.net, self.loss_fn, self.opt = net, loss, optimizer
        self.X, self.y = X.clone(), y.clone()
        net.train()
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
        self.opt.zero_grad(set_to_none=True)
        l = self.loss_fn(self.net(self.X), self.y)
        l.mean().backward()
        self.opt.step()
        return l

    def recapture(self):
        self.graph = tch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with tch.cuda.graph(self.graph):
            self.loss = self._step().detach()

    def __call__(self, X, y):
        self.X.copy_(X, non_blocking=True)
        self.y.copy_(y.view_as(self.y), non_blocking=True)
        self.graph.replay()
        return self.loss
