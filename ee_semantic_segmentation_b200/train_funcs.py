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
        return optim.SGD(params, lr=lr, momentum=.9, weight_decay=5e-4)
    return optim.SGD(net.parameters(), lr=lr, momentum=.9, weight_decay=5e-4)


def poly_scheduler(optimizer, num_epochs, lr=None, min_lr=None):
    """Per-epoch poly schedule (1 - k/num_epochs)^0.9 (deepv3_funcs.py:146-152)."""
    if min_lr and lr:
        w = (min_lr / lr) ** (1 / .9)
        n0 = num_epochs * w / (1 - w)
        return optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda k: (1 - k / (num_epochs + n0)) ** .9)
    return optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda k: (1 - k / num_epochs) ** .9)
