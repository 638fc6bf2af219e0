"""TEST INFRASTRUCTURE — CPU port of the reference's model + entropy-gated evaluation path, used as
the oracle for the model tests and as bench.py's `cpu_baseline` / `--impl reference` arm (the
reference itself is Python under /root/reference and cannot travel to the GPU box).

Restates, on the same third-party calls the reference makes (torchvision DeepLabHead / ResNet,
torch F.interpolate), with fp32 CPU tensors:
  * branchyDeepv3.__init__ section split + forward        (from_deepv3_new.py:57-97, 143-155)
  * br_evaluator's per-image gate and accumulation         (eval_br_ent.py:51-70)
Pinned against the real reference by tests/golden/model.npz (tests/test_model_port_cpu.py).
"""
import copy
import re

import numpy as np
import torch
import torchvision
from torch import nn
from torch.nn import functional as F
from torchvision.models.segmentation.deeplabv3 import DeepLabHead

from . import restate as R


def backbone_units(backbone):
    units, input_layers = [], True
    for name, mod in list(backbone.named_modules())[1:]:
        if input_layers and not re.match(r'layer', name):
            units.append(copy.deepcopy(mod))
        elif re.match(r'layer[0-9]+.[0-9]+$', name):
            units.append(mod)
        else:
            input_layers = False
    return units


class BranchyPort(nn.Module):
    """Same module tree / state-dict keys as the reference model for a given section split."""

    def __init__(self, base, sections, num_classes=21):
        super().__init__()
        self.classifier = copy.deepcopy(base.classifier)
        units = backbone_units(base.backbone)
        assert sum(sections) == len(units)
        secs, brs, pos = [], [], 0
        for k, ln in enumerate(sections):
            sec = units[pos:pos + ln]
            pos += ln
            secs.append(nn.Sequential(*sec))
            if k < len(sections) - 1:
                cout = [m for m in nn.Sequential(*sec).modules() if isinstance(m, nn.Conv2d)][-1].out_channels
                brs.append(DeepLabHead(cout, num_classes))
        self.base_model = nn.ModuleList(secs)
        self.branches = nn.ModuleList(brs)
        self.n_branches = len(brs)

    def forward(self, X):
        outputs = []
        inp_shape = X.shape[-2:]
        for i in range(self.n_branches):
            X = self.base_model[i](X)
            br = self.branches[i](X)
            outputs.append(F.interpolate(br, size=inp_shape, mode='bilinear', align_corners=False).unsqueeze(0))
        y = self.classifier(self.base_model[-1](X))
        outputs.append(F.interpolate(y, size=inp_shape, mode='bilinear', align_corners=False).unsqueeze(0))
        return torch.cat(outputs)


def reinit_branches(branches, seed):
    """Deterministic re-initialisation of the early-exit heads, applied identically to the reference
    model (oracle/make_golden.py) and to the port: the reference ctor draws torch.rand probes between
    seeding and branch creation (from_deepv3_new.py:105,121), so its default init cannot be replayed
    from a seed alone. Conv weights ~ N(0, 1/fan_in), BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running
    stats non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in branches.modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) / fan_in ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)


def build_port(sections, seed=0, branch_seed=None, arch='resnet50', num_classes=21):
    """Base model: torchvision default init under manual_seed(seed) (what the reference pickles at
    `base_name`); early-exit heads: reinit_branches(branch_seed) when given."""
    torch.manual_seed(seed)
    ctor = (torchvision.models.segmentation.deeplabv3_resnet50 if arch == 'resnet50'
            else torchvision.models.segmentation.deeplabv3_resnet101)
    base = ctor(weights=None, weights_backbone=None, num_classes=21, aux_loss=True)
    net = BranchyPort(base, sections, num_classes)
    if branch_seed is not None:
        reinit_branches(net.branches, branch_seed)
    return net


def evaluate_batch_cpu(net, X, y, n_classes, tau, skip=0):
    """One br_evaluator iteration per image (eval_br_ent.py:51-70) on CPU: returns
    (exit index per image, per-exit scores [n_br, N], confusion matrix [N, C+1, C] of the exit taken)."""
    with torch.no_grad():
        y_pred = net(X)
    E, N = y_pred.shape[:2]
    n_br = E - 1
    exits, cms, scores = [], [], np.zeros((n_br, N), np.float32)
    for k in range(N):
        ents = []
        for i in range(n_br):
            probs = F.softmax(y_pred[i, k:k + 1], 1).squeeze(0).numpy()
            ents.append(R.img_norm_entropy(probs, n_classes))
            scores[i, k] = ents[-1]
        ex = R.first_confident_exit(ents, tau, skip)
        exits.append(ex)
        pred = R.argmax_first(y_pred[ex, k].numpy().reshape(1, n_classes, -1), 1)
        cms.append(R.confusion_matrix(pred, y[k].numpy().reshape(1, -1), n_classes)[0])
    return np.array(exits), scores, np.stack(cms)
