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


def sharpen_heads(branches, factors):
    """Scales the final 1x1 classifier (weight and bias) of early-exit head i by factors[i]: larger logits = lower
    entropy, so fixtures can order the exits' confidences (applied identically to the reference model in
    oracle/make_golden_operator.py and to the port)."""
    with torch.no_grad():
        for br, f in zip(branches, factors):
            br[-1].weight.mul_(float(f))
            br[-1].bias.mul_(float(f))


def build_port(sections, seed=0, branch_seed=None, arch='resnet50', num_classes=21, sharpen=None):
    """Base model: torchvision default init under manual_seed(seed) (what the reference pickles at
    `base_name`); early-exit heads: reinit_branches(branch_seed) when given, then sharpen_heads(sharpen)."""
    torch.manual_seed(seed)
    ctor = (torchvision.models.segmentation.deeplabv3_resnet50 if arch == 'resnet50'
            else torchvision.models.segmentation.deeplabv3_resnet101)
    base = ctor(weights=None, weights_backbone=None, num_classes=21, aux_loss=True)
    net = BranchyPort(base, sections, num_classes)
    if branch_seed is not None:
        reinit_branches(net.branches, branch_seed)
    if sharpen is not None:
        sharpen_heads(net.branches, sharpen)
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


def operator_entropy_cpu(net, x, n_classes, th, less_than=True, ignore=()):
    """ee_dnn_op_ne.eval_ee_deeplabv3.__call__ (ee_dnn_op_ne.py:51-108) on CPU for one image x [3,H,W], without the
    FLOP bookkeeping: returns {'n', 'exit', 'last'} (int64 [H,W] maps) plus 'scores' (the metric value of every
    early exit that was evaluated, None for the others)."""
    out, scores, left = {}, [], False
    inp_shape = x.shape[-2:]
    X = x.unsqueeze(0)
    with torch.no_grad():
        for i in range(net.n_branches):
            X = net.base_model[i](X)
            scores.append(None)
            if i not in ignore and not left:
                br = F.interpolate(net.branches[i](X), size=inp_shape, mode='bilinear', align_corners=False)
                probs = F.softmax(br, 1).squeeze(0).numpy()
                t = R.img_norm_entropy(probs, n_classes)
                scores[-1] = float(t)
                if (t < th) if less_than else (t > th):
                    out['exit'] = br.argmax(dim=1).squeeze(0)
                    out['n'] = i + 1
                    left = True
        Y = F.interpolate(net.classifier(net.base_model[-1](X)), size=inp_shape, mode='bilinear', align_corners=False)
        Y = Y.argmax(dim=1).squeeze(0)
    out['last'] = Y
    if not left:
        out['exit'] = Y
        out['n'] = net.n_branches + 1
    out['scores'] = scores
    return out


def operator_similarity_cpu(net, x, metric, th, less_than=True, ignore=()):
    """ee_dnn_op.eval_ee_deeplabv3.__call__ (ee_dnn_op.py:51-118) on CPU for one image, without the FLOP bookkeeping:
    the first evaluated exit only provides the reference map; a later exit answers when
    metric(reference map, its map) crosses th, otherwise its map becomes the reference map."""
    out, left, y_ref = {}, False, None
    inp_shape = x.shape[-2:]
    X = x.unsqueeze(0)
    with torch.no_grad():
        for i in range(net.n_branches):
            X = net.base_model[i](X)
            if i not in ignore and not left:
                br = F.interpolate(net.branches[i](X), size=inp_shape, mode='bilinear', align_corners=False).argmax(dim=1)
                if y_ref is not None and ((metric(y_ref, br) < th) if less_than else (metric(y_ref, br) > th)):
                    out['exit'] = br.squeeze(0)
                    out['n'] = i + 1
                    left = True
                else:
                    y_ref = br
        Y = F.interpolate(net.classifier(net.base_model[-1](X)), size=inp_shape, mode='bilinear', align_corners=False)
        Y = Y.argmax(dim=1).squeeze(0)
    out['last'] = Y
    if not left:
        out['exit'] = Y
        out['n'] = net.n_branches + 1
    return out
