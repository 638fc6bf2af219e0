"""Drop-in for the reference's from_deepv3_new.py (and, through from_deepv3.py, its v1 twin):
`my_branch` (:15-39), `get_base_model` (:41-54) and `branchyDeepv3` (:56-155).

Same constructor arguments, attribute names (`base_model`, `branches`, `classifier`, `n_branches`,
`count_branches`) and state-dict layout as the reference, so its checkpoints load. What differs is
how the work runs on a B200:

* `forward(X)` in eval mode on CUDA returns the same `[E,N,C,H,W]` fp32 tensor, but each exit's head
  runs on the eeseg implicit-GEMM kernels (bf16 NHWC, BatchNorm folded; csrc/conv_igemm.cu) and the
  bilinear up-sampling kernel writes straight into the stacked output (no `unsqueeze`+`cat` copy,
  from_deepv3_new.py:150-155).
* `forward_lowres(X)` stops before the up-sampling and feeds the fused exit gate
  (`ee_semantic_segmentation_b200.ops.exit_gate`), which never materialises full-resolution logits.
* training mode runs the convolutions (forward, input and weight gradients), BatchNorm and up-sampling on the
  eeseg kernels behind autograd (head_train.py, backbone_train.py, bn_train.py).
* `strict_kernels = True` turns every library fall-back (a section or head layout the kernel plans do not cover)
  into an error; by default such a module runs on the PyTorch modules with ONE warning naming it.

Branch placement follows from_deepv3_new.py:75-89 literally; FLOPs are counted with
torch.utils.flop_counter on meta tensors (the reference's `pthflops` is not installed anywhere and
is unpinned, so placement parity is "unpinned" — pass `sections=[...]` to copy a reference split).
"""
import copy
import os
import re
import warnings

import torch as tch
import torchvision
from torch import nn
from torch.nn import functional as F
from torch.utils.flop_counter import FlopCounterMode
from torchvision.models.segmentation.deeplabv3 import ASPP, DeepLabHead

from . import backbone_train, head_train, ops
from .backbone_plan import SectionPlan, supported as _section_supported
from .head_plan import HeadPlan


class my_branch(nn.Sequential):
    def __init__(self, nin_channels, num_classes, atrous_rates, nout_channels, bottleneck=None, **kwargs):
        if bottleneck:
            super().__init__(
                nn.Conv2d(nin_channels, bottleneck, 1),
                ASPP(bottleneck, atrous_rates, nout_channels),
                nn.Conv2d(nout_channels, nout_channels, 3, padding=1, bias=False),
                nn.BatchNorm2d(nout_channels),
                nn.ReLU(),
                nn.Conv2d(nout_channels, num_classes, 1),
            )
        else:
            super().__init__(
                ASPP(nin_channels, atrous_rates, nout_channels),
                nn.Conv2d(nout_channels, nout_channels, 3, padding=1, bias=False),
                nn.BatchNorm2d(nout_channels),
                nn.ReLU(),
                nn.Conv2d(nout_channels, num_classes, 1),
            )


def get_base_model(name, model='deeplabv3_resnet101', pretrained=True):
    """from_deepv3_new.py:41-54: load the pickled whole module at `name`, else build the torchvision
    model and save it there. Without network access pretrained weights cannot be fetched; the
    model is then random-initialised (with a warning) instead of failing."""
    trained_model = None
    if name is not None and os.path.exists(name):
        try:
            trained_model = tch.load(name, weights_only=False)
        except Exception:
            trained_model = None
    if trained_model is None:
        if not re.search('deeplabv3', model):
            raise ValueError(f'unsupported base model {model}')
        ctor = (torchvision.models.segmentation.deeplabv3_resnet50 if re.search('resnet50', model)
                else torchvision.models.segmentation.deeplabv3_resnet101)
        if pretrained:
            try:
                trained_model = ctor(weights='DEFAULT')
            except Exception as err:  # no egress in this environment
                warnings.warn(f'pretrained weights unavailable ({type(err).__name__}); using random init')
        if trained_model is None:
            trained_model = ctor(weights=None, weights_backbone=None, num_classes=21, aux_loss=True)
        if name is not None and not os.path.exists(name):
            tch.save(trained_model, name)
    return trained_model


def _meta_flops(module, x):
    was = module.training
    module.eval()
    try:
        with tch.no_grad(), FlopCounterMode(display=False) as fc:
            y = module(x)
    finally:
        module.train(was)
    return fc.get_total_flops(), y


class branchyDeepv3(nn.Module):
    def __init__(self, base_name, base_type, n, img_dim, count_branches=True, skip=0,
                 branch_params=None, *, num_classes=21, sections=None, pretrained=True):
        super().__init__()
        aux_model = get_base_model(base_name, base_type, pretrained)
        self.classifier = copy.deepcopy(aux_model.classifier)
        self.count_branches = count_branches
        self.num_classes = num_classes
        if num_classes != self.classifier[-1].out_channels:
            # the reference hard-codes 21 classes (from_deepv3_new.py:86,131); other label sets get a
            # fresh final 1x1 classifier
            self.classifier[-1] = nn.Conv2d(self.classifier[-1].in_channels, num_classes, 1)
        self._branch_params = branch_params

        modules = self._backbone_units(aux_model.backbone)
        if sections is None:
            sections = self._place(aux_model.backbone, n, img_dim, count_branches, skip, branch_params)
        assert sum(sections) == len(modules), (sections, len(modules))
        base, branches, pos = [], [], 0
        for k, ln in enumerate(sections):
            sec = modules[pos:pos + ln]
            pos += ln
            base.append(nn.Sequential(*sec))
            if k < len(sections) - 1:
                branches.append(self._gen_branch(self._out_channels(sec), num_classes, branch_params))
        self.base_model = nn.ModuleList(base)
        self.branches = nn.ModuleList(branches)
        self.n_branches = len(self.branches)
        # __init_branches (from_deepv3_new.py:133-140) is a no-op in the reference (get_layers always
        # returns []), so the branches keep PyTorch's default initialisation.
        self._plans = {}
        self._section_plans = {}
        self.fast_inference = True     # eval-mode CUDA forward on the eeseg kernels
        self.fast_backbone = True      # ... including the ResNet bottlenecks of the sections
        self.fast_training_heads = True   # autograd forward: head convolutions (fwd/dgrad/wgrad) on eeseg kernels
        self.fast_training_backbone = True   # ... and the Bottleneck convolutions of the sections (bf16 activations)
        self.graph_inference = True    # forward_lowres replays one CUDA graph per input shape (see there)
        self._lowres_graphs = {}
        self._token_tensors = None
        self.weights_epoch = 0         # bumped by train() / load_state_dict(): captured graphs of older epochs are stale
        self.input_norm = None         # (mean[3], std[3]) applied inside the stem kernel to uint8 images (0..255)
        self.strict_kernels = False    # True: a module the eeseg plans do not cover raises instead of running on cuDNN

    _RUNTIME_DEFAULTS = dict(fast_inference=True, fast_backbone=True, fast_training_heads=True,
                             fast_training_backbone=True, graph_inference=True, weights_epoch=0, num_classes=21,
                             strict_kernels=False, input_norm=None)
    _warned_fallbacks = set()

    def _library_fallback(self, what):
        """A module the eeseg kernel plans do not cover: error in strict mode (bench.py and the GPU tests set it), one
        warning per module kind otherwise. There is never a silent switch of backend."""
        if self.strict_kernels:
            raise RuntimeError(f'strict_kernels: {what} is not covered by the eeseg kernel plans '
                               '(it would run on the PyTorch/cuDNN modules)')
        if what not in branchyDeepv3._warned_fallbacks:
            branchyDeepv3._warned_fallbacks.add(what)
            warnings.warn(f'{what} is not covered by the eeseg kernel plans: running it on the PyTorch modules')

    def __getstate__(self):
        """Pickles (tch.save(net), copy.deepcopy) carry parameters and flags, not the kernel plans / CUDA graphs."""
        st = dict(self.__dict__)
        st['_plans'], st['_section_plans'], st['_lowres_graphs'] = {}, {}, {}
        st['_token_tensors'] = None
        st.pop('_train_wcache', None)
        return st

    def __setstate__(self, st):
        """Also accepts whole-module pickles written by the reference class (eval_br_ent.py:146) or by an older torch:
        nn.Module.__setstate__ restores the hook dictionaries such pickles lack; the run-time attributes the reference
        class does not have take their defaults."""
        super().__setstate__(st)
        for k in ('_plans', '_section_plans', '_lowres_graphs'):
            self.__dict__[k] = {}
        self.__dict__['_token_tensors'] = None
        for k, v in self._RUNTIME_DEFAULTS.items():
            self.__dict__.setdefault(k, v)

    def _bump_epoch(self):
        """Drops everything derived from the parameters: CUDA graphs AND the folded kernel plans. The plans are keyed by
        the parameters' version counters, which a replayed training graph (train_funcs.GraphedTrainStep) does not advance —
        GraphedTrainStep therefore calls this after every replay, as do train(), load_state_dict() and .to()/.cuda()."""
        self.weights_epoch = getattr(self, 'weights_epoch', 0) + 1
        self._lowres_graphs = {}
        self._plans = {}
        self._section_plans = {}
        self._token_tensors = None

    def weights_token(self):
        """Cheap fingerprint of the weights every captured CUDA graph is tagged with (here, engine.EarlyExitEngine,
        ee_dnn_op_ne.eval_ee_deeplabv3) and compares before EACH replay: (weights_epoch, sum of the version counters of
        all parameters and buffers). The version counters catch what never reaches this class — a sub-module
        load_state_dict (net.branches.load_state_dict(...)), an optimizer step or any in-place edit in eval mode;
        weights_epoch catches what leaves the counters alone (graph-replayed training, .to()). ~30 us of host time."""
        ts = self.__dict__.get('_token_tensors')
        if ts is None:
            ts = list(self.parameters()) + list(self.buffers())
            self.__dict__['_token_tensors'] = ts
        v = 0
        for t in ts:
            v += t._version
        return (self.weights_epoch, v)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)        # .to() / .cuda() / .float(): new storages
        if hasattr(self, 'weights_epoch'):
            self._bump_epoch()
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        self._bump_epoch()                       # load_state_dict (also through a parent module / DDP wrapper)
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        """nn.Module.train; entering training mode announces parameter updates: CUDA graphs captured from the folded
        inference plans (here, in engine.EarlyExitEngine and ee_dnn_op_ne.eval_ee_deeplabv3) are keyed by
        `weights_epoch` and re-captured afterwards. Parameters edited in place while in eval mode need an explicit
        `_bump_epoch()`."""
        if mode and hasattr(self, 'weights_epoch'):
            self._bump_epoch()
        return super().train(mode)

    # ---- construction helpers ---------------------------------------------------------------------
    @staticmethod
    def _backbone_units(backbone):
        """Stem modules (deep-copied like :77-78) followed by every Bottleneck `layerX.Y` (:80-82)."""
        units, input_layers = [], True
        for name, mod in list(backbone.named_modules())[1:]:
            if input_layers and not re.match(r'layer', name):
                units.append(copy.deepcopy(mod))
            elif re.match(r'layer[0-9]+.[0-9]+$', name):
                units.append(mod)
            else:
                input_layers = False
        return units

    @staticmethod
    def _out_channels(section):
        for m in reversed(list(nn.Sequential(*section).modules())):
            if isinstance(m, nn.Conv2d):
                # last conv of a Bottleneck is conv3 unless a downsample conv follows it in module
                # order; both have the same out_channels
                return m.out_channels
        raise ValueError('section has no convolution')

    @staticmethod
    def _gen_branch(nin_channels, nout_channels=21, branch_params=None):
        if isinstance(branch_params, dict) and all(k in branch_params for k in ('nout_channels', 'atrous_rates')):
            return my_branch(nin_channels=nin_channels, num_classes=nout_channels, **branch_params)
        return DeepLabHead(nin_channels, nout_channels)

    @classmethod
    def _place(cls, backbone, n, img_dim, count_branches, skip, branch_params):
        """The FLOP-quantile walk of from_deepv3_new.py:66-91 on a meta-device copy of the backbone."""
        meta = copy.deepcopy(backbone).to('meta')
        units = cls._backbone_units(meta)
        x0 = tch.empty(1, 3, img_dim, img_dim, device='meta')
        tot_flops, _ = _meta_flops(meta, x0)
        flop_pos = tot_flops / (n + 1)
        sections, meta_sections, meta_branches, cur = [], [], [], []
        # stem modules never trigger a placement check in the reference (only Bottlenecks do)
        stem_len = 0
        for name, _ in list(meta.named_modules())[1:]:
            if re.match(r'layer', name):
                break
            stem_len += 1

        def check_flops(upto):
            fl, _ = _meta_flops(nn.Sequential(*units[:upto]), x0)
            if meta_branches and count_branches:
                t = x0
                for sec, br in zip(meta_sections, meta_branches):
                    with tch.no_grad():
                        t = sec.eval()(t)
                    bf, _ = _meta_flops(br, t)
                    fl += bf
            return fl

        for idx in range(len(units)):
            cur.append(units[idx])
            if idx < stem_len:
                continue
            nb = len(meta_branches)
            if n > nb:
                fl = check_flops(idx + 1)
                if tot_flops > fl > flop_pos * (nb + (1 + skip)):
                    sections.append(len(cur))
                    meta_sections.append(nn.Sequential(*cur))
                    br = cls._gen_branch(cls._out_channels(cur), 21, branch_params).to('meta')
                    meta_branches.append(br)
                    cur = []
        sections.append(len(cur))
        return sections

    # ---- forward ----------------------------------------------------------------------------------
    def _plan(self, i):
        """Folded-BN kernel plan of head i (i == n_branches -> classifier); rebuilt when parameters
        change (tracked through their version counters)."""
        head = self.classifier if i == self.n_branches else self.branches[i]
        key = self._state_key(head)
        ent = self._plans.get(i)
        if ent is None or ent[0] != key:
            ent = (key, HeadPlan(head))
            self._plans[i] = ent
        return ent[1]

    @staticmethod
    def _state_key(mod):
        return tuple((p.data_ptr(), p._version) for p in mod.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in mod.buffers())

    def run_section(self, i, X):
        """Inference forward of base_model[i]: Bottlenecks on the eeseg conv kernel (BN/ReLU/residual
        fused) when fast_backbone is set, else the PyTorch modules under bf16 autocast."""
        sec = self.base_model[i]
        if self.fast_backbone and _section_supported(sec):
            key = self._state_key(sec)
            ent = self._section_plans.get(i)
            if ent is None or ent[0] != key:
                ent = (key, SectionPlan(sec))
                self._section_plans[i] = ent
            return ent[1].run(X, getattr(self, 'input_norm', None))
        if self.fast_backbone:
            self._library_fallback(f'base_model[{i}] ({type(sec[0]).__name__} ...)')
        with tch.autocast('cuda', dtype=tch.bfloat16):
            return sec(X.contiguous(memory_format=tch.channels_last))

    def _head_autograd(self, head, X):
        """head(X) with autograd: the head's convolutions (forward, input and weight gradients) on the
        eeseg tcgen05 kernels when fast_training_heads is set and the head has the DeepLabHead layout,
        else the PyTorch modules (cuDNN)."""
        if self.fast_training_heads and self.training and X.is_cuda:
            if head_train.head_supported(head):
                return head_train.head_forward_train(head, X)
            self._library_fallback(f'exit head {type(head).__name__} (training)')
        return head(X)

    def _section_autograd(self, section, X):
        """base_model[i](X) with autograd: Bottleneck convolutions on the eeseg kernels in bf16 when
        fast_training_backbone is set (backbone_train.py), else the PyTorch modules in the input dtype."""
        if self.fast_training_backbone and self.training and X.is_cuda:
            if not backbone_train.section_supported(section, X):
                self._library_fallback(f'a unit of a backbone section ({type(section[0]).__name__} ...) in training')
            return backbone_train.section_forward_train(section, X)
        return section(X)

    def _upsample_autograd(self, y, size):
        """F.interpolate(bilinear, align_corners=False) with autograd: eeseg kernels (fused forward, gather-form
        deterministic backward) on CUDA when fast_training_heads is set, else ATen."""
        if self.fast_training_heads and self.training and y.is_cuda:
            return ops.upsample_bilinear_autograd(y, size)
        return F.interpolate(y, size=size, mode='bilinear', align_corners=False)

    def _forward_torch(self, X):
        """The reference data flow with autograd (training): backbone sections on the PyTorch modules,
        exit heads through _head_autograd."""
        import contextlib
        wcache = contextlib.nullcontext()
        if self.training and X.is_cuda and (self.fast_training_heads or self.fast_training_backbone):
            # bf16 copies of every conv weight in both layouts the step reads: one launch per step for all layers
            wc = self.__dict__.get('_train_wcache')
            if wc is None or not wc.valid():
                convs = head_train.trainable_convs(self)
                wc = head_train.TrainWeightCache(convs) if convs else None
                self.__dict__['_train_wcache'] = wc
            if wc is not None:
                wc.refresh()
                wcache = wc.active()
        outputs = []
        inp_shape = X.shape[-2:]
        with wcache:
            for i in range(self.n_branches):
                X = self._section_autograd(self.base_model[i], X)
                br = self._head_autograd(self.branches[i], X)
                outputs.append(self._upsample_autograd(br, inp_shape).unsqueeze(0))
            y = self._head_autograd(self.classifier, self._section_autograd(self.base_model[-1], X))
            outputs.append(self._upsample_autograd(y, inp_shape).unsqueeze(0))
        return tch.cat(outputs)

    def _use_fast(self, X):
        return self.fast_inference and X.is_cuda and not self.training and not tch.is_grad_enabled()

    def forward_lowres(self, X, active=None):
        """Low-resolution logits of every exit: list of E tensors [N,h,w,Cp] (fp32, NHWC, Cp >= C
        padded to the kernel's column multiple). Inference only."""
        if not X.is_cuda:
            raise RuntimeError('branchyDeepv3 fast path needs CUDA tensors (no CPU fallback)')
        if self.graph_inference and not tch.cuda.is_current_stream_capturing():
            return self._lowres_graph(X)
        return self._lowres_eager(X)

    def _lowres_eager(self, X):
        outs = []
        with tch.no_grad():
            for i in range(self.n_branches + 1):
                X = self.run_section(i, X)
                outs.append(self._plan(i).run(X))
        return outs

    def _lowres_graph(self, X):
        """The ~700 launches of a forward replayed as one CUDA graph per (input shape, device): evaluators that feed
        one image at a time (br_evaluator, mIoU_evaluator: batch-1 loaders as in the reference) are launch-bound
        otherwise. The returned tensors are the graph's static outputs: valid until the next call with that shape."""
        key = (tuple(X.shape), X.device, X.dtype)
        ent = self._lowres_graphs.get(key)
        token = self.weights_token()
        if ent is not None and ent[3] != token:
            # the weights moved since capture (sub-module load, in-place edit, optimizer step): the graph replays plans
            # folded from the old values — drop it and capture again from the current ones
            self._lowres_graphs.pop(key)
            ent = None
            self._lowres_graphs.setdefault('seen', {})[key] = 2
        if ent is None:
            # capture a shape the second time it shows up: a loader of varying image sizes must not pay a capture per image
            seen = self._lowres_graphs.setdefault('seen', {})
            seen[key] = seen.get(key, 0) + 1
            if seen[key] < 2:
                if len(seen) > 64:
                    seen.clear()
                return self._lowres_eager(X)
        with tch.cuda.device(X.device):
            if ent is None:
                xs = tch.zeros_like(X)
                xs.copy_(X)
                side = tch.cuda.Stream()
                side.wait_stream(tch.cuda.current_stream())
                with tch.cuda.stream(side):
                    for _ in range(2):
                        self._lowres_eager(xs)
                tch.cuda.current_stream().wait_stream(side)
                g = tch.cuda.CUDAGraph()
                with tch.cuda.graph(g):
                    outs = self._lowres_eager(xs)
                # the plans' folded weights are allocated outside the graph pool: the entry keeps them alive as long as
                # the graph can replay (a later _plan(i) may re-key and drop them from self._plans)
                ent = (g, xs, outs, token, (dict(self._plans), dict(self._section_plans)))
                graphs = [k for k in self._lowres_graphs if k != 'seen']
                if len(graphs) >= 8:                       # a few shapes at most: drop the oldest
                    self._lowres_graphs.pop(graphs[0])
                self._lowres_graphs[key] = ent
            g, xs, outs = ent[:3]
            xs.copy_(X, non_blocking=True)
            g.replay()
        return outs

    def forward(self, X):
        if not self._use_fast(X):
            if not X.is_cuda:
                raise RuntimeError('branchyDeepv3 runs on CUDA tensors only (no CPU fallback); '
                                   'the CPU oracle lives under oracle/')
            return self._forward_torch(X)
        H, W = X.shape[-2:]
        lows = self.forward_lowres(X)
        N = X.shape[0]
        out = tch.empty((len(lows), N, self.num_classes, H, W), dtype=tch.float32, device=X.device)
        for e, lo in enumerate(lows):
            ops.upsample_bilinear(lo, (H, W), out=out[e], layout='NHWC', n_classes=self.num_classes)
        return out
