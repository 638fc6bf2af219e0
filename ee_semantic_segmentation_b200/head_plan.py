"""Inference plan of one DeepLabHead / my_branch exit head on the eeseg kernels.

Takes the torchvision module (parameter container, reference layout `branches.{i}.0.convs.*`),
folds every BatchNorm into a per-channel scale/shift (projection and 3x3: scale into the bf16 weights, FOLD_SCALE), converts
weights once to bf16 [Cout][R][S][Cin] and runs the head as implicit-GEMM launches (csrc/conv_igemm.cu) on NHWC bf16 activations:

    ASPP 1x1 + three atrous 3x3 -> ONE grouped launch over a cost-sorted work list (CTA pairs for an even batch), written
    side by side into one [N,h,w,4*256] buffer (no concat),
    pooled branch -> global-avg-pool kernel + tiny matvec, folded into the projection as a
    per-image shift (the broadcast "bilinear" up-sampling of a 1x1 map is a constant),
    projection 1x1 (K = 4*256) -> 3x3 -> final 1x1 (+bias) to fp32 logits [N,h,w,Cp].

Mirrors torchvision.models.segmentation.deeplabv3.{DeepLabHead,ASPP,ASPPConv,ASPPPooling} as called
from from_deepv3_new.py:34,131,147,151. Dropout(0.5) in ASPP.project is the identity in eval mode.
"""
import os

import torch
from torch import nn
from torchvision.models.segmentation.deeplabv3 import ASPP

from . import _lib, torch_ops  # noqa: F401  (torch_ops registers torch.ops.eeseg.*)
from ._lib import check, lib


def _fold_bn(bn):
    """y = gamma*(x-mean)/sqrt(var+eps)+beta  ->  scale, shift (fp32)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _krsc(conv):
    """[Cout,Cin,R,S] -> bf16 [Cout,R,S,Cin] contiguous."""
    return conv.weight.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


PROFILE = None   # tools set this to a list: (start event, end event, nominal FLOPs, tag) per conv launch
PROFILE_TAG = "head"
FLOP_LOG = None  # bench.py sets this to a list: (nominal FLOPs, effective FLOPs, tag) per conv launch, no events
_LIVE_FRAC = {}


def live_tap_fraction(h, w, k, dil, stride=1, pad=None):
    """Fraction of the (output pixel, tap) pairs of a k x k convolution the kernel actually multiplies: taps that fall
    entirely into the zero padding for a whole tile are skipped (conv_igemm.cu live_taps; common for the ASPP dilations
    12/24/36 on a 65x65 map, never for 3x3 d <= 4). h, w: OUTPUT size. effective FLOPs = nominal x this."""
    import ctypes
    import numpy as np
    if k == 1:
        return 1.0
    key = (h, w, k, dil, stride, pad)
    if key not in _LIVE_FRAC:
        tx, ty, bw, bh, bn = (ctypes.c_int() for _ in range(5))
        check(lib().eeseg_conv_group_tiles(h, w, 256, ctypes.byref(tx), ctypes.byref(ty), ctypes.byref(bw),
                                           ctypes.byref(bh), ctypes.byref(bn)), "eeseg_conv_group_tiles")
        bw, bh = bw.value, bh.value
        p = dil * (k // 2) if pad is None or pad < 0 else pad
        hin, win = (h - 1) * stride + 1, (w - 1) * stride + 1       # smallest input that gives this output ('same' convs: equal)
        off = np.arange(k) * dil - p
        y0, x0 = np.arange(0, h, bh), np.arange(0, w, bw)
        y_hi, x_hi = np.minimum(y0 + bh, h), np.minimum(x0 + bw, w)
        live_y = (((y_hi[:, None] - 1) * stride + off[None, :] >= 0) & (y0[:, None] * stride + off[None, :] < hin)).sum(1)
        live_x = (((x_hi[:, None] - 1) * stride + off[None, :] >= 0) & (x0[:, None] * stride + off[None, :] < win)).sum(1)
        px = (y_hi - y0)[:, None] * (x_hi - x0)[None, :]
        _LIVE_FRAC[key] = float((live_y[:, None] * live_x[None, :] * px).sum() / (k * k * px.sum()))
    return _LIVE_FRAC[key]


def conv_igemm(x, wt, scale, shift, dilation, relu, out, out_dtype_code, ldo, shift_sn=0, stride=1,
               residual=None, pad=-1):
    """x bf16 NHWC [N,h,w,Cin]; wt bf16 [Cout,R,S,Cin]; out: tensor view whose data_ptr is the first
    output channel and whose pixel stride is ldo elements; residual: optional bf16 NHWC tensor of
    the output shape (added before the ReLU)."""
    N, h, w, Cin = x.shape
    Cout, R, S, _ = wt.shape
    with torch.cuda.device(x.device):
        if PROFILE is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        assert out_dtype_code == (_lib.F32 if out.dtype == torch.float32 else _lib.BF16)
        torch_ops.fast.conv_igemm_fwd(x, wt, scale, shift, shift_sn, dilation, stride, pad, bool(relu), residual, out, ldo)
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        if PROFILE is not None:
            b.record()
            PROFILE.append((a, b, 2 * N * ho * wo * Cout * R * S * Cin, PROFILE_TAG))
        if FLOP_LOG is not None:
            nominal = 2 * N * ho * wo * Cout * R * S * Cin
            frac = live_tap_fraction(ho, wo, R, dilation, stride, pad) if R == S else 1.0
            FLOP_LOG.append((nominal, nominal * frac, PROFILE_TAG))


def group_schedule(N, h, w, cin, cout, ksizes, dils, n_ctas=148, pairs=False, n_clusters=None):
    """Work list of a grouped conv launch: item = problem << 24 | tile, ordered in passes over IMAGES_PER_PASS images and,
    inside a pass, by decreasing cost (live taps x channel blocks — taps that fall entirely into the zero padding of a tile
    are skipped by the kernel), each round of n_ctas items dealt to the persistent CTAs by their load so far. Returns an
    int32 CPU tensor.

    pairs=True (N even): a PAIR list for clusters of two CTAs (eeseg_conv_igemm_grouped, cta_pairs) — an item is the same
    tile position of images 2k and 2k+1 (identical live taps, so the joint 256-row MMA skips exactly what each tile would
    skip alone); entries 2i, 2i+1 are its two tiles, rounds are n_clusters items."""
    import ctypes
    import numpy as np
    tx, ty, bw, bh, bn = (ctypes.c_int() for _ in range(5))
    check(lib().eeseg_conv_group_tiles(h, w, cout, ctypes.byref(tx), ctypes.byref(ty), ctypes.byref(bw),
                                       ctypes.byref(bh), ctypes.byref(bn)), "eeseg_conv_group_tiles")
    tx, ty, bw, bh, bn = tx.value, ty.value, bw.value, bh.value, bn.value
    # under-filled grid (one or two images): narrower channel tiles double the work items — twice the CTAs busy and a
    # finer-grained balance between the cheap 1x1 tiles and the 9-tap ones; the launcher reads the tile width off the list
    while bn > 64 and len(ksizes) * N * tx * ty * (cout // bn) <= n_ctas:
        bn //= 2
    n_tiles = cout // bn
    if pairs and (N % 2 or bn < 32):
        raise ValueError("group_schedule(pairs=True) needs an even number of images")
    y0 = np.arange(ty) * bh
    x0 = np.arange(tx) * bw
    step = 2 if pairs else 1                 # images per item
    items, costs, imgs = [], [], []
    for g, (k, d) in enumerate(zip(ksizes, dils)):
        pad = d * (k // 2)
        off = np.arange(k) * d - pad
        y_hi = np.minimum(y0 + bh, h)
        x_hi = np.minimum(x0 + bw, w)
        live_y = ((y_hi[:, None] - 1 + off[None, :] >= 0) & (y0[:, None] + off[None, :] < h)).sum(1)   # [ty]
        live_x = ((x_hi[:, None] - 1 + off[None, :] >= 0) & (x0[:, None] + off[None, :] < w)).sum(1)   # [tx]
        taps = live_y[:, None] * live_x[None, :]                                                      # [ty, tx]
        n_it = N // step
        cost = np.broadcast_to(taps.reshape(1, ty * tx, 1), (n_it, ty * tx, n_tiles)).reshape(-1) * (cin // 64)
        first = np.arange(n_it)[:, None, None] * step                                                  # first image of the item
        tile = ((first * (ty * tx) + np.arange(ty * tx)[None, :, None]) * n_tiles + np.arange(n_tiles)[None, None, :])
        items.append((g << 24) | tile.reshape(-1))
        costs.append(cost)
        imgs.append(np.broadcast_to(first, (n_it, ty * tx, n_tiles)).reshape(-1))
    items, costs, img = np.concatenate(items), np.concatenate(costs), np.concatenate(imgs)
    ipp = IMAGES_PER_PASS if IMAGES_PER_PASS else N
    # Passes over `ipp` images at a time: the CTAs then work on the same few images at any moment, so their activation
    # (17 MB per image at Cin = 2048) plus the weights of all problems (28 MB) stay L2-resident across the 27 taps that
    # re-read them, instead of every round of the list pulling all images through the L2 again (ncu: 291 MB of DRAM reads
    # for 98.6 MB of input + weights with the all-images order). Inside a pass: rounds of n_ctas items, longest first,
    # each round dealt to the CTAs in order of their load so far (CTA c walks positions c, c+G, c+2G, ...).
    flat = []
    for p0 in range(0, N, max(ipp, step)):
        idx = np.nonzero((img >= p0) & (img < p0 + max(ipp, step)))[0]
        flat.append(idx[np.argsort(-costs[idx], kind="stable")])
    flat = np.concatenate(flat)
    G = min((n_clusters if n_clusters else n_ctas // 2) if pairs else n_ctas, len(items))
    out = np.zeros(len(items), dtype=np.int64)
    load = np.zeros(G)
    for r in range((len(flat) + G - 1) // G):
        seg = flat[r * G:(r + 1) * G]
        seg = seg[np.argsort(-costs[seg], kind="stable")]
        ctas = np.argsort(load[:len(seg)], kind="stable")          # a short last round uses CTAs 0 .. len-1
        out[r * G + ctas] = items[seg]
        load[ctas] += costs[seg]
    if pairs:   # entry 2i = the item's tile in image 2k, entry 2i+1 = the same tile of image 2k+1
        out = np.stack([out, out + ty * tx * n_tiles], axis=1).reshape(-1)
    return torch.from_numpy(out.astype(np.int32))


def conv_igemm_grouped(x, wts, scales, shifts, ksizes, dils, ch_offs, relu, out, ldo, out_channels, schedule,
                       cta_pairs=False):
    """One persistent launch for several 'same' convolutions of x (see eeseg_conv_igemm_grouped); cta_pairs: `schedule`
    is a pair list (group_schedule(pairs=True))."""
    import ctypes
    N, h, w, Cin = x.shape
    n = len(wts)
    Cout = wts[0].shape[0]
    PA = ctypes.c_void_p * n
    IA = ctypes.c_int * n
    with torch.cuda.device(x.device):
        if PROFILE is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        torch_ops.fast.conv_igemm_grouped(x, list(wts), None if scales is None or scales[0] is None else list(scales),
                                           list(shifts), list(ksizes), list(dils),
                                           list(ch_offs), bool(relu), out, ldo, out_channels, schedule, bool(cta_pairs))
        if PROFILE is not None:
            b.record()
            PROFILE.append((a, b, sum(2 * N * h * w * Cout * k * k * Cin for k in ksizes), PROFILE_TAG))
        if FLOP_LOG is not None:
            noms = [2 * N * h * w * Cout * k * k * Cin for k in ksizes]
            FLOP_LOG.append((sum(noms), sum(f * live_tap_fraction(h, w, k, d) for f, k, d in zip(noms, ksizes, dils)),
                             PROFILE_TAG))


def dense_bn_act(x, W, scale, shift, relu):
    """fp32 [N,K] x [O,K]^T -> [N,O], * scale + shift, optional ReLU (eeseg_dense_bn_act)."""
    N, K = x.shape
    O = W.shape[0]
    y = torch.empty((N, O), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().eeseg_dense_bn_act(x.data_ptr(), W.data_ptr(), scale.data_ptr(), shift.data_ptr(), N, K, O,
                                       1 if relu else 0, y.data_ptr(),
                                       torch.cuda.current_stream(x.device).cuda_stream), "eeseg_dense_bn_act")
    return y


def global_avgpool_nhwc(xh):
    """bf16 NHWC [N,h,w,C] -> f32 [N,C] (AdaptiveAvgPool2d(1))."""
    N, h, w, C = xh.shape
    dev = xh.device
    pooled = torch.empty((N, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty((lib().eeseg_global_avgpool_workspace_bytes(N, C),), dtype=torch.uint8, device=dev)
        check(lib().eeseg_global_avgpool_nhwc(xh.data_ptr(), N, h * w, C, pooled.data_ptr(), ws.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream),
              "eeseg_global_avgpool_nhwc")
    return pooled


# BatchNorm scale of the inference convolutions: 2 = folded into the bf16 weights (one rounding of w * s instead of one
# rounding of w; the kernel's epilogue then only adds the shift: scale pointer NULL, half the shared-memory reads of its
# column loop; measured -1.1 % on the 513x513 step), 1 = only for the Bottleneck convolutions that take a residual (they
# need it: the residual is added inside the accumulator), 0 = as 1 but with an explicit tensor of ones (kept for A/B runs)
FOLD_SCALE = int(os.environ.get("EESEG_FOLD_SCALE", "2"))
GROUPED = True   # run the ASPP branch convolutions as one grouped launch
IMAGES_PER_PASS = int(os.environ.get("EESEG_GROUP_IMAGES_PER_PASS", "2"))   # grouped work list: images per pass (0 = all)
GROUP_PAIRS = os.environ.get("EESEG_GROUP_PAIRS", "1") != "0"   # grouped ASPP launch as CTA pairs when the batch is even
OVERLAP_POOLED = True   # pooled ASPP branch on a side stream, next to the grouped ASPP launch

_SIDE = {}


def _side_stream(dev):
    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


class HeadPlan:
    def __init__(self, head):
        self._schedules = {}
        mods = list(head.children())
        self.pre = None
        if isinstance(mods[0], nn.Conv2d):      # my_branch bottleneck 1x1 (+bias, no BN/ReLU)
            self.pre = mods[0]
            mods = mods[1:]
        aspp, conv3, bn3, _relu, last = mods
        assert isinstance(aspp, ASPP)
        dev = last.weight.device
        self.dev = dev
        self.n_classes = last.out_channels
        self.mid = aspp.project[0].out_channels
        self.branches = []
        for br in list(aspp.convs)[:-1]:
            conv, bn = br[0], br[1]
            s, b = _fold_bn(bn)
            self.branches.append((_krsc(conv), s.contiguous(), b.contiguous(), conv.dilation[0]))
        pool = aspp.convs[-1]
        self.pool_w = pool[1].weight.detach().float().flatten(1).contiguous()  # [mid, Cin]
        self.pool_s, self.pool_b = (t.contiguous() for t in _fold_bn(pool[2]))
        nb = len(self.branches)
        proj = aspp.project[0].weight.detach().float().flatten(1)          # [mid, (nb+1)*mid]
        self.proj_pool_w = proj[:, nb * self.mid:].contiguous()             # [mid, mid] fp32
        self.proj_s, self.proj_b = (t.contiguous() for t in _fold_bn(aspp.project[1]))
        self.c3_s, self.c3_b = _fold_bn(bn3)
        pw = proj[:, :nb * self.mid]
        if FOLD_SCALE >= 2:   # projection and 3x3: scale into the weights, NULL scale for the kernel (the pooled branch's
            # per-image shift keeps its own scale: dense_bn_act below)
            pw = pw * self.proj_s.view(-1, 1)
            self.c3_w = (conv3.weight.detach().float() * self.c3_s.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
            self.proj_cs = self.c3_s = None
        else:
            self.c3_w = _krsc(conv3)
            self.proj_cs = self.proj_s
        self.proj_w = pw.contiguous().view(self.mid, 1, 1, nb * self.mid).to(torch.bfloat16)
        # final classifier: pad Cout to a multiple of 16 for the MMA N dimension
        C = self.n_classes
        Cp = (C + 15) // 16 * 16
        self.Cp = Cp
        lw = torch.zeros((Cp, 1, 1, last.in_channels), dtype=torch.bfloat16, device=dev)
        lw[:C] = _krsc(last)
        self.last_w = lw
        self.last_s = torch.ones(Cp, dtype=torch.float32, device=dev)
        lb = torch.zeros(Cp, dtype=torch.float32, device=dev)
        if last.bias is not None:
            lb[:C] = last.bias.detach().float()
        self.last_b = lb
        if self.pre is not None:
            self.pre_w = _krsc(self.pre)
            self.pre_s = torch.ones(self.pre.out_channels, dtype=torch.float32, device=dev)
            self.pre_b = (self.pre.bias.detach().float() if self.pre.bias is not None
                          else torch.zeros(self.pre.out_channels, dtype=torch.float32, device=dev))

    def run(self, x):
        """x: [N,Cin,h,w] (any memory format / float dtype). Returns fp32 NHWC [N,h,w,Cp]."""
        N, Cin, h, w = x.shape
        dev = x.device
        xh = x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()       # NHWC bf16 (a view when channels_last)
        BF, F32 = _lib.BF16, _lib.F32
        if self.pre is not None:
            t = torch.empty((N, h, w, self.pre_w.shape[0]), dtype=torch.bfloat16, device=dev)
            conv_igemm(xh, self.pre_w, self.pre_s, self.pre_b, 1, False, t, BF, t.shape[-1])
            xh = t
        nb, mid = len(self.branches), self.mid
        cat = torch.empty((N, h, w, nb * mid), dtype=torch.bfloat16, device=dev)
        # The pooled branch (avg-pool -> 1x1 -> BN -> ReLU, then its share of the projection, per image: ~35 us of
        # four small launches) needs only the head input, like the ASPP convolutions: it runs on a side stream
        # (a parallel branch of the captured graph) while the grouped ASPP launch — whose CTAs leave ~28 KB of
        # shared memory and half of the register file free — occupies the main stream; joined before the projection.
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if OVERLAP_POOLED else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            pooled = global_avgpool_nhwc(xh)
            pv = dense_bn_act(pooled, self.pool_w, self.pool_s, self.pool_b, True)        # [N, mid]
            pshift = dense_bn_act(pv, self.proj_pool_w, self.proj_s, self.proj_b, False)    # [N, mid]
        if GROUPED and 2 <= nb <= 4 and all(wt.shape[1] in (1, 3) for wt, _, _, _ in self.branches) \
                and any(wt.shape[1] == 3 for wt, _, _, _ in self.branches):
            # the ASPP branches (1x1 + atrous 3x3) as ONE persistent launch over a cost-sorted work list
            key = (N, h, w, str(dev))
            # an even batch runs as CTA pairs (cta_group::2: the same tile of two images per cluster)
            pairs = GROUP_PAIRS and N % 2 == 0 and xh.shape[-1] >= 256
            sched = self._schedules.get(key)
            if sched is None:
                n_cl = lib().eeseg_conv_pair_clusters() if pairs else None
                pairs = pairs and bool(n_cl)
                sched = (group_schedule(N, h, w, xh.shape[-1], mid, [wt.shape[1] for wt, _, _, _ in self.branches],
                                        [d for _, _, _, d in self.branches], pairs=pairs, n_clusters=n_cl).to(dev), pairs)
                self._schedules[key] = sched
            sched, pairs = sched
            conv_igemm_grouped(xh, [b[0] for b in self.branches], [b[1] for b in self.branches],
                               [b[2] for b in self.branches], [b[0].shape[1] for b in self.branches],
                               [b[3] for b in self.branches], [k * mid for k in range(nb)], True, cat,
                               nb * mid, nb * mid, sched, cta_pairs=pairs)
        else:
            for k, (wt, s, b, dil) in enumerate(self.branches):
                conv_igemm(xh, wt, s, b, dil, True, cat[..., k * mid:], BF, nb * mid)
        if side is not main:
            main.wait_stream(side)
        y = torch.empty((N, h, w, mid), dtype=torch.bfloat16, device=dev)
        conv_igemm(cat, self.proj_w, self.proj_cs, pshift, 1, True, y, BF, mid, shift_sn=mid)
        z = torch.empty((N, h, w, mid), dtype=torch.bfloat16, device=dev)
        conv_igemm(y, self.c3_w, self.c3_s, self.c3_b, 1, True, z, BF, mid)
        out = torch.empty((N, h, w, self.Cp), dtype=torch.float32, device=dev)
        conv_igemm(z, self.last_w, self.last_s, self.last_b, 1, False, out, F32, self.Cp)
        return out
