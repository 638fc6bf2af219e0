"""TEST INFRASTRUCTURE — generates tests/golden/operator.npz by running the UNMODIFIED reference
operators `ee_dnn_op_ne.eval_ee_deeplabv3` (entropy gate, ee_dnn_op_ne.py:40-108) and
`ee_dnn_op.eval_ee_deeplabv3` (similarity gate, ee_dnn_op.py:40-118) from /root/reference (through
oracle/ref_import.py's stub set) on a seeded 3-exit BranchyDeepLabV3 ResNet-50.

Run in the build container only:  python -m oracle.make_golden_operator

What is stored (per image): the input, the reference model's per-exit scores (the reference's own
img_norm_entropy on softmax(net(x)[i])), for every case the operator's `n`, `exit` and `last` maps (uint8) and
its `*_flops` integers (FLOP counter = the pthflops stand-in of oracle/stubs: unpinned by the reference, kept
to check that the product's cached meta-tensor tables add up the same way), and per exit the top-1 minus top-2
logit gap of the reference's fp32 logits (float16) — a bf16 implementation can only be asked to reproduce the
argmax where that gap exceeds its logit error.

Cases (tau is placed between / around the reference's own scores s1 > s2 of the two early exits):
  ne_first   tau above s1            -> leaves at exit 1
  ne_second  s2 < tau < s1           -> leaves at exit 2
  ne_none    tau below both          -> final exit (n = 3)
  ne_ignore0 tau above s1, ignore=[0]-> exit 1 is not evaluated, leaves at exit 2
  sim_leave  similarity op, tau above d(exit1 map, exit2 map) -> leaves at exit 2 (exit 1 is only the reference map)
  sim_stay   similarity op, tau below it -> final exit
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import model_port, ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
BRANCH_SEED = 311
SHARPEN = (1.0, 6.0)          # see model_port.sharpen_heads: makes exit 2 more confident than exit 1
IMG_HW = (97, 129)
N_IMG = 2


def map_distance(a, b):
    """Similarity metric handed to both operators: fraction of pixels whose class differs."""
    a, b = torch.as_tensor(a).reshape(-1), torch.as_tensor(b).reshape(-1)
    return float((a.cpu() != b.cpu()).float().mean())


def main():
    os.makedirs(OUT, exist_ok=True)
    fd_old, op_ne, op_sim, ebe = ref_import.load("from_deepv3", "ee_dnn_op_ne", "ee_dnn_op", "eval_br_ent")
    import torchvision
    base_path = "/tmp/eeseg_oracle_base_r50.pth"
    torch.manual_seed(0)
    base = torchvision.models.segmentation.deeplabv3_resnet50(
        weights=None, weights_backbone=None, num_classes=21, aux_loss=True)
    torch.save(base, base_path)
    net = fd_old.branchyDeepv3(base_path, "deeplabv3_resnet50", 2, 513)
    model_port.reinit_branches(net.branches, BRANCH_SEED)
    model_port.sharpen_heads(net.branches, SHARPEN)
    net.eval()
    d = {"sections": np.array([len(s) for s in net.base_model]), "branch_seed": np.array(BRANCH_SEED),
         "sharpen": np.array(SHARPEN), "n_img": np.array(N_IMG)}
    g = torch.Generator().manual_seed(4242)
    metric = ebe.img_norm_entropy(21)
    cpu = torch.device("cpu")
    for k in range(N_IMG):
        x = torch.randn(3, *IMG_HW, generator=g)
        d[f"img{k}/x"] = x.numpy()
        with torch.no_grad():
            y = net(x.unsqueeze(0))                         # [3,1,21,H,W]
        s = [float(metric(F.softmax(y[i], 1).squeeze().cpu())) for i in range(2)]
        assert s[0] > s[1] + 0.02, s                        # the sharpened second head is the more confident one
        d[f"img{k}/scores"] = np.array(s, np.float32)
        top2 = y[:, 0].topk(2, dim=1).values                # [3,2,H,W]
        d[f"img{k}/gap"] = (top2[:, 0] - top2[:, 1]).numpy().astype(np.float16)
        d[f"img{k}/absmax"] = y[:, 0].abs().amax(dim=(1, 2, 3)).numpy()
        d[f"img{k}/argmax"] = y[:, 0].argmax(1).numpy().astype(np.uint8)
        dist12 = map_distance(y[0, 0].argmax(0), y[1, 0].argmax(0))
        d[f"img{k}/dist12"] = np.array(dist12, np.float32)
        cases = {
            "ne_first": (op_ne, dict(metric=metric, th=s[0] + 0.01)),
            "ne_second": (op_ne, dict(metric=metric, th=0.5 * (s[0] + s[1]))),
            "ne_none": (op_ne, dict(metric=metric, th=s[1] - 0.01)),
            "ne_ignore0": (op_ne, dict(metric=metric, th=s[0] + 0.01, ignore=[0])),
            "sim_leave": (op_sim, dict(metric=map_distance, th=dist12 * 1.5 + 1e-3)),
            "sim_stay": (op_sim, dict(metric=map_distance, th=dist12 * 0.5)),
        }
        for tag, (mod, kw) in cases.items():
            with torch.no_grad():
                out = mod.eval_ee_deeplabv3(net, device=cpu, **kw)(x)
            d[f"img{k}/{tag}/th"] = np.array(kw["th"], np.float64)
            d[f"img{k}/{tag}/n"] = np.array(out["n"])
            d[f"img{k}/{tag}/exit"] = out["exit"].numpy().astype(np.uint8)
            d[f"img{k}/{tag}/last"] = out["last"].numpy().astype(np.uint8)
            for key in out:
                if key.endswith("flops") or key.endswith("flops_2"):
                    d[f"img{k}/{tag}/{key}"] = np.array(int(out[key]), np.int64)
            print(k, tag, "n =", out["n"], "th =", kw["th"], sorted(out))
    d["cases"] = np.array(["ne_first", "ne_second", "ne_none", "ne_ignore0", "sim_leave", "sim_stay"])
    path = os.path.join(OUT, "operator.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    sys.exit(main())
