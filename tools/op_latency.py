"""Per-image latency of the reference-facing operator eval_ee_deeplabv3.__call__ (ee_dnn_op_ne.py:51-108): eagerly
launched stages against CUDA-graph stages, one 513x513 image at a time. Usage: python tools/op_latency.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ee_semantic_segmentation_b200.ee_dnn_op_ne import eval_ee_deeplabv3  # noqa: E402
from ee_semantic_segmentation_b200.eval_br_ent import img_norm_entropy  # noqa: E402
from ee_semantic_segmentation_b200.from_deepv3_new import branchyDeepv3  # noqa: E402


def main():
    bench.set_workload("voc513")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = branchyDeepv3(None, "deeplabv3_resnet50", 2, 513, sections=bench.SECTIONS, pretrained=False).to(dev).eval()
    X, _ = bench.synth_batch(0, 8)
    X = X.to(dev)
    rows = {}
    for th, tag in ((2.0, "exit1"), (-1.0, "no_exit")):
        for kw, name in ((dict(use_graph=False), "eager"), (dict(), "graph"), (dict(compute_last=False), "graph_no_tail")):
            op = eval_ee_deeplabv3(net, img_norm_entropy(21), th, device=dev, **kw)
            for k in range(3):
                op(X[k])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(8):
                op(X[k])
            torch.cuda.synchronize()
            rows[f"{tag}_{name}_ms_per_image"] = round((time.perf_counter() - t0) / 8 * 1e3, 3)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
