"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference for the GPU box's CPU arm.

    python -m oracle.build_ref            (run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no build step): its module files are copied verbatim from where they lie under
/root/reference into oracle/_ref/reference/, next to the stub set (oracle/stubs: the nine modules the published
reference imports but does not ship) in oracle/_ref/stubs/. oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so it travels to the GPU box, where /root/reference
does not exist. `bench.py --impl reference` and bench.py's `cpu_baseline` leg import the reference from there
(oracle/ref_import.py) and time its own `branchyDeepv3` + `br_evaluator` on the host cores; when oracle/_ref is
absent they fall back to the oracle port and say so (`cpu_baseline.kind = "port"`).
"""
import hashlib
import json
import os
import shutil
import sys

# SURVEY.md §8(a)'s hot-path modules and the reference modules they import (the rest of the reference is not staged)
FILES = [
    "from_deepv3.py", "from_deepv3_new.py", "ee_dnn_op.py", "ee_dnn_op_ne.py", "eval_br_ent.py", "eval_mIoU.py",
    "my_pixelwise_xentropy.py", "branchy_seg_losses.py", "lovaszsoftmax.py", "new_seg_losses.py", "seg_metrics.py",
    "compute_mIoU.py",
    "my_layers.py", "common_torch.py", "eval_flops.py", "funcs.py", "get_seg_datasets.py", "sim_metrics.py",
]

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("EESEG_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build():
    if not os.path.isdir(SRC):
        return None
    ref_dst, stub_dst = os.path.join(DST, "reference"), os.path.join(DST, "stubs")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(ref_dst)
    manifest = {}
    for f in FILES:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(ref_dst, f))
            with open(os.path.join(SRC, f), "rb") as fh:
                manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    shutil.copytree(os.path.join(HERE, "stubs"), stub_dst, ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    return DST


if __name__ == "__main__":
    print(build())
    sys.exit(0)
