"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference modules from /root/reference.

In the build container the modules come from ``/root/reference``; on the GPU box (where that path does not exist)
from the verbatim copy ``oracle/build_ref.py`` stages under ``oracle/_ref/`` (git-ignored). Used by
``oracle/make_golden*.py`` to generate the fixtures under ``tests/golden/``, by the ``not gpu`` tests that pin the
restatement in ``oracle/restate.py`` against the real thing, and by bench.py's CPU arm (``--impl reference`` /
``cpu_baseline``), which times the reference's own model + ``br_evaluator``. Nothing in the product package imports
this file.

The stub set (``oracle/stubs``) is the one SURVEY.md §8(c) lists: ``common_header``, a shadow
``module_variables``, a shadow ``allocate_cuda_device``, ``pthflops`` (FlopCounterMode stand-in),
``my_datahanddlers``, ``skimage.*`` (real ``measure.block_reduce``), ``matplotlib.*``.
"""
import contextlib
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("EESEG_REFERENCE_DIR", "/root/reference")
STUB_DIR = os.path.join(_HERE, "stubs")
STAGED = False
if not os.path.isdir(REFERENCE_DIR) and os.path.isdir(os.path.join(_HERE, "_ref", "reference")):
    # the GPU box: the verbatim copy staged by oracle/build_ref.py (git-ignored, travels with the snapshot)
    REFERENCE_DIR = os.path.join(_HERE, "_ref", "reference")
    STUB_DIR = os.path.join(_HERE, "_ref", "stubs")
    STAGED = True

# module names that exist both in the reference and (as drop-in shims) in this repo
_REF_MODULES = [
    "from_deepv3", "from_deepv3_new", "ee_dnn_op", "ee_dnn_op_ne", "eval_br_ent", "eval_mIoU",
    "my_pixelwise_xentropy", "branchy_seg_losses", "lovaszsoftmax", "new_seg_losses",
    "seg_metrics", "compute_mIoU", "eval_flops", "my_layers", "funcs", "get_seg_datasets",
    "sim_metrics", "common_torch",
    # stubs
    "common_header", "module_variables", "allocate_cuda_device", "pthflops", "my_datahanddlers",
    "skimage", "skimage.measure", "skimage.util", "skimage.metrics", "skimage.data",
    "matplotlib", "matplotlib.pyplot",
]


def available() -> bool:
    return os.path.isdir(REFERENCE_DIR)


@contextlib.contextmanager
def reference_on_path():
    """Put the stubs, then the reference, at the front of sys.path; restore afterwards."""
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k in _REF_MODULES}
    sys.path[:0] = [STUB_DIR, REFERENCE_DIR]
    try:
        yield
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k in _REF_MODULES:
                sys.modules.pop(k)
        sys.modules.update(saved_mods)


def load(*names):
    """Import reference modules by name; returns them in order. They stay alive after the
    context exits (their own sub-imports are already resolved)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    import torch
    out = []
    with reference_on_path():
        for n in names:
            m = importlib.import_module(n)
            # torch>=2.6 defaults weights_only=True, which breaks the reference's whole-module
            # pickles (from_deepv3_new.py:43 `load(name)`); patch the name it imported.
            if hasattr(m, "load") and getattr(m, "load") is torch.load:
                def _load(f, *a, **k):
                    k.setdefault("weights_only", False)
                    return torch.load(f, *a, **k)
                m.load = _load
            out.append(m)
    return out[0] if len(out) == 1 else out
