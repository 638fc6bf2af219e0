import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


# Logits of the bf16 network against fp32 arithmetic (measured at 4 x 513 x 513 in tests/test_parity_gpu.py:
# relative L2 6.9e-3 / 6.3e-3 / 1.09e-2 per exit, max-abs 1.2-1.8e-2 of the largest logit; torch's own bf16 path is worse on
# every exit). Model-level tests assert these two bounds, not a widened scalar.
BF16_MODEL_REL_L2 = 1.25e-2
BF16_MODEL_MAX_ABS = 2.5e-2


def assert_bf16_model_close(got, ref, what=""):
    import numpy as np
    import torch
    def as64(t):
        if isinstance(t, torch.Tensor):
            return t.detach().cpu().double()
        return torch.as_tensor(np.asarray(t)).double()
    g, r = as64(got), as64(ref)
    l2 = float((g - r).norm() / r.norm())
    mx = float((g - r).abs().max() / r.abs().max())
    assert l2 < BF16_MODEL_REL_L2 and mx < BF16_MODEL_MAX_ABS, (what, l2, mx)
    return l2, mx
