"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol that
include/eeseg.h declares, with a ctypes prototype for each. No kernel is launched."""
import ctypes
import os

import pytest

from ee_semantic_segmentation_b200 import _lib, build


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def test_library_builds_and_loads(libpath):
    assert os.path.exists(libpath)
    l = ctypes.CDLL(libpath)
    assert l.eeseg_abi_version() == 1


def test_every_header_symbol_is_exported_and_bound(libpath):
    names = _lib.header_symbols()
    assert len(names) >= 15
    l = ctypes.CDLL(libpath)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/eeseg.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in _lib.py"
    for n in _lib.PROTOTYPES:
        assert n in names, f"{n} bound in _lib.py but not declared in include/eeseg.h"


def test_pure_host_entry_points(libpath):
    l = _lib.lib()
    assert l.eeseg_exit_gate_num_partials(513, 513) == 17 * 65   # 32-column groups x 8-row strips
    assert l.eeseg_multi_exit_ce_workspace_bytes(3, 2, 513 * 513) > 0
    assert l.eeseg_lovasz_workspace_bytes(3, 1, 19, 768 * 768) >= 16 * 19 * 768 * 768
    assert _lib.launch_count() == 0


def test_argument_errors_are_reported_without_a_gpu(libpath):
    l = _lib.lib()
    rc = l.eeseg_confusion_hist(None, 0, 0, None, 1, 21, 10, None, 0, None)
    assert rc == 1
    assert b"null" in l.eeseg_last_error()


def test_ops_reject_cpu_tensors():
    import torch
    from ee_semantic_segmentation_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.confusion_hist(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.multi_exit_ce(torch.zeros(1, 1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.exit_gate(torch.zeros(1, 3, 4, 4))


def test_argument_errors_of_the_staged_engine_and_loss_entry_points(libpath):
    """Every launcher validates its arguments before touching CUDA: bad sizes / alignment / null pointers come back as
    return code 1 with a message, also on a machine without a GPU (fake non-null pointers are never dereferenced)."""
    l = _lib.lib()
    fake = ctypes.c_void_p(0x1000)                 # 16-byte aligned, never dereferenced on these paths
    odd = ctypes.c_void_p(0x1008)
    err = lambda: l.eeseg_last_error().decode()
    assert l.eeseg_compact_rows(None, fake, fake, None, 1, 1, 64, None) == 1 and "null" in err()
    assert l.eeseg_compact_rows(fake, fake, fake, None, 1, 1, 10, None) == 1 and "16-byte" in err()
    assert l.eeseg_compact_rows(fake, odd, fake, None, 1, 1, 64, None) == 1 and "16-byte" in err()
    assert l.eeseg_compact_rows(fake, fake, fake, None, 1, 0, 64, None) == 0           # nothing to move
    assert l.eeseg_exit_stage_commit(None, 0.5, 1, 0, 0, fake, fake, 2, 100, None, fake, fake, None, None, None, None,
                                     None, None) == 1 and "scores are required" in err()
    assert l.eeseg_exit_stage_commit(fake, 0.5, 1, 0, 0, fake, fake, 70000, 100, None, fake, fake, None, None, None,
                                     None, None, None) == 1 and "bad sizes" in err()
    assert l.eeseg_focal_fwd(fake, 0, 0, fake, 1, 1, 65, 10, 2.0, None, None, 0, 0, None, fake, None, None, fake,
                             None) == 1 and "C <= 64" in err()
    assert l.eeseg_focal_fwd(fake, 0, 0, fake, 1, 1, 21, 10, -1.0, None, None, 0, 0, None, fake, None, None, fake,
                             None) == 1 and "gamma" in err()
    assert l.eeseg_focal_fwd(fake, 7, 0, fake, 1, 1, 21, 10, 2.0, None, None, 0, 0, None, fake, None, None, fake,
                             None) == 1 and "dtype" in err()
    assert l.eeseg_maxpool3x3s2_nhwc_train(fake, 1, 8, 8, 12, fake, fake, None) == 1 and "C % 8" in err()
    assert l.eeseg_maxpool3x3s2_nhwc_bwd(fake, None, 1, 8, 8, 64, fake, None) == 1 and "null" in err()
    assert l.eeseg_focal_workspace_bytes(3, 2, 513 * 513) >= 3 * 2 * 1029 * 8
    assert _lib.launch_count() == 0
