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
