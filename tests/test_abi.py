"""CPU: the C-ABI library loads and exports exactly what include/pillarnet_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pillarnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_signatures_cover_the_header(lib):
    from pillarnet_lts_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_abi_version_and_pure_host_helpers(lib):
    assert lib.pn_abi_version() == 1
    assert lib.pn_mask_words(1, 1440, 1440) == 1440 * 1440 // 32
    assert lib.pn_mask_words(1, 3, 3) == 1
    assert lib.pn_pillarize_scratch_bytes(1, 1440, 1440) > 0
    assert lib.pn_nms_scratch_bytes(6, 1024) >= 6 * 1024 * 16 * 8


def test_struct_layouts_match_the_header(lib):
    from pillarnet_lts_b200._lib import ConvArgs, TaskArgs
    assert ctypes.sizeof(ConvArgs) == lib.pn_sizeof_conv_args()
    assert ctypes.sizeof(TaskArgs) == lib.pn_sizeof_task_args()


def test_no_cpu_fallback_message():
    from pillarnet_lts_b200 import _lib
    import torch
    import pytest
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.require_cuda(torch.zeros(1))
