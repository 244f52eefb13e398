"""The C-ABI library loads and exports every symbol include/bridges_b200.h declares; struct
layouts of the ctypes/numpy mirrors match the header (CPU only, no compute calls)."""
import ctypes as C
import os
import re

import pytest

from bridges_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bridges_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = L.load()
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.bw_abi_version() == L.BW_ABI_VERSION


def test_struct_layouts():
    dt = L.np_dtypes()
    assert C.sizeof(L.bw_action) == 40 and dt["action"].itemsize == 40
    assert C.sizeof(L.bw_block) == 40
    assert C.sizeof(L.bw_step_out) == 88
    assert C.sizeof(L.bw_interface) == 96
    assert C.sizeof(L.bw_task) == 16 + 16 * L.BW_MAX_OBSTACLES + 16 * L.BW_MAX_TARGETS + 40 * L.BW_MAX_BLOCKS
    assert L.bw_step_out.distance_to_targets.offset == 16 and L.bw_step_out.stable.offset == 72


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bridges_b200.envs.batched import BatchedAssemblyGym
    with pytest.raises(L.BridgesError):
        BatchedAssemblyGym(1, ["shapes/trapezoid.urdf"])
    # and the C entry point itself reports the missing device instead of computing on the CPU
    lib = L.load()
    cfg = L.bw_config()
    lib.bw_config_default(C.byref(cfg))
    handle = C.c_void_p()
    rc = lib.bw_create(C.byref(cfg), C.byref(handle))
    assert rc != 0
    assert b"no CUDA device" in lib.bw_last_error(handle)
    lib.bw_destroy(handle)
