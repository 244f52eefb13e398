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
    assert C.sizeof(L.bw_transition) == 1608 and dt["transition"].itemsize == 1608
    # bw_rollout_view (ABI 6): five device pointers, the store-slot pointer, amax + reserved
    assert C.sizeof(L.bw_rollout_view) == 56 and L.bw_rollout_view.slot.offset == 40 and L.bw_rollout_view.amax.offset == 48
    assert (L.BW_OK, L.BW_ERR_INVALID, L.BW_ERR_CUDA, L.BW_ERR_CAPACITY, L.BW_ERR_STATE) == (0, -1, -2, -3, -4)


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


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package names it, and importing the whole package
    (adapter, drop-in classes, rollout helpers) leaves it out of sys.modules."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "bridges-with-reinforcement-learning_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(d, f)
    code = ("import sys; import bridges_b200, bridges_b200.lib, bridges_b200.envs.batched, bridges_b200.envs.gym_env, "
            "bridges_b200.envs.assembly_env, bridges_b200.rollout, bridges_b200.sharding; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "print('LOADED', bad)")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "LOADED []" in out.stdout, out.stdout


def test_missing_library_fails_loudly(tmp_path):
    """No CPU implementation behind the package: a missing libbridges_b200.so is an error, not a fallback."""
    import subprocess
    import sys
    code = ("import os; os.environ['BRIDGES_B200_LIB'] = %r; from bridges_b200 import lib\n"
            "try:\n    lib.load()\n    print('LOADED')\nexcept Exception as e:\n    print('RAISED', type(e).__name__)\n"
            % str(tmp_path / "absent.so"))
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "RAISED" in out.stdout and "LOADED" not in out.stdout, out.stdout
