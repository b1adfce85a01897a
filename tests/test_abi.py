"""The C-ABI library builds, loads and exports every symbol include/kdot.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "kdot.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kdot_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_by_the_python_side():
    from kd_6d_pose_adlp_b200 import _lib

    assert set(_declared_symbols()) == set(_lib.EXPORTS)


def test_library_builds_and_exports_everything():
    import __graft_entry__

    __graft_entry__.build()
    from kd_6d_pose_adlp_b200 import _lib

    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert _lib.lib().kdot_version() == 100
    # pure host-side query, no device needed
    assert _lib.lib().kdot_workspace_bytes(64, 12, 12, 8, 2) == 0
    assert _lib.lib().kdot_workspace_bytes(32, 1360, 1364, 8, 2) > 32 * 1024 * 16


def test_sass_is_sm100a_with_packed_fp32():
    import shutil
    import subprocess

    from kd_6d_pose_adlp_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "FFMA2" in out and "MUFU.EX2" in out  # packed f32x2 math + SFU exp2 in the tiled kernel
