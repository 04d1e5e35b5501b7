"""The C-ABI library builds for sm_100a here (no GPU needed), loads, and exports every symbol include/mocap_b200.h
declares; the package refuses to run without CUDA instead of falling back to anything."""
import os
import re

import pytest
import torch

from mocapv2_b200 import _cabi, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _cabi.load(build.build())


def test_header_and_binding_agree(lib):
    header = open(os.path.join(REPO, "include", "mocap_b200.h")).read()
    declared = set(re.findall(r"\b(mocap_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    for macro, val in (("MOCAP_MAX_CAND", _cabi.MAX_CAND), ("MOCAP_MAX_CAMS", _cabi.MAX_CAMS), ("MOCAP_CAM_STRIDE", _cabi.CAM_STRIDE)):
        assert int(re.search(rf"#define {macro} (\d+)", header).group(1)) == val


def test_no_compute_entry_points_pure_host_calls(lib):
    assert lib.mocap_abi_version() == _cabi.ABI_VERSION
    assert lib.mocap_status_string(0) == b"ok" and b"workspace" in lib.mocap_status_string(-2)
    assert lib.mocap_undistort_table_bytes(480, 640) >= 480 * 640 * 4
    assert lib.mocap_undistort_table_bytes(0, 640) == 0
    small = lib.mocap_detect_workspace_bytes(1, 480, 640, 256, 512, 8192)
    big = lib.mocap_detect_workspace_bytes(4, 480, 640, 256, 512, 8192)
    assert 0 < small < big
    assert lib.mocap_detect_workspace_bytes(1, 20000, 640, 256, 512, 8192) == 0       # beyond the compiled limits
    assert lib.mocap_correspond_workspace_bytes(2, 6, 32, 64) > 0


def test_library_is_sm_100a_only():
    out = os.popen(f"cuobjdump -lelf {build.LIB} 2>/dev/null").read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the GPU-less behaviour")
def test_package_fails_loudly_without_cuda():
    from mocapv2_b200.engine import CaptureEngine
    with pytest.raises(_cabi.MocapError):
        CaptureEngine()
    from mocapv2_b200.lib import ImageOperations
    import numpy as np
    with pytest.raises(_cabi.MocapError):
        ImageOperations._find_dot(np.zeros((480, 640), np.uint8))


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(_cabi.MocapError):
        _cabi.load(str(tmp_path / "nope.so"))
