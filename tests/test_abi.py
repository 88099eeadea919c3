"""The C-ABI library loads and exports every symbol include/rgbd_b200.h declares (no compute)."""
import ctypes
import os
import re

import numpy as np

import rgbd_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rgbd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rgbd_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(built_lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/rgbd_b200.h but not exported"
    assert set(rgbd_b200.lib.EXPORTS) <= set(names) | {"rgbd_conv_validate"}


def test_version_and_struct_layout(built_lib):
    assert built_lib.rgbd_abi_version() == 1
    built_lib.rgbd_conv_desc_size.restype = ctypes.c_int
    assert built_lib.rgbd_conv_desc_size() == ctypes.sizeof(rgbd_b200.lib.ConvDesc)


def test_errors_are_codes_not_exceptions(built_lib):
    d = rgbd_b200.lib.ConvDesc()   # all zero -> invalid
    rc = built_lib.rgbd_conv_validate(ctypes.byref(d))
    assert rc == -1 and b"invalid argument" in built_lib.rgbd_last_error()


def test_host_pmf_to_quantized_cdf_matches_reference(built_lib, golden_dir):
    from rgbd_b200.entropy_models import pmf_to_quantized_cdf
    z = np.load(f"{golden_dir}/pmf_kat.npz")
    for n in sorted({k.rsplit(".", 1)[0] for k in z.files}):
        assert np.array_equal(pmf_to_quantized_cdf(z[n + ".pmf"]).numpy(), z[n + ".cdf"]), n


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    import pytest
    monkeypatch.setattr(rgbd_b200.lib, "_lib", None)
    monkeypatch.setattr(rgbd_b200.lib, "LIB_PATH", "/nonexistent/librgbd_b200.so")
    with pytest.raises(rgbd_b200.lib.RgbdError, match="no CPU fallback"):
        rgbd_b200.lib.load()
