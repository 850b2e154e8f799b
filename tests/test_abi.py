"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/lompc_b200.h declares, and fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

from chargingstation import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def _declared_functions():
    names = set()
    for f in sorted(os.listdir(INCLUDE)):
        if not f.endswith(".h"):
            continue
        src = open(os.path.join(INCLUDE, f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:lompc|price|bimpc|fleet)_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature in _native.py"


def test_version_and_strerror():
    lib = _native.load()
    assert b"sm_100a" in lib.lompc_version()
    assert b"lompc.py:87" in lib.lompc_strerror(_native.ERR_GAMMA)


def test_create_validates_constants_like_reference():
    """lompc.py:36-38: y_max in [0.75, 0.9], w_max in [0, 0.25], ev_type small/large."""
    lib = _native.load()
    h = C.c_void_p()
    assert lib.lompc_create(24, 0.05, 10.0, 0.95, 0.25, 0, 0, C.byref(h)) == _native.ERR_CONSTS
    assert lib.lompc_create(24, 0.05, 10.0, 0.9, 0.3, 0, 0, C.byref(h)) == _native.ERR_CONSTS
    assert lib.lompc_create(24, 0.05, 10.0, 0.9, 0.25, 7, 0, C.byref(h)) == _native.ERR_CONSTS
    assert lib.lompc_create(0, 0.05, 10.0, 0.9, 0.25, 0, 0, C.byref(h)) == _native.ERR_ARG


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run."""
    lib = _native.load()
    if lib.lompc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.lompc_create(24, 0.05, 10.0, 0.9, 0.25, 0, 0, C.byref(h)) == _native.ERR_NO_DEVICE
    from chargingstation.lompc import LoMPC, LoMPCConstants
    with pytest.raises(RuntimeError):
        LoMPC(24, LoMPCConstants(0.05, 10, 0.9, 0.25, "small"))


def test_python_mirror_asserts():
    from chargingstation.lompc import LoMPC, LoMPCConstants
    with pytest.raises(AssertionError):
        LoMPC(12, LoMPCConstants(0.05, 10, 0.95, 0.25, "small"))
    with pytest.raises(AssertionError):
        LoMPC(12, LoMPCConstants(0.05, 10, 0.9, 0.25, "medium"))


def test_product_does_not_import_oracle():
    """Nothing under the package may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "incentive-design-mpc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
