"""libmbt_b200.so loads on a CPU-only box and exports every symbol include/mbt_b200.h declares; no compute calls."""
import ctypes as C
import os
import re

from mbt_gym_b200 import _abi, _build, _lib
from tests.helpers import ROOT


def test_library_builds_and_exports_every_declared_symbol():
    _build.build_library()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "mbt_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|const char \*)\s*\*?(mbt_[a-z_0-9]+)\(", header, flags=re.M)))
    assert declared, "no declarations parsed from the header"
    assert sorted(_lib.ABI_SYMBOLS) == declared, (sorted(set(declared) ^ set(_lib.ABI_SYMBOLS)))
    for sym in declared:
        assert hasattr(lib, sym), f"libmbt_b200.so does not export {sym}"
    assert lib.mbt_abi_version() == _abi.MBT_ABI_VERSION


def test_struct_layout_matches_header():
    """sizeof(mbt_config) is checked by the library itself: a mismatching struct_size is rejected before any CUDA call."""
    lib = _lib.load()
    cfg = _abi.new_config(num_trajectories=4, n_steps=2, terminal_time=1.0, step_size=0.5,
                          dynamics=_abi.MBT_DYN_LIMIT, midprice=_abi.MBT_MID_BM, arrival=_abi.MBT_ARR_POISSON,
                          fill=_abi.MBT_FILL_EXPONENTIAL)
    a, d, s = _lib.config_dims(cfg)
    assert (a, d, s) == (2, 4, 3)
    cfg.arrival = _abi.MBT_ARR_HAWKES
    assert _lib.config_dims(cfg) == (2, 6, 5)
    bad = _abi.new_config(num_trajectories=4)
    bad.struct_size = 12
    h = C.c_void_p()
    rc = lib.mbt_create(C.byref(bad), 0, C.byref(h))
    assert rc == _abi.MBT_E_INVALID_ARG and b"struct_size" in lib.mbt_last_error()


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU, creating a handle fails loudly (MBT_E_CUDA) instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        return
    cfg = _abi.new_config(num_trajectories=4, n_steps=2, terminal_time=1.0, step_size=0.5,
                          dynamics=_abi.MBT_DYN_LIMIT, midprice=_abi.MBT_MID_BM, arrival=_abi.MBT_ARR_POISSON,
                          fill=_abi.MBT_FILL_EXPONENTIAL)
    try:
        _lib.NativeEnv(cfg)
    except _lib.MbtError as e:
        assert e.code == _abi.MBT_E_CUDA and "no CPU path" in str(e)
    else:
        raise AssertionError("mbt_create succeeded without a CUDA device")


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under mbt_gym_b200/ may import, load or include it."""
    pkg = os.path.join(ROOT, "mbt_gym_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|libmbt_oracle|#include\s+[\"<][^\">]*oracle|numpy_port|ref_shim", re.M)
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)
