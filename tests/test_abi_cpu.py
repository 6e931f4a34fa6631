"""CPU-only checks of the boundary: the shared library loads, exports every symbol include/gf2_abi.h declares, struct
sizes of the Python mirrors match the C records, and — with no GPU — compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gf2_abi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gf2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(gf2):
    lib = gf2.lib()
    decl = _declared_symbols()
    assert len(decl) >= 25
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in gf2_abi.h but not exported"
    assert set(decl) == set(gf2.ABI_SYMBOLS)
    assert lib.gf2_abi_version() == 1


def test_record_sizes_match(gf2, oracle):
    abi = gf2.abi
    assert oracle.lib.gf2o_sizeof(0) == abi.IMU_PREINT.itemsize
    assert oracle.lib.gf2o_sizeof(1) == abi.WHEEL_PREINT.itemsize
    assert oracle.lib.gf2o_sizeof(2) == abi.PLANE.itemsize == 72
    assert oracle.lib.gf2o_sizeof(3) == abi.PRIOR_BLOCK.itemsize
    assert oracle.lib.gf2o_sizeof(6) == C.sizeof(abi.SolveOpts)
    assert oracle.lib.gf2o_sizeof(7) == abi.SUMMARY.itemsize
    assert abi.OBS.itemsize == 16


def test_no_cpu_fallback(gf2):
    if gf2.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(gf2.Gf2Error, match="no CPU fallback"):
        gf2.Solver(1, 11, 100, 750)


def test_product_never_imports_oracle():
    """A product path that routes through the oracle voids parity: the package must not reference oracle/."""
    pkg = os.path.join(ROOT, "ground-fusion2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "gf2_oracle" not in txt and "gf2o_" not in txt and "oracle/" not in txt.replace("no oracle/", ""), f


def test_synth_is_deterministic_and_fp32_exact(gf2):
    import importlib
    synth = importlib.import_module("gf2_b200.synth")
    a = synth.make_windows(2, n_landmarks=50); b = synth.make_windows(2, n_landmarks=50)
    assert np.array_equal(a["obs"], b["obs"]) and np.array_equal(a["para_pose"], b["para_pose"])
    assert a["obs"].dtype == gf2.abi.OBS
    # W10-F1000 shape: 6,500 projection factors, 7,500 records
    w = synth.make_windows(1, n_landmarks=1000)
    assert int(w["track_len"].sum()) == 7500 and int((w["track_len"] - 1).sum()) == 6500
    assert (np.diff(w["start_frame"][0]) >= 0).all()
