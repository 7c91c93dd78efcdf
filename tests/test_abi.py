"""C-ABI surface: the library loads, exports every symbol include/kmc_b200.h declares, the Python signature
table covers the header, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kmcb200_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(kmc):
    lib = C.CDLL(kmc.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in kmc_b200.h but not exported: {missing}"


def test_python_signatures_cover_header(kmc):
    assert sorted(kmc.SIGNATURES.keys()) == header_symbols()
    assert kmc.load_library().kmcb200_version() == 100


def test_no_cpu_fallback(kmc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = kmc.load_library()
    h = C.c_void_p()
    rc = lib.kmcb200_create(C.byref(h), 0, None)
    assert rc == -3  # KMCB200_E_NOGPU
    assert b"no CPU fallback" in lib.kmcb200_last_error()
    with pytest.raises(kmc.KMCB200Error):
        kmc.Context(0)


def test_product_does_not_touch_oracle():
    """the product package must never import / link / load anything under oracle/"""
    pkg = os.path.join(ROOT, "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "kmc_oracle" not in txt and "oracle.binding" not in txt and "from oracle" not in txt, f
    for hdr in os.listdir(os.path.join(ROOT, "include")):
        assert "kmc_oracle" not in open(os.path.join(ROOT, "include", hdr)).read()


def test_cutoff_d2max_is_the_exact_squared_distance_threshold(kmc):
    """kmcb200_cutoff_d2max (host-only): d2 <= d2max must select exactly the pairs with sqrt(d2) < cutoff, for the
    reference's 20 A cutoff (src/potential_solver_gpu.cu:1549-1551) and for awkward cutoffs, checked on the doubles
    around the boundary and on random squared distances (numpy's sqrt is IEEE correctly rounded, like the device's)."""
    import numpy as np
    lib = kmc.load_library()
    rng = np.random.default_rng(7)
    cutoffs = [20.0, 3.5, 1e-3, 12.5, np.nextafter(20.0, 0.0), np.nextafter(20.0, 100.0), float(np.sqrt(2.0)), 1e8] + \
        list(rng.uniform(0.1, 50.0, 20))
    for cutoff in cutoffs:
        d2max = lib.kmcb200_cutoff_d2max(float(cutoff))
        assert np.sqrt(d2max) < cutoff <= np.sqrt(np.nextafter(d2max, np.inf))
        near = [d2max]
        for _ in range(50):
            near.append(np.nextafter(near[-1], 0.0))
        up = np.nextafter(d2max, np.inf)
        for _ in range(50):
            near.append(up)
            up = np.nextafter(up, np.inf)
        d2 = np.concatenate([np.array(near), rng.uniform(0.0, 2.0 * cutoff * cutoff, 20000)])
        assert ((d2 <= d2max) == (np.sqrt(d2) < cutoff)).all()
    assert lib.kmcb200_cutoff_d2max(0.0) < 0.0 and lib.kmcb200_cutoff_d2max(-1.0) < 0.0
