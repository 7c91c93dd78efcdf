"""The CPU arm builds its workloads without the product package (oracle/workload.py): they must be identical to the
product's own structures, array by array, and the two host-model restatements must agree with each other."""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLD, PKG

PARAM_5NM = os.path.join(GOLD, "5nm_device", "parameters.txt")
FIELDS = ("element", "x", "y", "z", "layer")
SCALARS = ("pbc", "nn_dist", "N_left", "N_right", "sigma", "k", "T_bg", "freq", "high_G", "low_G", "Vd", "t_switch")


def _same(a, b):
    for f in FIELDS:
        assert (getattr(a, f) == getattr(b, f)).all(), f
    for f in SCALARS:
        assert getattr(a, f) == getattr(b, f), f
    assert list(a.metals) == list(b.metals) and tuple(a.lattice) == tuple(b.lattice)
    for k in ("E_gen", "E_rec", "E_Vdiff", "E_Odiff"):
        assert (np.asarray(a.E[k]) == np.asarray(b.E[k])).all(), k


def test_5nm_identical(kmc, orc):
    from oracle import workload
    _same(workload.load_5nm(), kmc.load_structure(PARAM_5NM))


@pytest.mark.parametrize("name", ["standin2x2_brick", "standin2x2", "highvac7x7_brick"])
def test_synthetic_identical(kmc, orc, name):
    from oracle import workload
    syn = importlib.import_module(PKG + ".synthetic")
    if name.startswith("highvac"):
        w = workload.crossbar_standin(2, 2, order="brick", Vd=5.0, rnd_seed=5, vacancy_concentration=0.25)
        s = syn.crossbar_standin(PARAM_5NM, 2, 2, order="brick", Vd=5.0, rnd_seed=5, vacancy_concentration=0.25)
    else:
        w, _ = workload.build(name)
        order = name.partition("_")[2] or "file"
        s = syn.crossbar_standin(PARAM_5NM, 2, 2, order=order, Vd=15.0, rnd_seed=32)
    _same(w, s)


def test_workload_module_does_not_load_the_product():
    """bench.py --impl reference must not pull libkmc_b200.so into the process"""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import workload; w, d = workload.build('5nm');"
            "maps = open('/proc/self/maps').read(); assert 'libkmc_b200' not in maps, 'product library loaded';"
            "assert not any('b200' in m for m in sys.modules), 'product package imported'; print(w.N)"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "37650"


def test_workload_names(kmc, orc):
    """bench.py and the CPU arm parse 'standinTxT[_order]' the same way and reject anything else"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import workload
    assert bench.standin_tiles("standin24x24") == 24 and bench.standin_tiles("standin2x2") == 2
    for bad in ("standin2x3", "standin", "standinx", "other4x4"):
        with pytest.raises(ValueError):
            bench.standin_tiles(bad)
    with pytest.raises(ValueError):
        workload.build("standin2x3_brick")
    w, desc = workload.build("standin3x3_brick")
    s, desc2 = bench.build_workload(kmc, "standin3x3_brick")
    _same(w, s)
    assert "3x3" in desc2 and str(s.N) in desc2
