"""Synthetic workloads and the site-reordering helper (host logic, no GPU)."""
import importlib
import os
import subprocess
import sys

import numpy as np

from conftest import GOLD, PKG, ROOT

PARAM = os.path.join(GOLD, "5nm_device", "parameters.txt")


def test_brick_permutation_keeps_contacts_and_sites(kmc, orc):
    syn = importlib.import_module(PKG + ".synthetic")
    s = kmc.load_structure(PARAM, apply_vacancies=False)
    perm = syn.brick_permutation(s.x, s.y, s.z, s.N_left, s.N_right, 12.5)
    assert np.array_equal(np.sort(perm), np.arange(s.N))
    assert np.array_equal(perm[:s.N_left], np.arange(s.N_left))
    assert np.array_equal(perm[-s.N_right:], np.arange(s.N - s.N_right, s.N))
    # the point of the order: neighbours sit close in index space (3.5 A neighbour pairs)
    def index_locality(x, y, z):
        nb = orc.neighbor_list(np.ascontiguousarray(x), np.ascontiguousarray(y), np.ascontiguousarray(z), 3.5, 52, use_cells=True)
        i = np.repeat(np.arange(len(x)), nb.shape[1]).reshape(nb.shape)
        m = nb >= 0
        d = np.abs(nb[m] - i[m])
        return float(d.mean()), float((d < 256).mean())
    mean_file, near_file = index_locality(s.x, s.y, s.z)
    mean_brick, near_brick = index_locality(s.x[perm], s.y[perm], s.z[perm])
    assert mean_brick < 0.35 * mean_file, (mean_file, mean_brick)
    assert near_brick > 0.6 and near_brick > 2 * near_file, (near_file, near_brick)


def test_tiled_standin_orders_are_permutations_of_each_other(kmc):
    syn = importlib.import_module(PKG + ".synthetic")
    ref = syn.crossbar_standin(PARAM, 2, 1, order="file", vacancy_concentration=0.0)
    key = lambda t: np.lexsort((t.z, t.y, t.x))
    for order in ("xsorted", "lex", "brick", "brick10"):
        t = syn.crossbar_standin(PARAM, 2, 1, order=order, vacancy_concentration=0.0)
        assert t.N == ref.N and t.N_left == ref.N_left and t.N_right == ref.N_right
        assert np.array_equal(t.x[:t.N_left], ref.x[:ref.N_left]) and np.array_equal(t.z[-t.N_right:], ref.z[-ref.N_right:])
        kt, kr = key(t), key(ref)
        assert np.array_equal(t.x[kt], ref.x[kr]) and np.array_equal(t.y[kt], ref.y[kr]) and np.array_equal(t.z[kt], ref.z[kr])
        assert np.array_equal(t.element[kt], ref.element[kr])


def test_reorder_xyz_tool_round_trip(kmc, tmp_path):
    out = tmp_path / "brick.xyz"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reorder_xyz.py"), PARAM, str(out)], check=True,
                   capture_output=True)
    s = kmc.load_structure(PARAM, apply_vacancies=False)
    el, x, y, z = kmc.read_xyz(str(out))
    assert len(x) == s.N
    key = lambda a, b, c: np.lexsort((c, b, a))
    k1, k2 = key(x, y, z), key(s.x, s.y, s.z)
    assert np.allclose(x[k1], s.x[k2], rtol=0, atol=1e-6) and np.allclose(z[k1], s.z[k2], rtol=0, atol=1e-6)
    assert np.array_equal(el[k1], s.element[k2])
    assert np.allclose(x[:s.N_left], s.x[:s.N_left], atol=1e-6) and np.allclose(x[-s.N_right:], s.x[-s.N_right:], atol=1e-6)
