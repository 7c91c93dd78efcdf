"""Host model (parameters.txt, xyz, makeSubstoichiometric, layers, partition) against the REFERENCE's own
host code: live through oracle/_ref (compiled from /root/reference/src/input_parser.cpp + utils.cpp) when it is
present, and always against the committed dumps of it (tests/golden/ref_parser_*.json) and the shipped
snapshot_init.xyz."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLD

FILES = {"5nm": os.path.join(GOLD, "5nm_device", "parameters.txt"), "40nm": os.path.join(GOLD, "40nm_parameters.txt")}


def params_as_dict(p):
    out = {}
    for fname, _ in p._fields_:
        v = getattr(p, fname)
        if isinstance(v, bytes):
            v = v.decode()
        elif hasattr(v, "__len__"):
            v = list(v)
        out[fname] = v
    return out


@pytest.mark.parametrize("name", ["5nm", "40nm"])
def test_parser_matches_reference_dump(kmc, name):
    mine = params_as_dict(kmc.parse_parameters(FILES[name]))
    ref = json.load(open(os.path.join(GOLD, f"ref_parser_{name}.json")))
    assert mine.keys() == ref.keys()
    for k in ref:
        assert mine[k] == ref[k], (k, mine[k], ref[k])  # exact: same doubles, same strings


@pytest.mark.parametrize("name", ["5nm", "40nm"])
def test_parser_matches_reference_live(kmc, orc, name):
    L = orc.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    ref = kmc.Params()
    L.ref_parse_params(FILES[name].encode(), C.byref(ref))
    assert params_as_dict(kmc.parse_parameters(FILES[name])) == params_as_dict(ref)


def test_parameter_vectors(kmc):
    V = kmc.parse_parameter_vector(FILES["5nm"], 0)
    t = kmc.parse_parameter_vector(FILES["5nm"], 1)
    assert list(V) == [5.0] and list(t) == [1e-12]  # everything after '//' is a comment


def test_xyz_matches_reference_reader(kmc, orc):
    path = os.path.join(GOLD, "5nm_device", "reordered_device_5.xyz")
    el, x, y, z = kmc.read_xyz(path)
    assert len(el) == 37650
    L = orc.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built")
    el2 = np.zeros(37650, dtype=np.int32); x2 = np.zeros(37650); y2 = np.zeros(37650); z2 = np.zeros(37650)
    n = L.ref_read_xyz(path.encode(), 37650, el2.ctypes.data_as(C.c_void_p), x2.ctypes.data_as(C.c_void_p),
                       y2.ctypes.data_as(C.c_void_p), z2.ctypes.data_as(C.c_void_p))
    assert n == 37650
    assert (el == el2).all() and (x == x2).all() and (y == y2).all() and (z == z2).all()


def test_make_substoichiometric_reproduces_snapshot_init(kmc, s5):
    """golden (1): Results_5.000000/snapshot_init.xyz pins xyz parsing + Device RNG (seed 5) + conversion rule"""
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_init.xyz.gz"), "rt") as f:
        lines = f.read().split("\n")
    assert int(lines[0]) == s5.N
    gold = [l.split()[0] for l in lines[2:2 + s5.N]]
    mine = [kmc.ELEMENT_NAMES[e] for e in s5.element]
    assert mine == gold
    assert int((s5.element == kmc.VACANCY).sum()) == 400


def test_rng_stream_matches_reference(orc):
    L = orc.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built")
    r = C.c_void_p(L.ref_rng_create(C.c_uint(1)))
    mine = orc.Rng(1)
    for _ in range(2000):
        assert L.ref_rng_next(r) == mine.next()


def test_site_dist_and_v_solve_match_reference(orc):
    """oracle distance / potential helpers vs the reference's host site_dist / v_solve (utils.cpp:100-137, utils.h:102)"""
    L = orc.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    lat = np.array([100.0, 51.15, 51.15])
    for pbc in (0, 1):
        for _ in range(200):
            a = rng.uniform(0, 51.15, 3); b = rng.uniform(0, 51.15, 3)
            ref = L.ref_site_dist(*[C.c_double(v) for v in (*a, *b)], lat.ctypes.data_as(C.c_void_p), C.c_int(pbc))
            x = np.array([a[0], b[0]]); y = np.array([a[1], b[1]]); z = np.array([a[2], b[2]])
            rp, col = orc.block_sparsity(x, y, z, lat, pbc, ref * (1 + 1e-12) + 1e-300, 1, 1, 0, 1)
            rp2, _ = orc.block_sparsity(x, y, z, lat, pbc, ref, 1, 1, 0, 1)
            assert rp[1] == 1 and rp2[1] == 0  # dist < cutoff flips exactly at the reference's distance
    # v_solve through a 2-site Coulomb sum
    for r in (1.0, 2.5, 7.0, 19.9):
        x = np.array([0.0, r]); y = np.zeros(2); z = np.zeros(2)
        pot = orc.coulomb(x, y, z, [3, 2], [0, 2], 3.5e-10, 8.987552e9 / 23.0)
        ref = L.ref_v_solve(C.c_double(1e-10 * r), C.c_int(2), C.c_double(3.5e-10), C.c_double(8.987552e9 / 23.0),
                            C.c_double(1.60217663e-19))
        assert pot[0] == ref


def test_layers_and_partition(kmc, s5):
    t = kmc.layer_table()
    assert list(t["E_gen"]) == [0.0, 3.93, 3.93, 1.66, 1.73]
    assert list(t["E_Odiff"]) == [0.76, 0.76, 0.76, 0.76, 2.8]
    lay = s5.layer
    assert lay.min() == 0 and lay.max() == 4
    x = s5.x
    assert ((lay == 2) == ((x > 3.0) & (x < 48.1431))).all()  # oxide slab (boundaries belong to the later layer)
    for n, P in ((36498, 1), (36498, 8), (10, 3), (7, 8)):
        c, d = kmc.partition(n, P)
        assert c.sum() == n and d[0] == 0 and (np.diff(d) == c[:-1]).all()
        assert c.max() - c.min() <= 1 and (np.diff(c) <= 0).all()  # KMC_comm.h:249-263: the first n%P ranks get +1
        ca, da = kmc.partition(n, P, aligned=True)
        assert ca.sum() == n and ((da % 256 == 0) | (da == n)).all()


def test_reference_cpu_charge_sum_matches_formula(orc):
    """oracle/_ref drives the reference's surviving CPU charge sum (Device::poisson_gridless,
    src/potential_solver.cpp:74-94) over the reference's own compiled site_dist / v_solve; check it against the
    formula written out in numpy (non-periodic and periodic)."""
    import math
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(0)
    N = 80
    L = [30.0, 25.0, 20.0]
    x, y, z = rng.uniform(0, L[0], N), rng.uniform(0, L[1], N), rng.uniform(0, L[2], N)
    q = rng.choice([0, 0, 2, -2], N).astype(np.int32)
    sigma, k, e = 3.5e-10, 8.987552e9 / 23, 1.60217663e-19
    for pbc in (0, 1):
        out = orc.ref_poisson_gridless_rows(x, y, z, q, L, pbc, sigma, k, 7, 41)
        ref = np.zeros(34)
        for a, i in enumerate(range(7, 41)):
            for j in range(N):
                if i != j and q[j] != 0:
                    dx, dy, dz = abs(x[i] - x[j]), abs(y[i] - y[j]), abs(z[i] - z[j])
                    if pbc:  # site_dist, src/utils.cpp: minimum image in y and z only
                        dy, dz = min(dy, L[1] - dy), min(dz, L[2] - dz)
                    r = 1e-10 * math.sqrt(dx * dx + dy * dy + dz * dz)
                    ref[a] += q[j] * math.erfc(r / (sigma * math.sqrt(2))) * k * e / r
        assert np.allclose(out, ref, rtol=1e-13, atol=0), pbc
