"""The CPU oracle itself: pinned to the reference's shipped golden run, and internally consistent
(cell-list enumeration == the reference's brute-force loops, summation spec sane, selection rule ==
thrust::upper_bound semantics)."""
import gzip
import json
import os

import numpy as np

from conftest import GOLD, make_synthetic


def test_oracle_reproduces_reference_golden_run(kmc, orc, s5):
    """golden (2)+(3) of SURVEY.md section 8c: structures/5nm_device/expected_output."""
    sim = orc.OracleSim(s5)
    assert int(sim.sp["row_ptr"][-1]) == 940008
    times, events, cg = [], [], []
    while sim.kmc_time < s5.t_switch:
        r = sim.superstep()
        times.append(r["kmc_time"]); cg.append(r["cg_iterations"])
        events += [tuple(int(v) for v in e[:3]) for e in r["events"]]
    # six "KMC time is:" lines of output1_0.txt
    gold = [float(l.split(":")[1]) for l in open(os.path.join(GOLD, "5nm_device", "output1_0.txt")) if l.startswith("KMC time is")]
    assert len(gold) == 6 and len(times) == 6
    assert np.allclose(times, gold, rtol=1e-3)
    assert np.abs(np.array(times) / np.array(gold) - 1).max() < 1e-4   # observed 5e-5
    # the 8 events (all VACANCY_DIFFUSION), identical (i, j) and order
    assert events == [(14620, 14113, 2), (8420, 7760, 2), (8261, 7557, 2), (10445, 9726, 2), (14607, 13989, 2),
                      (9291, 9180, 2), (9790, 9138, 2), (14545, 13865, 2)]
    assert cg[0] > 300 and cg[1:] == [0] * 5   # cold PCG, then warm starts that already satisfy the test
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_6.xyz.gz"), "rt") as f:
        rows = [l.split() for l in f.read().split("\n")[2:2 + s5.N]]
    assert [kmc.ELEMENT_NAMES[e] for e in sim.element] == [r[0] for r in rows]
    gp = np.array([float(r[4]) for r in rows])
    assert np.abs(gp - sim.pot_total).max() < 5e-4          # reference's own PCG-order noise floor (observed 2.6e-5)
    # contact-layer sites: no boundary part in this code path and > 20 A from every charged defect
    assert (gp[:s5.N_left] == 0).all() and (sim.pot_total[:s5.N_left] == 0).all()
    assert (gp[-s5.N_right:] == 0).all() and (sim.pot_total[-s5.N_right:] == 0).all()
    assert all(float(r[5]) == 0 for r in rows)   # site_power column


def test_trajectory_fixture_prefix(kmc, orc, s5):
    """the committed 1000-step fixture is what the oracle produces (first 25 steps re-run here)"""
    traj = json.load(open(os.path.join(GOLD, "traj_5nm.json")))
    sim = orc.OracleSim(s5)
    for st in traj["steps"][:25]:
        r = sim.superstep()
        assert r["n_events"] == st["n_events"] and r["cg_iterations"] == st["cg"]
        assert [[int(v) for v in e[:3]] for e in r["events"]] == st["events"]
        assert r["event_time"] == float.fromhex(st["event_time"])


def test_cell_enumeration_equals_brute_force(kmc, orc, s_small):
    s = s_small
    a = orc.neighbor_list(s.x, s.y, s.z, 3.5, 52, use_cells=False)
    b = orc.neighbor_list(s.x, s.y, s.z, 3.5, 52, use_cells=True)
    assert (a == b).all()
    a = orc.neighbor_list(s.x, s.y, s.z, 3.5, 5, 10, 50, use_cells=False)
    b = orc.neighbor_list(s.x, s.y, s.z, 3.5, 5, 10, 50, use_cells=True)
    assert (a == b).all() and (a[:, -1] >= 0).any()
    for pbc in (0, 1):
        k1 = orc.sparsity_K(s.x, s.y, s.z, s.lattice, pbc, 3.5, s.N_left, s.N_right, False)
        k2 = orc.sparsity_K(s.x, s.y, s.z, s.lattice, pbc, 3.5, s.N_left, s.N_right, True)
        assert all((k1[k] == k2[k]).all() for k in k1)
    k0 = orc.sparsity_K(s.x, s.y, s.z, s.lattice, 0, 3.5, s.N_left, s.N_right)
    k1 = orc.sparsity_K(s.x, s.y, s.z, s.lattice, 1, 3.5, s.N_left, s.N_right)
    assert k1["row_ptr"][-1] > k0["row_ptr"][-1]   # periodic images add neighbours


def test_cutoff_list_equivalent_to_inline_predicate(kmc, orc, s_small):
    """the Coulomb oracle (inline membership test) == summing over the reference's materialised cutoff list"""
    s = s_small
    charge = np.zeros(s.N, np.int32); charge[s.element == kmc.VACANCY] = 2; charge[s.element == kmc.OXYGEN_DEFECT] = -2
    cnt = orc.cutoff_count(s.element, s.x, s.y, s.z, 6.0)
    lst = orc.cutoff_list(s.element, s.x, s.y, s.z, int(cnt.max()), 6.0)
    assert ((lst >= 0).sum(1) == cnt).all()
    pot = orc.coulomb(s.x, s.y, s.z, s.element, charge, s.sigma, s.k, cutoff=6.0)
    from math import erfc, sqrt
    for i in (0, 77, 200, s.N - 1):
        acc = 0.0
        for j in lst[i]:
            if j >= 0 and j != i and charge[j] != 0:
                d = 1e-10 * sqrt((s.x[j] - s.x[i]) ** 2 + (s.y[j] - s.y[i]) ** 2 + (s.z[j] - s.z[i]) ** 2)
                acc += float(charge[j]) * erfc(d / (s.sigma * sqrt(2.0))) * s.k * 1.60217663e-19 / d
        assert abs(acc - pot[i]) <= 1e-13 * max(1e-30, abs(acc))


def test_coulomb_cell_variant_is_bit_identical(kmc, orc, s_small, s5):
    """orc_coulomb_cells (what orc_superstep / the CPU baseline run) == the all-sources loop, bit for bit"""
    for s, cut in ((s_small, 6.0), (s_small, 20.0), (s5, 20.0)):
        charge = np.zeros(s.N, np.int32); charge[s.element == kmc.VACANCY] = 2; charge[s.element == kmc.OXYGEN_DEFECT] = -2
        a = orc.coulomb(s.x, s.y, s.z, s.element, charge, s.sigma, s.k, cutoff=cut)
        b = orc.coulomb(s.x, s.y, s.z, s.element, charge, s.sigma, s.k, cutoff=cut, use_cells=True)
        assert (a == b).all() and np.abs(a).max() > 0
    part = orc.coulomb(s5.x, s5.y, s5.z, s5.element, charge, s5.sigma, s5.k, row_start=9000, row_count=500, use_cells=True)
    assert (part[9000:9500] == a[9000:9500]).all() and (part[:9000] == 0).all()


def test_summation_spec(orc):
    rng = np.random.default_rng(0)
    for n in (1, 100, 256, 257, 5000):
        u, v = rng.standard_normal(n), rng.standard_normal(n)
        assert abs(orc.dot(u, v) - float(np.dot(u, v))) < 1e-12 * n
    v = rng.uniform(0, 1, 256)
    incl = orc.block_scan_256(v)
    assert np.allclose(incl, np.cumsum(v), rtol=1e-14)
    assert (np.diff(incl) >= 0).all()


def test_selection_rule_is_upper_bound(orc):
    rng = np.random.default_rng(4)
    N, nn = 1000, 52
    prob = np.zeros(N * nn)
    idx = rng.choice(N * nn, 300, replace=False)
    prob[idx] = 10.0 ** rng.uniform(-30, 12, 300)
    cum = np.cumsum(prob)
    agree = 0
    for u in rng.uniform(0, 1, 200):
        slot, psum = orc.select_event(prob, N, nn, u * cum[-1])
        assert prob[slot] > 0
        agree += int(slot == int(np.searchsorted(cum, u * cum[-1], side="right")))
        assert abs(psum - cum[-1]) < 1e-12 * cum[-1]
    assert agree >= 198   # differs from a flat scan only when u*Psum lands within rounding of a boundary
    slot, psum = orc.select_event(np.zeros(N * nn), N, nn, 0.0)
    assert slot == -1 and psum == 0.0


def test_pcg_oracle_solves(orc):
    import scipy.sparse as sp
    import scipy.sparse.linalg
    n = 2000
    A = sp.random(n, n, density=10.0 / n, random_state=1, format="csr")
    A = (A + A.T + sp.diags(np.full(n, 30.0))).tocsr(); A.sort_indices()
    b = np.random.default_rng(1).standard_normal(n)
    x, r, it, stats = orc.pcg_jacobi(A.indptr, A.indices, A.data, 1.0 / A.diagonal(), b, np.zeros(n), 1e-12, 500)
    assert 0 < it < 200
    assert np.abs(x - scipy.sparse.linalg.spsolve(A.tocsc(), b)).max() < 1e-9
    x2, _, it2, _ = orc.pcg_jacobi(A.indptr, A.indices, A.data, 1.0 / A.diagonal(), b, x, 1e-12, 500)
    assert it2 == 0 and (x2 == x).all()
    x3, _, it3, _ = orc.pcg_jacobi(A.indptr, A.indices, A.data, 1.0 / A.diagonal(), np.zeros(n), np.zeros(n), 1e-12, 500)
    assert it3 == 0   # b = 0: 0/0 comparison is false, like the reference's while condition
