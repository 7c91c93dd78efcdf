"""GPU parity: every stage of the hot path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact for integer / index work; FP64 tolerances are written next to each check."""
import gzip
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLD, make_synthetic

pytestmark = pytest.mark.gpu


def to_np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def dev5(kmc, ctx, s5):
    return kmc.DeviceKMC(s5, ctx=ctx)


@pytest.fixture(scope="module")
def orc5(orc, s5):
    return orc.OracleSim(s5)


# ---------------------------------------------------------------- a1 / a2 / a3: index structures, bit exact
def test_neighbor_list_bit_exact(dev5, orc5):
    assert (to_np(dev5.neigh) == orc5.neigh).all()


def test_neighbor_list_row_range_and_small_nn(kmc, ctx, orc, s_small):
    s = s_small
    x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
    for nn, rs, rc in ((52, 0, s.N), (6, 17, 200), (3, s.N - 5, 5)):
        got = to_np(ctx.compute_neighbor_list(x, y, z, 3.5, nn, rs, rc))
        want = orc.neighbor_list(s.x, s.y, s.z, 3.5, nn, rs, rc)
        assert (got == want).all()   # cap applied after ascending ordering


def test_cutoff_size_and_list(kmc, ctx, orc, s_small):
    s = s_small
    x, y, z, el = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z), ctx.dev_i(s.element)
    want = orc.cutoff_count(s.element, s.x, s.y, s.z, 20.0)
    mx, counts = ctx.cutoff_size(el, x, y, z, 20.0, want_counts=True)
    assert (to_np(counts) == want).all() and mx == want.max()
    got = to_np(ctx.cutoff_list(el, x, y, z, mx))
    assert (got == orc.cutoff_list(s.element, s.x, s.y, s.z, mx)).all()


def test_cutoff_size_5nm(dev5, ctx, s5):
    mx, _ = ctx.cutoff_size(dev5.element, dev5.x, dev5.y, dev5.z, 20.0)
    assert mx == 4217   # N_cutoff of the shipped device


def test_sparsity_K_bit_exact_5nm(dev5, orc5):
    K = dev5.K.to_host()
    assert dev5.K.nnz == 940008 and dev5.K.left_nnz == 2784 and dev5.K.right_nnz == 2784
    for k in ("row_ptr", "col", "left_row_ptr", "left_col", "right_row_ptr", "right_col"):
        assert (K[k] == orc5.sp[k]).all(), k


@pytest.mark.parametrize("pbc", [0, 1])
def test_sparsity_K_small_pbc_and_blocks(kmc, ctx, orc, pbc):
    s = make_synthetic(kmc, pbc=pbc, seed=5)
    x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
    want = orc.sparsity_K(s.x, s.y, s.z, s.lattice, pbc, s.nn_dist, s.N_left, s.N_right)
    K = ctx.initialize_sparsity_K(x, y, z, s.lattice, pbc, s.nn_dist, s.N_left, s.N_right)
    got = K.to_host()
    for k in want:
        assert (got[k] == want[k]).all(), k
    # the reference's Distributed_matrix layout: per (rank, neighbour rank) sub-CSR with block-local columns
    n = s.N - s.N_left - s.N_right
    P = 3
    counts, displs = kmc.partition(n, P)
    for r in range(P):
        Kr = ctx.initialize_sparsity_K(x, y, z, s.lattice, pbc, s.nn_dist, s.N_left, s.N_right, int(displs[r]), int(counts[r]))
        for q in range(P):
            rp, col = Kr.block_view(int(displs[q]), int(counts[q]))
            wrp, wcol = orc.block_sparsity(s.x, s.y, s.z, s.lattice, pbc, s.nn_dist, int(counts[r]), int(counts[q]),
                                           s.N_left + int(displs[r]), s.N_left + int(displs[q]))
            assert (rp == wrp).all() and (col == wcol).all()
        Kr.close()
    K.close()


# ---------------------------------------------------------------- a5: charges, exact
def test_update_charge_exact(dev5, orc, s5, ctx):
    ch = ctx.empty_i(s5.N, 0)
    ctx.update_charge(dev5.element, ch, dev5.neigh, s5.metals)
    want = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), to_np(dev5.neigh), s5.metals)
    got = to_np(ch)
    assert (got == want).all()
    assert 300 < (got != 0).sum() < 400


# ---------------------------------------------------------------- a6: assembly, bit exact (values are +-{1,1e-8} sums)
def test_assemble_K_bit_exact(kmc, ctx, orc, s5, dev5, orc5):
    charge = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), orc5.neigh, s5.metals)
    # make some vacancies uncharged neighbours so the cvacancy branch is exercised
    data, dinv, rhs = orc.assemble_K(s5.N, s5.N_left, s5.N_right, s5.element, charge, s5.metals, orc5.sp, s5.Vd,
                                     s5.high_G, s5.low_G)
    ctx.assemble_K(dev5.K, s5.N, s5.N_left, s5.N_right, ctx.dev_i(s5.element), ctx.dev_i(charge), s5.metals, s5.Vd,
                   s5.high_G, s5.low_G)
    K = dev5.K.to_host()
    assert (K["val"] == data).all()
    assert (K["inv_diag"] == dinv).all()
    assert (K["rhs"] == rhs).all()
    # algebraic invariants of the reference's postprocessing/test_matrices.py:38-48
    import scipy.sparse as sp
    A = sp.csr_matrix((K["val"], K["col"], K["row_ptr"]))
    assert abs(A - A.T).max() == 0.0
    assert (A.diagonal() > 0).all()


# ---------------------------------------------------------------- summation spec primitives, bit exact
def test_dot_and_spmv_bit_exact(ctx, orc, dev5, orc5, s5):
    rng = np.random.default_rng(1)
    for n in (1, 255, 256, 257, 36498, 100003):
        u = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        v = rng.standard_normal(n)
        assert ctx.dot(ctx.dev_d(u), ctx.dev_d(v)) == orc.dot(u, v)
    K = dev5.K.to_host()
    xv = rng.standard_normal(dev5.K.rows)
    y = ctx.empty_d(dev5.K.rows)
    ctx.spmv(dev5.K, ctx.dev_d(xv), y)
    want = orc.spmv(K["row_ptr"], K["col"], K["val"], xv)
    assert (to_np(y) == want).all()
    y2 = ctx.empty_d(dev5.K.rows)
    d = ctx.spmv_dot(dev5.K, ctx.dev_d(xv), y2)          # fused SpMV + x.(Ax), the PCG's hot kernel
    assert (to_np(y2) == want).all() and d == orc.dot(xv, want)


# ---------------------------------------------------------------- a7: PCG, same iteration count, bit-identical iterates
def test_pcg_cold_start_identical(kmc, ctx, orc, s5, dev5, orc5):
    charge = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), orc5.neigh, s5.metals)
    data, dinv, rhs = orc.assemble_K(s5.N, s5.N_left, s5.N_right, s5.element, charge, s5.metals, orc5.sp, s5.Vd,
                                     s5.high_G, s5.low_G)
    n = s5.N - 2 * s5.N_left
    tol = 1e-14 * n
    xo, ro, it_o, stats = orc.pcg_jacobi(orc5.sp["row_ptr"], orc5.sp["col"], data, dinv, rhs, np.zeros(n), tol)
    pot = ctx.empty_d(s5.N, 0.0)
    it_g = ctx.background_potential(dev5.K, s5.N, s5.N_left, s5.N_right, ctx.dev_i(s5.element), ctx.dev_i(charge),
                                    s5.metals, s5.Vd, s5.high_G, s5.low_G, pot)
    assert it_g == it_o and it_o > 100
    got = to_np(pot)
    assert (got[:s5.N_left] == 0).all() and (got[-s5.N_right:] == 0).all()
    # north_star tolerance 1e-10 relative; the spec makes it bit-identical
    assert np.abs(got[s5.N_left:-s5.N_right] - xo).max() <= 1e-10 * np.abs(xo).max()
    assert (got[s5.N_left:-s5.N_right] == xo).all()
    # warm start: 0 iterations
    it2 = ctx.background_potential(dev5.K, s5.N, s5.N_left, s5.N_right, ctx.dev_i(s5.element), ctx.dev_i(charge),
                                   s5.metals, s5.Vd, s5.high_G, s5.low_G, pot)
    assert it2 == 0


def test_pcg_generic_csr_random_spd(ctx, orc):
    """dist_iterative_test shape: exported CSR + rhs, solve and compare with the oracle (main_test_cg.cpp:125-135)"""
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    n = 7302
    A = sp.random(n, n, density=25.0 / n, random_state=3, format="csr")
    A = A + A.T + sp.diags(np.full(n, 60.0))
    A = A.tocsr(); A.sort_indices()
    b = rng.standard_normal(n)
    dinv = 1.0 / A.diagonal()
    rp = A.indptr.astype(np.int32); col = A.indices.astype(np.int32); val = A.data
    xo, ro, it_o, _ = orc.pcg_jacobi(rp, col, val, dinv, b, np.zeros(n), 1e-11, 500)
    K = ctx.kmat_from_csr(ctx.dev_i(rp), ctx.dev_i(col), ctx.dev_d(val))
    r = ctx.dev_d(b); x = ctx.empty_d(n, 0.0)
    it_g = ctx.pcg_jacobi(K, r, x, ctx.dev_d(dinv), 1e-11, 500)
    assert it_g == it_o
    assert (to_np(x) == xo).all()
    ref = sp.linalg.spsolve(A.tocsc(), b)
    assert np.abs(to_np(x) - ref).max() < 1e-8
    K.close()


# ---------------------------------------------------------------- a8 / a9: Coulomb sum
def test_coulomb_matches_oracle(ctx, orc, s5, dev5, orc5):
    charge = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), orc5.neigh, s5.metals)
    want = orc.coulomb(s5.x, s5.y, s5.z, s5.element, charge, s5.sigma, s5.k)
    pot = ctx.empty_d(s5.N, -1.0)
    ctx.poisson_gridless(dev5.x, dev5.y, dev5.z, dev5.element, ctx.dev_i(charge), s5.sigma, s5.k, pot)
    got = to_np(pot)
    # same ascending-j summation order; only erfc differs (CUDA vs glibc, <= 2 ulp per term)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    q, tests, inrange = ctx.poisson_stats()
    assert q == int((charge != 0).sum())
    # the in-range pair count the Coulomb roofline of bench.py is built on, against a KD-tree count
    from scipy.spatial import cKDTree
    pts = np.stack([s5.x, s5.y, s5.z], 1)
    src = np.nonzero(charge)[0]
    nb = cKDTree(pts[src]).query_ball_point(pts, 20.0, return_length=True)
    exact = int(nb.sum()) - len(src)           # minus the i == j pairs of the charged sites themselves
    assert abs(inrange - exact) <= 1e-6 * exact  # (KD-tree uses <=, the kernel <: identical unless a pair sits at 20.0 A)
    assert tests >= inrange
    # row sub-range leaves other rows untouched
    pot2 = ctx.empty_d(s5.N, -7.0)
    ctx.poisson_gridless(dev5.x, dev5.y, dev5.z, dev5.element, ctx.dev_i(charge), s5.sigma, s5.k, pot2, row_start=1000, row_count=500)
    g2 = to_np(pot2)
    assert (g2[:1000] == -7.0).all() and (g2[1500:] == -7.0).all() and (g2[1000:1500] == got[1000:1500]).all()
    b = ctx.dev_d(np.arange(s5.N, dtype=float))
    ctx.sum_potential(pot, b)
    assert (to_np(pot) == got + np.arange(s5.N)).all()


def test_coulomb_no_charges_and_mixed_signs(kmc, ctx, orc, s_small):
    s = s_small
    x, y, z, el = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z), ctx.dev_i(s.element)
    pot = ctx.empty_d(s.N, 5.0)
    ctx.poisson_gridless(x, y, z, el, ctx.empty_i(s.N, 0), s.sigma, s.k, pot)
    assert (to_np(pot) == 0).all()
    charge = np.zeros(s.N, np.int32)
    charge[s.element == kmc.VACANCY] = 2
    charge[s.element == kmc.OXYGEN_DEFECT] = -2
    want = orc.coulomb(s.x, s.y, s.z, s.element, charge, s.sigma, s.k)
    ctx.poisson_gridless(x, y, z, el, ctx.dev_i(charge), s.sigma, s.k, pot)
    assert np.abs(to_np(pot) - want).max() <= 1e-12 * np.abs(want).max()


# ---------------------------------------------------------------- a10 / a11: events
def test_rng_stream_identical(ctx, orc, dev5):
    dev5.ev.rng_seed(1)
    got = dev5.ev.rng_draw(3000)
    r = orc.Rng(1)
    want = np.array([r.next() for _ in range(3000)])
    assert (got == want).all()
    mt, pos = r.state()
    mt2, pos2 = dev5.ev.rng_get_state()
    assert pos == pos2 and (mt == mt2).all()
    dev5.ev.rng_seed(1)


def test_event_rates_match_oracle(kmc, ctx, orc, s_small):
    s = s_small
    rng = np.random.default_rng(2)
    sim = orc.OracleSim(s)
    charge = orc.update_charge(s.element, np.zeros(s.N, np.int32), sim.neigh, s.metals)
    pot = rng.uniform(-2.5, 2.5, s.N)
    typ, prob = orc.build_events(sim.neigh, s.layer, s.T_bg, s.freq, s.sigma, s.k, s.x, s.y, s.z, pot, s.element, charge, s.E)
    neigh = ctx.dev_i(sim.neigh.ravel()).view(s.N, 52)
    ev = ctx.events_create(neigh)
    ev.set_activation_energies(s.E["E_gen"], s.E["E_rec"], s.E["E_Vdiff"], s.E["E_Odiff"])
    ev.build_event_list(neigh, ctx.dev_i(s.layer), s.T_bg, s.freq, s.sigma, s.k, ctx.dev_d(s.x), ctx.dev_d(s.y),
                        ctx.dev_d(s.z), ctx.dev_d(pot), ctx.dev_i(s.element), ctx.dev_i(charge))
    gp, gt = ev.event_arrays()
    assert (gt == typ).all()
    assert set(np.unique(typ)) >= {0, 1, 2, 3, 4}  # all four event classes + NULL are present
    nz = prob > 0
    assert (gp[~nz] == 0).all()
    # exp/erfc differ by a few ulp between CUDA and glibc; rates span > 40 decades
    assert np.abs(gp[nz] / prob[nz] - 1).max() < 1e-12
    ev.close()


def test_superstep_sequence_5nm_golden_and_oracle(kmc, ctx, orc, s5):
    """6 supersteps of the shipped 5 nm run: identical events to the oracle AND to the reference's golden
    output (snapshot_6.xyz elements, 'KMC time is:' lines to 1e-3, potentials to 5e-4 V)."""
    dev = kmc.DeviceKMC(s5, ctx=ctx)
    sim = orc.OracleSim(s5)
    times = []
    while dev.kmc_time < s5.t_switch:
        et, ne = dev.superstep()
        log, psum = dev.ev.log()
        r = sim.superstep()
        assert ne == r["n_events"] and dev.last_cg_iterations == r["cg_iterations"]
        assert (log[:, :3] == r["events"][:, :3]).all()
        assert abs(et - r["event_time"]) <= 1e-12 * r["event_time"]
        times.append(dev.kmc_time)
    assert dev.step_count == 6
    gold_t = [2.91075e-14, 5.12158e-14, 9.36848e-14, 2.6667e-13, 9.45779e-13, 1.06019e-12]  # output1_0.txt
    assert np.allclose(times, gold_t, rtol=1e-3)
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_6.xyz.gz"), "rt") as f:
        rows = [l.split() for l in f.read().split("\n")[2:2 + s5.N]]
    assert [kmc.ELEMENT_NAMES[e] for e in to_np(dev.element)] == [r[0] for r in rows]
    gp = np.array([float(r[4]) for r in rows])
    pot = to_np(dev.pot_charge)
    assert np.abs(pot - gp).max() < 5e-4
    assert np.abs(pot - sim.pot_total).max() <= 1e-10 * np.abs(sim.pot_total).max()
    assert (to_np(dev.charge) == sim.charge).all()


def test_trajectory_1000_steps_matches_fixture(kmc, ctx, s5):
    """north_star: identical KMC event sequence over the first 1000 steps of the 5 nm device at fixed seed
    (fixture generated by the oracle: tests/golden/make_golden.py)."""
    traj = json.load(open(os.path.join(GOLD, "traj_5nm.json")))
    dev = kmc.DeviceKMC(s5, ctx=ctx)
    for k, st in enumerate(traj["steps"]):
        et, ne = dev.superstep()
        log, _ = dev.ev.log()
        assert ne == st["n_events"], f"first divergent step {k}: event count {ne} vs {st['n_events']}"
        assert log[:, :3].tolist() == st["events"], f"first divergent step {k}"
        assert dev.last_cg_iterations == st["cg"], f"step {k}: PCG iterations {dev.last_cg_iterations} vs {st['cg']}"
        want = float.fromhex(st["event_time"])
        assert abs(et - want) <= 1e-12 * want
    fin = traj["final"]
    assert abs(dev.kmc_time - float.fromhex(fin["kmc_time"])) <= 1e-11 * dev.kmc_time
    assert int((to_np(dev.element) == kmc.VACANCY).sum()) == fin["n_vacancy"]
    assert int((to_np(dev.charge) != 0).sum()) == fin["n_charged"]


def test_event_loop_all_event_types_small(kmc, ctx, orc):
    """synthetic device where generation / recombination / both diffusions all fire; capped event count"""
    s = make_synthetic(kmc, seed=11, vac=0.15)
    dev = kmc.DeviceKMC(s, ctx=ctx)
    sim = orc.OracleSim(s)
    seen = set()
    for step in range(12):
        dev.field_solve()
        et, ne = dev.ev.execute_kmc_step(dev.neigh, dev.layer, s.T_bg, s.freq, s.sigma, s.k, dev.x, dev.y, dev.z,
                                         dev.pot_charge, dev.element, dev.charge, max_events=40)
        log, psum = dev.ev.log()
        # oracle: same stages, same cap
        sim.charge = orc.update_charge(sim.element, sim.charge, sim.neigh, s.metals)
        data, dinv, rhs = orc.assemble_K(s.N, s.N_left, s.N_right, sim.element, sim.charge, s.metals, sim.sp, s.Vd, s.high_G, s.low_G)
        n = s.N - s.N_left - s.N_right
        xo, _, it, _ = orc.pcg_jacobi(sim.sp["row_ptr"], sim.sp["col"], data, dinv, rhs, sim.pot_boundary[s.N_left:s.N_left + n], 1e-14 * n)
        sim.pot_boundary[s.N_left:s.N_left + n] = xo
        assert it == dev.last_cg_iterations
        pot = orc.coulomb(s.x, s.y, s.z, sim.element, sim.charge, s.sigma, s.k) + sim.pot_boundary
        typ, prob = orc.build_events(sim.neigh, s.layer, s.T_bg, s.freq, s.sigma, s.k, s.x, s.y, s.z, pot, sim.element, sim.charge, s.E)
        r = orc.event_loop(sim.neigh, typ, prob, sim.element, sim.charge, s.freq, sim.rng, max_events=40)
        sim.element, sim.charge = r["element"], r["charge"]
        assert ne == r["n_events"]
        assert (log == r["log"]).all(), f"step {step}"
        assert np.allclose(psum, r["psum"], rtol=1e-11, atol=0)
        assert abs(et - r["event_time"]) <= 1e-11 * abs(r["event_time"])
        assert (to_np(dev.element) == sim.element).all() and (to_np(dev.charge) == sim.charge).all()
        seen |= set(log[:, 2].tolist())
    assert seen >= {1, 2, 3}  # (generation fires in the 5 nm 1000-step fixture: 12 events)


def test_cpp_host_driver_reproduces_golden_output(kmc, tmp_path):
    """The C++ host (host/kmc_main.cpp: reference main() call order through include/gpu_solvers_b200.hpp) run on the
    shipped 5 nm inputs writes the reference's output files: 'KMC time is:' lines within 1e-3 of output1_0.txt and the
    final snapshot's elements identical to snapshot_6.xyz."""
    import subprocess
    exe = os.path.join(os.path.dirname(kmc.LIB_PATH), "kmc_b200_run")
    if not os.path.exists(exe):
        pytest.skip("kmc_b200_run not built")
    r = subprocess.run([exe, os.path.join(GOLD, "5nm_device", "parameters.txt")], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = open(tmp_path / "output1_0.txt").read()
    mine = [float(l.split(":")[1]) for l in out.split("\n") if l.startswith("KMC time is")]
    gold = [float(l.split(":")[1]) for l in open(os.path.join(GOLD, "5nm_device", "output1_0.txt")) if l.startswith("KMC time is")]
    assert len(mine) == len(gold) == 6 and np.allclose(mine, gold, rtol=1e-3)
    snap = open(tmp_path / "Results_5.000000" / "snapshot_6.xyz").read().split("\n")
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_6.xyz.gz"), "rt") as f:
        gsnap = f.read().split("\n")
    assert [l.split()[0] for l in snap[2:37652]] == [l.split()[0] for l in gsnap[2:37652]]
    pm = np.array([float(l.split()[4]) for l in snap[2:37652]]); pg = np.array([float(l.split()[4]) for l in gsnap[2:37652]])
    assert np.abs(pm - pg).max() < 5e-4
    init = open(tmp_path / "Results_5.000000" / "snapshot_init.xyz").read().split("\n")
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_init.xyz.gz"), "rt") as f:
        ginit = f.read().split("\n")
    assert init[:37652] == ginit[:37652]   # byte-identical initial snapshot (same ostream formatting)


def test_cpp_host_driver_runs_the_current_solver(kmc, orc, s5, tmp_path):
    """With comm_T kept alive (KMCB200_ENABLE_CURRENT=1; the reference forces it to MPI_COMM_NULL, src/KMC_comm.h:243) the C++
    host runs setLaplacePotential -> initialize_sparsity_T -> update_power_gpu_sparse_dist through the reference-named
    entry points and logs the macroscopic current: the first superstep's value against the oracle chain."""
    import subprocess
    exe = os.path.join(os.path.dirname(kmc.LIB_PATH), "kmc_b200_run")
    if not os.path.exists(exe):
        pytest.skip("kmc_b200_run not built")
    env = dict(os.environ, KMCB200_ENABLE_CURRENT="1")
    r = subprocess.run([exe, os.path.join(GOLD, "5nm_device", "parameters.txt"), "2"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    out = open(tmp_path / "output1_0.txt").read()
    im = [float(l.split(":")[1]) for l in out.split("\n") if l.startswith("I_macro")]
    assert len(im) == 2
    sim = orc.OracleSim(s5, use_cells=True)
    ch = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), sim.neigh, s5.metals)
    ko = orc.KirchhoffOracle(s5, sim.sp, 10, site_charge=ch)
    ko.solve()
    assert abs(im[0] - ko.imacro * 1e6) <= 2e-5 * abs(ko.imacro * 1e6)    # printed with 6 significant digits
    # the KMC trajectory is unchanged by the current solver
    mine = [float(l.split(":")[1]) for l in out.split("\n") if l.startswith("KMC time is")]
    gold = [float(l.split(":")[1]) for l in open(os.path.join(GOLD, "5nm_device", "output1_0.txt")) if l.startswith("KMC time is")]
    assert np.allclose(mine, gold[:2], rtol=1e-3)


def test_cpp_host_driver_bias_sweep(kmc, ctx, s5, tmp_path):
    """f-4: the bias loop over the V_switch / t_switch vectors (src/kmc_main.cpp:255-575): two bias points; the second
    starts from the structure the first one left, the warm PCG start and the KMC generator carry over.  The C++ host's
    'KMC time is:' lines against the Python mirror of the same loop (same library, independent host code)."""
    import shutil, subprocess
    exe = os.path.join(os.path.dirname(kmc.LIB_PATH), "kmc_b200_run")
    if not os.path.exists(exe):
        pytest.skip("kmc_b200_run not built")
    src = os.path.join(GOLD, "5nm_device")
    lines = open(os.path.join(src, "parameters.txt")).read().split("\n")
    out = []
    for l in lines:
        if l.startswith("V_switch"):
            l = "V_switch = 5 6 // two bias points"
        if l.startswith("t_switch"):
            l = "t_switch = 6e-14 8e-14 // [s]"
        out.append(l)
    open(tmp_path / "parameters.txt", "w").write("\n".join(out))
    shutil.copy(os.path.join(src, "reordered_device_5.xyz"), tmp_path / "reordered_device_5.xyz")
    r = subprocess.run([exe, str(tmp_path / "parameters.txt")], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    txt = open(tmp_path / "output1_0.txt").read()
    assert "Applied Voltage = 5 V" in txt and "Applied Voltage = 6 V" in txt
    mine = [float(l.split(":")[1]) for l in txt.split("\n") if l.startswith("KMC time is")]
    # Python mirror of the bias loop
    dev = kmc.DeviceKMC(s5, ctx=ctx)
    want = []
    for Vd, t in ((5.0, 6e-14), (6.0, 8e-14)):
        dev.s.Vd = Vd
        dev.kmc_time = 0.0
        while dev.kmc_time < t:
            dev.superstep()
            want.append(dev.kmc_time)
    dev.s.Vd = 5.0
    assert len(mine) == len(want) >= 3 and np.allclose(mine, want, rtol=2e-5)    # 6 printed digits
    assert os.path.exists(tmp_path / "Results_5.000000" / "snapshot_init.xyz")
    snaps6 = sorted(os.listdir(tmp_path / "Results_6.000000"))
    assert "snapshot_init.xyz" in snaps6 and len(snaps6) == 2
    end5 = [f for f in os.listdir(tmp_path / "Results_5.000000") if f != "snapshot_init.xyz"][0]
    a = [l.split()[0] for l in open(tmp_path / "Results_5.000000" / end5).read().split("\n")[2:37652]]
    b = [l.split()[0] for l in open(tmp_path / "Results_6.000000" / "snapshot_init.xyz").read().split("\n")[2:37652]]
    assert a == b    # the second bias point starts from the structure the first one left


def test_cpp_host_driver_two_ranks(kmc, tmp_path):
    """The compiled host at N > 1 (VERDICT r1 missing #3): two processes of kmc_b200_run -- one rank per process, bootstrap
    (CUDA-IPC handles, halo need maps, barriers) over a shared directory (kmcb200_rdv_*), K solve row-sharded with halo /
    dot exchanges in peer memory, Coulomb rows sharded, potentials all-gathered over peer memory, events replicated.
    Both ranks must log the golden run's KMC times and rank 0 must write the golden final structure.  Uses two GPUs when
    the box has them, else both ranks share GPU 0 (CUDA IPC works between processes on one device; slower, same code)."""
    import subprocess
    import torch
    exe = os.path.join(os.path.dirname(kmc.LIB_PATH), "kmc_b200_run")
    if not os.path.exists(exe):
        pytest.skip("kmc_b200_run not built")
    ngpu = torch.cuda.device_count()
    rdv = tmp_path / "rdv"
    procs = []
    for r in range(2):
        env = dict(os.environ, KMCB200_RANK=str(r), KMCB200_WORLD_SIZE="2", KMCB200_RENDEZVOUS=str(rdv),
                   KMCB200_DEVICE=str(r if ngpu >= 2 else 0), KMCB200_COMM_TIMEOUT_MS="120000")
        procs.append(subprocess.Popen([exe, os.path.join(GOLD, "5nm_device", "parameters.txt")], cwd=tmp_path, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    gold = [float(l.split(":")[1]) for l in open(os.path.join(GOLD, "5nm_device", "output1_0.txt")) if l.startswith("KMC time is")]
    times = []
    for r in range(2):
        out = open(tmp_path / f"output2_{r}.txt").read()
        times.append([l for l in out.split("\n") if l.startswith("KMC time is")])
        mine = [float(l.split(":")[1]) for l in times[-1]]
        assert len(mine) == 6 and np.allclose(mine, gold, rtol=1e-3)
        times[-1] += [l for l in out.split("\n") if l.startswith("PCG iterations")]
    assert times[0] == times[1]                      # replicated event selection: identical trajectories
    # ... and identical to the single-rank run of the same binary: same PCG iteration counts (cold solve included), same
    # printed KMC times
    one = tmp_path / "one"
    os.makedirs(one)
    r1 = subprocess.run([exe, os.path.join(GOLD, "5nm_device", "parameters.txt")], cwd=one, capture_output=True, text=True,
                        timeout=600)
    assert r1.returncode == 0, r1.stderr[-2000:]
    out1 = open(one / "output1_0.txt").read().split("\n")
    ref = [l for l in out1 if l.startswith("KMC time is")] + [l for l in out1 if l.startswith("PCG iterations")]
    assert times[0] == ref and int(ref[6].split(":")[1]) > 100
    snap = open(tmp_path / "Results_5.000000" / "snapshot_6.xyz").read().split("\n")
    with gzip.open(os.path.join(GOLD, "5nm_device", "snapshot_6.xyz.gz"), "rt") as f:
        gsnap = f.read().split("\n")
    assert [l.split()[0] for l in snap[2:37652]] == [l.split()[0] for l in gsnap[2:37652]]
    pm = np.array([float(l.split()[4]) for l in snap[2:37652]]); pg = np.array([float(l.split()[4]) for l in gsnap[2:37652]])
    assert np.abs(pm - pg).max() < 5e-4


# ---------------------------------------------------------------- edge cases
def test_edge_empty_ranges_and_bad_arguments(kmc, ctx, s_small):
    s = s_small
    x, y, z, el = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z), ctx.dev_i(s.element)
    # empty row ranges are no-ops
    assert ctx.compute_neighbor_list(x, y, z, 3.5, 52, 10, 0).shape == (0, 52)
    ch = ctx.empty_i(s.N, 7)
    nb = ctx.compute_neighbor_list(x, y, z)
    ctx.update_charge(el, ch, nb[:0], s.metals, row_start=5, row_count=0)
    assert (to_np(ch) == 7).all()
    pot = ctx.empty_d(s.N, 3.0)
    ctx.poisson_gridless(x, y, z, el, ch, s.sigma, s.k, pot, row_start=0, row_count=0)
    assert (to_np(pot) == 3.0).all()
    # invalid arguments come back as errors with a message, never as a crash
    with pytest.raises(kmc.KMCB200Error, match="row range"):
        ctx.compute_neighbor_list(x, y, z, 3.5, 52, s.N - 1, 5)
    with pytest.raises(kmc.KMCB200Error, match="nn"):
        ctx.compute_neighbor_list(x, y, z, 3.5, 65)
    with pytest.raises(kmc.KMCB200Error):
        ctx.initialize_sparsity_K(x, y, z, s.lattice, 0, 3.5, s.N, s.N)       # no interior rows
    K = ctx.initialize_sparsity_K(x, y, z, s.lattice, 0, 3.5, s.N_left, s.N_right)
    with pytest.raises(kmc.KMCB200Error, match="do not match"):
        ctx.assemble_K(K, s.N - 1, s.N_left, s.N_right, el, ch, s.metals, 1.0, 1.0, 1e-8)
    K.close()


def test_edge_no_possible_events_and_zero_bias(kmc, ctx, orc):
    """a device without vacancies / ions / interstitial pairs has Psum = 0: the loop draws its two numbers, executes
    nothing and returns an infinite residence time; Vd = 0 gives b = 0 and the PCG's 0/0 test stops at 0 iterations."""
    s = make_synthetic(kmc, seed=2)
    s.element = np.where(np.isin(s.element, [kmc.VACANCY, kmc.OXYGEN_DEFECT, kmc.DEFECT]), kmc.Hf_EL, s.element).astype(np.int32)
    s.Vd = 0.0
    dev = kmc.DeviceKMC(s, ctx=ctx)
    et, ne = dev.superstep()
    assert ne == 1 and np.isinf(et) and et > 0
    log, _ = dev.ev.log()
    assert len(log) in (0, 1)
    assert dev.last_cg_iterations == 0
    assert (to_np(dev.pot_charge) == 0).all()
    mt, pos = dev.ev.rng_get_state()
    r = orc.Rng(1); r.next(); r.next()
    mt2, pos2 = r.state()
    assert pos == pos2 and (mt == mt2).all()      # exactly two draws were consumed


def test_edge_ragged_sizes_pbc_and_small_nn(kmc, ctx, orc):
    """row count not a multiple of 256, periodic K, nn < 52: full supersteps against the oracle"""
    s = make_synthetic(kmc, nx=9, ny=5, nz=7, seed=4, pbc=1, vac=0.12)
    assert (s.N - 2 * s.N_left) % 256 != 0
    dev = kmc.DeviceKMC(s, ctx=ctx)
    sim = orc.OracleSim(s)
    K = dev.K.to_host()
    assert (K["col"] == sim.sp["col"]).all() and (K["row_ptr"] == sim.sp["row_ptr"]).all()
    for _ in range(4):
        et, ne = dev.superstep(max_events=25)
        # oracle superstep has no event cap: emulate with the staged calls
        sim.charge = orc.update_charge(sim.element, sim.charge, sim.neigh, s.metals)
        data, dinv, rhs = orc.assemble_K(s.N, s.N_left, s.N_right, sim.element, sim.charge, s.metals, sim.sp, s.Vd, s.high_G, s.low_G)
        n = s.N - s.N_left - s.N_right
        xo, _, it, _ = orc.pcg_jacobi(sim.sp["row_ptr"], sim.sp["col"], data, dinv, rhs, sim.pot_boundary[s.N_left:s.N_left + n], 1e-14 * n)
        sim.pot_boundary[s.N_left:s.N_left + n] = xo
        pot = orc.coulomb(s.x, s.y, s.z, sim.element, sim.charge, s.sigma, s.k) + sim.pot_boundary
        typ, prob = orc.build_events(sim.neigh, s.layer, s.T_bg, s.freq, s.sigma, s.k, s.x, s.y, s.z, pot, sim.element, sim.charge, s.E)
        r = orc.event_loop(sim.neigh, typ, prob, sim.element, sim.charge, s.freq, sim.rng, max_events=25)
        sim.element, sim.charge = r["element"], r["charge"]
        log, _ = dev.ev.log()
        assert it == dev.last_cg_iterations and ne == r["n_events"] and (log == r["log"]).all()
        assert (to_np(dev.pot_boundary)[s.N_left:s.N_left + n] == xo).all()
        assert (to_np(dev.element) == sim.element).all()


def test_superstep_brick_ordered_standin_matches_oracle(kmc, ctx, orc):
    """The bench workload family (synthetic.crossbar_standin, bandwidth-minimised 'brick' site order) at 1x1 tiles:
    contacts stay first/last, 3 supersteps identical to the oracle (sparsity, PCG iteration count, events, potentials)."""
    syn = importlib.import_module(kmc.__name__ + ".synthetic")
    s = syn.crossbar_standin(os.path.join(GOLD, "5nm_device", "parameters.txt"), 1, 1, order="brick", Vd=5.0, rnd_seed=5)
    s5 = kmc.load_structure(os.path.join(GOLD, "5nm_device", "parameters.txt"), apply_vacancies=False)
    assert s.N == s5.N and s.N_left == s5.N_left
    # same sites, permuted: contacts untouched, interior is a permutation
    assert np.array_equal(s.x[:s.N_left], s5.x[:s.N_left]) and np.array_equal(s.x[-s.N_right:], s5.x[-s.N_right:])
    key = lambda t: np.lexsort((t.z, t.y, t.x))
    assert np.array_equal(s.x[key(s)], s5.x[key(s5)]) and np.array_equal(s.z[key(s)], s5.z[key(s5)])
    dev = kmc.DeviceKMC(s, ctx=ctx)
    sim = orc.OracleSim(s)
    h = dev.K.to_host()
    assert np.array_equal(h["row_ptr"], sim.sp["row_ptr"]) and np.array_equal(h["col"], sim.sp["col"])
    for _ in range(3):
        et, ne = dev.superstep()
        log, psum = dev.ev.log()
        r = sim.superstep()
        assert ne == r["n_events"] and dev.last_cg_iterations == r["cg_iterations"]
        assert (log[:, :3] == r["events"][:, :3]).all()
        assert abs(et - r["event_time"]) <= 1e-12 * r["event_time"]
    assert np.abs(to_np(dev.pot_charge) - sim.pot_total).max() <= 1e-10 * np.abs(sim.pot_total).max()
    assert (to_np(dev.element) == sim.element).all() and (to_np(dev.charge) == sim.charge).all()


def test_event_loop_large_device_path(kmc, ctx, orc, monkeypatch):
    """Devices above ~3.2 M sites keep the chunk sums / stored prefixes in global memory instead of shared memory
    (event_loop_kernel<false>); force that path on a small device and compare the event log with the oracle."""
    monkeypatch.setenv("KMCB200_EV_NO_SMEM", "1")
    s = make_synthetic(kmc, seed=21, vac=0.15)
    dev = kmc.DeviceKMC(s, ctx=ctx)
    sim = orc.OracleSim(s)
    for _ in range(4):
        et, ne = dev.superstep()
        log, psum = dev.ev.log()
        r = sim.superstep()
        assert ne == r["n_events"] and dev.last_cg_iterations == r["cg_iterations"]
        assert (log[:, :3] == r["events"][:, :3]).all()
        assert abs(et - r["event_time"]) <= 1e-12 * r["event_time"]
    assert (to_np(dev.element) == sim.element).all() and (to_np(dev.charge) == sim.charge).all()
