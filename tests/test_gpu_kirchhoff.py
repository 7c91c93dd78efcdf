"""GPU parity of the Kirchhoff / current chain against the oracle restatement, through the C ABI.
Everything is bit-exact: atom compaction, T_neighbor CSR + values + diagonal, tunnel points, tunnel CSR, CB-edge solve
(identical iteration count), and -- because the two transcendental calls of the WKB coefficients go through routines
shared operation by operation with the oracle -- the tunnel values, the preconditioner, the split-sparse PCG iterates and
the macroscopic current (north_star: "currents within 1e-10 relative")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def setup5(kmc, ctx, orc, s5):
    dev = kmc.DeviceKMC(s5, ctx=ctx)
    sim = orc.OracleSim(s5, use_cells=True)
    ch = orc.update_charge(s5.element, np.zeros(s5.N, np.int32), sim.neigh, s5.metals)
    ko = orc.KirchhoffOracle(s5, sim.sp, 10, site_charge=ch)
    return dev, sim, ch, ko


def test_cb_edge_bit_identical(kmc, ctx, s5, setup5):
    dev, sim, ch, ko = setup5
    cb = ctx.empty_d(s5.N, 0.0)
    it = ctx.update_CB_edge(dev.K, s5.N, s5.N_left, s5.N_right, dev.element, s5.metals, s5.Vd, s5.high_G, s5.low_G, cb)
    assert it == ko.cb_iterations
    assert (to_np(cb) == ko.site_cb).all()


def _assembled(kmc, ctx, s5, setup5):
    dev, sim, ch, ko = setup5
    T = ctx.initialize_sparsity_T(dev.element, dev.x, dev.y, dev.z, s5.nn_dist, s5.N_left, s5.N_left, 10)
    cb = ctx.dev_d(ko.site_cb)
    ctx.assemble_T(T, dev.element, ctx.dev_i(ch), cb, s5.metals, s5.Vd, ko.high_G, ko.low_G, ko.loop_G, ko.m_e, ko.V0)
    return T, cb


def test_T_assembly_vs_oracle(kmc, ctx, s5, setup5):
    dev, sim, ch, ko = setup5
    T, cb = _assembled(kmc, ctx, s5, setup5)
    h = T.to_host()
    assert T.N_atom == ko.N_atom and (h["atom_ind"] == ko.atom_ind).all()
    assert (h["row_ptr"] == ko.row_ptr).all() and (h["col"] == ko.col).all()            # sparsity bit-exact
    assert (h["val"] == ko.data).all()                                                    # values are sums of +-{G}: exact
    assert (h["tunnel_atoms"] == ko.tunnel_atoms).all()
    assert (h["t_row_ptr"] == ko.t_row_ptr).all() and (h["t_col"] == ko.t_col).all()      # tunnel sparsity bit-exact
    # WKB coefficients: exp / x^1.5 go through the deterministic routines shared with the oracle -> bit-identical
    assert (h["t_val"] == ko.t_data).all() and (h["t_diag"] == ko.t_diag).all()
    assert (h["inv_diag"] == ko.inv_diag).all()
    assert (h["rhs"] == ko.rhs).all()
    # split-sparse SpMV
    xv = np.random.default_rng(2).standard_normal(ko.N_atom + 1)
    y = ctx.empty_d(ko.N_atom + 1, 0.0)
    ctx.tmat_spmv(T, ctx.dev_d(xv), y)
    want = orc_split(ko, xv, setup5)
    assert (to_np(y) == want).all()
    T.close()


def orc_split(ko, xv, setup5):
    from oracle import binding as orc
    return orc.split_spmv(ko.row_ptr, ko.col, ko.data, ko.t_row_ptr, ko.t_col, ko.t_data, ko.tunnel_rows, xv, lanes=8, t_lanes=32)


def test_split_sparse_solve_and_current_vs_oracle(kmc, ctx, orc, s5, setup5):
    dev, sim, ch, ko = setup5
    T = ctx.initialize_sparsity_T(dev.element, dev.x, dev.y, dev.z, s5.nn_dist, s5.N_left, s5.N_left, 10)
    cb = ctx.dev_d(ko.site_cb)
    V = ctx.empty_d(ko.N_atom + 1, 0.0)
    ko.x = np.zeros(ko.N_atom + 1)
    for step in range(2):          # second call: warm start from the first solution
        im, it = ctx.update_power_sparse(T, dev.element, ctx.dev_i(ch), cb, s5.metals, s5.Vd, ko.high_G, ko.low_G, ko.loop_G,
                                         ko.G0, ko.m_e, ko.V0, V)
        it_o = ko.solve()
        assert it == it_o == 100   # the reference's max_iterations (current_solver_gpu.cu:1456)
        got = to_np(V)
        assert (got == ko.x).all()                                          # virtual potentials: bit-identical iterates
        assert abs(im - ko.imacro) <= 1e-10 * abs(ko.imacro)                # macroscopic current (north_star: 1e-10)
        assert im == ko.imacro
    assert im > 0
    T.close()


def test_chain_after_kmc_events(kmc, ctx, orc, s5):
    """the atoms' elements / charges change with the KMC events: assemble on the evolved state (vacancies moved,
    tunnel points moved) must still match the oracle"""
    dev = kmc.DeviceKMC(s5, ctx=ctx)
    sim = orc.OracleSim(s5, use_cells=True)
    for _ in range(3):
        dev.superstep(); sim.superstep()
    assert (to_np(dev.element) == sim.element).all()
    ko = orc.KirchhoffOracle(s5, sim.sp, 10, site_charge=sim.charge)
    ko.assemble(sim.element, sim.charge)
    T = ctx.initialize_sparsity_T(dev.element, dev.x, dev.y, dev.z, s5.nn_dist, s5.N_left, s5.N_left, 10)
    ctx.assemble_T(T, dev.element, dev.charge, ctx.dev_d(ko.site_cb), s5.metals, s5.Vd, ko.high_G, ko.low_G, ko.loop_G, ko.m_e, ko.V0)
    h = T.to_host()
    assert (h["tunnel_atoms"] == ko.tunnel_atoms).all() and (h["t_col"] == ko.t_col).all() and (h["val"] == ko.data).all()
    T.close()


def test_global_temperature_recurrence(ctx, orc):
    """f-4: update_temperatureglobal_gpu (heat_solver_gpu.cu:43-69): exact sum of the site powers (summation spec) and the
    scalar recurrence; several calls chained like a bias sweep with heating"""
    import ctypes as C
    L = orc.lib(); L.orc_update_temperature_global.restype = C.c_double
    rng = np.random.default_rng(3)
    for N in (1, 300, 37650, 100003):
        power = np.abs(rng.standard_normal(N)) * 1e-9
        T_dev = ctx.dev_d(np.array([300.0]))
        T = 300.0
        for step in range(3):
            a, b, steps, Cth, small = 0.99, 3.0 + step, 50.0, 1e-15, 1e-16
            ctx.update_temperature_global(ctx.dev_d(power), T_dev, a, b, steps, Cth, small)
            T = L.orc_update_temperature_global(orc._p(power), C.c_int(N), C.c_double(T), C.c_double(a), C.c_double(b),
                                                C.c_double(steps), C.c_double(Cth), C.c_double(small))
            assert to_np(T_dev)[0] == T
