"""ctypes binding of the CPU oracle (oracle/_build/libkmc_oracle.so) and of oracle/_ref/libref_host.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libkmc_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_host.so")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
_vp = C.c_void_p


class OrcParams(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("nn", C.c_int), ("N_left", C.c_int), ("N_right", C.c_int), ("pbc", C.c_int),
        ("num_metals", C.c_int), ("metals", C.c_int * 4),
        ("lattice", C.c_double * 3),
        ("nn_dist", C.c_double), ("sigma", C.c_double), ("k", C.c_double), ("T_bg", C.c_double),
        ("freq", C.c_double), ("high_G", C.c_double), ("low_G", C.c_double), ("cutoff_radius", C.c_double),
        ("Vd", C.c_double),
        ("E_gen", C.c_double * 5), ("E_rec", C.c_double * 5), ("E_Vdiff", C.c_double * 5), ("E_Odiff", C.c_double * 5),
        ("cg_tol_per_row", C.c_double), ("cg_max_it", C.c_int), ("spmv_lanes", C.c_int),
    ]


class OrcStepInfo(C.Structure):
    _fields_ = [("cg_iterations", C.c_int), ("n_events", C.c_int), ("event_time", C.c_double),
                ("t_charge", C.c_double), ("t_boundary", C.c_double), ("t_coulomb", C.c_double),
                ("t_events", C.c_double)]


def build(force: bool = False):
    """compile the oracle (and oracle/_ref when /root/reference exists) with oracle/Makefile"""
    srcs = [os.path.join(HERE, f) for f in ("kmc_oracle.cpp", "kirchhoff_oracle.cpp", "kmc_oracle.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-s", "-C", HERE, os.path.join(HERE, "_build", "libkmc_oracle.so")], check=True)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_LIB)):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_dot.restype = C.c_double
        L.orc_rng_create.restype = _vp
        L.orc_rng_next.restype = C.c_double
        L.orc_block_sparsity.restype = C.c_long
        L.orc_select_event.restype = C.c_long
        _lib = L
    return _lib


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_vp)


def neighbor_list(x, y, z, nn_dist=3.5, nn=52, row_start=0, row_count=None, use_cells=False):
    x, y, z = _d(x), _d(y), _d(z)
    N = len(x)
    row_count = N - row_start if row_count is None else row_count
    out = np.empty((row_count, nn), dtype=np.int32)
    lib().orc_neighbor_list(C.c_int(N), _p(x), _p(y), _p(z), C.c_double(nn_dist), C.c_int(nn), C.c_int(row_start),
                            C.c_int(row_count), C.c_int(int(use_cells)), _p(out))
    return out


def cutoff_count(element, x, y, z, cutoff=20.0, row_start=0, row_count=None):
    element, x, y, z = _i(element), _d(x), _d(y), _d(z)
    N = len(x)
    row_count = N - row_start if row_count is None else row_count
    out = np.empty(row_count, dtype=np.int32)
    lib().orc_cutoff_count(C.c_int(N), _p(element), _p(x), _p(y), _p(z), C.c_double(cutoff), C.c_int(row_start),
                           C.c_int(row_count), _p(out))
    return out


def cutoff_list(element, x, y, z, max_num_cutoff, cutoff=20.0, row_start=0, row_count=None):
    element, x, y, z = _i(element), _d(x), _d(y), _d(z)
    N = len(x)
    row_count = N - row_start if row_count is None else row_count
    out = np.empty((row_count, max_num_cutoff), dtype=np.int32)
    lib().orc_cutoff_list(C.c_int(N), _p(element), _p(x), _p(y), _p(z), C.c_double(cutoff), C.c_int(max_num_cutoff),
                          C.c_int(row_start), C.c_int(row_count), _p(out))
    return out


def block_sparsity(x, y, z, lattice, pbc, cutoff, size_i, size_j, start_i, start_j, use_cells=False):
    x, y, z, lattice = _d(x), _d(y), _d(z), _d(lattice)
    rp = np.zeros(size_i + 1, dtype=np.int32)
    args = [_p(x), _p(y), _p(z), _p(lattice), C.c_int(int(pbc)), C.c_double(cutoff), C.c_int(size_i), C.c_int(size_j),
            C.c_int(start_i), C.c_int(start_j), C.c_int(int(use_cells)), C.c_int(len(x))]
    nnz = lib().orc_block_sparsity(*args, _p(rp), None)
    col = np.zeros(max(nnz, 1), dtype=np.int32)
    lib().orc_block_sparsity(*args, _p(rp), _p(col))
    return rp, col[:nnz]


def sparsity_K(x, y, z, lattice, pbc, nn_dist, N_left, N_right, use_cells=False):
    """global (1-rank) K sparsity + contact blocks as initialize_sparsity_K builds them"""
    N = len(x)
    n = N - N_left - N_right
    rp, col = block_sparsity(x, y, z, lattice, pbc, nn_dist, n, n, N_left, N_left, use_cells)
    lrp, lcol = block_sparsity(x, y, z, lattice, pbc, nn_dist, n, N_left, N_left, 0, use_cells)
    rrp, rcol = block_sparsity(x, y, z, lattice, pbc, nn_dist, n, N_right, N_left, N_left + n, use_cells)
    return dict(row_ptr=rp, col=col, left_row_ptr=lrp, left_col=lcol, right_row_ptr=rrp, right_col=rcol)


def update_charge(element, charge, neigh, metals, row_start=0, row_end=None):
    element, neigh, metals = _i(element), _i(neigh), _i(metals)
    charge = _i(charge).copy()
    N = len(element)
    row_end = N if row_end is None else row_end
    lib().orc_update_charge(C.c_int(N), C.c_int(neigh.shape[-1]), _p(element), _p(charge), _p(neigh), _p(metals),
                            C.c_int(len(metals)), C.c_int(row_start), C.c_int(row_end))
    return charge


def assemble_K(N, N_left, N_right, element, charge, metals, sp, Vd, high_G, low_G):
    element, charge, metals = _i(element), _i(charge), _i(metals)
    n = N - N_left - N_right
    nnz = int(sp["row_ptr"][n])
    data = np.zeros(nnz); inv_diag = np.zeros(n); rhs = np.zeros(n)
    lib().orc_assemble_K(C.c_int(N), C.c_int(N_left), C.c_int(N_right), _p(element), _p(charge), _p(metals),
                         C.c_int(len(metals)), _p(sp["row_ptr"]), _p(sp["col"]), _p(sp["left_row_ptr"]),
                         _p(sp["left_col"]), _p(sp["right_row_ptr"]), _p(sp["right_col"]), C.c_double(Vd),
                         C.c_double(high_G), C.c_double(low_G), _p(data), _p(inv_diag), _p(rhs))
    return data, inv_diag, rhs


def dot(u, v):
    u, v = _d(u), _d(v)
    return lib().orc_dot(_p(u), _p(v), C.c_long(len(u)))


def spmv(row_ptr, col, data, x, lanes=8):
    row_ptr, col, data, x = _i(row_ptr), _i(col), _d(data), _d(x)
    n = len(row_ptr) - 1
    y = np.zeros(n)
    lib().orc_spmv(C.c_long(n), _p(row_ptr), _p(col), _p(data), _p(x), _p(y), C.c_int(lanes))
    return y


def pcg_jacobi(row_ptr, col, data, inv_diag, rhs, x0, tol, max_it=10000, lanes=8):
    row_ptr, col, data, inv_diag = _i(row_ptr), _i(col), _d(data), _d(inv_diag)
    r = _d(rhs).copy(); x = _d(x0).copy()
    stats = np.zeros(2)
    n = len(row_ptr) - 1
    it = lib().orc_pcg_jacobi(C.c_long(n), _p(row_ptr), _p(col), _p(data), _p(inv_diag), _p(r), _p(x),
                              C.c_double(tol), C.c_int(max_it), C.c_int(lanes), _p(stats))
    return x, r, it, stats


def coulomb(x, y, z, element, charge, sigma, k, cutoff=20.0, row_start=0, row_count=None, use_cells=False):
    x, y, z, element, charge = _d(x), _d(y), _d(z), _i(element), _i(charge)
    N = len(x)
    row_count = N - row_start if row_count is None else row_count
    pot = np.zeros(N)
    (lib().orc_coulomb_cells if use_cells else lib().orc_coulomb)(C.c_int(N), _p(x), _p(y), _p(z), _p(element), _p(charge), C.c_double(sigma), C.c_double(k),
                      C.c_double(cutoff), C.c_int(row_start), C.c_int(row_count), _p(pot))
    return pot


def build_events(neigh, layer, T_bg, freq, sigma, k, x, y, z, pot, element, charge, E, row_start=0, row_count=None):
    neigh, layer, element, charge = _i(neigh), _i(layer), _i(element), _i(charge)
    x, y, z, pot = _d(x), _d(y), _d(z), _d(pot)
    N = len(x); nn = neigh.shape[-1]
    row_count = N - row_start if row_count is None else row_count
    typ = np.zeros(row_count * nn, dtype=np.int32)
    prob = np.zeros(row_count * nn)
    Es = [_d(E[k_]) for k_ in ("E_gen", "E_rec", "E_Vdiff", "E_Odiff")]
    lib().orc_build_events(C.c_int(N), C.c_int(nn), _p(neigh), _p(layer), C.c_double(T_bg), C.c_double(freq),
                           C.c_double(sigma), C.c_double(k), _p(x), _p(y), _p(z), _p(pot), _p(element), _p(charge),
                           _p(Es[0]), _p(Es[1]), _p(Es[2]), _p(Es[3]), C.c_int(row_start), C.c_int(row_count),
                           _p(typ), _p(prob))
    return typ, prob


class Rng:
    def __init__(self, seed):
        self.h = _vp(lib().orc_rng_create(C.c_uint(seed)))

    def next(self):
        return lib().orc_rng_next(self.h)

    def state(self):
        mt = np.zeros(624, dtype=np.uint32)
        pos = C.c_int(0)
        lib().orc_rng_get_state(self.h, _p(mt), C.byref(pos))
        return mt, pos.value

    def __del__(self):
        try:
            lib().orc_rng_destroy(self.h)
        except Exception:
            pass


def event_loop(neigh, typ, prob, element, charge, freq, rng: Rng, max_events=0, max_log=4096):
    neigh = _i(neigh)
    N, nn = neigh.shape
    typ = _i(typ).copy(); prob = _d(prob).copy(); element = _i(element).copy(); charge = _i(charge).copy()
    log = np.zeros((max_log, 4), dtype=np.int32)
    psum = np.zeros(max_log)
    et = C.c_double(0)
    n = lib().orc_event_loop(C.c_int(N), C.c_int(nn), _p(neigh), _p(typ), _p(prob), _p(element), _p(charge),
                             C.c_double(freq), rng.h, C.c_int(max_events), C.c_int(max_log), _p(log), _p(psum),
                             C.byref(et))
    return dict(n_events=n, event_time=et.value, log=log[:min(n, max_log)], psum=psum[:min(n, max_log)],
                element=element, charge=charge, prob=prob, type=typ)


def block_scan_256(v):
    v = _d(v)
    out = np.zeros(256)
    lib().orc_block_scan_256(_p(v), _p(out))
    return out


def select_event(prob, N, nn, number):
    prob = _d(prob)
    ps = C.c_double(0)
    s = lib().orc_select_event(C.c_int(N), C.c_int(nn), _p(prob), C.c_double(number), C.byref(ps))
    return s, ps.value


def make_params(s, nn=52, cutoff=20.0, lanes=8) -> OrcParams:
    """OrcParams from a product-side Structure-like object (duck typed)"""
    p = OrcParams()
    p.N, p.nn, p.N_left, p.N_right, p.pbc = s.N, nn, s.N_left, s.N_right, int(s.pbc)
    p.num_metals = len(s.metals)
    for i, m in enumerate(s.metals):
        p.metals[i] = m
    for i in range(3):
        p.lattice[i] = s.lattice[i]
    p.nn_dist, p.sigma, p.k, p.T_bg, p.freq = s.nn_dist, s.sigma, s.k, s.T_bg, s.freq
    p.high_G, p.low_G, p.cutoff_radius, p.Vd = s.high_G, s.low_G, cutoff, s.Vd
    for i in range(5):
        p.E_gen[i], p.E_rec[i], p.E_Vdiff[i], p.E_Odiff[i] = (s.E["E_gen"][i], s.E["E_rec"][i], s.E["E_Vdiff"][i],
                                                             s.E["E_Odiff"][i])
    p.cg_tol_per_row, p.cg_max_it, p.spmv_lanes = 1e-14, 10000, lanes
    return p


class OracleSim:
    """1-rank KMC simulation driven entirely by the oracle (the reference's main loop order)."""

    def __init__(self, s, use_cells=False, nn=52):
        self.s = s
        self.p = make_params(s, nn=nn)
        self.x, self.y, self.z = _d(s.x), _d(s.y), _d(s.z)
        self.layer = _i(s.layer)
        self.element = _i(s.element).copy()
        self.charge = np.zeros(s.N, dtype=np.int32)
        self.pot_boundary = np.zeros(s.N)
        self.pot_total = np.zeros(s.N)
        self.neigh = neighbor_list(self.x, self.y, self.z, 3.5, nn, use_cells=use_cells)
        self.sp = sparsity_K(self.x, self.y, self.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, use_cells)
        self.rng = Rng(1)
        self.kmc_time = 0.0
        self.step_count = 0

    def superstep(self, max_log=4096):
        info = OrcStepInfo()
        log = np.zeros((max_log, 4), dtype=np.int32)
        sp = self.sp
        lib().orc_superstep(C.byref(self.p), _p(self.x), _p(self.y), _p(self.z), _p(self.layer), _p(self.neigh),
                            _p(sp["row_ptr"]), _p(sp["col"]), _p(sp["left_row_ptr"]), _p(sp["left_col"]),
                            _p(sp["right_row_ptr"]), _p(sp["right_col"]), _p(self.element), _p(self.charge),
                            _p(self.pot_boundary), _p(self.pot_total), self.rng.h, C.c_int(max_log), _p(log),
                            C.byref(info))
        self.kmc_time += info.event_time
        self.step_count += 1
        return dict(step=self.step_count, event_time=info.event_time, kmc_time=self.kmc_time,
                    n_events=info.n_events, cg_iterations=info.cg_iterations, events=log[:info.n_events].copy(),
                    t=(info.t_charge, info.t_boundary, info.t_coulomb, info.t_events))


# ---- reference host code (oracle/_ref) -----------------------------------------------------------
def ref_lib():
    build()
    if not os.path.exists(REF_LIB):
        return None
    L = C.CDLL(REF_LIB)
    L.ref_rng_create.restype = _vp
    L.ref_rng_next.restype = C.c_double
    L.ref_site_dist.restype = C.c_double
    L.ref_v_solve.restype = C.c_double
    if hasattr(L, "ref_poisson_gridless_rows"):
        L.ref_poisson_gridless_rows.restype = C.c_double
    return L


def ref_poisson_gridless_rows(x, y, z, charge, lattice, pbc, sigma, k, row_begin, row_end):
    """The reference's CPU charge sum (Device::poisson_gridless, src/potential_solver.cpp:74-94) for rows
    [row_begin, row_end), evaluated with the reference's own compiled site_dist / v_solve (oracle/_ref).
    All-to-all, PBC-aware, no cutoff: a timing baseline, not the live algorithm.  Returns None without oracle/_ref."""
    L = ref_lib()
    if L is None or not hasattr(L, "ref_poisson_gridless_rows"):
        return None
    x, y, z, charge = _d(x), _d(y), _d(z), _i(charge)
    lat = _d(np.asarray(lattice, dtype=np.float64))
    out = np.zeros(max(row_end - row_begin, 0))
    L.ref_poisson_gridless_rows(C.c_int(len(x)), _p(x), _p(y), _p(z), _p(charge), _p(lat), C.c_int(int(pbc)),
                                C.c_double(sigma), C.c_double(k), C.c_int(row_begin), C.c_int(row_end), _p(out))
    return out


# ---- Kirchhoff / current chain (oracle/kirchhoff_oracle.cpp) ---------------------------------------
M_0 = 9.11e-31   # src/input_parser.h (electron rest mass used by the reference)


def atoms_compact(element):
    element = _i(element)
    out = np.zeros(len(element), dtype=np.int32)
    n = lib().orc_atoms_compact(C.c_int(len(element)), _p(element), _p(out))
    return out[:n].copy()


def T_sparsity(ax, ay, az, nn_dist, num_source_inj, num_ground_ext, row_start=0, row_count=None):
    ax, ay, az = _d(ax), _d(ay), _d(az)
    N_atom = len(ax)
    row_count = N_atom + 1 - row_start if row_count is None else row_count
    L = lib(); L.orc_T_sparsity.restype = C.c_long
    rp = np.zeros(row_count + 1, dtype=np.int32)
    args = [C.c_int(N_atom), _p(ax), _p(ay), _p(az), C.c_double(nn_dist), C.c_int(num_source_inj), C.c_int(num_ground_ext),
            C.c_int(row_start), C.c_int(row_count)]
    nnz = L.orc_T_sparsity(*args, _p(rp), None)
    col = np.zeros(max(nnz, 1), dtype=np.int32)
    L.orc_T_sparsity(*args, _p(rp), _p(col))
    return rp, col[:nnz]


def T_values(ax, ay, az, a_element, a_charge, metals, nn_dist, high_G, low_G, loop_G, num_source_inj, num_ground_ext,
             row_ptr, col, row_start=0):
    ax, ay, az, a_element, a_charge, metals = _d(ax), _d(ay), _d(az), _i(a_element), _i(a_charge), _i(metals)
    rows = len(row_ptr) - 1
    data = np.zeros(max(len(col), 1)); diag = np.zeros(rows)
    lib().orc_T_values(C.c_int(len(ax)), _p(ax), _p(ay), _p(az), _p(a_element), _p(a_charge), _p(metals),
                       C.c_int(len(metals)), C.c_double(nn_dist), C.c_double(high_G), C.c_double(low_G), C.c_double(loop_G),
                       C.c_int(num_source_inj), C.c_int(num_ground_ext), C.c_int(row_start), C.c_int(rows),
                       _p(_i(row_ptr)), _p(_i(col)), _p(data), _p(diag))
    return data[:len(col)], diag


def tunnel_points(a_element, ax):
    a_element, ax = _i(a_element), _d(ax)
    out = np.zeros(len(ax), dtype=np.int32)
    n = lib().orc_tunnel_points(C.c_int(len(ax)), _p(a_element), _p(ax), _p(out))
    if n < 0:
        raise ValueError("atom 0 qualifies as a tunnel point (the reference cannot list it)")
    return out[:n].copy()


def tunnel_block(ax, ay, az, a_element, a_cb, metals, nn_dist, num_layers_contact, num_source_inj, num_ground_ext, m_e, V0,
                 tunnel_atoms, t_start=0, t_count=None):
    ax, ay, az, a_element, a_cb, metals, tunnel_atoms = _d(ax), _d(ay), _d(az), _i(a_element), _d(a_cb), _i(metals), _i(tunnel_atoms)
    nt = len(tunnel_atoms)
    t_count = nt - t_start if t_count is None else t_count
    L = lib(); L.orc_tunnel_block.restype = C.c_long
    rp = np.zeros(t_count + 1, dtype=np.int32)
    args = [C.c_int(len(ax)), _p(ax), _p(ay), _p(az), _p(a_element), _p(a_cb), _p(metals), C.c_int(len(metals)),
            C.c_double(nn_dist), C.c_int(num_layers_contact), C.c_int(num_source_inj), C.c_int(num_ground_ext),
            C.c_double(m_e), C.c_double(V0), C.c_int(nt), _p(tunnel_atoms), C.c_int(t_start), C.c_int(t_count)]
    nnz = L.orc_tunnel_block(*args, _p(rp), None, None, None)
    col = np.zeros(max(nnz, 1), dtype=np.int32); data = np.zeros(max(nnz, 1)); diag = np.zeros(max(t_count, 1))
    L.orc_tunnel_block(*args, _p(rp), _p(col), _p(data), _p(diag))
    return rp, col[:nnz], data[:nnz], diag[:t_count]


def split_spmv(row_ptr, col, data, t_row_ptr, t_col, t_data, tunnel_rows, x, lanes=8, t_lanes=32):
    n = len(row_ptr) - 1
    y = np.zeros(n)
    lib().orc_split_spmv(C.c_int(n), _p(_i(row_ptr)), _p(_i(col)), _p(_d(data)), C.c_int(len(tunnel_rows)), _p(_i(t_row_ptr)),
                         _p(_i(t_col)), _p(_d(t_data)), _p(_i(tunnel_rows)), _p(_d(x)), _p(y), C.c_int(lanes), C.c_int(t_lanes))
    return y


def pcg_jacobi_split_sparse(row_ptr, col, data, t_row_ptr, t_col, t_data, tunnel_rows, inv_diag, rhs, x0, tol, max_it=100,
                            lanes=8, t_lanes=32):
    n = len(row_ptr) - 1
    r = _d(rhs).copy(); x = _d(x0).copy(); stats = np.zeros(2)
    it = lib().orc_pcg_jacobi_split_sparse(C.c_int(n), _p(_i(row_ptr)), _p(_i(col)), _p(_d(data)), C.c_int(len(tunnel_rows)),
                                           _p(_i(t_row_ptr)), _p(_i(t_col)), _p(_d(t_data)), _p(_i(tunnel_rows)),
                                           _p(_d(inv_diag)), _p(r), _p(x), C.c_double(tol), C.c_int(max_it), C.c_int(lanes),
                                           C.c_int(t_lanes), _p(stats))
    return x, r, it, stats


def imacro(row_ptr, col, data, virtual_potentials, G0):
    L = lib(); L.orc_imacro.restype = C.c_double
    return L.orc_imacro(_p(_i(row_ptr)), _p(_i(col)), _p(_d(data)), _p(_d(virtual_potentials)), C.c_double(G0))


def update_CB_edge(s, sp, site_cb0=None, max_it=50000, high_G=None, low_G=None, Vd=None):
    """update_CB_edge_gpu_sparse on structure s with K sparsity sp; returns (site_CB_edge [J], CG iterations)"""
    cb = np.zeros(s.N) if site_cb0 is None else _d(site_cb0).copy()
    el, metals = _i(s.element), _i(s.metals)
    it = lib().orc_update_CB_edge(C.c_int(s.N), C.c_int(s.N_left), C.c_int(s.N_right), _p(el), _p(metals), C.c_int(len(metals)),
                                  _p(sp["row_ptr"]), _p(sp["col"]), _p(sp["left_row_ptr"]), _p(sp["left_col"]),
                                  _p(sp["right_row_ptr"]), _p(sp["right_col"]), C.c_double(s.Vd if Vd is None else Vd),
                                  C.c_double(s.high_G if high_G is None else high_G),
                                  C.c_double(s.low_G if low_G is None else low_G), _p(cb), C.c_int(max_it), C.c_int(8))
    return cb, it


class KirchhoffOracle:
    """The reference's sparse_dist current solver on one rank, driven by the oracle:
    setLaplacePotential (src/kmc_main.cpp:272) -> initialize_sparsity_T (:273) -> update_power_gpu_sparse_dist (:467)
    with the constants of src/kmc_main.cpp:294-301."""

    def __init__(self, s, sp, num_layers_contact, m_r=0.85, V0=1.6, site_charge=None, site_cb=None, cb_max_it=50000):
        self.s = s
        self.loop_G, self.high_G, self.low_G = s.high_G * 10000000, s.high_G * 100000, s.low_G   # kmc_main.cpp:294-296
        self.G0 = 2 * 3.8612e-5 * 1e-5                                                              # :297-298
        self.nsi = self.nge = s.N_left                                                              # :300-301
        self.nlc, self.m_e, self.V0 = num_layers_contact, m_r * M_0, V0
        self.site_cb, self.cb_iterations = (update_CB_edge(s, sp, max_it=cb_max_it) if site_cb is None else (site_cb, 0))
        self.atom_ind = atoms_compact(s.element)
        a = self.atom_ind
        self.ax, self.ay, self.az = _d(s.x[a]), _d(s.y[a]), _d(s.z[a])
        self.N_atom = len(a)
        self.row_ptr, self.col = T_sparsity(self.ax, self.ay, self.az, s.nn_dist, self.nsi, self.nge)
        self.x = np.zeros(self.N_atom + 1)      # gpubuf.atom_virtual_potentials (warm start of the next solve)
        self.assemble(s.element, np.zeros(s.N, np.int32) if site_charge is None else site_charge)

    def assemble(self, site_element, site_charge):
        s, a = self.s, self.atom_ind
        self.a_el, self.a_ch, self.a_cb = _i(np.asarray(site_element)[a]), _i(np.asarray(site_charge)[a]), _d(self.site_cb[a])
        self.data, self.diag = T_values(self.ax, self.ay, self.az, self.a_el, self.a_ch, s.metals, s.nn_dist, self.high_G,
                                        self.low_G, self.loop_G, self.nsi, self.nge, self.row_ptr, self.col)
        self.tunnel_atoms = tunnel_points(self.a_el, self.ax)
        metals2 = list(s.metals)[:2]                                        # num_metals = 2 hard-coded (init...T.cu:800)
        self.t_row_ptr, self.t_col, self.t_data, self.t_diag = tunnel_block(
            self.ax, self.ay, self.az, self.a_el, self.a_cb, metals2, s.nn_dist, self.nlc, self.nsi, self.nge, self.m_e,
            self.V0, self.tunnel_atoms)
        self.tunnel_rows = (self.tunnel_atoms + 2).astype(np.int32)
        d = self.diag.copy()
        d[self.tunnel_rows] += self.t_diag                                   # assemble_preconditioner
        self.inv_diag = 1.0 / d                                              # invert_diag
        self.rhs = np.zeros(self.N_atom + 1)
        self.rhs[0], self.rhs[1] = -self.loop_G * s.Vd, self.loop_G * s.Vd   # current_solver_gpu.cu:1626-1631

    def solve(self, max_it=100):
        tol = 1e-30 * self.N_atom                                            # current_solver_gpu.cu:1455
        self.x, r, it, stats = pcg_jacobi_split_sparse(self.row_ptr, self.col, self.data, self.t_row_ptr, self.t_col,
                                                       self.t_data, self.tunnel_rows, self.inv_diag, self.rhs, self.x, tol,
                                                       max_it)
        self.imacro = imacro(self.row_ptr, self.col, self.data, self.x, self.G0)
        return it
