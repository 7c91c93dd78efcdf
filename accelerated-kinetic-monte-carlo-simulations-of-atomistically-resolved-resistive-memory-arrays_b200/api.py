"""ctypes binding of libkmc_b200.so + a Python mirror of the reference's host flow.

The product path is the CUDA library: importing this module loads
``libkmc_b200.so`` (built in-tree by ``build.sh``) and FAILS LOUDLY if it is missing.
There is no CPU fallback and nothing here imports ``oracle/``.

Torch is used for plumbing only (device buffers, the current CUDA stream, torch.distributed).

Reference host flow mirrored by :class:`DeviceKMC`: ``src/kmc_main.cpp:117-239`` (setup) and
``:328-540`` (one KMC superstep).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmc_b200.so")

# element / event enums (reference src/utils.h:37-60)
DEFECT, OXYGEN_DEFECT, VACANCY, O_EL, Hf_EL, Ni_EL, Ti_EL, Pt_EL, N_EL, NULL_ELEMENT = range(10)
VACANCY_GENERATION, VACANCY_RECOMBINATION, VACANCY_DIFFUSION, ION_DIFFUSION, NULL_EVENT = range(5)
ELEMENT_NAMES = ["d", "Od", "V", "O", "Hf", "Ni", "Ti", "Pt", "N"]

CHUNK = 256
SPMV_LANES = 8
MAX_NUM_NEIGHBORS = 52      # reference src/Device.cpp:59, src/neighbor_lists_gpu.cu:265
NEIGHBOR_NN_DIST = 3.5      # reference src/neighbor_lists_gpu.cu:266 (hard-coded, overrides parameters.txt)
CUTOFF_RADIUS = 20.0        # reference src/neighbor_lists_gpu.cu:262
RND_SEED_KMC = 1            # reference src/structure_input.h:8


class KMCB200Error(RuntimeError):
    pass


class Params(C.Structure):
    """POD mirror of kmcb200_params (subset of the reference's KMCParameters, src/input_parser.h)."""
    _fields_ = [
        ("rnd_seed", C.c_uint),
        ("restart", C.c_int), ("pristine", C.c_int), ("shift", C.c_int), ("pbc", C.c_int),
        ("solve_potential", C.c_int), ("solve_current", C.c_int), ("solve_heating_global", C.c_int),
        ("solve_heating_local", C.c_int), ("perturb_structure", C.c_int),
        ("log_freq", C.c_int), ("output_freq", C.c_int),
        ("num_atoms_first_layer", C.c_int), ("num_layers_contact", C.c_int), ("num_atoms_contact", C.c_int),
        ("num_atoms_reservoir", C.c_int),
        ("num_metals", C.c_int), ("metals", C.c_int * 8),
        ("n_V_switch", C.c_int), ("n_t_switch", C.c_int), ("n_lattice", C.c_int), ("n_shifts", C.c_int),
        ("V_switch0", C.c_double), ("t_switch0", C.c_double),
        ("lattice", C.c_double * 3), ("shifts", C.c_double * 3),
        ("initial_vacancy_concentration", C.c_double), ("freq", C.c_double), ("nn_dist", C.c_double),
        ("sigma", C.c_double), ("epsilon", C.c_double), ("k", C.c_double), ("high_G", C.c_double),
        ("low_G", C.c_double),
        ("background_temp", C.c_double), ("m_r", C.c_double), ("V0", C.c_double), ("Icc", C.c_double),
        ("Rs", C.c_double), ("t_ox", C.c_double), ("A", C.c_double),
        ("restart_xyz_file", C.c_char * 512), ("atom_xyz_file", C.c_char * 512),
        ("interstitial_xyz_file", C.c_char * 512),
    ]


_vp = C.c_void_p
_i = C.c_int
_d = C.c_double
_ll = C.c_longlong
_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)
_pll = C.POINTER(C.c_longlong)
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/kmc_b200.h declares
SIGNATURES = {
    "kmcb200_last_error": (C.c_char_p, []),
    "kmcb200_version": (_i, []),
    "kmcb200_launch_count": (_ll, []),
    "kmcb200_create": (_i, [_pvp, _i, _vp]),
    "kmcb200_destroy": (_i, [_vp]),
    "kmcb200_set_stream": (_i, [_vp, _vp]),
    "kmcb200_synchronize": (_i, [_vp]),
    "kmcb200_device_info": (_i, [_vp, _pi, _pi, _pi, C.POINTER(C.c_size_t)]),
    "kmcb200_fp64_peak": (_i, [_vp, _pd]),
    "kmcb200_malloc": (_i, [_vp, _pvp, C.c_size_t]),
    "kmcb200_free": (_i, [_vp, _vp]),
    "kmcb200_memcpy_h2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "kmcb200_memcpy_d2h": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "kmcb200_memcpy_d2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "kmcb200_memset": (_i, [_vp, _vp, _i, C.c_size_t]),
    "kmcb200_host_alloc_pinned": (_i, [_pvp, C.c_size_t]),
    "kmcb200_host_free_pinned": (_i, [_vp]),
    "kmcb200_compute_neighbor_list": (_i, [_vp, _i, _vp, _vp, _vp, _d, _i, _i, _i, _vp]),
    "kmcb200_cutoff_size": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _i, _i, _vp, _pi]),
    "kmcb200_cutoff_list": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _i, _i, _i, _vp]),
    "kmcb200_initialize_sparsity_K": (_i, [_vp, _i, _vp, _vp, _vp, _pd, _i, _d, _i, _i, _i, _i, _pvp]),
    "kmcb200_kmat_from_csr": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _pvp]),
    "kmcb200_kmat_destroy": (_i, [_vp]),
    "kmcb200_kmat_info": (_i, [_vp, _pi, _pll, _pll, _pll]),
    "kmcb200_kmat_pointers": (_i, [_vp] + [_pvp] * 9),
    "kmcb200_kmat_block_view": (_i, [_vp, _i, _i, _vp, _vp, _pll]),
    "kmcb200_update_charge": (_i, [_vp, _vp, _vp, _vp, _i, _i, _pi, _i, _i, _i]),
    "kmcb200_assemble_K": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _pi, _i, _d, _d, _d]),
    "kmcb200_pcg_jacobi": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _pi]),
    "kmcb200_spmv": (_i, [_vp, _vp, _vp, _vp]),
    "kmcb200_spmv_dot": (_i, [_vp, _vp, _vp, _vp, _pd]),
    "kmcb200_dot": (_i, [_vp, _vp, _vp, _ll, _pd]),
    "kmcb200_background_potential": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _pi, _i, _d, _d, _d, _vp, _pi]),
    "kmcb200_sparsity_K_row_counts": (_i, [_vp, _i, _vp, _vp, _vp, _pd, _i, _d, _i, _i, _vp]),
    "kmcb200_comm_create": (_i, [_vp, _i, _i, _i, _pi, _pi, _pvp]),
    "kmcb200_comm_create_ex": (_i, [_vp, _i, _i, _i, _pi, _pi, _ll, _pvp]),
    "kmcb200_comm_destroy": (_i, [_vp]),
    "kmcb200_comm_allgather": (_i, [_vp, _vp, _pi, _pi]),
    "kmcb200_rdv_open": (_i, [C.c_char_p, _i, _i, _pvp]),
    "kmcb200_rdv_allgather": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "kmcb200_rdv_barrier": (_i, [_vp]),
    "kmcb200_rdv_close": (_i, [_vp]),
    "kmcb200_comm_ipc_handle": (_i, [_vp, _vp]),
    "kmcb200_comm_open_peers": (_i, [_vp, _vp]),
    "kmcb200_kmat_attach_comm": (_i, [_vp, _vp]),
    "kmcb200_kmat_need_map": (_i, [_vp, _vp]),
    "kmcb200_comm_set_send_masks": (_i, [_vp, _vp]),
    "kmcb200_comm_info": (_i, [_vp, _pi, _pi, C.POINTER(C.c_uint), _pll]),
    "kmcb200_poisson_gridless": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _i, _i, _vp]),
    "kmcb200_poisson_stats": (_i, [_vp, _pll, _pll, _pll]),
    "kmcb200_cutoff_d2max": (C.c_double, [C.c_double]),
    "kmcb200_sum_potential": (_i, [_vp, _i, _vp, _vp]),
    "kmcb200_events_create": (_i, [_vp, _i, _i, _vp, _pvp]),
    "kmcb200_events_destroy": (_i, [_vp]),
    "kmcb200_set_activation_energies": (_i, [_vp, _i, _pd, _pd, _pd, _pd]),
    "kmcb200_rng_seed": (_i, [_vp, C.c_uint]),
    "kmcb200_rng_set_state": (_i, [_vp, C.POINTER(C.c_uint), _i]),
    "kmcb200_rng_get_state": (_i, [_vp, C.POINTER(C.c_uint), _pi]),
    "kmcb200_rng_draw": (_i, [_vp, _i, _pd]),
    "kmcb200_execute_kmc_step": (_i, [_vp, _vp, _i, _i, _vp, _vp, _d, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _i,
                                      _pd, _pi]),
    "kmcb200_build_event_list": (_i, [_vp, _vp, _i, _i, _vp, _vp, _d, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kmcb200_events_pointers": (_i, [_vp, _pvp, _pvp]),
    "kmcb200_events_log": (_i, [_vp, _i, _pi, _pd, _pi]),
    "kmcb200_update_CB_edge": (_i, [_vp, _vp, _i, _i, _i, _vp, _pi, _i, _d, _d, _d, _vp, _i, _pi]),
    "kmcb200_initialize_sparsity_T": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _i, _i, _i, _pvp]),
    "kmcb200_tmat_destroy": (_i, [_vp]),
    "kmcb200_tmat_info": (_i, [_vp, _pi, _pll, _pi, _pll]),
    "kmcb200_tmat_pointers": (_i, [_vp] + [_pvp] * 11),
    "kmcb200_assemble_T": (_i, [_vp, _vp, _vp, _vp, _vp, _pi, _i, _d, _d, _d, _d, _d, _d]),
    "kmcb200_tmat_spmv": (_i, [_vp, _vp, _vp, _vp]),
    "kmcb200_pcg_jacobi_split_sparse": (_i, [_vp, _vp, _vp, _vp, _d, _i, _pi]),
    "kmcb200_imacro": (_i, [_vp, _vp, _vp, _d, _pd]),
    "kmcb200_update_power_sparse": (_i, [_vp, _vp, _vp, _vp, _vp, _pi, _i, _d, _d, _d, _d, _d, _d, _d, _vp, _pd, _pi]),
    "kmcb200_update_temperature_global": (_i, [_vp, _vp, _vp, _i, _d, _d, _d, _d, _d]),
    "kmcb200_parse_parameters": (_i, [C.c_char_p, C.POINTER(Params)]),
    "kmcb200_parse_parameter_vector": (_i, [C.c_char_p, _i, _i, _pd]),
    "kmcb200_xyz_count": (_i, [C.c_char_p]),
    "kmcb200_read_xyz": (_i, [C.c_char_p, _i, _pi, _pd, _pd, _pd]),
    "kmcb200_make_substoichiometric": (_i, [_i, _pi, _d, C.c_uint]),
    "kmcb200_num_layers": (_i, []),
    "kmcb200_layer_table": (_i, [_pd] * 6),
    "kmcb200_assign_layers": (_i, [_i, _pd, _pi]),
    "kmcb200_partition": (None, [_i, _i, _pi, _pi]),
    "kmcb200_partition_aligned": (None, [_i, _i, _pi, _pi]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libkmc_b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KMCB200Error(
            f"{LIB_PATH} is missing: build it with build.sh (or __graft_entry__.build()); "
            "this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def launch_count() -> int:
    """kernels launched by libkmc_b200 in this process so far"""
    return int(load_library().kmcb200_launch_count())


def _check(rc: int):
    if rc != 0:
        msg = load_library().kmcb200_last_error()
        raise KMCB200Error(f"libkmc_b200 error {rc}: {msg.decode() if msg else ''}")


def _np_i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _np_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------------
# host model (no GPU needed)
# ------------------------------------------------------------------------------------------------
def parse_parameters(path: str) -> Params:
    lib = load_library()
    p = Params()
    _check(lib.kmcb200_parse_parameters(path.encode(), C.byref(p)))
    return p


def parse_parameter_vector(path: str, which: int) -> np.ndarray:
    lib = load_library()
    n = lib.kmcb200_parse_parameter_vector(path.encode(), which, 0, None)
    if n < 0:
        _check(n)
    out = np.zeros(max(n, 1), dtype=np.float64)
    lib.kmcb200_parse_parameter_vector(path.encode(), which, n, out.ctypes.data_as(_pd))
    return out[:n]


def read_xyz(path: str):
    lib = load_library()
    n = lib.kmcb200_xyz_count(path.encode())
    if n < 0:
        _check(n)
    el = np.zeros(n, dtype=np.int32)
    x = np.zeros(n); y = np.zeros(n); z = np.zeros(n)
    got = lib.kmcb200_read_xyz(path.encode(), n, el.ctypes.data_as(_pi), x.ctypes.data_as(_pd),
                               y.ctypes.data_as(_pd), z.ctypes.data_as(_pd))
    if got < 0:
        _check(got)
    return el, x, y, z


def make_substoichiometric(element: np.ndarray, concentration: float, seed: int) -> int:
    lib = load_library()
    assert element.dtype == np.int32 and element.flags.c_contiguous
    r = lib.kmcb200_make_substoichiometric(len(element), element.ctypes.data_as(_pi), concentration, seed)
    if r < 0:
        _check(r)
    return r


def layer_table():
    lib = load_library()
    arrs = [np.zeros(5) for _ in range(6)]
    lib.kmcb200_layer_table(*[a.ctypes.data_as(_pd) for a in arrs])
    return dict(zip(["E_gen", "E_rec", "E_Vdiff", "E_Odiff", "start_x", "end_x"], arrs))


def assign_layers(x: np.ndarray) -> np.ndarray:
    lib = load_library()
    x = _np_d(x)
    out = np.zeros(len(x), dtype=np.int32)
    _check(lib.kmcb200_assign_layers(len(x), x.ctypes.data_as(_pd), out.ctypes.data_as(_pi)))
    return out


def partition(nrows: int, nranks: int, aligned: bool = False):
    lib = load_library()
    counts = np.zeros(nranks, dtype=np.int32)
    displs = np.zeros(nranks, dtype=np.int32)
    fn = lib.kmcb200_partition_aligned if aligned else lib.kmcb200_partition
    fn(nrows, nranks, counts.ctypes.data_as(_pi), displs.ctypes.data_as(_pi))
    return counts, displs


# ------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------
class Context:
    """kmcb200_ctx bound to a torch CUDA device and torch's current stream."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        import torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise KMCB200Error("no CUDA device: libkmc_b200 has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream else None
        h = C.c_void_p()
        _check(self.lib.kmcb200_create(C.byref(h), device, C.c_void_p(stream) if stream is not None else None))
        self.h = h

    def close(self):
        if self.h:
            self.lib.kmcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.lib.kmcb200_synchronize(self.h))

    def fp64_peak_tflops(self) -> float:
        out = C.c_double(0)
        _check(self.lib.kmcb200_fp64_peak(self.h, C.byref(out)))
        return out.value

    # -- tensor helpers
    def dev_i(self, a):
        return self.torch.as_tensor(_np_i(a), device=self.device)

    def dev_d(self, a):
        return self.torch.as_tensor(_np_d(a), device=self.device)

    def empty_i(self, n, fill=None):
        t = self.torch.empty(int(n), dtype=self.torch.int32, device=self.device)
        if fill is not None:
            t.fill_(fill)
        return t

    def empty_d(self, n, fill=None):
        t = self.torch.empty(int(n), dtype=self.torch.float64, device=self.device)
        if fill is not None:
            t.fill_(fill)
        return t

    # -- a1
    def compute_neighbor_list(self, x, y, z, nn_dist=NEIGHBOR_NN_DIST, nn=MAX_NUM_NEIGHBORS, row_start=0,
                              row_count=None):
        N = x.numel()
        row_count = N - row_start if row_count is None else row_count
        out = self.empty_i(row_count * nn)
        _check(self.lib.kmcb200_compute_neighbor_list(self.h, N, _ptr(x), _ptr(y), _ptr(z), nn_dist, nn, row_start,
                                                      row_count, _ptr(out)))
        return out.view(row_count, nn)

    # -- a2
    def cutoff_size(self, element, x, y, z, cutoff=CUTOFF_RADIUS, row_start=0, row_count=None, want_counts=False):
        N = x.numel()
        row_count = N - row_start if row_count is None else row_count
        counts = self.empty_i(row_count, 0) if want_counts else None
        mx = C.c_int(0)
        _check(self.lib.kmcb200_cutoff_size(self.h, N, _ptr(element), _ptr(x), _ptr(y), _ptr(z), cutoff, row_start,
                                            row_count, _ptr(counts), C.byref(mx)))
        return mx.value, counts

    def cutoff_list(self, element, x, y, z, max_num_cutoff, cutoff=CUTOFF_RADIUS, row_start=0, row_count=None):
        N = x.numel()
        row_count = N - row_start if row_count is None else row_count
        out = self.empty_i(row_count * max_num_cutoff)
        _check(self.lib.kmcb200_cutoff_list(self.h, N, _ptr(element), _ptr(x), _ptr(y), _ptr(z), cutoff,
                                            max_num_cutoff, row_start, row_count, _ptr(out)))
        return out.view(row_count, max_num_cutoff)

    # -- a3
    def initialize_sparsity_K(self, x, y, z, lattice, pbc, nn_dist, N_left, N_right, row_start=0, row_count=None):
        N = x.numel()
        n_int = N - N_left - N_right
        row_count = n_int - row_start if row_count is None else row_count
        lat = (C.c_double * 3)(*[float(v) for v in lattice])
        h = C.c_void_p()
        _check(self.lib.kmcb200_initialize_sparsity_K(self.h, N, _ptr(x), _ptr(y), _ptr(z), lat, int(pbc), nn_dist,
                                                      N_left, N_right, row_start, row_count, C.byref(h)))
        return KMatrix(self, h)

    def sparsity_K_row_counts(self, x, y, z, lattice, pbc, nn_dist, N_left, N_right):
        N = x.numel()
        out = self.empty_i(N - N_left - N_right, 0)
        lat = (C.c_double * 3)(*[float(v) for v in lattice])
        _check(self.lib.kmcb200_sparsity_K_row_counts(self.h, N, _ptr(x), _ptr(y), _ptr(z), lat, int(pbc), nn_dist, N_left,
                                                      N_right, _ptr(out)))
        return out

    def kmat_from_csr(self, row_ptr, col, val, cols_global=None, row_start=0):
        rows = row_ptr.numel() - 1
        cols_global = rows if cols_global is None else cols_global
        h = C.c_void_p()
        _check(self.lib.kmcb200_kmat_from_csr(self.h, rows, cols_global, row_start, _ptr(row_ptr), _ptr(col),
                                              _ptr(val), C.byref(h)))
        K = KMatrix(self, h)
        K._keep = (row_ptr, col, val)
        return K

    # -- a5
    def update_charge(self, element, charge, neigh, metals, row_start=0, row_count=None):
        N = element.numel()
        nn = neigh.shape[-1]
        row_count = N - row_start if row_count is None else row_count
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        _check(self.lib.kmcb200_update_charge(self.h, _ptr(element), _ptr(charge), _ptr(neigh), N, nn, m, len(metals),
                                              row_start, row_count))

    # -- a6 / a7
    def assemble_K(self, K, N, N_left, N_right, element, charge, metals, Vd, high_G, low_G):
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        _check(self.lib.kmcb200_assemble_K(self.h, K.h, N, N_left, N_right, _ptr(element), _ptr(charge), m,
                                           len(metals), Vd, high_G, low_G))

    def pcg_jacobi(self, K, r, x, dinv, tol, max_it) -> int:
        it = C.c_int(0)
        _check(self.lib.kmcb200_pcg_jacobi(self.h, K.h, _ptr(r), _ptr(x), _ptr(dinv), tol, max_it, C.byref(it)))
        return it.value

    def spmv(self, K, x, y):
        _check(self.lib.kmcb200_spmv(self.h, K.h, _ptr(x), _ptr(y)))

    def spmv_dot(self, K, x, y, want_scalar=True):
        out = C.c_double(0)
        _check(self.lib.kmcb200_spmv_dot(self.h, K.h, _ptr(x), _ptr(y), C.byref(out) if want_scalar else None))
        return out.value if want_scalar else None

    def dot(self, u, v) -> float:
        out = C.c_double(0)
        _check(self.lib.kmcb200_dot(self.h, _ptr(u), _ptr(v), u.numel(), C.byref(out)))
        return out.value

    def background_potential(self, K, N, N_left, N_right, element, charge, metals, Vd, high_G, low_G,
                             pot_boundary) -> int:
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        it = C.c_int(0)
        _check(self.lib.kmcb200_background_potential(self.h, K.h, N, N_left, N_right, _ptr(element), _ptr(charge), m,
                                                     len(metals), Vd, high_G, low_G, _ptr(pot_boundary),
                                                     C.byref(it)))
        return it.value

    # -- a8 / a9
    def poisson_gridless(self, x, y, z, element, charge, sigma, k, pot_charge, cutoff=CUTOFF_RADIUS, row_start=0,
                         row_count=None):
        N = x.numel()
        row_count = N - row_start if row_count is None else row_count
        _check(self.lib.kmcb200_poisson_gridless(self.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(element), _ptr(charge),
                                                 sigma, k, cutoff, row_start, row_count, _ptr(pot_charge)))

    def poisson_stats(self):
        """(charged sources, distance tests, pairs inside the cutoff) of the last poisson_gridless call"""
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        _check(self.lib.kmcb200_poisson_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sum_potential(self, pot_charge, pot_boundary):
        _check(self.lib.kmcb200_sum_potential(self.h, pot_charge.numel(), _ptr(pot_charge), _ptr(pot_boundary)))

    # -- a12: Kirchhoff / current chain
    def update_CB_edge(self, K, N, N_left, N_right, element, metals, Vd, high_G, low_G, site_cb, max_it=50000) -> int:
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        it = C.c_int(0)
        _check(self.lib.kmcb200_update_CB_edge(self.h, K.h, N, N_left, N_right, _ptr(element), m, len(metals), Vd, high_G,
                                               low_G, _ptr(site_cb), max_it, C.byref(it)))
        return it.value

    def initialize_sparsity_T(self, element, x, y, z, nn_dist, num_source_inj, num_ground_ext, num_layers_contact):
        h = C.c_void_p()
        _check(self.lib.kmcb200_initialize_sparsity_T(self.h, element.numel(), _ptr(element), _ptr(x), _ptr(y), _ptr(z),
                                                      nn_dist, num_source_inj, num_ground_ext, num_layers_contact,
                                                      C.byref(h)))
        return TMatrix(self, h)

    def assemble_T(self, T, element, charge, site_cb, metals, Vd, high_G, low_G, loop_G, m_e, V0):
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        _check(self.lib.kmcb200_assemble_T(self.h, T.h, _ptr(element), _ptr(charge), _ptr(site_cb), m, len(metals), Vd,
                                           high_G, low_G, loop_G, m_e, V0))
        T.refresh()

    def tmat_spmv(self, T, x, y):
        _check(self.lib.kmcb200_tmat_spmv(self.h, T.h, _ptr(x), _ptr(y)))

    def pcg_jacobi_split_sparse(self, T, r, x, tol, max_it=100) -> int:
        it = C.c_int(0)
        _check(self.lib.kmcb200_pcg_jacobi_split_sparse(self.h, T.h, _ptr(r), _ptr(x), tol, max_it, C.byref(it)))
        return it.value

    def imacro(self, T, V, G0) -> float:
        out = C.c_double(0)
        _check(self.lib.kmcb200_imacro(self.h, T.h, _ptr(V), G0, C.byref(out)))
        return out.value

    def update_power_sparse(self, T, element, charge, site_cb, metals, Vd, high_G, low_G, loop_G, G0, m_e, V0, V):
        """-> (I_macro, PCG iterations); V = atom_virtual_potentials (N_atom + 1), warm start in / solution out"""
        m = (C.c_int * max(len(metals), 1))(*[int(v) for v in metals])
        im, it = C.c_double(0), C.c_int(0)
        _check(self.lib.kmcb200_update_power_sparse(self.h, T.h, _ptr(element), _ptr(charge), _ptr(site_cb), m, len(metals),
                                                    Vd, high_G, low_G, loop_G, G0, m_e, V0, _ptr(V), C.byref(im),
                                                    C.byref(it)))
        T.refresh()
        return im.value, it.value

    # -- f-4
    def update_temperature_global(self, site_power, T_bg, a_coeff, b_coeff, number_steps, C_thermal, small_step):
        _check(self.lib.kmcb200_update_temperature_global(self.h, _ptr(site_power), _ptr(T_bg), site_power.numel(), a_coeff,
                                                          b_coeff, number_steps, C_thermal, small_step))

    # -- a10
    def events_create(self, neigh):
        N, nn = neigh.shape
        h = C.c_void_p()
        _check(self.lib.kmcb200_events_create(self.h, N, nn, _ptr(neigh), C.byref(h)))
        return Events(self, h, N, nn)


M_0 = 9.11e-31   # electron rest mass the reference uses (src/input_parser.h:99)


class TMatrix:
    """kmcb200_tmat handle: the Kirchhoff system (reference GPUBuffers::T_distributed + tunnel sub-block + atom arrays)."""

    def __init__(self, ctx: "Context", h):
        self.ctx, self.h = ctx, h
        self.refresh()

    def refresh(self):
        a, b, c, d = C.c_int(0), C.c_longlong(0), C.c_int(0), C.c_longlong(0)
        _check(self.ctx.lib.kmcb200_tmat_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.N_atom, self.nnz, self.n_tunnel, self.tunnel_nnz = a.value, b.value, c.value, d.value

    def _d2h(self, ptr, n, dtype):
        out = np.zeros(max(int(n), 1), dtype=dtype)
        if n > 0:
            _check(self.ctx.lib.kmcb200_memcpy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), int(n) * out.itemsize))
            self.ctx.sync()
        return out[:int(n)]

    def to_host(self):
        self.refresh()
        ps = [C.c_void_p() for _ in range(11)]
        _check(self.ctx.lib.kmcb200_tmat_pointers(self.h, *[C.byref(p) for p in ps]))
        v = [p.value for p in ps]
        n, nt = self.N_atom + 1, self.n_tunnel
        out = {"atom_ind": self._d2h(v[0], self.N_atom, np.int32), "row_ptr": self._d2h(v[1], n + 1, np.int32),
               "col": self._d2h(v[2], self.nnz, np.int32), "val": self._d2h(v[3], self.nnz, np.float64),
               "inv_diag": self._d2h(v[4], n, np.float64), "rhs": self._d2h(v[5], n, np.float64),
               "tunnel_atoms": self._d2h(v[6], nt, np.int32), "t_row_ptr": self._d2h(v[7], nt + 1 if nt else 0, np.int32)}
        if nt:
            out.update({"t_col": self._d2h(v[8], self.tunnel_nnz, np.int32), "t_val": self._d2h(v[9], self.tunnel_nnz, np.float64),
                        "t_diag": self._d2h(v[10], nt, np.float64)})
        return out

    def close(self):
        if self.h:
            self.ctx.lib.kmcb200_tmat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KMatrix:
    """kmcb200_kmat handle (reference: GPUBuffers::K_distributed + contact CSR blocks)."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        rows, nnz, lnnz, rnnz = C.c_int(0), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        _check(ctx.lib.kmcb200_kmat_info(h, C.byref(rows), C.byref(nnz), C.byref(lnnz), C.byref(rnnz)))
        self.rows, self.nnz, self.left_nnz, self.right_nnz = rows.value, nnz.value, lnnz.value, rnnz.value

    def _raw_pointers(self):
        ps = [C.c_void_p() for _ in range(9)]
        _check(self.ctx.lib.kmcb200_kmat_pointers(self.h, *[C.byref(p) for p in ps]))
        return [p.value for p in ps]

    def _copy_out(self, ptr, n, dtype):
        """D2H copy of n elements at raw device pointer ptr"""
        out = np.zeros(max(int(n), 1), dtype=dtype)
        if n > 0:
            _check(self.ctx.lib.kmcb200_memcpy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr),
                                                   int(n) * out.itemsize))
            self.ctx.sync()
        return out[:int(n)]

    def to_host(self):
        """dict of numpy arrays: row_ptr, col, val, left_*, right_*, inv_diag, rhs"""
        rp, col, val, lrp, lcol, rrp, rcol, dinv, rhs = self._raw_pointers()
        out = {
            "row_ptr": self._copy_out(rp, self.rows + 1, np.int32),
            "col": self._copy_out(col, self.nnz, np.int32),
            "val": self._copy_out(val, self.nnz, np.float64),
        }
        if lrp:
            out.update({
                "left_row_ptr": self._copy_out(lrp, self.rows + 1, np.int32),
                "left_col": self._copy_out(lcol, self.left_nnz, np.int32),
                "right_row_ptr": self._copy_out(rrp, self.rows + 1, np.int32),
                "right_col": self._copy_out(rcol, self.right_nnz, np.int32),
                "inv_diag": self._copy_out(dinv, self.rows, np.float64),
                "rhs": self._copy_out(rhs, self.rows, np.float64),
            })
        return out

    def device_vectors(self):
        """(inv_diag, rhs) of the last assemble_K as torch tensors (device-to-device copies of the kmat's arrays)"""
        ctx = self.ctx
        ptrs = self._raw_pointers()
        out = []
        for ptr in (ptrs[7], ptrs[8]):
            t = ctx.empty_d(self.rows)
            _check(ctx.lib.kmcb200_memcpy_d2d(ctx.h, _ptr(t), C.c_void_p(ptr), self.rows * 8))
            out.append(t)
        return out

    def block_view(self, col_start, col_count):
        ctx = self.ctx
        rp = ctx.empty_i(self.rows + 1)
        nnz = C.c_longlong(0)
        _check(ctx.lib.kmcb200_kmat_block_view(self.h, col_start, col_count, _ptr(rp), None, C.byref(nnz)))
        col = ctx.empty_i(max(nnz.value, 1))
        _check(ctx.lib.kmcb200_kmat_block_view(self.h, col_start, col_count, _ptr(rp), _ptr(col), C.byref(nnz)))
        ctx.sync()
        return rp.cpu().numpy(), col.cpu().numpy()[:nnz.value]

    def close(self):
        if self.h:
            self.ctx.lib.kmcb200_kmat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Events:
    """kmcb200_events handle: event list + hierarchical sums + device MT19937."""

    def __init__(self, ctx: Context, h, N, nn):
        self.ctx, self.h, self.N, self.nn = ctx, h, N, nn

    def set_activation_energies(self, E_gen, E_rec, E_Vdiff, E_Odiff):
        arrs = [_np_d(a) for a in (E_gen, E_rec, E_Vdiff, E_Odiff)]
        _check(self.ctx.lib.kmcb200_set_activation_energies(self.h, len(arrs[0]),
                                                            *[a.ctypes.data_as(_pd) for a in arrs]))

    def rng_seed(self, seed: int):
        _check(self.ctx.lib.kmcb200_rng_seed(self.h, seed))

    def rng_set_state(self, mt624: np.ndarray, pos: int):
        mt = np.ascontiguousarray(mt624, dtype=np.uint32)
        _check(self.ctx.lib.kmcb200_rng_set_state(self.h, mt.ctypes.data_as(C.POINTER(C.c_uint)), pos))

    def rng_get_state(self):
        mt = np.zeros(624, dtype=np.uint32)
        pos = C.c_int(0)
        _check(self.ctx.lib.kmcb200_rng_get_state(self.h, mt.ctypes.data_as(C.POINTER(C.c_uint)), C.byref(pos)))
        return mt, pos.value

    def rng_draw(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        _check(self.ctx.lib.kmcb200_rng_draw(self.h, n, out.ctypes.data_as(_pd)))
        return out

    def build_event_list(self, neigh, layer, T_bg, freq, sigma, k, x, y, z, pot, element, charge):
        _check(self.ctx.lib.kmcb200_build_event_list(self.ctx.h, self.h, self.N, self.nn, _ptr(neigh), _ptr(layer),
                                                     T_bg, freq, sigma, k, _ptr(x), _ptr(y), _ptr(z), _ptr(pot),
                                                     _ptr(element), _ptr(charge)))

    def event_arrays(self):
        """(event_prob float64[N*nn], event_type uint8[N*nn]) copied to host"""
        pp, pt = C.c_void_p(), C.c_void_p()
        _check(self.ctx.lib.kmcb200_events_pointers(self.h, C.byref(pp), C.byref(pt)))
        n = self.N * self.nn
        prob = np.zeros(n, dtype=np.float64)
        typ = np.zeros(n, dtype=np.uint8)
        lib = self.ctx.lib
        _check(lib.kmcb200_memcpy_d2h(self.ctx.h, prob.ctypes.data_as(C.c_void_p), pp, n * 8))
        _check(lib.kmcb200_memcpy_d2h(self.ctx.h, typ.ctypes.data_as(C.c_void_p), pt, n))
        self.ctx.sync()
        return prob, typ

    def execute_kmc_step(self, neigh, layer, T_bg, freq, sigma, k, x, y, z, pot, element, charge, max_events=0):
        et, ne = C.c_double(0), C.c_int(0)
        _check(self.ctx.lib.kmcb200_execute_kmc_step(self.ctx.h, self.h, self.N, self.nn, _ptr(neigh), _ptr(layer),
                                                     T_bg, freq, sigma, k, _ptr(x), _ptr(y), _ptr(z), _ptr(pot),
                                                     _ptr(element), _ptr(charge), max_events, C.byref(et),
                                                     C.byref(ne)))
        return et.value, ne.value

    def log(self, max_rows=65536):
        log = np.zeros((max_rows, 4), dtype=np.int32)
        psum = np.zeros(max_rows, dtype=np.float64)
        rows = C.c_int(0)
        _check(self.ctx.lib.kmcb200_events_log(self.h, max_rows, log.ctypes.data_as(_pi), psum.ctypes.data_as(_pd),
                                               C.byref(rows)))
        return log[:rows.value].copy(), psum[:rows.value].copy()

    def close(self):
        if self.h:
            self.ctx.lib.kmcb200_events_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Host-side structure container + the reference's main-loop order
# ------------------------------------------------------------------------------------------------
@dataclass
class Structure:
    """What Device + KMCProcess hold on the host (reference src/Device.h, src/KMCProcess.h)."""
    element: np.ndarray
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    lattice: Tuple[float, float, float]
    pbc: int
    nn_dist: float
    N_left: int            # num_atoms_first_layer
    N_right: int
    metals: List[int]
    sigma: float
    k: float
    T_bg: float
    freq: float
    high_G: float
    low_G: float
    Vd: float
    t_switch: float = 1e-12
    layer: Optional[np.ndarray] = None
    E: dict = field(default_factory=dict)

    @property
    def N(self):
        return len(self.element)


def load_structure(param_file: str, base_dir: Optional[str] = None, apply_vacancies: bool = True) -> Structure:
    """Parse parameters.txt + xyz exactly as reference src/kmc_main.cpp:117-155 does (restart or atom+interstitial
    files, makeSubstoichiometric when pristine) and attach the KMCProcess layer data."""
    p = parse_parameters(param_file)
    base = base_dir or os.path.dirname(os.path.abspath(param_file))
    files = [p.restart_xyz_file.decode()] if p.restart else [p.atom_xyz_file.decode(),
                                                              p.interstitial_xyz_file.decode()]
    els, xs, ys, zs = [], [], [], []
    for f in files:
        el, x, y, z = read_xyz(os.path.join(base, f))
        els.append(el); xs.append(x); ys.append(y); zs.append(z)
    el = np.ascontiguousarray(np.concatenate(els)); x = np.concatenate(xs); y = np.concatenate(ys)
    z = np.concatenate(zs)
    if p.pristine and apply_vacancies:
        make_substoichiometric(el, p.initial_vacancy_concentration, p.rnd_seed)
    V = parse_parameter_vector(param_file, 0)
    t = parse_parameter_vector(param_file, 1)
    s = Structure(element=el, x=x, y=y, z=z, lattice=tuple(p.lattice), pbc=p.pbc, nn_dist=p.nn_dist,
                  N_left=p.num_atoms_first_layer, N_right=p.num_atoms_first_layer,
                  metals=[p.metals[i] for i in range(p.num_metals)], sigma=p.sigma, k=p.k,
                  T_bg=p.background_temp, freq=p.freq, high_G=p.high_G, low_G=p.low_G,
                  Vd=float(V[0]) if len(V) else 0.0, t_switch=float(t[0]) if len(t) else 0.0)
    s.layer = assign_layers(s.x)
    s.E = layer_table()
    return s


class DeviceKMC:
    """GPU-resident KMC simulation: GPUBuffers + the setup and superstep call order of the reference main().

    setup  = src/kmc_main.cpp:187-239 (GPUBuffers, compute_neighbor_list, initialize_sparsity_K,
             copytoConstMemory); the 20 A cutoff list of :205 is not materialised (see kmc_b200.h a2/a8).
    superstep = src/kmc_main.cpp:328-540 (update_charge_gpu -> background_potential_gpu_sparse ->
             poisson_gridless_gpu -> sum_and_gather_potential -> execute_kmc_step_mpi).
    """

    def __init__(self, s: Structure, device: int = 0, ctx: Optional[Context] = None):
        self.s = s
        self.ctx = ctx or Context(device)
        c = self.ctx
        self.N = s.N
        self.x, self.y, self.z = c.dev_d(s.x), c.dev_d(s.y), c.dev_d(s.z)
        self.element = c.dev_i(s.element)
        self.charge = c.empty_i(self.N, 0)
        self.layer = c.dev_i(s.layer)
        self.pot_boundary = c.empty_d(self.N, 0.0)
        self.pot_charge = c.empty_d(self.N, 0.0)
        self.neigh = c.compute_neighbor_list(self.x, self.y, self.z)
        self.K = c.initialize_sparsity_K(self.x, self.y, self.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
        self.ev = c.events_create(self.neigh)
        self.ev.set_activation_energies(s.E["E_gen"], s.E["E_rec"], s.E["E_Vdiff"], s.E["E_Odiff"])
        self.ev.rng_seed(RND_SEED_KMC)
        self.kmc_time = 0.0
        self.step_count = 0
        self.last_cg_iterations = 0
        self.last_n_events = 0

    def sync_host_to_gpu(self, element: np.ndarray, charge: np.ndarray):
        """GPUBuffers::sync_HostToGPU (src/gpu_buffers.cpp:10-34) for the mutable site state"""
        self.element.copy_(self.ctx.torch.from_numpy(element), non_blocking=True)
        self.charge.copy_(self.ctx.torch.from_numpy(charge), non_blocking=True)

    def field_solve(self):
        s, c = self.s, self.ctx
        c.update_charge(self.element, self.charge, self.neigh, s.metals)
        self.last_cg_iterations = c.background_potential(self.K, self.N, s.N_left, s.N_right, self.element,
                                                         self.charge, s.metals, s.Vd, s.high_G, s.low_G,
                                                         self.pot_boundary)
        c.poisson_gridless(self.x, self.y, self.z, self.element, self.charge, s.sigma, s.k, self.pot_charge)
        c.sum_potential(self.pot_charge, self.pot_boundary)

    def superstep(self, max_events: int = 0):
        s = self.s
        torch = self.ctx.torch
        if not hasattr(self, "_tev"):
            self._tev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            self.field_ms_total = 0.0
            self.events_ms_total = 0.0
        self._tev[0].record()
        self.field_solve()
        self._tev[1].record()
        et, ne = self.ev.execute_kmc_step(self.neigh, self.layer, s.T_bg, s.freq, s.sigma, s.k, self.x, self.y,
                                          self.z, self.pot_charge, self.element, self.charge, max_events)
        self._tev[2].record()
        self._tev[2].synchronize()   # (execute_kmc_step already synchronised to return the event time)
        self.field_ms_total += self._tev[0].elapsed_time(self._tev[1])
        self.events_ms_total += self._tev[1].elapsed_time(self._tev[2])
        self.kmc_time += et
        self.step_count += 1
        self.last_n_events = ne
        return et, ne

    def snapshot(self):
        """Mutable simulation state (site elements/charges, PCG warm start, generator, clocks): restoring it replays
        the same supersteps bit for bit (the restart the reference gets from its snapshot files)."""
        mt, pos = self.ev.rng_get_state()
        return {"element": self.element.clone(), "charge": self.charge.clone(),
                "pot_boundary": self.pot_boundary.clone(), "pot_charge": self.pot_charge.clone(),
                "mt": mt.copy(), "pos": int(pos), "kmc_time": self.kmc_time, "step_count": self.step_count}

    def restore(self, snap):
        self.element.copy_(snap["element"]); self.charge.copy_(snap["charge"])
        self.pot_boundary.copy_(snap["pot_boundary"]); self.pot_charge.copy_(snap["pot_charge"])
        self.ev.rng_set_state(snap["mt"], snap["pos"])
        self.kmc_time, self.step_count = snap["kmc_time"], snap["step_count"]

    def run(self, t_switch: Optional[float] = None, max_steps: Optional[int] = None):
        """while (kmc_time < t) superstep  (src/kmc_main.cpp:328); returns the per-step records"""
        t = self.s.t_switch if t_switch is None else t_switch
        rec = []
        while self.kmc_time < t and (max_steps is None or self.step_count < max_steps):
            et, ne = self.superstep()
            log, _ = self.ev.log()
            rec.append({"step": self.step_count, "event_time": et, "kmc_time": self.kmc_time, "n_events": ne,
                        "cg_iterations": self.last_cg_iterations, "events": log})
        return rec
