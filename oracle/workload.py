"""Workload construction for the CPU arm WITHOUT the product package (TEST / BENCH INFRASTRUCTURE).

bench.py --impl reference and the oracle-side parity checks build their inputs here, so that nothing of
libkmc_b200.so is loaded on that path.  Everything restates the reference's host side:

  * parameters: the values the REFERENCE's own parser (src/input_parser.cpp, compiled into oracle/_ref) produced for
    the shipped files, committed as tests/golden/ref_parser_*.json by tests/golden/make_golden.py;
  * xyz reader: src/utils.cpp:72-98 (line 1 = N, line 2 ignored, then "El x y z"), element strings src/utils.cpp:7-29;
  * Device::makeSubstoichiometric: src/Device.cpp:180-211 (oracle: orc_make_substoichiometric);
  * KMCProcess layer assignment: src/KMCProcess.cpp:34-50 with the table of src/structure_input.h:10-50;
  * the lateral tilings of SURVEY.md section 8(d) (the 40 nm xyz files are not shipped with the reference).

tests/test_workload.py checks that these structures are identical to the product's own (api.load_structure /
synthetic.crossbar_standin), array by array.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ELEMENTS = {"d": 0, "Od": 1, "V": 2, "O": 3, "Hf": 4, "Ni": 5, "Ti": 6, "Pt": 7, "N": 8}   # src/utils.h:37-44
PITCH = 51.15   # lateral pitch of the shipped 5 nm cell [A] (24 x 2.13125)

# src/structure_input.h:10-50
LAYERS = {
    "E_gen":   np.array([0.0, 3.93, 3.93, 1.66, 1.73]),
    "E_rec":   np.array([0.0, 0.0, 0.0, 0.0, 0.0]),
    "E_Vdiff": np.array([0.0, 1.09, 1.09, 1.09, 0.0]),
    "E_Odiff": np.array([0.76, 0.76, 0.76, 0.76, 2.8]),
    "start_x": np.array([-22.0, 0.0, 3.0, 48.1431, 52.6431]),
    "end_x":   np.array([0.0, 3.0, 48.1431, 52.6431, 90.0]),
}


@dataclass
class Workload:
    """duck-typed like the product's api.Structure (what oracle.binding.make_params / OracleSim read)"""
    element: np.ndarray
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    lattice: Tuple[float, float, float]
    pbc: int
    nn_dist: float
    N_left: int
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


def read_xyz(path):
    """src/utils.cpp:72-98"""
    with open(path) as f:
        n = int(f.readline().split()[0])
        f.readline()
        el = np.empty(n, dtype=np.int32)
        xyz = np.empty((n, 3))
        for i in range(n):
            t = f.readline().split()
            el[i] = ELEMENTS[t[0]]
            xyz[i] = (float(t[1]), float(t[2]), float(t[3]))
    return el, np.ascontiguousarray(xyz[:, 0]), np.ascontiguousarray(xyz[:, 1]), np.ascontiguousarray(xyz[:, 2])


def assign_layers(x):
    """src/KMCProcess.cpp:34-50: the LAST layer l with start_x <= x <= end_x"""
    layer = np.full(len(x), 1000, dtype=np.int32)
    for l in range(5):
        layer[(LAYERS["start_x"][l] <= x) & (x <= LAYERS["end_x"][l])] = l
    if (layer == 1000).any():
        raise ValueError("a site is not inside the device (KMCProcess.cpp:45-49)")
    return layer


def make_substoichiometric(element, concentration, seed):
    assert element.dtype == np.int32 and element.flags.c_contiguous
    return orc.lib().orc_make_substoichiometric(C.c_int(len(element)), element.ctypes.data_as(C.c_void_p),
                                                C.c_double(concentration), C.c_uint(seed))


def load_5nm(apply_vacancies=True) -> Workload:
    """the reference's shipped structures/5nm_device, initialised as its main() does (src/kmc_main.cpp:117-155)"""
    p = json.load(open(os.path.join(GOLD, "ref_parser_5nm.json")))
    el, x, y, z = read_xyz(os.path.join(GOLD, "5nm_device", p["restart_xyz_file"]))
    if p["pristine"] and apply_vacancies:
        make_substoichiometric(el, p["initial_vacancy_concentration"], p["rnd_seed"])
    w = Workload(element=el, x=x, y=y, z=z, lattice=tuple(p["lattice"]), pbc=p["pbc"], nn_dist=p["nn_dist"],
                 N_left=p["num_atoms_first_layer"], N_right=p["num_atoms_first_layer"],
                 metals=p["metals"][:p["num_metals"]], sigma=p["sigma"], k=p["k"], T_bg=p["background_temp"],
                 freq=p["freq"], high_G=p["high_G"], low_G=p["low_G"], Vd=p["V_switch0"], t_switch=p["t_switch0"])
    w.layer = assign_layers(w.x)
    w.E = dict(LAYERS)
    return w


def brick_permutation(x, y, z, n_left, n_right, B=12.5):
    """bandwidth-minimising 3-D blocking of the interior sites: cubes of edge B, cubes ordered y-major, (x, y, z)
    lexicographic inside a cube; contacts keep their places (src/potential_solver_gpu.cu:855-861)"""
    n = len(x)
    interior = np.arange(n_left, n - n_right)
    bx, by, bz = (np.floor(x[interior] / B), np.floor(y[interior] / B), np.floor(z[interior] / B))
    perm = interior[np.lexsort((z[interior], y[interior], x[interior], bx, bz, by))]
    return np.concatenate([np.arange(n_left), perm, np.arange(n - n_right, n)])


def crossbar_standin(ty=8, tz=8, order="file", Vd=15.0, vacancy_concentration=0.05, rnd_seed=32) -> Workload:
    """ty x tz lateral tiling of the PRISTINE 5 nm cell (site-major images), optional re-ordering of the interior
    sites, then makeSubstoichiometric on the tiled device (stand-in for structures/40nm_crossbar)."""
    base = load_5nm(apply_vacancies=False)
    nimg = ty * tz
    oy = np.repeat(np.arange(ty), tz) * PITCH
    oz = np.tile(np.arange(tz), ty) * PITCH
    x = np.repeat(base.x, nimg)
    y = (base.y[:, None] + oy[None, :]).ravel()
    z = (base.z[:, None] + oz[None, :]).ravel()
    el = np.repeat(base.element, nimg).astype(np.int32)
    NL, NR = base.N_left * nimg, base.N_right * nimg
    n = len(x)
    interior = np.arange(NL, n - NR)
    full = None
    if order == "xsorted":
        full = np.concatenate([np.arange(NL), interior[np.argsort(x[interior], kind="stable")], np.arange(n - NR, n)])
    elif order == "lex":
        full = np.concatenate([np.arange(NL), interior[np.lexsort((z[interior], y[interior], x[interior]))],
                               np.arange(n - NR, n)])
    elif order.startswith("brick"):
        full = brick_permutation(x, y, z, NL, NR, float(order[5:]) if len(order) > 5 else 12.5)
    elif order != "file":
        raise ValueError(order)
    if full is not None:
        x, y, z, el = x[full], y[full], z[full], el[full]
    el = np.ascontiguousarray(el)
    if vacancy_concentration > 0:
        make_substoichiometric(el, vacancy_concentration, rnd_seed)
    w = Workload(element=el, x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), z=np.ascontiguousarray(z),
                 lattice=(base.lattice[0], PITCH * ty, PITCH * tz), pbc=base.pbc, nn_dist=base.nn_dist, N_left=NL,
                 N_right=NR, metals=list(base.metals), sigma=base.sigma, k=base.k, T_bg=base.T_bg, freq=base.freq,
                 high_G=base.high_G, low_G=base.low_G, Vd=Vd, t_switch=base.t_switch)
    w.layer = assign_layers(w.x)
    w.E = dict(LAYERS)
    return w


def build(name: str):
    """the named bench workloads (same names and parameters as bench.py's GPU arm)"""
    if name == "5nm":
        return load_5nm(), "structures/5nm_device (shipped), N=37650"
    base, _, order = name.partition("_")
    order = order or "file"
    if base == "highvac7x7":
        w = crossbar_standin(7, 7, order=order, Vd=5.0, rnd_seed=5, vacancy_concentration=0.25)
        return w, (f"synthetic high-vacancy lattice: 7x7 lateral tiling of the shipped 5nm cell, N={w.N}, 25 % oxygen "
                   f"vacancies, Vd=5, site order '{order}'")
    import re
    m = re.fullmatch(r"standin(\d+)x(\d+)", base)
    if not m or m.group(1) != m.group(2):
        raise ValueError(f"unknown workload '{name}'")
    t = int(m.group(1))
    w = crossbar_standin(t, t, order=order, Vd=15.0, rnd_seed=32)
    note = {"file": "site order 'file' (tile images site-major: the 5nm file's block structure, wide K bandwidth)",
            "brick": "site order 'brick' (bandwidth-minimised: interior sites grouped in 12.5 A cubes, contacts "
                     "first/last -- the layout the reference's crossbar_40_bwmin.xyz input is named for)"}
    return w, (f"40nm_crossbar stand-in: {t}x{t} lateral tiling of the shipped 5nm cell, N={w.N}, "
               f"num_atoms_first_layer={w.N_left}, Vd=15, " + note.get(order, f"site order '{order}'"))
