"""Synthetic replicated crossbar lattices (BASELINE.json configs 2-5).

The reference's 40 nm structure files are not shipped (structures/40nm_crossbar/*.xyz are stripped blobs), so the
larger workloads are lateral (y, z) tilings of the shipped 5 nm TiN/HfO2/Ti/TiN cell, which is a wrapped periodic
cell of pitch 51.15 A.  Site order contract kept from the reference: the first / last `N_left` sites are the
contact layers (src/potential_solver_gpu.cu:855-861).

  order="file"    : images emitted site-major -> the 5 nm file's block structure is preserved (wide K bandwidth)
  order="xsorted" : interior sites stably sorted by x (narrow halo for row-sharded solves); contacts stay first/last
  order="lex"     : interior sites sorted by (x, y, z) lexicographically, the scheme of the shipped file's crystalline
                    blocks (narrow halo AND neighbouring rows share neighbours)
  order="brick[B]": interior sites grouped into cubes of edge B angstrom (default 12.5, about one 256-row chunk); the
                    bandwidth-minimised layout the reference's 40 nm input (crossbar_40_bwmin.xyz) is named for
"""
from __future__ import annotations

import os

import numpy as np

from .api import Structure, assign_layers, layer_table, load_structure, make_substoichiometric

PITCH = 51.15  # lateral pitch of the 5 nm cell [A]  (24 x 2.13125)


def brick_permutation(x, y, z, n_left: int, n_right: int, B: float = 12.5) -> np.ndarray:
    """Bandwidth-minimising 3-D blocking of the interior sites (what the reference's crossbar_40_bwmin.xyz is for):
    interior sites are grouped into cubes of edge B angstrom (12.5 A ~ 256 sites, one SpMV / event chunk); cubes are
    ordered y-major (then z, then x) so that contiguous row blocks are slabs across the long lateral axis and a
    row-sharded solve exchanges only thin x-z interfaces; sites inside a cube in (x, y, z) lexicographic order.  The first
    n_left and last n_right sites (the contact layers, src/potential_solver_gpu.cu:855-861) keep their places.
    Returns the permutation `perm` with new_site[k] = old_site[perm[k]]."""
    n = len(x)
    interior = np.arange(n_left, n - n_right)
    bx, by, bz = (np.floor(x[interior] / B), np.floor(y[interior] / B), np.floor(z[interior] / B))
    perm = interior[np.lexsort((z[interior], y[interior], x[interior], bx, bz, by))]
    return np.concatenate([np.arange(n_left), perm, np.arange(n - n_right, n)])


def tile_structure(base: Structure, ty: int, tz: int, order: str = "file", vacancy_concentration: float = 0.05,
                   rnd_seed: int = 32, Vd: float | None = None) -> Structure:
    """ty x tz lateral tiling of a PRISTINE base cell, then Device::makeSubstoichiometric on the tiled device."""
    nimg = ty * tz
    oy = np.repeat(np.arange(ty), tz) * PITCH
    oz = np.tile(np.arange(tz), ty) * PITCH
    # site-major: site s -> images 0..nimg-1
    x = np.repeat(base.x, nimg)
    y = (base.y[:, None] + oy[None, :]).ravel()
    z = (base.z[:, None] + oz[None, :]).ravel()
    el = np.repeat(base.element, nimg).astype(np.int32)
    NL = base.N_left * nimg
    NR = base.N_right * nimg
    if order == "xsorted":
        n = len(x)
        interior = np.arange(NL, n - NR)
        perm = interior[np.argsort(x[interior], kind="stable")]
        full = np.concatenate([np.arange(NL), perm, np.arange(n - NR, n)])
        x, y, z, el = x[full], y[full], z[full], el[full]
    elif order == "lex":
        # (x, y, z) lexicographic order of the interior sites -- how the crystalline parts of the shipped 5 nm file are
        # ordered (z fastest, then y, then x); contacts stay first / last
        n = len(x)
        interior = np.arange(NL, n - NR)
        perm = interior[np.lexsort((z[interior], y[interior], x[interior]))]
        full = np.concatenate([np.arange(NL), perm, np.arange(n - NR, n)])
        x, y, z, el = x[full], y[full], z[full], el[full]
    elif order.startswith("brick"):
        B = float(order[5:]) if len(order) > 5 else 12.5
        full = brick_permutation(x, y, z, NL, NR, B)
        x, y, z, el = x[full], y[full], z[full], el[full]
    elif order != "file":
        raise ValueError(order)
    el = np.ascontiguousarray(el)
    if vacancy_concentration > 0:
        make_substoichiometric(el, vacancy_concentration, rnd_seed)
    s = Structure(element=el, x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), z=np.ascontiguousarray(z),
                  lattice=(base.lattice[0], PITCH * ty, PITCH * tz), pbc=base.pbc, nn_dist=base.nn_dist, N_left=NL,
                  N_right=NR, metals=list(base.metals), sigma=base.sigma, k=base.k, T_bg=base.T_bg, freq=base.freq,
                  high_G=base.high_G, low_G=base.low_G, Vd=base.Vd if Vd is None else Vd, t_switch=base.t_switch)
    s.layer = assign_layers(s.x)
    s.E = layer_table()
    return s


def crossbar_standin(param_5nm: str, ty: int = 8, tz: int = 8, order: str = "file", Vd: float = 15.0,
                     vacancy_concentration: float = 0.05, rnd_seed: int = 32) -> Structure:
    """Stand-in for structures/40nm_crossbar (8x8 tiling: N = 2 409 600, num_atoms_first_layer = 36 864) with that
    file's V_switch = 15, rnd_seed = 32, pbc = 0.  The real crossbar has patterned electrodes (33 600 first-layer
    atoms); this is a full slab."""
    base = load_structure(param_5nm, apply_vacancies=False)
    return tile_structure(base, ty, tz, order=order, vacancy_concentration=vacancy_concentration, rnd_seed=rnd_seed,
                          Vd=Vd)
