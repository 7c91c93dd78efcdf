import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kmc():
    """the product package (loads libkmc_b200.so; fails loudly if it was not built)"""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle binding (test infrastructure)"""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def s5(kmc):
    """the reference's shipped 5 nm device, initialised exactly as its main() does"""
    return kmc.load_structure(os.path.join(GOLD, "5nm_device", "parameters.txt"))


def make_synthetic(kmc, nx=10, ny=6, nz=6, a=2.2, seed=3, vac=0.08, pbc=0, jitter=0.25):
    """small TiN | HfO2-like | TiN sandwich on a jittered cubic grid with interstitial 'd' sites;
    contacts first/last like the reference's ordering contract (first/last N_left sites = contact layers)."""
    rng = np.random.default_rng(seed)
    pts = []
    # x layout: layer 0 = left contact plane, then oxide planes, last = right contact plane
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                pts.append((ix, iy, iz))
    pts = np.array(pts, dtype=float)
    ncontact = ny * nz
    xyz = pts * a
    xyz[:, 0] += 0.05  # oxide starts just inside layer 1
    xyz += rng.uniform(-jitter, jitter, xyz.shape) * (pts[:, :1] > 0) * (pts[:, :1] < nx - 1)
    el = np.full(len(pts), kmc.O_EL, dtype=np.int32)
    el[(pts[:, 0] + pts[:, 1] + pts[:, 2]) % 2 == 0] = kmc.Hf_EL
    is_left = pts[:, 0] == 0
    is_right = pts[:, 0] == nx - 1
    el[is_left | is_right] = np.where((pts[is_left | is_right, 1] + pts[is_left | is_right, 2]) % 2 == 0, kmc.Ti_EL, kmc.N_EL)
    # second metal plane on each side (interior metal rows exercise the high_G branch)
    m2 = (pts[:, 0] == 1) | (pts[:, 0] == nx - 2)
    el[m2] = np.where((pts[m2, 1] + pts[m2, 2]) % 2 == 0, kmc.Ti_EL, kmc.N_EL)
    # interstitial defect sites at cell centres of the oxide
    inter = []
    for ix in range(2, nx - 3):
        for iy in range(ny - 1):
            for iz in range(nz - 1):
                inter.append(((ix + 0.5) * a + 0.05, (iy + 0.5) * a, (iz + 0.5) * a))
    inter = np.array(inter)
    left = np.where(is_left)[0]
    right = np.where(is_right)[0]
    mid = np.where(~(is_left | is_right))[0]
    order_xyz = np.concatenate([xyz[left], xyz[mid], inter, xyz[right]])
    order_el = np.concatenate([el[left], el[mid], np.full(len(inter), kmc.DEFECT, dtype=np.int32), el[right]])
    # vacancies + a few oxygen ions
    ox = np.where(order_el == kmc.O_EL)[0]
    nv = max(2, int(vac * len(ox)))
    order_el[rng.choice(ox, nv, replace=False)] = kmc.VACANCY
    dd = np.where(order_el == kmc.DEFECT)[0]
    order_el[rng.choice(dd, max(2, len(dd) // 20), replace=False)] = kmc.OXYGEN_DEFECT
    x = np.ascontiguousarray(order_xyz[:, 0]); y = np.ascontiguousarray(order_xyz[:, 1]); z = np.ascontiguousarray(order_xyz[:, 2])
    lattice = (float(nx * a), float(ny * a), float(nz * a))
    s = kmc.Structure(element=np.ascontiguousarray(order_el, dtype=np.int32), x=x, y=y, z=z, lattice=lattice, pbc=pbc,
                      nn_dist=3.5, N_left=ncontact, N_right=ncontact, metals=[kmc.Ti_EL, kmc.N_EL], sigma=3.5e-10,
                      k=8.987552e9 / 23.0, T_bg=300.0, freq=1e14, high_G=1.0, low_G=1e-8, Vd=5.0, t_switch=1e-12)
    s.layer = kmc.assign_layers(np.clip(s.x, -21.9, 89.9))
    s.E = kmc.layer_table()
    return s


@pytest.fixture(scope="session")
def s_small(kmc):
    return make_synthetic(kmc)


@pytest.fixture(scope="session")
def ctx(kmc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = kmc.Context(0)
    yield c
    c.close()
