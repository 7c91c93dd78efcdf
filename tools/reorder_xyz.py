"""Write a bandwidth-minimised ('brick') copy of a DeviceKMC restart xyz file.

The reference keeps the site order of its input file (src/Device.cpp:40-41, "DO NOT SORT") and ships pre-reordered
inputs (reordered_device_5.xyz, crossbar_40_bwmin.xyz) because the K-matrix bandwidth -- and with it the locality of the
SpMV gather and of the event-list repairs -- is a property of that order.  This tool produces such an input: the first /
last `num_atoms_first_layer` sites (contact layers) stay in place, interior sites are grouped into cubes of edge B.

    python tools/reorder_xyz.py parameters.txt out.xyz [B=12.5]

Reordering sites changes event-slot order and therefore the KMC trajectory: use it for production /
throughput runs, not for parity runs against a trajectory recorded in another order."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"


def main():
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    kmc = importlib.import_module(PKG)
    syn = importlib.import_module(PKG + ".synthetic")
    B = float(sys.argv[3]) if len(sys.argv) > 3 else 12.5
    s = kmc.load_structure(sys.argv[1], apply_vacancies=False)
    perm = syn.brick_permutation(s.x, s.y, s.z, s.N_left, s.N_right, B)
    names = kmc.ELEMENT_NAMES
    with open(sys.argv[2], "w") as f:
        f.write(f"{s.N}\n\n")
        for k in perm:
            f.write(f"{names[int(s.element[k])]}   {s.x[k]:.10g}   {s.y[k]:.10g}   {s.z[k]:.10g}\n")
    print(f"wrote {s.N} sites to {sys.argv[2]} (cube edge {B} A, contacts {s.N_left}/{s.N_right} kept in place)")


if __name__ == "__main__":
    main()
