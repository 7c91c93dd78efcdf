"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference for the 5 nm data, which was copied verbatim to
tests/golden/5nm_device/ by the first section below, and the CPU oracle):

    python tests/golden/make_golden.py [--steps 1000]

Outputs
  tests/golden/5nm_device/{parameters.txt,reordered_device_5.xyz,output1_0.txt,snapshot_*.xyz.gz}
      verbatim copies of the reference's shipped input + expected output (structures/5nm_device/).
  tests/golden/40nm_parameters.txt   verbatim copy of structures/40nm_crossbar/parameters.txt
  tests/golden/ref_parser_5nm.json, ref_parser_40nm.json
      every field the REFERENCE's own parser (oracle/_ref, compiled from src/input_parser.cpp) returns.
  tests/golden/traj_5nm.json
      oracle trajectory of the 5 nm device (seeds 5 / 1) for --steps supersteps: per step the events
      (i, j, type), event count, PCG iterations and the event time as a hex float.
"""
import argparse
import ctypes as C
import gzip
import importlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"


def copy_reference_data():
    if not os.path.isdir(REF):
        return
    d = os.path.join(GOLD, "5nm_device")
    os.makedirs(d, exist_ok=True)
    src = os.path.join(REF, "structures", "5nm_device")
    for f in ("parameters.txt", "reordered_device_5.xyz"):
        shutil.copyfile(os.path.join(src, f), os.path.join(d, f))
    shutil.copyfile(os.path.join(src, "expected_output", "output1_0.txt"), os.path.join(d, "output1_0.txt"))
    for f in ("snapshot_init.xyz", "snapshot_6.xyz"):
        with open(os.path.join(src, "expected_output", "Results_5.000000", f), "rb") as fi, \
                gzip.GzipFile(os.path.join(d, f + ".gz"), "wb", 9, mtime=0) as fo:
            fo.write(fi.read())
    shutil.copyfile(os.path.join(REF, "structures", "40nm_crossbar", "parameters.txt"),
                    os.path.join(GOLD, "40nm_parameters.txt"))


def dump_reference_parser():
    from oracle import binding as orc
    kmc = importlib.import_module(PKG)
    L = orc.ref_lib()
    if L is None:
        return
    for name, path in (("5nm", os.path.join(GOLD, "5nm_device", "parameters.txt")),
                       ("40nm", os.path.join(GOLD, "40nm_parameters.txt"))):
        p = kmc.Params()  # same POD layout as ref_params in oracle/ref_host_shim.cpp
        L.ref_parse_params(path.encode(), C.byref(p))
        out = {}
        for fname, ftype in p._fields_:
            v = getattr(p, fname)
            if isinstance(v, bytes):
                v = v.decode()
            elif hasattr(v, "__len__"):
                v = list(v)
            out[fname] = v
        with open(os.path.join(GOLD, f"ref_parser_{name}.json"), "w") as f:
            json.dump(out, f, indent=1)


def make_trajectory(steps):
    from oracle import binding as orc
    kmc = importlib.import_module(PKG)
    s = kmc.load_structure(os.path.join(GOLD, "5nm_device", "parameters.txt"))
    sim = orc.OracleSim(s)
    rec = []
    t0 = time.time()
    for _ in range(steps):
        r = sim.superstep()
        rec.append({"n_events": int(r["n_events"]), "cg": int(r["cg_iterations"]),
                    "event_time": float(r["event_time"]).hex(),
                    "events": [[int(e[0]), int(e[1]), int(e[2])] for e in r["events"]]})
        if len(rec) % 100 == 0:
            print(len(rec), "steps", round(time.time() - t0, 1), "s", flush=True)
    import numpy as np
    final = {"steps": steps, "kmc_time": float(sim.kmc_time).hex(),
             "n_vacancy": int((sim.element == 2).sum()), "n_charged": int((sim.charge != 0).sum()),
             "pot_abs_sum": float(np.abs(sim.pot_total).sum()).hex()}
    with open(os.path.join(GOLD, "traj_5nm.json"), "w") as f:
        json.dump({"final": final, "steps": rec}, f, separators=(",", ":"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    a = ap.parse_args()
    copy_reference_data()
    dump_reference_parser()
    make_trajectory(a.steps)
