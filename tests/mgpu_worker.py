"""Multi-GPU worker (launched by torchrun from tests/test_multigpu.py and usable stand-alone):
row-sharded K solve + sharded Coulomb vs the single-GPU result, bit for bit."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kmc = importlib.import_module(PKG)
    mg = importlib.import_module(PKG + ".multigpu")
    syn = importlib.import_module(PKG + ".synthetic")
    which = sys.argv[1] if len(sys.argv) > 1 else "5nm"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if which == "5nm":
        s = kmc.load_structure(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"))
    else:
        s = syn.crossbar_standin(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"), 2, 2, order=which)
    ctx = kmc.Context(local)
    sim = mg.DistributedDeviceKMC(s, ctx, rank, world)
    ref = kmc.DeviceKMC(s, ctx=ctx) if rank == 0 else None   # single-GPU run of the same device on rank 0
    ok = True
    report = {"world": world, "which": which, "comm": sim.comm.info(), "steps": []}
    for k in range(steps):
        et, ne = sim.superstep()
        if rank == 0:
            et0, ne0 = ref.superstep()
            same_pot = bool((sim.pot_charge == ref.pot_charge).all().item())
            same_el = bool((sim.element == ref.element).all().item())
            rec = {"cg": sim.last_cg_iterations, "cg_ref": ref.last_cg_iterations, "ne": ne, "ne_ref": ne0,
                   "et_equal": et == et0, "pot_bit_identical": same_pot, "elements_identical": same_el}
            report["steps"].append(rec)
            ok = ok and same_pot and same_el and et == et0 and ne == ne0 and sim.last_cg_iterations == ref.last_cg_iterations
    # every rank must hold the same state
    chk = torch.tensor([float(sim.pot_charge.abs().sum().item()), float(sim.element.sum().item())], device="cuda",
                       dtype=torch.float64)
    parts = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(parts, chk)
    same_everywhere = all(bool((p == parts[0]).all().item()) for p in parts)
    if rank == 0:
        report["ranks_agree"] = same_everywhere
        report["ok"] = bool(ok and same_everywhere)
        print("MGPU_REPORT " + json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (rank != 0 or (ok and same_everywhere)) else 1)


if __name__ == "__main__":
    main()
