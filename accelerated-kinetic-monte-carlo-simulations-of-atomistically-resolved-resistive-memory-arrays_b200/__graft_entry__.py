"""Driver entry points: build() compiles every native component, smoke() runs one small hot-path invocation on
cuda:0 and checks it against the CPU oracle."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build() -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo for every .cu (in-tree libkmc_b200.so), g++ for the
    host model, the oracle's C++ restatement, and -- only when /root/reference is present -- oracle/_ref from the
    reference's own host sources.  Then imports the package (loads the .so, checks every declared symbol)."""
    subprocess.run(["bash", os.path.join(ROOT, PKG, "build.sh")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"),
                    os.path.join(ROOT, "oracle", "_build", "libkmc_oracle.so")], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    kmc = importlib.import_module(PKG)
    kmc.load_library()


def smoke() -> None:
    """Two KMC supersteps of the shipped 5 nm device on cuda:0 through the C ABI, checked against the oracle:
    CSR sparsity bit-exact, same PCG iteration count, potentials within 1e-10 relative, identical events."""
    import numpy as np
    import torch
    kmc = importlib.import_module(PKG)
    from oracle import binding as orc
    assert torch.cuda.is_available(), "smoke() needs a CUDA device (no CPU fallback)"
    s = kmc.load_structure(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"))
    dev = kmc.DeviceKMC(s, device=0)
    sim = orc.OracleSim(s, use_cells=True)
    K = dev.K.to_host()
    assert (K["row_ptr"] == sim.sp["row_ptr"]).all() and (K["col"] == sim.sp["col"]).all()
    for _ in range(2):
        et, ne = dev.superstep()
        log, _ = dev.ev.log()
        r = sim.superstep()
        assert dev.last_cg_iterations == r["cg_iterations"], (dev.last_cg_iterations, r["cg_iterations"])
        assert ne == r["n_events"] and (log[:, :3] == r["events"][:, :3]).all()
        pot = dev.pot_charge.cpu().numpy()
        assert np.abs(pot - sim.pot_total).max() <= 1e-10 * np.abs(sim.pot_total).max()
        assert abs(et - r["event_time"]) <= 1e-12 * r["event_time"]
    print(f"smoke ok: 2 supersteps, {kmc.launch_count()} kernel launches, kmc_time={dev.kmc_time:.6g}")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
