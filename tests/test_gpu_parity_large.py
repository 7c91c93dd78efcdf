"""GPU parity on the kernel instantiations bench.py actually runs.

The 5 nm device (143 dot chunks, 1 event super) only reaches the FUSE=true PCG kernels and the single-super event
hierarchy.  The lattices here are large enough for
  * nchunks > 1024  -> spmv_kernel<..,FUSE=false> + dot_finalize_kernel + cg_update_kernel<false> + cg_init_kernel<false>
  * nsuper  > 1     -> the top-level scan over several supers, super_list / super_flag repair, cross-super prefix
                       subtraction in the selector (reference semantics: kmc_events.cu:448-516, upper_bound on the scan)
and are still small enough for the CPU oracle (seconds per superstep).  Every comparison is against the ORACLE.
Bit-exact: CSR sparsity, PCG iteration count, boundary potential (no transcendental functions), event log, elements,
charges.  Total potential / event time: 1e-10 relative (erfc / exp / log ulps), as north_star states.
"""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLD, PKG

pytestmark = pytest.mark.gpu
PARAM_5NM = os.path.join(GOLD, "5nm_device", "parameters.txt")


def to_np(t):
    return t.detach().cpu().numpy()


def _standin(kmc, t, **kw):
    syn = importlib.import_module(PKG + ".synthetic")
    return syn.crossbar_standin(PARAM_5NM, t, t, order="brick", **kw)


def _check_supersteps(kmc, ctx, orc, s, nsteps, expect_types=None):
    dev = kmc.DeviceKMC(s, ctx=ctx)
    sim = orc.OracleSim(s, use_cells=True)
    K = dev.K.to_host()
    for k in ("row_ptr", "col", "left_row_ptr", "left_col", "right_row_ptr", "right_col"):
        assert (K[k] == sim.sp[k]).all(), k
    nchunks = (dev.K.rows + 255) // 256
    nsuper = ((s.N + 255) // 256 + 255) // 256
    assert nchunks > 1024, "this test must reach the FUSE=false PCG kernels (what bench.py runs)"
    assert nsuper > 1, "this test must reach the multi-super event hierarchy (what bench.py runs)"
    seen = np.zeros(5, dtype=np.int64)
    for step in range(nsteps):
        et, ne = dev.superstep()
        log, psum = dev.ev.log()
        r = sim.superstep(max_log=1 << 16)
        assert dev.last_cg_iterations == r["cg_iterations"], (step, dev.last_cg_iterations, r["cg_iterations"])
        assert r["cg_iterations"] > 0
        assert (to_np(dev.pot_boundary) == sim.pot_boundary).all(), f"step {step}: boundary potential not bit-identical"
        assert ne == r["n_events"], (step, ne, r["n_events"])
        ev = r["events"]
        bad = np.nonzero((log != ev).any(axis=1))[0]
        assert bad.size == 0, f"step {step}: first divergent event {bad[0]}: gpu {log[bad[0]]} oracle {ev[bad[0]]}"
        assert (to_np(dev.element) == sim.element).all() and (to_np(dev.charge) == sim.charge).all()
        pot = to_np(dev.pot_charge)
        assert np.abs(pot - sim.pot_total).max() <= 1e-10 * np.abs(sim.pot_total).max()
        assert abs(et - r["event_time"]) <= 1e-10 * r["event_time"]
        seen += np.bincount(ev[:, 2], minlength=5)
        # events landed in more than one super: the cross-super selection path was exercised
        assert len(np.unique(ev[:, 0] >> 16)) > 1
    if expect_types is not None:
        assert (seen[:4] > 0).sum() >= expect_types, seen
    mt, pos = dev.ev.rng_get_state()
    omt, opos = sim.rng.state()
    assert pos == opos and (mt == omt).all()
    dev.ev.close(); dev.K.close()
    return seen


@pytest.mark.parametrize("no_smem", [False, True])
def test_standin4x4_brick_two_supersteps_vs_oracle(kmc, ctx, orc, no_smem):
    """4x4 'brick' stand-in (602 400 sites, 2 282 dot chunks, 10 supers): cold + warm superstep, once with
    event_loop_kernel<true> (chunk sums in shared memory) and once with event_loop_kernel<false> (KMCB200_EV_NO_SMEM)."""
    s = _standin(kmc, 4, Vd=15.0, rnd_seed=32)
    assert s.N == 602400
    old = os.environ.pop("KMCB200_EV_NO_SMEM", None)
    try:
        if no_smem:
            os.environ["KMCB200_EV_NO_SMEM"] = "1"
        _check_supersteps(kmc, ctx, orc, s, 2)
    finally:
        os.environ.pop("KMCB200_EV_NO_SMEM", None)
        if old is not None:
            os.environ["KMCB200_EV_NO_SMEM"] = old


def test_highvac3x3_brick_supersteps_vs_oracle(kmc, ctx, orc):
    """BASELINE config 5 at 3x3 (338 850 sites, 25 % oxygen vacancies, Vd = 5): thousands of PCG iterations on the
    badly conditioned vacancy-rich K, ~1e4 charged sources in the Coulomb sum, all four event classes active."""
    s = _standin(kmc, 3, Vd=5.0, rnd_seed=5, vacancy_concentration=0.25)
    seen = _check_supersteps(kmc, ctx, orc, s, 2, expect_types=3)   # recombinations start in the second superstep
    assert seen[kmc.VACANCY_GENERATION] > 0 and seen[kmc.VACANCY_RECOMBINATION] > 0


def test_persistent_pcg_kernel_vs_oracle(kmc, ctx, orc):
    """The opt-in one-kernel PCG loop (pcg_loop_kernel: grid barriers, ticketed chunks, group stage after the barrier,
    KMCB200_PCG_PERSISTENT=1) against the ORACLE on the 4x4 stand-in: same iteration count, boundary potential bit for bit."""
    s = _standin(kmc, 4, Vd=15.0, rnd_seed=32)
    old = os.environ.get("KMCB200_PCG_PERSISTENT")
    os.environ["KMCB200_PCG_PERSISTENT"] = "1"
    try:
        _check_supersteps(kmc, ctx, orc, s, 2)
    finally:
        if old is None:
            os.environ.pop("KMCB200_PCG_PERSISTENT", None)
        else:
            os.environ["KMCB200_PCG_PERSISTENT"] = old


@pytest.mark.parametrize("mode", ["cells", "flags", "persistent"])
def test_sharded_solve_two_processes_vs_oracle(kmc, tmp_path, mode):
    """Row-sharded K solve + sharded Coulomb sum on 2 ranks against the ORACLE, on the kernel paths
    bench.py runs at N > 1: 4x4 stand-in = 2 282 dot chunks -> FUSE=false kernels + dot_finalize with the two-level
    (group-total) exchange, halo pushes, NVLink all-gather of the potentials.  Driven through the C ABI only (file
    rendezvous + CUDA IPC, no NCCL), so both ranks can share GPU 0 on a one-GPU box: the test never skips.
    mode "cells" (the default protocol): z halo + locally formed p halo, dot contributions as self-validating 16-byte
    cells, no fence / flag in the loop; "flags" (KMCB200_COMM_LL=0): p halo pushes + fence + flag exchanges;
    "persistent": the opt-in one-kernel PCG loop (grid barriers + peer flags inside the kernel)."""
    persistent = "1" if mode == "persistent" else "0"
    ll = "0" if mode == "flags" else "1"
    import json
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu_cabi_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path / "rdv"), str(r if ngpu >= 2 else 0), "4"],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(os.environ, KMCB200_COMM_TIMEOUT_MS="120000", KMCB200_PCG_PERSISTENT=persistent,
                                       KMCB200_COMM_LL=ll))
             for r in range(2)]
    outs = [p.communicate(timeout=1500) for p in procs]
    reps = []
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
        reps.append(json.loads([l for l in so.split("\n") if l.startswith("CABI_REPORT ")][-1][len("CABI_REPORT "):]))
    r0 = [r for r in reps if r["rank"] == 0][0]
    r1 = [r for r in reps if r["rank"] == 1][0]
    assert r0["rows"] % 16384 == 0 and r0["rows"] + r1["rows"] == 583968
    for a, b in zip(r0["steps"], r1["steps"]):
        assert a["cg"] == b["cg"] == a["cg_oracle"] > 0                      # same iteration count as the 1-rank oracle
        assert a["pot_boundary_bit_identical_to_oracle"]                      # sharded dots / SpMV reproduce it bit for bit
        assert a["pot_sum"] == b["pot_sum"] and a["coul_sum"] == b["coul_sum"]  # every rank holds the same full vectors
        assert a["coulomb_max_rel_err"] <= 1e-12
    assert r0["steps"][0]["cg"] != r0["steps"][1]["cg"]                      # the second (warm, changed) solve really iterated
