#!/usr/bin/env python
"""bench.py -- KMC steps/s of the field-solve + event-selection hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: real supersteps of the oracle port on the host cores

Workload (config.workload): BASELINE.json configs[1], "structures/40nm_crossbar full KMC run after
initialization, 1 B200".  The 40 nm xyz files are not shipped with the reference, so the stand-in of
SURVEY.md section 8(d) is used: 8x8 lateral tiling of the shipped 5 nm cell (N = 2 409 600 sites,
num_atoms_first_layer = 36 864, V = 15 V, Device seed 32, KMC seed 1).  A "step" is one KMC superstep:
update_charge -> K assembly + Jacobi-PCG -> screened Coulomb sum -> potential sum -> event rates +
residence-time loop (reference src/kmc_main.cpp:328-540).  Inputs (~0.75 GB matrix + 1 GB event list per step)
are far larger than L2 (126 MB), so no L2 flush is needed between timed steps.

Both arms run the SAME trajectory (same seeds; the GPU path is bit-compatible with the oracle), so their per-step
counters (PCG iterations, events) agree step by step; the GPU arm checks its first superstep against the oracle
(`parity`, outside the timed region) at every GPU count.

One JSON line on stdout (rank 0).  DESIGN.md section 7 documents every key.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"
PARAM_5NM = os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt")
PROFILES = os.path.join(ROOT, "profiles")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi SM clocks + throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def standin_tiles(base):
    """'standinTxT' -> T (lateral tiling factor of the shipped 5 nm cell)"""
    import re
    m = re.fullmatch(r"standin(\d+)x(\d+)", base)
    if not m or m.group(1) != m.group(2):
        raise ValueError(f"unknown workload '{base}'")
    return int(m.group(1))


def workload_desc(name, N, N_left):
    if name == "5nm":
        return "structures/5nm_device (shipped), N=37650"
    base, _, order = name.partition("_")
    order = order or "file"
    if base == "highvac7x7":
        return (f"synthetic high-vacancy lattice: 7x7 lateral tiling of the shipped 5nm cell, N={N}, 25 % oxygen "
                f"vacancies, Vd=5, site order '{order}'")
    t = standin_tiles(base)
    note = {"file": "site order 'file' (tile images site-major: the 5nm file's block structure, wide K bandwidth)",
            "brick": "site order 'brick' (bandwidth-minimised: interior sites grouped in 12.5 A cubes, contacts "
                     "first/last -- the layout the reference's crossbar_40_bwmin.xyz input is named for)"}
    return (f"40nm_crossbar stand-in: {t}x{t} lateral tiling of the shipped 5nm cell, N={N}, "
            f"num_atoms_first_layer={N_left}, Vd=15, " + note.get(order, f"site order '{order}'"))


def build_workload(kmc, name):
    """product-side construction (host model of libkmc_b200.so); oracle/workload.py builds the identical structures for
    the CPU arm without the product (tests/test_workload.py)"""
    syn = importlib.import_module(PKG + ".synthetic")
    if name == "5nm":
        s = kmc.load_structure(PARAM_5NM)
    else:
        base, _, order = name.partition("_")
        order = order or "file"
        if base == "highvac7x7":
            s = syn.crossbar_standin(PARAM_5NM, 7, 7, order=order, Vd=5.0, rnd_seed=5, vacancy_concentration=0.25)
        else:
            t = standin_tiles(base)
            s = syn.crossbar_standin(PARAM_5NM, t, t, order=order, Vd=15.0, rnd_seed=32)
    return s, workload_desc(name, s.N, s.N_left)


def use_all_host_threads():
    """The CPU arm uses every host core it may run on.  torchrun exports OMP_NUM_THREADS=1 for its workers; the OpenMP
    runtime of the oracle library reads the variable when it is loaded, so it is reset before the first import."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


# ------------------------------------------------------------------------------------------------
# CPU arm: real, state-advancing supersteps of the oracle port (the reference has no CPU implementation of this
# path and its GPU build needs hipcc + ROCm + MPI: SURVEY.md 8c)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = use_all_host_threads()
    from oracle import binding as orc
    from oracle import workload
    orc.lib().orc_set_num_threads(nthreads)
    w, desc = workload.build(args.workload)
    t0 = time.perf_counter()
    sim = orc.OracleSim(w, use_cells=True)
    t_setup = time.perf_counter() - t0
    recs, stage = [], np.zeros(4)
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = sim.superstep(max_log=16)
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            recs.append((dt, r["cg_iterations"], r["n_events"]))
            stage += np.array(r["t"])
    ms = 1e3 * float(np.mean([r[0] for r in recs]))
    val = 1e3 / ms
    cores = orc.lib().orc_num_threads()
    sample = (f"{args.steps} full, state-advancing supersteps of the same trajectory the GPU arm runs (after {args.warmup} "
              f"warm-up supersteps): update_charge + K assembly + Jacobi-PCG to the reference tolerance + cell-binned "
              f"Coulomb sum + rate list + residence-time loop; nothing sampled or extrapolated")
    line = {"impl": "reference", "metric": "kmc_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic" if args.workload != "5nm" else "shipped structure",
            "config": {"workload": desc, "cg_iterations_per_step": float(np.mean([r[1] for r in recs])),
                       "events_per_step": float(np.mean([r[2] for r in recs])), "setup_s": round(t_setup, 2)},
            "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "stage_seconds_per_step": dict(zip(["charge", "K_assembly_plus_pcg", "coulomb", "rates_plus_events"],
                                                                [round(float(v) / len(recs), 4) for v in stage]))},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def parity_first_superstep(s, sim, first, world):
    """CHECKER (outside every timed region): the GPU path's first superstep against the CPU oracle on the same inputs.
    `first` holds what the GPU produced in superstep 1.  Returns (report, OracleSim)."""
    nthreads = use_all_host_threads()
    from oracle import binding as orc
    orc.lib().orc_set_num_threads(nthreads)   # (an OpenMP runtime loaded earlier -- torch's -- has already read the env)
    t0 = time.perf_counter()
    osim = orc.OracleSim(s, use_cells=True)
    r = osim.superstep(max_log=1 << 16)
    NL, n = s.N_left, s.N - s.N_left - s.N_right
    log = first["log"]
    ne = min(len(log), len(r["events"]))
    rep = {"checked_against": "oracle (oracle/kmc_oracle.cpp), superstep 1 (cold PCG), same inputs",
           "cg_iterations": [int(first["cg"]), int(r["cg_iterations"])],
           "cg_iterations_equal": int(first["cg"]) == int(r["cg_iterations"]),
           "pot_boundary_bit_identical": bool((first["pot_boundary"][NL:NL + n] == osim.pot_boundary[NL:NL + n]).all()),
           "n_events": [int(first["ne"]), int(r["n_events"])],
           "event_log_identical": bool(first["ne"] == r["n_events"] and (log[:ne] == r["events"][:ne]).all()),
           "elements_identical": bool((first["element"] == osim.element).all()),
           "total_potential_max_rel_err": float(np.abs(first["pot_total"] - osim.pot_total).max() /
                                                max(np.abs(osim.pot_total).max(), 1e-300)),
           "seconds": round(time.perf_counter() - t0, 1)}
    rep["ok"] = bool(rep["cg_iterations_equal"] and rep["pot_boundary_bit_identical"] and rep["event_log_identical"] and
                     rep["elements_identical"] and rep["total_potential_max_rel_err"] <= 1e-10)
    if world > 1:
        rep["sharded_bit_identical"] = rep["ok"]   # the row-sharded solve reproduces the 1-rank oracle bit for bit
    return rep, osim


def time_ms(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def pcg_fixed_iterations(kmc, ctx, s, rank, world, dist, iters=60, solves=3):
    """cold Jacobi-PCG with a FIXED iteration count (tolerance 0) on workload s, row-sharded over `world` ranks:
    ms per iteration (max over ranks) and algorithmic GB/s per GPU"""
    import torch
    mg = importlib.import_module(PKG + ".multigpu")
    c = ctx
    x, y, z = c.dev_d(s.x), c.dev_d(s.y), c.dev_d(s.z)
    element, charge = c.dev_i(s.element), c.empty_i(s.N, 0)
    n = s.N - s.N_left - s.N_right
    if world > 1:
        w = c.sparsity_K_row_counts(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
        counts, displs = mg.balanced_partition(w.cpu().numpy(), world)
        del w
    else:
        counts, displs = kmc.partition(n, world, aligned=True)
    comm = mg.Comm(c, rank, world, n, counts, displs, dist if world > 1 else None)
    K = c.initialize_sparsity_K(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, int(displs[rank]), int(counts[rank]))
    comm.attach(K, dist if world > 1 else None)
    neigh = c.compute_neighbor_list(x, y, z)
    c.update_charge(element, charge, neigh, s.metals)
    del neigh
    c.assemble_K(K, s.N, s.N_left, s.N_right, element, charge, s.metals, s.Vd, s.high_G, s.low_G)
    rows = K.rows
    dinv, rhs0 = K.device_vectors()
    rhs_t = torch.empty_like(rhs0)
    xsol = torch.zeros(rows, dtype=torch.float64, device=c.device)
    times = []
    for k in range(solves + 1):
        rhs_t.copy_(rhs0)
        xsol.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        it = c.pcg_jacobi(K, rhs_t, xsol, dinv, 0.0, iters)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=c.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if k > 0:
            times.append(float(t.item()) / max(it, 1))
    import hashlib
    x_sha1 = hashlib.sha1(xsol.cpu().numpy().tobytes()).hexdigest()[:16]   # this rank's slice of the iterate (A/B runs must agree bit for bit)
    nnz_local = K.nnz
    tot = torch.tensor([float(nnz_local)], device=c.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    nnz = float(tot.item())
    ms_it = float(np.mean(times))
    out = {"N": int(s.N), "rows": int(n), "nnz": int(nnz), "gpus": world, "iterations_per_solve": iters,
           "ms_per_pcg_iteration": ms_it,
           "GBs_per_gpu": (12.0 * nnz + 108.0 * n) / world / (ms_it * 1e-3) / 1e9,
           "partition_rows": [int(v) for v in counts], "x_sha1_rank0": x_sha1}
    K.close(); comm.close()
    return out


def kirchhoff_chain(kmc, ctx, name, peak):
    """BASELINE config 3's split-sparse leg (dist_iterative_test/main_test_cg_split.cpp shape): the Kirchhoff / current
    chain on a stand-in lattice -- CB-edge solve, T_neighbor + WKB tunnel block assembly, 100-iteration split-sparse
    Jacobi-PCG (the reference's fixed count, current_solver_gpu.cu:1456) and I_macro -- plus the split SpMV alone.
    Constants as src/kmc_main.cpp:294-302."""
    import torch
    s, desc = build_workload(kmc, name)
    p = kmc.parse_parameters(PARAM_5NM)
    dev = kmc.DeviceKMC(s, ctx=ctx)
    dev.field_solve()
    cb = ctx.empty_d(s.N, 0.0)
    t0 = time.perf_counter()
    it_cb = ctx.update_CB_edge(dev.K, s.N, s.N_left, s.N_right, dev.element, s.metals, s.Vd, s.high_G, s.low_G, cb)
    torch.cuda.synchronize()
    cb_ms = 1e3 * (time.perf_counter() - t0)
    T = ctx.initialize_sparsity_T(dev.element, dev.x, dev.y, dev.z, s.nn_dist, s.N_left, s.N_left, int(p.num_layers_contact))
    loop_G, high_G, low_G = s.high_G * 10000000, s.high_G * 100000, s.low_G
    G0, m_e, V0 = 2 * 3.8612e-5 * 1e-5, float(p.m_r) * 9.11e-31, float(p.V0)
    V = ctx.empty_d(T.N_atom + 1, 0.0)
    chain = lambda: ctx.update_power_sparse(T, dev.element, dev.charge, cb, s.metals, s.Vd, high_G, low_G, loop_G, G0, m_e, V0, V)
    im, it = chain()
    chain_ms = time_ms(torch, chain, 2)
    asm_ms = time_ms(torch, lambda: ctx.assemble_T(T, dev.element, dev.charge, cb, s.metals, s.Vd, high_G, low_G, loop_G, m_e, V0), 2)
    T.refresh()
    xv, yv = ctx.empty_d(T.N_atom + 1, 1.0), ctx.empty_d(T.N_atom + 1, 0.0)
    spmv_ms = time_ms(torch, lambda: ctx.tmat_spmv(T, xv, yv), 10)
    spmv_bytes = 12.0 * (T.nnz + T.tunnel_nnz) + 20.0 * (T.N_atom + 1) + 16.0 * T.n_tunnel
    out = {"workload": desc, "N_atom": int(T.N_atom), "T_neighbor_nnz": int(T.nnz), "tunnel_points": int(T.n_tunnel),
           "tunnel_nnz": int(T.tunnel_nnz), "cb_edge_solve_ms": cb_ms, "cb_edge_iterations": int(it_cb),
           "assemble_T_ms": asm_ms, "chain_ms_assemble_plus_100_iterations_plus_imacro": chain_ms,
           "ms_per_split_sparse_pcg_iteration": (chain_ms - asm_ms) / max(it, 1), "pcg_iterations": int(it),
           "split_spmv_ms": spmv_ms, "split_spmv_GBs": spmv_bytes / (spmv_ms * 1e-3) / 1e9,
           "split_spmv_frac_of_hbm_peak": spmv_bytes / (spmv_ms * 1e-3) / 1e9 / peak, "I_macro_A": float(im)}
    T.close(); dev.ev.close(); dev.K.close()
    return out


def run_extras(args, kmc, ctx, rank, world, dist, peak):
    """Other BASELINE.json configurations, measured outside the headline region.  Rank 0 alone runs
    the single-GPU ones; the field-solve scaling entry uses all ranks."""
    import torch
    extras = {}
    t_begin = time.perf_counter()
    if rank == 0:
        def guarded(key, fn):
            """an auxiliary measurement must never cost the headline line (rank 0 always reaches the barrier below)"""
            try:
                torch.cuda.empty_cache()
                extras[key] = fn()
            except Exception as e:
                extras[key] = {"error": repr(e)[:300]}

        # ---- config 1: the shipped 5 nm device, steady-state KMC steps/s (the only number the reference publishes: ~87/s)
        def cfg_5nm():
            s5, _ = build_workload(kmc, "5nm")
            d5 = kmc.DeviceKMC(s5, ctx=ctx)
            for _ in range(3):
                d5.superstep()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nst = 200
            for _ in range(nst):
                d5.superstep()
            torch.cuda.synchronize()
            ms5 = 1e3 * (time.perf_counter() - t0) / nst
            d5.ev.close(); d5.K.close()
            return {"workload": workload_desc("5nm", s5.N, s5.N_left), "kmc_steps_per_sec": 1e3 / ms5,
                    "ms_per_step": ms5, "supersteps_timed": nst, "vs_baseline": (1e3 / ms5) / 87.0,
                    "baseline": "87 steps/s steady state, 1 MI250X GCD (reference output1_0.txt)"}
        guarded("5nm_device", cfg_5nm)

        # ---- config 3: exported Poisson CSR + Jacobi-PCG at the dist_iterative_test/main_test_cg.cpp:182-207 shapes
        #      (sub-tilings of the stand-in: K after step-0 assembly, x0 = 0, fixed 60 iterations)
        def cfg_exported_csr():
            shapes = []
            for name, want in (("5nm", "7 302 x 186 684 .. closest shipped: 36 498 rows"), ("standin2x2_brick", "70 630 / 1 719 652"),
                               ("standin4x4_brick", "403 605 / 10 007 089"), ("standin8x8_brick", "1 632 355 / 41 208 963")):
                if name == "standin8x8_brick" and args.workload == "standin8x8_brick":
                    continue   # that shape is the headline's own roofline block
                sw, _ = build_workload(kmc, name)
                r = pcg_fixed_iterations(kmc, ctx, sw, 0, 1, None)
                r["reference_shape_rows_nnz"] = want
                r["frac_of_hbm_peak"] = r["GBs_per_gpu"] / peak
                shapes.append(r)
            return shapes
        guarded("exported_csr_pcg", cfg_exported_csr)

        # ---- config 5: ~2 M-site high-vacancy lattice (charge sum + rate list + event selection stressed)
        def cfg_highvac():
            sh, desc = build_workload(kmc, "highvac7x7_brick")
            dh = kmc.DeviceKMC(sh, ctx=ctx)
            dh.superstep()
            dh.field_ms_total = 0.0; dh.events_ms_total = 0.0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            et, ne = dh.superstep()
            torch.cuda.synchronize()
            msh = 1e3 * (time.perf_counter() - t0)
            q, tests, inr = ctx.poisson_stats()
            out = {"workload": desc, "ms_per_step": msh, "kmc_steps_per_sec": 1e3 / msh,
                   "field_solve_ms": dh.field_ms_total, "events_ms": dh.events_ms_total,
                   "cg_iterations": dh.last_cg_iterations, "events": ne, "charged_sources": q,
                   "us_per_event": 1e3 * dh.events_ms_total / max(ne, 1)}
            dh.ev.close(); dh.K.close()
            return out
        if args.workload != "highvac7x7_brick":
            guarded("highvac7x7", cfg_highvac)
        # ---- config 3, split-sparse leg: Kirchhoff chain with its tunnel block (O(tunnel points^2) non-zeros)
        for wname in ("5nm", "standin4x4_brick"):
            try:
                torch.cuda.empty_cache()
                extras.setdefault("split_sparse_pcg", []).append(kirchhoff_chain(kmc, ctx, wname, peak))
            except Exception as e:  # never lose the headline line to an auxiliary measurement
                extras.setdefault("split_sparse_pcg", []).append({"workload": wname, "error": repr(e)[:300]})
    if world > 1:
        dist.barrier()
    # ---- config 4: >= 10 M-site lattice, strong scaling of the field solve's PCG (the north_star's 0.7 target)
    #      (two lattices: 9.6 M sites -- the figure kept since round 1 -- and 21.7 M sites, strictly above 10 M)
    if not args.no_scaling_extra:
        for key, wname in (("field_solve_scaling", "standin16x16_brick"), ("field_solve_scaling_21M", "standin24x24_brick")):
            torch.cuda.empty_cache()
            sw, desc = build_workload(kmc, wname)
            r = pcg_fixed_iterations(kmc, ctx, sw, rank, world, dist)
            r["workload"] = desc
            r["frac_of_hbm_peak_per_gpu"] = r["GBs_per_gpu"] / peak
            extras[key] = r
            del sw
    extras["seconds"] = round(time.perf_counter() - t_begin, 1)
    return extras


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kmc = importlib.import_module(PKG)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(minutes=30))
    s, desc = build_workload(kmc, args.workload)
    ctx = kmc.Context(local_rank)
    t_setup0 = time.perf_counter()
    if world > 1:
        par = importlib.import_module(PKG + ".multigpu")
        sim = par.DistributedDeviceKMC(s, ctx=ctx, rank=rank, world=world)
    else:
        sim = kmc.DeviceKMC(s, ctx=ctx)
    ctx.sync()
    t_setup = time.perf_counter() - t_setup0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first superstep = cold PCG from a zero guess); superstep 1 is also the parity sample ----------
    per_step, first = [], None
    for k in range(args.warmup):
        et, ne = sim.superstep()
        per_step.append((sim.last_cg_iterations, ne))
        if k == 0 and rank == 0 and not args.no_parity_check:
            first = {"cg": sim.last_cg_iterations, "ne": ne, "log": sim.ev.log()[0],
                     "pot_boundary": sim.pot_boundary.cpu().numpy(), "pot_total": sim.pot_charge.cpu().numpy(),
                     "element": sim.element.cpu().numpy()}
    # ---- timed region: K supersteps, device resident ----------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    snap = sim.snapshot()   # the e2e leg below replays exactly these supersteps
    launches0 = kmc.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    timed = []
    sim.field_ms_total = 0.0
    sim.events_ms_total = 0.0
    for _ in range(args.steps):
        et, ne = sim.superstep()
        timed.append((sim.last_cg_iterations, ne))
    e1.record()
    field_ms = sim.field_ms_total / args.steps      # sharded part: charge + K assembly + PCG + Coulomb (+ all-gathers)
    events_ms = sim.events_ms_total / args.steps    # replicated part: rate list + residence-time loop
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = kmc.launch_count() - launches0
    if world > 1:
        tt = torch.tensor([ms_total, field_ms, events_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, field_ms, events_ms = (float(v) for v in tt.tolist())
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- e2e: same supersteps through the host-buffer API (H2D of the mutable site state from pinned host memory,
    #      D2H of element/charge/potential every step), reference GPUBuffers::sync_HostToGPU/GPUToHost -------------
    N = s.N
    h_el = torch.empty(N, dtype=torch.int32).pin_memory()
    h_ch = torch.empty(N, dtype=torch.int32).pin_memory()
    h_pot = torch.empty(N, dtype=torch.float64).pin_memory()
    sim.restore(snap)
    h_el.copy_(sim.element); h_ch.copy_(sim.charge)
    torch.cuda.synchronize()
    e2e_steps = max(1, args.steps)
    e2e_rec = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.element.copy_(h_el, non_blocking=True)
        sim.charge.copy_(h_ch, non_blocking=True)
        _, ne2 = sim.superstep()
        e2e_rec.append((sim.last_cg_iterations, ne2))
        h_el.copy_(sim.element, non_blocking=True)
        h_ch.copy_(sim.charge, non_blocking=True)
        h_pot.copy_(sim.pot_charge, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    if rank == 0:
        sampler.stop_flag = True
    clocks = sampler.summary() if rank == 0 else None

    cg_per_step = float(np.mean([c for c, _ in timed]))
    ev_per_step = float(np.mean([e for _, e in timed]))
    peak, peak_src = peaks()

    # ---- stage breakdown + rooflines, CUDA events on the library stream (N = 1) ----------------------------------
    stages, roof, roof_ev, roof_coul = {}, None, None, None
    if world == 1:
        K = sim.K
        n = K.rows
        xv = ctx.empty_d(n, 1.0)
        yv = ctx.empty_d(n, 0.0)
        t_spmv_plain = time_ms(torch, lambda: ctx.spmv(K, xv, yv), 20)
        t_spmv = time_ms(torch, lambda: ctx.spmv_dot(K, xv, yv, want_scalar=False), 20)   # fused SpMV + p.Ap (+ finalize)
        spmv_bytes = 12.0 * K.nnz + 20.0 * n           # val 8 + col 4 per nnz; row_ptr 4 + y 8 + x 8 per row
        ach = spmv_bytes / (t_spmv * 1e-3) / 1e9
        roof = {"kernel": "spmv_kernel<8,DOT> (CSR SpMV with the p.Ap chunk partials fused; + dot_finalize)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "frac_of_nominal_8TBs": ach / 8000.0,   # report both denominators
                "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": t_spmv,
                "share_of_step": cg_per_step * t_spmv / ms_per_step}
        try:
            prof = json.load(open(os.path.join(PROFILES, "spmv_traffic.json")))
            roof["traffic"] = prof.get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        el2, ch2 = sim.element.clone(), sim.charge.clone()
        stages["update_charge_ms"] = time_ms(torch, lambda: ctx.update_charge(el2, ch2, sim.neigh, s.metals), 5)
        stages["assemble_K_ms"] = time_ms(torch, lambda: ctx.assemble_K(K, N, s.N_left, s.N_right, el2, ch2, s.metals, s.Vd,
                                                                        s.high_G, s.low_G), 5)
        stages["assemble_K_GBs"] = (12.0 * K.nnz + 36.0 * n) / (stages["assemble_K_ms"] * 1e-3) / 1e9
        pc = ctx.empty_d(N, 0.0)
        stages["coulomb_ms"] = time_ms(torch, lambda: ctx.poisson_gridless(sim.x, sim.y, sim.z, el2, ch2, s.sigma, s.k, pc), 3)
        q, tests, inrange = ctx.poisson_stats()
        stages["coulomb_charged_sources"] = q
        stages["coulomb_pair_tests_per_s"] = tests / (stages["coulomb_ms"] * 1e-3)
        # an in-range pair = 60 FP64 flop-equivalents, a predicate-only test = 8
        fp64_peak = ctx.fp64_peak_tflops()
        flops = 60.0 * inrange + 8.0 * (tests - inrange)
        ach_tf = flops / (stages["coulomb_ms"] * 1e-3) / 1e12
        roof_coul = {"kernel": "coulomb_cell_kernel (+ source compaction and per-cell lists; whole poisson_gridless call)",
                     "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                     "peak_source": "kmcb200_fp64_peak: 8 independent DFMA chains per thread, measured in this run",
                     "pairs_in_range": inrange, "pair_tests": tests, "flop_model": "60 per in-range pair, 8 per predicate-only test",
                     "ms_per_call": stages["coulomb_ms"], "share_of_step": stages["coulomb_ms"] / ms_per_step}
        stages["build_event_list_ms"] = time_ms(torch, lambda: sim.ev.build_event_list(sim.neigh, sim.layer, s.T_bg, s.freq, s.sigma,
                                                                                       s.k, sim.x, sim.y, sim.z, sim.pot_charge,
                                                                                       el2, ch2), 3)
        stages["build_event_list_GBs"] = N * 52.0 * 13.0 / (stages["build_event_list_ms"] * 1e-3) / 1e9
        stages["spmv_dot_ms"] = t_spmv
        stages["spmv_dot_GBs"] = ach
        stages["spmv_plain_ms"] = t_spmv_plain
        stages["spmv_plain_GBs"] = spmv_bytes / (t_spmv_plain * 1e-3) / 1e9
        u, v = ctx.empty_d(n, 1.0), ctx.empty_d(n, 2.0)
        stages["dot_ms"] = time_ms(torch, lambda: ctx.dot(u, v), 10)
        if cg_per_step > 0:
            # one Jacobi-PCG iteration = SpMV + 11 vector streams (12 nnz + 108 n bytes); time derived from
            # the timed supersteps: field solve minus the separately timed charge / assembly / Coulomb stages
            pcg_ms = (field_ms - stages["update_charge_ms"] - stages["assemble_K_ms"] - stages["coulomb_ms"]) / cg_per_step
            stages["pcg_iteration_ms"] = pcg_ms
            stages["pcg_iteration_GBs"] = (12.0 * K.nnz + 108.0 * n) / (pcg_ms * 1e-3) / 1e9
            stages["pcg_iteration_frac_of_hbm_peak"] = stages["pcg_iteration_GBs"] / peak
        # the time-dominant kernel: the persistent event loop.  It is a chain of DEPENDENT memory round trips and
        # single-warp scans (one event cannot start before the previous one has repaired the rate sums), so its bound
        # is latency, not bandwidth: see profiles/r2_event_loop.md for the measured chain.
        loop_ms = events_ms - stages["build_event_list_ms"]
        us_ev = 1e3 * loop_ms / max(ev_per_step, 1.0)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        try:
            evp = json.load(open(os.path.join(PROFILES, "r2_event_loop.json")))
        except Exception:
            evp = {}
        floor_us = evp.get("floor_us_per_event", 1.2)
        roof_ev = {"kernel": "event_loop_kernel (one persistent CTA; the residence-time loop of kmc_events.cu:448-516)",
                   "bound": "latency", "us_per_event": us_ev, "cycles_per_event": us_ev * sm_mhz,
                   "events_per_step": ev_per_step, "share_of_step": loop_ms / ms_per_step,
                   "floor_us_per_event": floor_us, "floor_model": evp.get("floor_model"), "frac": floor_us / us_ev,
                   "dram_bytes_per_event": evp.get("dram_bytes_per_event"), "evidence": "profiles/r2_event_loop.md"}

    # ---- parity vs the oracle (rank 0; every GPU count) and the CPU baseline (N = 1), outside the timed region --------
    parity, cpu = None, None
    if rank == 0 and first is not None:
        parity, osim = parity_first_superstep(s, sim, first, world)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding as orc
            nb = 2
            t0 = time.perf_counter()
            recs = [osim.superstep(max_log=16) for _ in range(nb)]   # supersteps 2.. of the same trajectory
            dt = (time.perf_counter() - t0) / nb
            cpu = {"value": 1.0 / dt, "unit": "steps/s", "cores": orc.lib().orc_num_threads(), "kind": "port",
                   "sample": (f"{nb} full, state-advancing oracle supersteps (supersteps 2..{nb + 1} of the same trajectory; the "
                              f"reference has no CPU code for this path): nothing sampled or extrapolated"),
                   "cg_iterations": [int(r["cg_iterations"]) for r in recs], "events": [int(r["n_events"]) for r in recs],
                   "gpu_counters_same_steps": per_step[1:1 + nb]}
        del osim
    if world > 1:
        dist.barrier()

    extras = None
    if not args.no_extras:
        sim.ev.close(); sim.K.close()
        Krows, Knnz = int(sim.K.rows), int(sim.K.nnz)
        del sim
        torch.cuda.empty_cache()
        extras = run_extras(args, kmc, ctx, rank, world, dist, peak)
    else:
        Krows, Knnz = int(sim.K.rows), int(sim.K.nnz)

    if rank == 0:
        line = {"metric": "kmc_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                # BASELINE.md publishes one number for this metric: ~87 KMC steps/s, steady state, shipped 5 nm device
                "vs_baseline": (value / 87.0 if args.workload == "5nm" else None), "dtype": "f64",
                "data": ("shipped structure" if args.workload == "5nm" else "synthetic"),
                "config": {"workload": desc, "l2_policy": "inputs (matrix + event list) >> L2; no flush needed",
                           "cg_iterations_per_step": cg_per_step, "events_per_step": ev_per_step,
                           "warmup_counters": per_step, "setup_s": round(t_setup, 2), "K_rows": Krows, "K_nnz": Knnz},
                "e2e": {"value": 1e3 / e2e_ms, "unit": "steps/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 16 * N,
                        "replays_timed_supersteps": e2e_rec == timed[:len(e2e_rec)]},
                "field_solve_ms_per_step": field_ms, "events_ms_per_step": events_ms,
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_events": roof_ev,
                "roofline_coulomb": roof_coul, "stages": stages, "parity": parity, "cpu_baseline": cpu,
                "other_configs": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("KMC_BENCH_WORKLOAD", "standin8x8_brick"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison of superstep 1")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (other_configs)")
    ap.add_argument("--no-scaling-extra", action="store_true", help="skip the 9.6 M-site field-solve scaling entry")
    args = ap.parse_args()
    if args.warmup < 1 and not args.no_parity_check:
        args.no_parity_check = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
