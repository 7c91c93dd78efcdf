"""Strong scaling of the field solve (K assembly + Jacobi-PCG + Coulomb) on a replicated lattice (BASELINE config 4).
torchrun --nproc-per-node P tools/field_scaling.py [tiles=16] [order=lex] [solves=3]
Prints one JSON line on rank 0."""
import importlib, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"

def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    order = sys.argv[2] if len(sys.argv) > 2 else "lex"
    solves = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kmc = importlib.import_module(PKG); mg = importlib.import_module(PKG + ".multigpu"); syn = importlib.import_module(PKG + ".synthetic")
    s = syn.crossbar_standin(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"), tiles, tiles, order=order, Vd=5.0, rnd_seed=5)
    ctx = kmc.Context(local)
    c = ctx
    x, y, z = c.dev_d(s.x), c.dev_d(s.y), c.dev_d(s.z)
    element = c.dev_i(s.element); charge = c.empty_i(s.N, 0)
    pot_b = c.empty_d(s.N, 0.0); pot_c = c.empty_d(s.N, 0.0)
    neigh = c.compute_neighbor_list(x, y, z)
    n = s.N - s.N_left - s.N_right
    if world > 1:
        w = c.sparsity_K_row_counts(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
        counts, displs = mg.balanced_partition(w.cpu().numpy(), world)
    else:
        counts, displs = kmc.partition(n, world, aligned=True)
    cN, dN = kmc.partition(s.N, world)
    comm = mg.Comm(c, rank, world, n, counts, displs, dist if world > 1 else None)
    K = c.initialize_sparsity_K(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, int(displs[rank]), int(counts[rank]))
    comm.attach(K, dist if world > 1 else None)
    c.update_charge(element, charge, neigh, s.metals)
    def bar():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    res = []
    for it in range(solves + 1):
        pot_b.zero_()
        bar()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        c.assemble_K(K, s.N, s.N_left, s.N_right, element, charge, s.metals, s.Vd, s.high_G, s.low_G)
        e[1].record()
        iters = c.background_potential(K, s.N, s.N_left, s.N_right, element, charge, s.metals, s.Vd, s.high_G, s.low_G, pot_b)
        e[2].record()
        c.poisson_gridless(x, y, z, element, charge, s.sigma, s.k, pot_c, row_start=int(dN[rank]), row_count=int(cN[rank]))
        e[3].record()
        bar()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])], device="cuda", dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0: res.append((iters, t.cpu().tolist()))
    chk = float(pot_b.abs().sum().item())
    if rank == 0:
        iters = res[0][0]
        asm = np.mean([r[1][0] for r in res]); solve = np.mean([r[1][1] for r in res]); coul = np.mean([r[1][2] for r in res])
        nnz_local = K.nnz
        print("FIELD_SCALING " + json.dumps({"gpus": world, "N": s.N, "n": n, "order": order, "pcg_iterations": iters,
              "assemble_ms": asm, "assemble_plus_pcg_ms": solve, "ms_per_pcg_iteration": (solve - asm) / max(iters, 1),
              "coulomb_ms": coul, "field_solve_ms": solve + coul, "rank0_rows": int(counts[0]), "rank0_nnz": int(nnz_local),
              "recv_mask": comm.info()["recv_mask"], "pot_boundary_abs_sum_local": chk}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
