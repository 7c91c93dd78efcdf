"""One rank of a row-sharded K solve driven through the C ABI ONLY (no torch.distributed / NCCL): bootstrap over the
library's file rendezvous, halo / dot exchanges and the potential all-gather over CUDA-IPC peer memory.  Launched twice by
tests/test_gpu_parity_large.py (both ranks may share one GPU).  Rank 0 compares the sharded result with the CPU ORACLE.
usage: mgpu_cabi_worker.py <rank> <size> <rendezvous dir> <device> <tiles>"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"


def main():
    rank, size, rdv_dir, device, tiles = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    kmc = importlib.import_module(PKG)
    syn = importlib.import_module(PKG + ".synthetic")
    api = kmc.api
    s = syn.crossbar_standin(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"), tiles, tiles, order="brick")
    ctx = kmc.Context(device)
    lib, chk = ctx.lib, api._check
    n = s.N - s.N_left - s.N_right
    counts, displs = kmc.partition(n, size, aligned=True)
    cN, dN = kmc.partition(s.N, size)
    # ---- bootstrap: rendezvous, exchange plan, IPC handles ------------------------------------------------------
    rdv = C.c_void_p()
    chk(lib.kmcb200_rdv_open(rdv_dir.encode(), rank, size, C.byref(rdv)))
    comm = C.c_void_p()
    chk(lib.kmcb200_comm_create_ex(ctx.h, rank, size, n, counts.ctypes.data_as(api._pi), displs.ctypes.data_as(api._pi),
                                   int(max(counts.max(), cN.max())), C.byref(comm)))
    mine, allh = np.zeros(64, np.uint8), np.zeros(64 * size, np.uint8)
    chk(lib.kmcb200_comm_ipc_handle(comm, mine.ctypes.data_as(C.c_void_p)))
    chk(lib.kmcb200_rdv_allgather(rdv, mine.ctypes.data_as(C.c_void_p), 64, allh.ctypes.data_as(C.c_void_p)))
    chk(lib.kmcb200_comm_open_peers(comm, allh.ctypes.data_as(C.c_void_p)))
    # ---- this rank's rows of K, halo need maps --------------------------------------------------------------------
    x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
    el, ch = ctx.dev_i(s.element), ctx.empty_i(s.N, 0)
    K = ctx.initialize_sparsity_K(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, int(displs[rank]), int(counts[rank]))
    chk(lib.kmcb200_kmat_attach_comm(K.h, comm))
    need = ctx.torch.zeros(n, dtype=ctx.torch.uint8, device=ctx.device)
    chk(lib.kmcb200_kmat_need_map(K.h, api._ptr(need)))
    need_h = need.cpu().numpy()
    all_need = np.zeros(n * size, np.uint8)
    chk(lib.kmcb200_rdv_allgather(rdv, need_h.ctypes.data_as(C.c_void_p), n, all_need.ctypes.data_as(C.c_void_p)))
    all_d = ctx.torch.from_numpy(all_need).to(ctx.device)
    chk(lib.kmcb200_comm_set_send_masks(comm, api._ptr(all_d)))
    chk(lib.kmcb200_rdv_barrier(rdv))
    # ---- two field solves (cold, then warm after a structural change), sharded ---------------------------------
    neigh = ctx.compute_neighbor_list(x, y, z)
    pot_b = ctx.empty_d(s.N, 0.0)
    pot_c = ctx.empty_d(s.N, 0.0)
    report = {"rank": rank, "size": size, "N": int(s.N), "rows": int(counts[rank]), "steps": []}
    osim = None
    if rank == 0:
        from oracle import binding as orc
        osp = orc.sparsity_K(s.x, s.y, s.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, use_cells=True)
        oneigh = orc.neighbor_list(s.x, s.y, s.z, 3.5, 52, use_cells=True)
        o_el, o_ch, o_x = s.element.copy(), np.zeros(s.N, np.int32), np.zeros(n)
    for step in range(2):
        if step == 1:   # move some vacancies: conductances change, the warm start is no longer a solution
            e = s.element.copy()
            vac = np.flatnonzero(e == kmc.VACANCY)[:200]
            e[vac] = kmc.O_EL
            e[np.flatnonzero(e == kmc.O_EL)[5000:5200]] = kmc.VACANCY
            el.copy_(ctx.dev_i(e))
            if rank == 0:
                o_el = e
        ctx.update_charge(el, ch, neigh, s.metals)
        it = ctx.background_potential(K, s.N, s.N_left, s.N_right, el, ch, s.metals, s.Vd, s.high_G, s.low_G, pot_b)
        interior = pot_b[s.N_left:s.N - s.N_right]
        chk(lib.kmcb200_comm_allgather(comm, api._ptr(interior), counts.ctypes.data_as(api._pi), displs.ctypes.data_as(api._pi)))
        ctx.poisson_gridless(x, y, z, el, ch, s.sigma, s.k, pot_c, row_start=int(dN[rank]), row_count=int(cN[rank]))
        chk(lib.kmcb200_comm_allgather(comm, api._ptr(pot_c), cN.ctypes.data_as(api._pi), dN.ctypes.data_as(api._pi)))
        rec = {"cg": it, "pot_sum": float(pot_b.abs().sum().item()), "coul_sum": float(pot_c.abs().sum().item())}
        if rank == 0:
            o_ch = orc.update_charge(o_el, o_ch, oneigh, s.metals)
            data, dinv, rhs = orc.assemble_K(s.N, s.N_left, s.N_right, o_el, o_ch, s.metals, osp, s.Vd, s.high_G, s.low_G)
            o_x, _, oit, _ = orc.pcg_jacobi(osp["row_ptr"], osp["col"], data, dinv, rhs, o_x, 1e-14 * n)
            want_c = orc.coulomb(s.x, s.y, s.z, o_el, o_ch, s.sigma, s.k, use_cells=True)
            got_b, got_c = interior.cpu().numpy(), pot_c.cpu().numpy()
            rec.update({"cg_oracle": int(oit), "pot_boundary_bit_identical_to_oracle": bool((got_b == o_x).all()),
                        "coulomb_max_rel_err": float(np.abs(got_c - want_c).max() / max(np.abs(want_c).max(), 1e-300))})
        report["steps"].append(rec)
    chk(lib.kmcb200_rdv_barrier(rdv))
    print("CABI_REPORT " + json.dumps(report), flush=True)
    K.close()
    lib.kmcb200_comm_destroy(comm)
    lib.kmcb200_rdv_close(rdv)


if __name__ == "__main__":
    main()
