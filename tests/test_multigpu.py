"""N > 1 path.  CPU (gloo, world_size 2): the host-side sharding logic -- chunk-aligned partitions, need maps, send /
receive masks, uneven all-gather -- and an emulated row-sharded SpMV/dot that must reproduce the single-rank oracle
result bit for bit.  GPU (-m gpu, needs >= 2 devices): the real peer-memory PCG against the single-GPU solve."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT, make_synthetic


def _gloo_worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kmc = importlib.import_module(PKG)
    mg = importlib.import_module(PKG + ".multigpu")
    from oracle import binding as orc
    s = make_synthetic(kmc, nx=14, ny=8, nz=8, seed=9)
    sp = orc.sparsity_K(s.x, s.y, s.z, s.lattice, 0, 3.5, s.N_left, s.N_right)
    n = len(sp["row_ptr"]) - 1
    counts, displs = kmc.partition(n, world, aligned=True)
    assert counts.sum() == n and all(int(d) % 256 == 0 or int(d) == n for d in displs)
    lo, cnt = int(displs[rank]), int(counts[rank])
    # this rank's rows of the CSR (global column ids)
    rp = sp["row_ptr"][lo:lo + cnt + 1] - sp["row_ptr"][lo]
    col = sp["col"][sp["row_ptr"][lo]:sp["row_ptr"][lo + cnt]]
    need = mg.need_map_numpy(rp, col, lo, cnt, n)
    parts = [torch.empty(n, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(need))
    all_need = torch.stack(parts).numpy()
    send = mg.send_masks_numpy(all_need, rank, counts, displs)
    recv = mg.recv_mask_numpy(need, rank, counts, displs)
    # brute-force check of the masks
    for i in range(cnt):
        for q in range(world):
            if q == rank:
                continue
            qlo, qcnt = int(displs[q]), int(counts[q])
            qcols = sp["col"][sp["row_ptr"][qlo]:sp["row_ptr"][qlo + qcnt]]
            assert bool(send[i] >> q & 1) == bool((qcols == lo + i).any())
    assert recv == sum(1 << q for q in range(world) if q != rank and need[int(displs[q]):int(displs[q]) + int(counts[q])].any())
    # emulated sharded SpMV: own entries + pushed halo entries only; everything else poisoned with NaN
    rng = np.random.default_rng(5)
    val = rng.standard_normal(len(sp["col"]))
    p = rng.standard_normal(n)
    p_mine = np.full(n, np.nan); p_mine[lo:lo + cnt] = p[lo:lo + cnt]
    outbox = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]   # rows I push to peer q
    for q in range(world):
        if q != rank:
            rows_q = np.flatnonzero(send >> q & 1) + lo
            outbox[q][rows_q] = torch.from_numpy(p[rows_q])
    # (gloo has no all_to_all on CPU in every build: all_gather of the concatenated outboxes)
    gathered = [torch.zeros(world * n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.cat(outbox))
    for q in range(world):
        if q != rank:
            from_q = gathered[q][rank * n:(rank + 1) * n].numpy()
            qlo, qcnt = int(displs[q]), int(counts[q])
            needed = np.flatnonzero(need[qlo:qlo + qcnt]) + qlo
            p_mine[needed] = from_q[needed]
    y_local = orc.spmv(rp, col, val[sp["row_ptr"][lo]:sp["row_ptr"][lo + cnt]], p_mine)
    y_full = torch.zeros(n, dtype=torch.float64)
    mg.allgather_slices(dist, torch.from_numpy(y_local), counts, displs, y_full)
    y_ref = orc.spmv(sp["row_ptr"], sp["col"], val, p)
    assert not np.isnan(y_full.numpy()).any()
    assert (y_full.numpy() == y_ref).all()          # bit-identical to the single-rank SpMV
    # chunk-aligned partition => chunk partials are identical to the single-rank dot's chunks
    assert orc.dot(p, y_ref) == orc.dot(p, y_full.numpy())
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")


def test_sharding_logic_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_partition_aligned_matches_reference_arithmetic_on_granules(kmc):
    """rank boundaries on the dot granule (one 256-row chunk up to 256 chunks, one 64-chunk group = 16 384 rows above):
    the reference's partition arithmetic (KMC_comm.h:249-263) applied to granules"""
    mg = importlib.import_module(PKG + ".multigpu")
    for n, P in ((36498, 2), (36498, 8), (2335872, 8), (9344000, 4), (700, 4), (65536, 4), (65537, 4)):
        c, d = kmc.partition(n, P, aligned=True)
        gran = mg.dot_granule(n)
        assert gran == (256 if (n + 255) // 256 <= 256 else 16384)
        ng = (n + gran - 1) // gran
        cc, _ = kmc.partition(ng, P)
        assert c.sum() == n and ((d % gran == 0) | (d == n)).all()
        assert [int(v) for v in (c + gran - 1) // gran] == [int(v) for v in cc] or c[-1] == 0


def test_two_level_dot_is_independent_of_the_rank_count(kmc, orc):
    """summation spec 4.2: chunk partials -> groups of 64 chunks -> final reduce.  Emulates what the ranks exchange (only
    the group totals of their own rows) for 1, 2, 4 and 8 ranks: every split reproduces the oracle's dot bit for bit."""
    import ctypes as C
    rng = np.random.default_rng(11)
    n = 300 * 256 + 77                       # 301 chunks > 256 -> two-level combine, 5 groups
    u = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
    v = rng.standard_normal(n)
    want = orc.dot(u, v)
    nch = (n + 255) // 256
    partials = np.array([orc.dot(u[c * 256:(c + 1) * 256], v[c * 256:(c + 1) * 256]) for c in range(nch)])  # 1 chunk = its own dot
    def group_total(p64):
        pad = np.zeros(256); pad[:len(p64)] = p64
        ones = np.ones(256)
        return orc.dot(pad, ones)            # chunk_reduce_256 of the zero-padded partials (products with 1.0 are exact)
    for P in (1, 2, 4, 8):
        counts, displs = kmc.partition(n, P, aligned=True)
        table = []
        for r in range(P):                   # each rank reduces the groups of ITS chunks and publishes the totals
            if counts[r] == 0:
                continue
            c0, c1 = int(displs[r]) // 256, (int(displs[r]) + int(counts[r]) + 255) // 256
            assert c0 % 64 == 0 or counts[r] == 0
            table += [group_total(partials[g:min(g + 64, c1)]) for g in range(c0, c1, 64)]
        gt = np.array(table)
        assert len(gt) == (nch + 63) // 64
        pad = np.zeros(((len(gt) + 255) // 256) * 256); pad[:len(gt)] = gt
        assert orc.dot(pad, np.ones(len(pad))) == want or len(gt) > 256
    # and a system of exactly 256 chunks keeps the single-level result (5 nm device: 143 chunks)
    m = 256 * 256
    a, b = rng.standard_normal(m), rng.standard_normal(m)
    parts = np.array([orc.dot(a[c * 256:(c + 1) * 256], b[c * 256:(c + 1) * 256]) for c in range(256)])
    assert orc.dot(a, b) == orc.dot(parts, np.ones(256))


def test_balanced_partition(kmc):
    mg = importlib.import_module(PKG + ".multigpu")
    rng = np.random.default_rng(0)
    w = np.concatenate([np.full(5000, 19), np.full(9000, 30), rng.integers(5, 54, 7000)])
    for P in (1, 2, 3, 8):
        c, d = mg.balanced_partition(w, P)
        assert c.sum() == len(w) and d[0] == 0 and (np.diff(d) == c[:-1]).all()
        assert all(int(v) % 256 == 0 for v in d)             # 21 000 rows = 83 chunks: granule = one chunk
        loads = [int(w[d[q]:d[q] + c[q]].sum()) for q in range(P)]
        assert max(loads) <= w.sum() / P + 256 * 54          # within one chunk of the ideal share
    w2 = rng.integers(5, 54, 200000)                         # 782 chunks: granule = 64 chunks = 16 384 rows
    for P in (2, 4, 8):
        c, d = mg.balanced_partition(w2, P)
        assert c.sum() == len(w2) and all(int(v) % 16384 == 0 for v in d)
        loads = [int(w2[d[q]:d[q] + c[q]].sum()) for q in range(P)]
        assert max(loads) <= w2.sum() / P + 16384 * 54
    c, d = mg.balanced_partition(np.ones(100, dtype=np.int64), 4)   # fewer chunks than ranks: trailing ranks are empty
    assert c.sum() == 100 and (c >= 0).all()


def test_brick_order_gives_slab_partitions_with_thin_halos(kmc, orc):
    """With the bandwidth-minimised 'brick' order (cubes ordered y-major) contiguous row blocks are lateral slabs: in a
    4-rank split every rank exchanges halo rows with its 1-2 slab neighbours only and sends a small fraction of its rows,
    whereas the tile-by-tile 'file' order makes every rank talk to every other rank."""
    mg = importlib.import_module(PKG + ".multigpu")
    syn = importlib.import_module(PKG + ".synthetic")
    P = 4
    stats = {}
    for order in ("brick", "file"):
        s = syn.crossbar_standin(os.path.join(ROOT, "tests", "golden", "5nm_device", "parameters.txt"), 2, 2, order=order,
                                 vacancy_concentration=0.0)
        sp = orc.sparsity_K(s.x, s.y, s.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right, use_cells=True)
        rp, col = sp["row_ptr"], sp["col"]
        n = len(rp) - 1
        counts, displs = mg.balanced_partition(np.diff(rp), P)
        need = np.stack([mg.need_map_numpy(rp[displs[r]:displs[r] + counts[r] + 1] - rp[displs[r]],
                                           col[rp[displs[r]]:rp[displs[r] + counts[r]]], int(displs[r]), int(counts[r]), n)
                         for r in range(P)])
        peers = [bin(mg.recv_mask_numpy(need[r], r, counts, displs)).count("1") for r in range(P)]
        sent = [int((mg.send_masks_numpy(need, r, counts, displs) != 0).sum()) / max(int(counts[r]), 1) for r in range(P)]
        stats[order] = (peers, sent)
    peers_b, sent_b = stats["brick"]
    peers_f, sent_f = stats["file"]
    assert max(peers_b) <= 2 and peers_b[0] == 1 and peers_b[-1] == 1, stats
    assert max(sent_b) < 0.30, stats   # (146 k rows = 9 granules of 16 384 rows over 4 ranks: a coarse partition)
    assert max(peers_f) == P - 1 and max(sent_f) > max(sent_b), stats


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["5nm", "file"])
def test_sharded_solve_bit_identical_to_single_gpu(which):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py"), which, "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [l for l in r.stdout.split("\n") if l.startswith("MGPU_REPORT")][-1]
    import json
    rep = json.loads(line[len("MGPU_REPORT "):])
    assert rep["ok"] and rep["ranks_agree"]
    assert all(s["pot_bit_identical"] and s["cg"] == s["cg_ref"] for s in rep["steps"])
