"""Row-sharded multi-GPU driver: one process per GPU (torch.distributed for the bootstrap and the per-superstep
all-gathers; the PCG's halo pushes and dot-product exchanges run inside the CUDA kernels over NVLink peer memory).

What shards: K assembly + Jacobi-PCG (interior rows, boundaries multiples of 256) and the
Coulomb sum (site rows).  Charge update, rate list and the event loop are replicated on every rank (deterministic,
same RNG seed), so no per-event broadcast exists.  Reference counterpart: KMC_comm row partitions
(src/KMC_comm.h:245-391) + the MPI_Gatherv / MPI_Bcast of the potentials (src/kmc_main.cpp:367-384,411-427,
src/potential_solver_gpu.cu:1133-1142).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api
from .api import Context, DeviceKMC, KMatrix, Structure, _check, _ptr, partition


# ---- host-side plan logic (pure numpy; unit-tested on CPU with gloo) ---------------------------------------------
def need_map_numpy(row_ptr: np.ndarray, col: np.ndarray, row_start: int, rows: int, n_global: int) -> np.ndarray:
    """need[j] = 1 iff one of this rank's rows references global column j that it does not own"""
    need = np.zeros(n_global, dtype=np.uint8)
    c = col[: row_ptr[rows]]
    other = c[(c < row_start) | (c >= row_start + rows)]
    need[other] = 1
    return need


def send_masks_numpy(all_need: np.ndarray, rank: int, counts, displs) -> np.ndarray:
    """send_mask[i] bit q set iff rank q needs this rank's local row i"""
    lo, cnt = int(displs[rank]), int(counts[rank])
    m = np.zeros(cnt, dtype=np.uint8)
    for q in range(all_need.shape[0]):
        if q != rank:
            m |= (all_need[q, lo:lo + cnt].astype(np.uint8) << q).astype(np.uint8)
    return m


def recv_mask_numpy(my_need: np.ndarray, rank: int, counts, displs) -> int:
    mask = 0
    for q in range(len(counts)):
        if q != rank and my_need[int(displs[q]):int(displs[q]) + int(counts[q])].any():
            mask |= 1 << q
    return mask


def dot_granule(n_rows: int) -> int:
    """rows between allowed rank boundaries: one 256-row dot chunk, or one 64-chunk dot group (16 384 rows) for systems
    whose dot products are combined in two levels (more than 256 chunks), kmc_b200.h KMCB200_DOT_GROUP"""
    return 256 if (n_rows + 255) // 256 <= 256 else 64 * 256


def balanced_partition(row_weight: np.ndarray, nranks: int, chunk: int = 0):
    """Contiguous row blocks with boundaries on multiples of the dot granule and (nearly) equal total weight (non-zeros).
    Any granule-aligned partition gives bit-identical results; this one also balances the SpMV."""
    n = len(row_weight)
    chunk = chunk or dot_granule(n)
    nch = (n + chunk - 1) // chunk
    pad = np.zeros(nch * chunk, dtype=np.int64)
    pad[:n] = row_weight
    cw = np.concatenate([[0], np.cumsum(pad.reshape(nch, chunk).sum(1))])
    total = cw[-1]
    bounds = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        b = int(np.searchsorted(cw, target))
        if b > 0 and abs(cw[b - 1] - target) <= abs(cw[min(b, nch)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), nch))
    bounds.append(nch)
    displs = np.minimum(np.array(bounds[:-1], dtype=np.int64) * chunk, n)
    ends = np.minimum(np.array(bounds[1:], dtype=np.int64) * chunk, n)
    return (ends - displs).astype(np.int32), displs.astype(np.int32)


def allgather_slices(dist, local, counts, displs, out):
    """out[displs[q] : displs[q]+counts[q]] = rank q's `local` (uneven slices; equal-size padded all_gather)."""
    import torch
    world = len(counts)
    maxc = int(max(counts))
    send = torch.zeros(maxc, dtype=local.dtype, device=local.device)
    send[: local.numel()] = local
    parts = [torch.empty(maxc, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, send)
    for q in range(world):
        out[int(displs[q]): int(displs[q]) + int(counts[q])] = parts[q][: int(counts[q])]
    return out


# ---- device side --------------------------------------------------------------------------------------------------
class Comm:
    """kmcb200_comm handle: peer-memory exchange plan of one rank"""

    def __init__(self, ctx: Context, rank: int, size: int, n_global: int, counts, displs, dist=None, gather_capacity=0):
        import torch
        self.ctx, self.rank, self.size = ctx, rank, size
        self.counts = np.ascontiguousarray(counts, dtype=np.int32)
        self.displs = np.ascontiguousarray(displs, dtype=np.int32)
        self.n_global = n_global
        h = C.c_void_p()
        _check(ctx.lib.kmcb200_comm_create_ex(ctx.h, rank, size, n_global, self.counts.ctypes.data_as(api._pi),
                                              self.displs.ctypes.data_as(api._pi), int(gather_capacity), C.byref(h)))
        self.h = h
        if size > 1:
            handle = np.zeros(64, dtype=np.uint8)
            _check(ctx.lib.kmcb200_comm_ipc_handle(h, handle.ctypes.data_as(C.c_void_p)))
            mine = torch.from_numpy(handle).to(ctx.device)
            parts = [torch.empty(64, dtype=torch.uint8, device=ctx.device) for _ in range(size)]
            dist.all_gather(parts, mine)
            allh = np.ascontiguousarray(torch.stack(parts).cpu().numpy())
            _check(ctx.lib.kmcb200_comm_open_peers(h, allh.ctypes.data_as(C.c_void_p)))

    def attach(self, K: KMatrix, dist=None):
        import torch
        ctx = self.ctx
        _check(ctx.lib.kmcb200_kmat_attach_comm(K.h, self.h))
        if self.size > 1:
            need = torch.zeros(self.n_global, dtype=torch.uint8, device=ctx.device)
            _check(ctx.lib.kmcb200_kmat_need_map(K.h, _ptr(need)))
            parts = [torch.empty_like(need) for _ in range(self.size)]
            dist.all_gather(parts, need)
            all_need = torch.stack(parts).contiguous()
            _check(ctx.lib.kmcb200_comm_set_send_masks(self.h, _ptr(all_need)))
            dist.barrier()
        K._comm = self
        return self

    def allgather(self, vec, counts, displs):
        """in-place all-gather of row slices of a device vector over NVLink peer memory (kmcb200_comm_allgather)"""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        displs = np.ascontiguousarray(displs, dtype=np.int32)
        _check(self.ctx.lib.kmcb200_comm_allgather(self.h, _ptr(vec), counts.ctypes.data_as(api._pi),
                                                   displs.ctypes.data_as(api._pi)))

    def info(self):
        r, s, m, b = C.c_int(0), C.c_int(0), C.c_uint(0), C.c_longlong(0)
        _check(self.ctx.lib.kmcb200_comm_info(self.h, C.byref(r), C.byref(s), C.byref(m), C.byref(b)))
        return {"rank": r.value, "size": s.value, "recv_mask": m.value, "arena_bytes": b.value}

    def close(self):
        if self.h:
            self.ctx.lib.kmcb200_comm_destroy(self.h)
            self.h = None


class DistributedDeviceKMC(DeviceKMC):
    """DeviceKMC with the K solve and the Coulomb sum sharded over `world` ranks (strong scaling of one device)."""

    def __init__(self, s: Structure, ctx: Context, rank: int, world: int):
        import torch.distributed as dist
        self.dist, self.rank, self.world = dist, rank, world
        self.s, self.ctx = s, ctx
        c = ctx
        self.N = s.N
        self.x, self.y, self.z = c.dev_d(s.x), c.dev_d(s.y), c.dev_d(s.z)
        self.element = c.dev_i(s.element)
        self.charge = c.empty_i(self.N, 0)
        self.layer = c.dev_i(s.layer)
        self.pot_boundary = c.empty_d(self.N, 0.0)
        self.pot_charge = c.empty_d(self.N, 0.0)
        self.neigh = c.compute_neighbor_list(self.x, self.y, self.z)
        n = s.N - s.N_left - s.N_right
        if world > 1:   # nnz-balanced, chunk-aligned row blocks (every rank computes the same boundaries)
            w = c.sparsity_K_row_counts(self.x, self.y, self.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
            self.counts_K, self.displs_K = balanced_partition(w.cpu().numpy(), world)
        else:
            self.counts_K, self.displs_K = partition(n, world, aligned=True)
        self.counts_N, self.displs_N = partition(s.N, world)
        self.comm = Comm(c, rank, world, n, self.counts_K, self.displs_K, dist,
                         gather_capacity=int(max(self.counts_K.max(), self.counts_N.max())))
        self.K = c.initialize_sparsity_K(self.x, self.y, self.z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right,
                                         int(self.displs_K[rank]), int(self.counts_K[rank]))
        self.comm.attach(self.K, dist)
        self.ev = c.events_create(self.neigh)
        self.ev.set_activation_energies(s.E["E_gen"], s.E["E_rec"], s.E["E_Vdiff"], s.E["E_Odiff"])
        self.ev.rng_seed(api.RND_SEED_KMC)
        self.kmc_time = 0.0
        self.step_count = 0
        self.last_cg_iterations = 0
        self.last_n_events = 0

    def field_solve(self):
        s, c, r = self.s, self.ctx, self.rank
        c.update_charge(self.element, self.charge, self.neigh, s.metals)          # replicated (no Allgatherv)
        self.last_cg_iterations = c.background_potential(self.K, self.N, s.N_left, s.N_right, self.element, self.charge,
                                                         s.metals, s.Vd, s.high_G, s.low_G, self.pot_boundary)
        # potentials of the other ranks' rows: pulled out of the owners' arenas over NVLink (replaces the reference's
        # MPI_Gatherv + MPI_Bcast, src/kmc_main.cpp:367-384,411-427, src/potential_solver_gpu.cu:1133-1142)
        interior = self.pot_boundary[s.N_left: self.N - s.N_right]
        if self.world > 1:
            self.comm.allgather(interior, self.counts_K, self.displs_K)
        c.poisson_gridless(self.x, self.y, self.z, self.element, self.charge, s.sigma, s.k, self.pot_charge,
                           row_start=int(self.displs_N[r]), row_count=int(self.counts_N[r]))
        if self.world > 1:
            self.comm.allgather(self.pot_charge, self.counts_N, self.displs_N)
        c.sum_potential(self.pot_charge, self.pot_boundary)
