// pcg.cu -- a7: CSR SpMV with the p.Ap dot fused into its epilogue, fused Jacobi-PCG vector passes, the deterministic
// dot product of the summation spec, and their row-sharded multi-GPU form (comm.cuh).
// Reference: dist_iterative/dist_conjugate_gradient.cpp:149-276 (update order kept exactly),
// dist_iterative/dist_spmv_gpu_packing.cpp:106-228 (SpMV + halo), dist_iterative/utils_cg.cu:323-371 (elementwise).
// The reference runs per iteration: one rocsparse_spmv per neighbour block with pack/Isend/Irecv/unpack, 2 hipblasDdot
// (each a blocking host round trip + MPI_Allreduce), 3 daxpy, 1 dscal, 1 elementwise.  Here: 3 kernels per iteration
// (5 above 1024 chunks, where a 1-CTA kernel completes each dot), chained by programmatic dependent launch; all scalars
// stay on the device; on several GPUs the z halo and the dot contributions are delivered to the peers by the kernels
// that produce them, two exchanges per iteration, no fence or flag on the critical path (comm.cuh).
#include <stdlib.h>
#include <string.h>

#include "comm.cuh"
#include "kmat.cuh"
#include "pcg.cuh"

struct CgState {
    double bb, rz, rz_old, pAp, tol2, scalar_out;
    int k, max_it, done, iters, comm_error;  // comm_error: a peer-flag wait timed out (the solve is abandoned)
    unsigned cnt[8];
};

namespace {

constexpr int CH = KMCB200_CHUNK;  // 256 rows per CTA == dot chunk
constexpr int FIN_WIDE = 1024;     // threads of the 1-CTA dot finalize when it runs the two-level combine

// "last CTA done" election (threadFenceReduction pattern).  Returns true in every thread of the last CTA.
// remote: this CTA issued stores to peer memory that the elected CTA's flag must cover.
__device__ __forceinline__ bool last_cta(unsigned *counter, int *sm_flag, bool remote) {
    if (threadIdx.x == 0) {
        if (remote) __threadfence_system(); else __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        *sm_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    bool last = (*sm_flag != 0);
    if (last) __threadfence();
    return last;
}

// chunk partial of global chunk (chunk_start + local_chunk): stored locally; the rank's whole slice is pushed to the
// peers by the last CTA (exchange_partials), so ordinary CTAs issue no remote traffic and need no system fence
__device__ __forceinline__ void publish_partial(const CommDev &cm, int slot, int local_chunk, double v) {
    cm.partials[(size_t)slot * cm.nchunks_global + cm.chunk_start + local_chunk] = v;
}
// last CTA (all threads): copy this rank's partial slice(s) into every peer's array, raise the flag, wait for theirs
__device__ __forceinline__ void exchange_partials(const CommDev &cm, int slot0, int nslots, int local_chunks,
                                                  unsigned long long seq) {
    if (cm.size > 1) {
        for (int sidx = 0; sidx < nslots; ++sidx) {
            const size_t base = (size_t)(slot0 + sidx) * cm.nchunks_global + cm.chunk_start;
            for (int i0 = threadIdx.x; i0 < local_chunks; i0 += 8 * blockDim.x) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * blockDim.x;
                    v[u] = (i < local_chunks) ? __ldcg(cm.partials + base + i) : 0.0;
                }
                for (int q = 0; q < cm.size; ++q) {
                    if (q == cm.rank) continue;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * blockDim.x;
                        if (i < local_chunks) cm.peer_partials[q][base + i] = v[u];
                    }
                }
            }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < cm.size; ++q)
                if (q != cm.rank) kmc_store_relaxed_sys(cm.peer_flag_dot[q] + cm.rank, seq);
            kmc_wait_flags(cm.flag_dot, (1u << cm.size) - 1u, cm.rank, seq, cm.timeout_ns, cm.err);
            __threadfence_system();
        }
        __syncthreads();
    }
}
// Level 1 of the two-level dot combine (systems with more than 256 chunks): each warp reduces whole groups of 64 local
// chunk partials (group = chunk_reduce_256 of the 64 values padded with zeros = the two warp butterflies added, then six
// exact "+ 0.0"), stores the total in this rank's table of ALL group totals and in every peer's table.
__device__ __forceinline__ void reduce_and_push_groups(const CommDev &cm, int slot0, int nslots, int local_chunks) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int local_groups = (local_chunks + KMCB200_DOT_GROUP - 1) / KMCB200_DOT_GROUP;
    for (int sidx = 0; sidx < nslots; ++sidx) {
        const double *part = cm.partials + (size_t)(slot0 + sidx) * cm.nchunks_global + cm.chunk_start;
        const size_t gbase = (size_t)(slot0 + sidx) * cm.ngroups_global + cm.group_start;
        // two groups per trip: four loads in flight, four butterflies interleaved (the phase is latency, not work)
        for (int g = w; g < local_groups; g += 2 * nw) {
            const int g2 = g + nw;
            const int c0 = g * KMCB200_DOT_GROUP + lane, d0 = g2 * KMCB200_DOT_GROUP + lane;
            const bool two = g2 < local_groups;
            double v0 = (c0 < local_chunks) ? __ldcg(part + c0) : 0.0;
            double v1 = (c0 + 32 < local_chunks) ? __ldcg(part + c0 + 32) : 0.0;
            double u0 = (two && d0 < local_chunks) ? __ldcg(part + d0) : 0.0;
            double u1 = (two && d0 + 32 < local_chunks) ? __ldcg(part + d0 + 32) : 0.0;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const double a0 = __shfl_xor_sync(KMC_FULL_MASK, v0, off), a1 = __shfl_xor_sync(KMC_FULL_MASK, v1, off);
                const double b0 = __shfl_xor_sync(KMC_FULL_MASK, u0, off), b1 = __shfl_xor_sync(KMC_FULL_MASK, u1, off);
                v0 = v0 + a0; v1 = v1 + a1; u0 = u0 + b0; u1 = u1 + b1;
            }
            double tot = v0 + v1, tot2 = u0 + u1;
#pragma unroll
            for (int q = 0; q < 6; ++q) { tot = tot + 0.0; tot2 = tot2 + 0.0; }  // the six empty warps of chunk_reduce_256
            if (lane == 0) {
                cm.gtotals[gbase + g] = tot;
                if (two) cm.gtotals[gbase + g2] = tot2;
                for (int q = 0; q < cm.size; ++q)
                    if (q != cm.rank) {
                        cm.peer_gtotals[q][gbase + g] = tot;
                        if (two) cm.peer_gtotals[q][gbase + g2] = tot2;
                    }
            }
        }
    }
}
// final_reduce when the CTA is wider than the spec's 256 threads: threads 0..255 reduce, every thread takes the barriers
__device__ __forceinline__ double final_reduce_wide(const double *partials, long long n, double *sm) {
    if (blockDim.x == CH) return kmc_final_reduce(partials, n, sm);
    double acc = 0.0;
    if (threadIdx.x < CH)
        for (long long k = threadIdx.x; k < n; k += CH) acc = acc + __ldcg(partials + k);
    acc = kmc_warp_xor_sum(acc);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < 8) sm[w] = acc;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
        r = sm[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) r = r + sm[q];
    }
    __syncthreads();
    return r;
}
// completes `nslots` dot products at once: exchange (chunk partials for small systems, group totals for large ones), one
// flag round trip, then the fixed-order reduction.  All threads of a 256-thread CTA; out[] valid in thread 0.
// ---- the same completion with the fence-free exchange (CommDev::ll_mode): this rank's contributions (chunk partials of a
// single-level system, group totals otherwise) are stored as {lo, seq, hi, seq} cells into every peer's cell array, then
// the cells of all other ranks' index ranges are polled out of the local cell array into the plain table the fixed-order
// reduction reads.  No flag, no fence: a cell is valid as soon as both halves carry this dot's sequence number.
__device__ __forceinline__ void finish_dots_ll(const CommDev &cm, int slot0, int nslots, int local_chunks,
                                               unsigned long long seq, double *red, double *out) {
    const bool grp = cm.group_chunks != 0;
    const unsigned s32 = (unsigned)seq;
    const int nglob = cm.ll_vals;
    const int nloc = grp ? (local_chunks + KMCB200_DOT_GROUP - 1) / KMCB200_DOT_GROUP : local_chunks;
    const int first = grp ? cm.group_start : cm.chunk_start;
    double *table = grp ? cm.gtotals : cm.partials;
    const int tstride = grp ? cm.ngroups_global : cm.nchunks_global;
    const size_t cell0 = (size_t)(seq & 1ull) * 2u * (size_t)nglob;
    if (grp) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int sidx = 0; sidx < nslots; ++sidx) {
            const double *part = cm.partials + (size_t)(slot0 + sidx) * cm.nchunks_global + cm.chunk_start;
            // two groups per trip: four loads in flight, four butterflies interleaved (the phase is latency, not work)
            for (int g = w; g < nloc; g += 2 * nw) {
                const int g2 = g + nw;
                const int c0 = g * KMCB200_DOT_GROUP + lane, d0 = g2 * KMCB200_DOT_GROUP + lane;
                const bool two = g2 < nloc;
                double v0 = (c0 < local_chunks) ? __ldcg(part + c0) : 0.0;
                double v1 = (c0 + 32 < local_chunks) ? __ldcg(part + c0 + 32) : 0.0;
                double u0 = (two && d0 < local_chunks) ? __ldcg(part + d0) : 0.0;
                double u1 = (two && d0 + 32 < local_chunks) ? __ldcg(part + d0 + 32) : 0.0;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const double a0 = __shfl_xor_sync(KMC_FULL_MASK, v0, off), a1 = __shfl_xor_sync(KMC_FULL_MASK, v1, off);
                    const double b0 = __shfl_xor_sync(KMC_FULL_MASK, u0, off), b1 = __shfl_xor_sync(KMC_FULL_MASK, u1, off);
                    v0 = v0 + a0; v1 = v1 + a1; u0 = u0 + b0; u1 = u1 + b1;
                }
                double tot = v0 + v1, tot2 = u0 + u1;
#pragma unroll
                for (int q = 0; q < 6; ++q) { tot = tot + 0.0; tot2 = tot2 + 0.0; }  // the six empty warps of chunk_reduce_256
                // lanes 0..size-1 deliver: lane == rank keeps the plain copy, every other lane serves one peer
                if (lane < cm.size) {
                    if (lane == cm.rank) {
                        table[(size_t)(slot0 + sidx) * tstride + first + g] = tot;
                        if (two) table[(size_t)(slot0 + sidx) * tstride + first + g2] = tot2;
                    } else {
                        uint4 *cells = cm.peer_ll[lane] + cell0 + (size_t)sidx * nglob + first;
                        kmc_ll_store(cells + g, tot, s32);
                        if (two) kmc_ll_store(cells + g2, tot2, s32);
                    }
                }
            }
        }
    } else {
        for (int sidx = 0; sidx < nslots; ++sidx) {
            const double *part = cm.partials + (size_t)(slot0 + sidx) * cm.nchunks_global + cm.chunk_start;
            for (int i = threadIdx.x; i < nloc; i += blockDim.x) {
                const double v = __ldcg(part + i);
                for (int q = 0; q < cm.size; ++q)
                    if (q != cm.rank) kmc_ll_store(cm.peer_ll[q] + cell0 + (size_t)sidx * nglob + first + i, v, s32);
            }
        }
    }
    // collect the other ranks' cells
    unsigned long long deadline = 0;
    for (int sidx = 0; sidx < nslots; ++sidx) {
        const uint4 *cells = cm.ll + cell0 + (size_t)sidx * nglob;
        double *trow = table + (size_t)(slot0 + sidx) * tstride;
        for (int idx = threadIdx.x; idx < nglob; idx += blockDim.x) {
            if (idx >= first && idx < first + nloc) continue;
            double v = 0.0;
            unsigned spins = 0;
            while (!kmc_ll_try_load(cells + idx, s32, &v)) {
                if ((++spins & 1023u) == 0) {
                    const unsigned long long now = kmc_globaltimer_ns();
                    if (deadline == 0) deadline = now + cm.timeout_ns;
                    else if (now > deadline || (cm.err && *(volatile int *)cm.err)) {
                        if (cm.err) *(volatile int *)cm.err = 1;
                        break;
                    }
                }
            }
            trow[idx] = v;
        }
    }
    __syncthreads();
    for (int sidx = 0; sidx < nslots; ++sidx) {
        const double *trow = table + (size_t)(slot0 + sidx) * tstride;
        out[sidx] = grp ? final_reduce_wide(trow, nglob, red) : kmc_final_reduce(trow, nglob, red);
    }
}
__device__ __forceinline__ void finish_dots(const CommDev &cm, int slot0, int nslots, int local_chunks,
                                            unsigned long long seq, double *red, double *out) {
    if (cm.ll_mode) {
        finish_dots_ll(cm, slot0, nslots, local_chunks, seq, red, out);
        return;
    }
    if (cm.group_chunks == 0) {
        exchange_partials(cm, slot0, nslots, local_chunks, seq);
        for (int sidx = 0; sidx < nslots; ++sidx)
            out[sidx] = kmc_final_reduce(cm.partials + (size_t)(slot0 + sidx) * cm.nchunks_global, cm.nchunks_global, red);
        return;
    }
    reduce_and_push_groups(cm, slot0, nslots, local_chunks);
    if (cm.size > 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < cm.size; ++q)
                if (q != cm.rank) kmc_store_relaxed_sys(cm.peer_flag_dot[q] + cm.rank, seq);
            kmc_wait_flags(cm.flag_dot, (1u << cm.size) - 1u, cm.rank, seq, cm.timeout_ns, cm.err);
            __threadfence_system();
        }
    } else {
        __threadfence();
    }
    __syncthreads();
    for (int sidx = 0; sidx < nslots; ++sidx)
        out[sidx] = final_reduce_wide(cm.gtotals + (size_t)(slot0 + sidx) * cm.ngroups_global, cm.ngroups_global, red);
}
// z halo (CommDev::ll_mode): the kernel that computes z[i] stores it into the z_full of every peer whose matrix block
// references global row g.  Returns whether a remote store was issued (the thread fences those before the kernel ends).
__device__ __forceinline__ bool push_z_halo(const CommDev &cm, int i, int g, double zi) {
    unsigned m = cm.send_mask[i];
    const bool any = m != 0;
    while (m) {
        const int q = __ffs(m) - 1;
        m &= m - 1;
        cm.peer_z_full[q][g] = zi;
    }
    return any;
}

// ---- completion of a dot product: exchange with the peers, reduce in the fixed order, update the PCG state.  Runs either
// in the last CTA of the producing kernel (small problems: saves a launch) or in the 1-CTA dot_finalize_kernel (large
// problems: the producing kernel's CTAs then need no fence / atomic at all).
__device__ __forceinline__ void finish_pap(const CommDev &cm, int local_chunks, unsigned long long seq, CgState *st,
                                           double *red) {
    double tot;
    finish_dots(cm, 0, 1, local_chunks, seq, red, &tot);
    if (threadIdx.x == 0) {
        st->pAp = tot;
        if (cm.err && *(volatile int *)cm.err) st->done = 1;
    }
}
__device__ __forceinline__ void finish_rz(const CommDev &cm, int local_chunks, unsigned long long seq, CgState *st,
                                          double *red) {
    double rz;
    finish_dots(cm, 1, 1, local_chunks, seq, red, &rz);
    if (threadIdx.x == 0) {
        st->rz_old = st->rz;
        st->rz = rz;
        int k = st->k + 1;
        st->k = k;
        st->iters = k - 1;
        st->done = !(rz / st->bb > st->tol2 && k <= st->max_it) || (cm.err && *(volatile int *)cm.err);
    }
}
__device__ __forceinline__ void finish_init(const CommDev &cm, int local_chunks, unsigned long long seq, CgState *st,
                                            double *red) {
    double o[2];
    finish_dots(cm, 2, 2, local_chunks, seq, red, o);
    if (threadIdx.x == 0) {
        const double bb = o[0], rz = o[1];
        st->bb = bb;
        st->rz = rz;
        st->rz_old = 0.0;
        st->k = 1;
        st->iters = 0;
        st->done = !(rz / bb > st->tol2 && 1 <= st->max_it) || (cm.err && *(volatile int *)cm.err);
    }
}
// kind: 0 = p.Ap, 1 = r.z (+ iteration bookkeeping), 2 = setup (b.b and r.z)
// (launched with FIN_WIDE threads for the two-level combine -- 32 warps for the group level -- and CH otherwise)
__global__ void __launch_bounds__(FIN_WIDE) dot_finalize_kernel(CommDev cm, int kind, int local_chunks, unsigned long long seq,
                                                         CgState *__restrict__ st) {
    __shared__ double red[8];
    kmc_pdl_trigger();
    kmc_pdl_wait();
    if (kind != 2 && st->done) return;
    if (kind == 0) finish_pap(cm, local_chunks, seq, st, red);
    else if (kind == 1) finish_rz(cm, local_chunks, seq, st, red);
    else finish_init(cm, local_chunks, seq, st, red);
}

// y = A x (x indexed by GLOBAL column), optional fused partial of  x[row].y[row]  (p.Ap)
// Row reduction spec: L lanes per row, lane l accumulates entries l, l+L, ... with fma in increasing k,
// then a butterfly over the L lanes.
// FUSE: the last CTA completes the dot product (small problems).  Keeping that cold path out of the FUSE=false
// instantiation matters: its register pressure would otherwise cap the occupancy of the hot loop (48-64 vs 32 regs).
template <int L, bool DOT, int DEPTH, bool FUSE>
__global__ void __launch_bounds__(CH) spmv_kernel(int rows, const int *__restrict__ row_ptr,
                                                 const int *__restrict__ col, const double *__restrict__ val,
                                                 const double *__restrict__ xg, double *__restrict__ y, CommDev cm,
                                                 unsigned long long dot_seq, CgState *__restrict__ st) {
    kmc_pdl_trigger();
    kmc_pdl_wait();
    if (DOT && st->done) return;
    __shared__ double prod[DOT ? CH : 1];
    __shared__ double red[8];
    __shared__ int flag;
    (void)flag;
    constexpr int GROUPS = CH / L;  // rows per pass
    const int lane = threadIdx.x % L;
    const int grp = threadIdx.x / L;
    const int row0 = blockIdx.x * CH;
#pragma unroll 1
    for (int pass = 0; pass < L; ++pass) {
        int rl = pass * GROUPS + grp;
        int r = row0 + rl;
        double acc = 0.0;
        if (r < rows) {
            const int s = row_ptr[r], e = row_ptr[r + 1];
            // all of this lane's entries (k = s+lane, s+lane+L, ...) are loaded before the FMA chain starts, so a lane
            // keeps up to DEPTH (val, col, x) triples in flight instead of one; the FMA order is unchanged.
            for (int k0 = s + lane; k0 < e; k0 += L * DEPTH) {
                double v[DEPTH], xv[DEPTH];
                int c[DEPTH];
#pragma unroll
                for (int t = 0; t < DEPTH; ++t) {
                    const int k = k0 + t * L;
                    const bool ok = k < e;
                    v[t] = ok ? __ldcs(val + k) : 0.0;
                    c[t] = ok ? __ldcs(col + k) : -1;
                }
#pragma unroll
                for (int t = 0; t < DEPTH; ++t) xv[t] = (c[t] >= 0) ? __ldg(xg + c[t]) : 0.0;
#pragma unroll
                for (int t = 0; t < DEPTH; ++t)
                    if (c[t] >= 0) acc = fma(v[t], xv[t], acc);
            }
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(KMC_FULL_MASK, acc, off);
        if (lane == 0) {
            double pv = 0.0;
            if (r < rows) {
                y[r] = acc;
                if (DOT) pv = xg[cm.row_start + r] * acc;
            }
            if (DOT) prod[rl] = pv;
        }
    }
    if (DOT) {
        __syncthreads();
        double c = kmc_chunk_reduce_256(prod[threadIdx.x], red);
        if (threadIdx.x == 0) publish_partial(cm, 0, blockIdx.x, c);
        if (FUSE) {
            if (last_cta(&st->cnt[0], &flag, false)) {
                finish_pap(cm, gridDim.x, dot_seq, st, red);
                if (threadIdx.x == 0) st->cnt[0] = 0;
            }
        }
    }
}

// Shared-memory staged variant (single GPU, opt-in): the chunk's unique columns (u_col) are gathered ONCE into shared
// memory, every non-zero then reads x through its 16-bit chunk-local id.  Same row reduction spec.
template <int L, bool DOT>
__global__ void __launch_bounds__(CH) spmv_staged_kernel(int rows, const int *__restrict__ row_ptr,
                                                        const unsigned short *__restrict__ lcol,
                                                        const double *__restrict__ val,
                                                        const int *__restrict__ u_ptr, const int *__restrict__ u_col,
                                                        const double *__restrict__ xg, double *__restrict__ y,
                                                        CommDev cm, CgState *__restrict__ st) {
    if (DOT && st->done) return;
    extern __shared__ double xs[];  // the chunk's x window
    __shared__ double prod[CH];
    __shared__ double red[8];
    __shared__ int flag;
    constexpr int GROUPS = CH / L;
    const int lane = threadIdx.x % L;
    const int grp = threadIdx.x / L;
    const int row0 = blockIdx.x * CH;
    {
        const int ub = u_ptr[blockIdx.x], nu = u_ptr[blockIdx.x + 1] - ub;
        for (int i = threadIdx.x; i < nu; i += CH) xs[i] = __ldg(xg + __ldcs(u_col + ub + i));
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < L; ++pass) {
        int rl = pass * GROUPS + grp;
        int r = row0 + rl;
        double acc = 0.0;
        if (r < rows) {
            const int s = row_ptr[r], e = row_ptr[r + 1];
            constexpr int DEPTH = 7;
            for (int k0 = s + lane; k0 < e; k0 += L * DEPTH) {
                double v[DEPTH];
                int c[DEPTH];
#pragma unroll
                for (int t = 0; t < DEPTH; ++t) {
                    const int k = k0 + t * L;
                    const bool ok = k < e;
                    v[t] = ok ? __ldcs(val + k) : 0.0;
                    c[t] = ok ? (int)__ldcs(lcol + k) : -1;
                }
#pragma unroll
                for (int t = 0; t < DEPTH; ++t)
                    if (c[t] >= 0) acc = fma(v[t], xs[c[t]], acc);
            }
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(KMC_FULL_MASK, acc, off);
        if (lane == 0) {
            double pv = 0.0;
            if (r < rows) {
                y[r] = acc;
                if (DOT) pv = xg[cm.row_start + r] * acc;
            }
            if (DOT) prod[rl] = pv;
        }
    }
    if (DOT) {
        __syncthreads();
        double c = kmc_chunk_reduce_256(prod[threadIdx.x], red);
        if (threadIdx.x == 0) publish_partial(cm, 0, blockIdx.x, c);
        if (last_cta(&st->cnt[0], &flag, false)) {
            double tot = kmc_final_reduce(cm.partials, cm.nchunks_global, red);
            if (threadIdx.x == 0) {
                st->pAp = tot;
                st->cnt[0] = 0;
            }
        }
    }
}

// The three vector kernels below are persistent-style: a CTA walks chunks c = blockIdx.x, blockIdx.x + gridDim.x, ...
// (one 256-row dot chunk per trip), so there is one "last CTA" election per CTA instead of one per chunk.

// r = b - A x0 ; z = M^-1 r ; bb = b.b ; rz = r.z     (dist_conjugate_gradient.cpp:187-213)
template <bool FUSE>
__global__ void __launch_bounds__(CH) cg_init_kernel(int rows, int nchunks, double *__restrict__ r,
                                                    const double *__restrict__ Ap, const double *__restrict__ dinv,
                                                    double *__restrict__ z, CommDev cm,
                                                    unsigned long long dot_seq, CgState *__restrict__ st) {
    __shared__ double red[8];
    __shared__ int flag;
    const bool zh = cm.ll_mode != 0;
    bool remote = false;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int i = c * CH + threadIdx.x;
        double vbb = 0.0, vrz = 0.0;
        if (i < rows) {
            double b = __ldcs(r + i);
            double ri = b - __ldcs(Ap + i);
            double zi = ri * __ldcs(dinv + i);
            __stcs(r + i, ri);
            __stcs(z + i, zi);
            if (zh) remote |= push_z_halo(cm, i, cm.row_start + i, zi);
            vbb = b * b;
            vrz = ri * zi;
        }
        double cbb = kmc_chunk_reduce_256(vbb, red);
        double crz = kmc_chunk_reduce_256(vrz, red);
        if (threadIdx.x == 0) {
            publish_partial(cm, 2, c, cbb);
            publish_partial(cm, 3, c, crz);
        }
    }
    if (remote) __threadfence_system();  // z halo stores performed at the peers before this rank's dot cells can leave
    if (FUSE) {
        if (zh) __syncthreads();
        if (last_cta(&st->cnt[1], &flag, false)) {
            finish_init(cm, nchunks, dot_seq, st, red);
            if (threadIdx.x == 0) st->cnt[1] = 0;
        }
    }
}

// MODE 1: p = z (k == 1)  or  p = z + (rz/rz_old) p      (dist_conjugate_gradient.cpp:218-227)
// MODE 0: p = src (the copy of x0 at :178)
// The new entries are written to this rank's p vector and, for rows a peer's block references, straight into the
// peer's p vector (halo push).  The last CTA raises the halo flag at the peers and waits for theirs, so when this kernel
// has finished every halo entry this rank needs has arrived and the SpMV that follows needs no synchronisation.
template <int MODE>
__global__ void __launch_bounds__(CH) cg_pupdate_kernel(int rows, int nchunks, const double *__restrict__ src,
                                                       const double *__restrict__ p_old, double *__restrict__ p_new,
                                                       CommDev cm, int buf, unsigned long long halo_seq,
                                                       CgState *__restrict__ st) {
    kmc_pdl_trigger();
    kmc_pdl_wait();
    if (MODE == 1 && st->done) return;
    __shared__ int flag;
    const bool first = (MODE == 0) || (st->k == 1);
    const double beta = first ? 0.0 : st->rz / st->rz_old;
    // ll_mode: p is not exchanged inside the loop -- the halo entries of p are formed here from the z halo the peers
    // delivered (same operands, same operations as on the owner)
    const bool zh = (MODE == 1) && cm.ll_mode != 0;
    int remote = 0;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int i = c * CH + threadIdx.x;
        if (i < rows) {
            const int g = cm.row_start + i;
            double v;
            // streaming (evict-first) reads: only the new p must stay L2 resident for the SpMV gathers
            if (first) {
                v = __ldcs(src + i);
            } else {
                double t = beta * __ldcs(p_old + g);
                v = __ldcs(src + i) + t;
            }
            p_new[g] = v;
            if (cm.size > 1 && !zh) {
                unsigned m = cm.send_mask[i];
                remote |= (m != 0);
                while (m) {
                    int q = __ffs(m) - 1;
                    m &= m - 1;
                    cm.peer_p_full[q][buf][g] = v;
                }
            }
        }
    }
    if (zh) {
        for (int h = blockIdx.x * CH + threadIdx.x; h < cm.nhalo; h += gridDim.x * CH) {
            const int g = __ldg(cm.halo_rows + h);
            double v;
            if (first) {
                v = __ldcs(cm.z_full + g);
            } else {
                double t = beta * __ldcs(p_old + g);
                v = __ldcs(cm.z_full + g) + t;
            }
            p_new[g] = v;
        }
    }
    if (cm.size > 1 && !zh) {
        remote = __syncthreads_or(remote);
        if (last_cta(&st->cnt[4], &flag, remote != 0)) {
            if (threadIdx.x == 0) {
                __threadfence_system();
                for (int q = 0; q < cm.size; ++q)
                    if (q != cm.rank) kmc_store_relaxed_sys(cm.peer_flag_halo[q] + cm.rank, halo_seq);
                if (!kmc_wait_flags(cm.flag_halo, cm.recv_mask, cm.rank, halo_seq, cm.timeout_ns, cm.err)) st->done = 1;
                __threadfence_system();
                st->cnt[4] = 0;
            }
        }
    }
}

// a = rz / p.Ap ; x += a p ; r -= a Ap ; z = M^-1 r ; rz' = r.z ; k++   (dist_conjugate_gradient.cpp:243-266)
template <bool FUSE>
__global__ void __launch_bounds__(CH) cg_update_kernel(int rows, int nchunks, const double *__restrict__ p_full,
                                                      const double *__restrict__ Ap, const double *__restrict__ dinv,
                                                      double *__restrict__ x, double *__restrict__ r,
                                                      double *__restrict__ z, CommDev cm,
                                                      unsigned long long dot_seq, CgState *__restrict__ st) {
    kmc_pdl_trigger();
    kmc_pdl_wait();
    if (st->done) return;
    __shared__ double red[8];
    __shared__ int flag;
    const double a = st->rz / st->pAp;
    const double na = -a;
    const bool zh = cm.ll_mode != 0;
    bool remote = false;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int i = c * CH + threadIdx.x;
        double v = 0.0;
        if (i < rows) {
            double pi = p_full[cm.row_start + i];
            double xi = fma(a, pi, __ldcs(x + i));
            double ri = fma(na, __ldcs(Ap + i), __ldcs(r + i));
            double zi = ri * __ldcs(dinv + i);
            __stcs(x + i, xi);
            __stcs(r + i, ri);
            __stcs(z + i, zi);
            if (zh) remote |= push_z_halo(cm, i, cm.row_start + i, zi);
            v = ri * zi;
        }
        double cv = kmc_chunk_reduce_256(v, red);
        if (threadIdx.x == 0) publish_partial(cm, 1, c, cv);
    }
    if (remote) __threadfence_system();  // z halo stores performed at the peers before this rank's dot cells can leave
    if (FUSE) {
        if (zh) __syncthreads();
        if (last_cta(&st->cnt[2], &flag, false)) {
            finish_rz(cm, nchunks, dot_seq, st, red);
            if (threadIdx.x == 0) st->cnt[2] = 0;
        }
    }
}

// ================================ the whole PCG loop as ONE persistent cooperative kernel =============================
// (opt-in.)  grid = as many 256-thread CTAs as are co-resident; a CTA walks the dot chunks c = blockIdx.x,
// blockIdx.x + gridDim.x, ... in every phase.  Per iteration: p update (+ halo push) | grid barrier (+ halo flags) |
// SpMV + p.Ap chunk partials | grid barrier (+ dot flags) | x, r, z update + r.z chunk partials | grid barrier (+ dot
// flags).  The arithmetic, its order and the summation spec are those of the kernels above (same device functions), so
// the iterates are bit-identical to the multi-kernel path and to the oracle.
//  * two-level dots: the CTA that publishes the LAST chunk partial of a 64-chunk group (per-group counter) reduces the
//    group, stores the total and pushes it to the peers; after the barrier EVERY CTA reduces the group totals itself, so
//    alpha / beta / the convergence test are computed redundantly and identically everywhere -- no scalar broadcast.
//  * every wait (grid barrier, peer flags) is bounded; a time-out raises CgState::comm_error and all CTAs leave.
//  * the barrier's gpu-scope fence invalidates L1, so p entries written by other CTAs / peers are re-read from L2; the
//    gathers use plain loads (not ld.global.nc: the vector changes during the kernel).
struct PcgLoopArgs {
    int rows, nchunks;
    const int *row_ptr, *col;
    const double *val, *dinv;
    double *x, *r, *z, *Ap;
    CgState *st;
    unsigned long long *bar;   // grid barrier counter (zeroed by the host before the launch)
    unsigned long long *work;  // chunk tickets of the SpMV phase (zeroed by the host before the launch)
    unsigned long long *gdone; // group totals stored so far (monotone; zeroed by the host before the launch)
    unsigned long long halo_seq0, dot_seq0;
    unsigned long long *prof;  // KMCB200_PCG_LOOP_PROFILE: 8 accumulated phase times (ns) of CTA 0, else NULL
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// All threads of all CTAs call; returns false (in every thread of the CTA) when the wait timed out / the solve was aborted.
// Two-level arrival (32 sub-counters, the last arrival of each goes on to the top counter) so that the arrivals do not
// serialise on one L2 address; everybody spins on the top counter.  bar[0] = top, bar[16 + j] = sub-counter j (both
// monotone over the whole solve).
constexpr int BAR_SUB = 32;
__device__ __forceinline__ bool grid_barrier(unsigned long long *bar, unsigned long long &gen, const CommDev &cm, int *sm_ok) {
    __syncthreads();
    ++gen;
    if (threadIdx.x == 0) {
        const unsigned nsub = gridDim.x < BAR_SUB ? gridDim.x : BAR_SUB;
        const unsigned j = blockIdx.x % nsub;
        const unsigned members = gridDim.x / nsub + (j < gridDim.x % nsub ? 1u : 0u);
        __threadfence();
        if (atomicAdd(bar + 16 + j, 1ull) + 1ull == gen * members) atomicAdd(bar, 1ull);
        const unsigned long long target = gen * nsub;
        int ok = 1;
        unsigned spins = 0;
        unsigned long long deadline = 0;
        while (ld_acquire_gpu_u64(bar) < target) {
            if ((++spins & 255u) == 0) {
                const unsigned long long now = kmc_globaltimer_ns();
                if (deadline == 0) deadline = now + cm.timeout_ns;
                else if (now > deadline || *(volatile int *)cm.err) { *(volatile int *)cm.err = 1; ok = 0; break; }
            }
        }
        __threadfence();
        *sm_ok = ok;
    }
    __syncthreads();
    return *sm_ok != 0;
}
// after a grid barrier: CTA 0 tells the peers that this rank's data of step `seq` is complete; every CTA waits for the
// peers named in `mask`
__device__ __forceinline__ bool peer_sync(const CommDev &cm, unsigned long long *const *peer_flags,
                                          const unsigned long long *my_flags, unsigned mask, unsigned long long seq,
                                          int *sm_ok) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int q = 0; q < cm.size; ++q)
                if (q != cm.rank) kmc_store_relaxed_sys(peer_flags[q] + cm.rank, seq);
        }
        const bool ok = kmc_wait_flags(my_flags, mask, cm.rank, seq, cm.timeout_ns, cm.err);
        __threadfence_system();
        *sm_ok = ok ? 1 : 0;
    }
    __syncthreads();
    return *sm_ok != 0;
}
// chunk partial of slot `slot`, published by thread 0 (no fence, no atomic in the hot phase: the grid barrier orders it)
__device__ __forceinline__ void publish_chunk(const CommDev &cm, int slot, int c, double v) {
    if (threadIdx.x == 0) {
        const size_t idx = (size_t)slot * cm.nchunks_global + cm.chunk_start + c;
        cm.partials[idx] = v;
        if (cm.group_chunks == 0 && cm.size > 1) {
            for (int q = 0; q < cm.size; ++q)
                if (q != cm.rank) cm.peer_partials[q][idx] = v;
            __threadfence_system();
        }
    }
}
// Group level of the two-level dot, after the grid barrier that made every chunk partial visible: warp w of CTA b reduces
// group 8 b + w (spec: chunk_reduce_256 of the 64 partials padded with zeros), stores the total locally and at the peers and
// counts it in; then EVERY CTA waits until all local groups are in (gdone is monotone: `expect` = groups x dot steps).
__device__ __forceinline__ bool group_stage(const CommDev &cm, const PcgLoopArgs &a, int slot, int local_groups,
                                            unsigned long long expect, int *sm_ok) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = blockIdx.x * (CH / 32) + w; g < local_groups; g += gridDim.x * (CH / 32)) {
        const double *part = cm.partials + (size_t)slot * cm.nchunks_global + cm.chunk_start;
        const int c0 = g * KMCB200_DOT_GROUP + lane;
        const double v0 = (c0 < a.nchunks) ? __ldcg(part + c0) : 0.0;
        const double v1 = (c0 + 32 < a.nchunks) ? __ldcg(part + c0 + 32) : 0.0;
        double tot = kmc_warp_xor_sum(v0);
        tot = tot + kmc_warp_xor_sum(v1);
#pragma unroll
        for (int q = 0; q < 6; ++q) tot = tot + 0.0;  // the six empty warps of chunk_reduce_256
        if (lane == 0) {
            const size_t gi = (size_t)slot * cm.ngroups_global + cm.group_start + g;
            cm.gtotals[gi] = tot;
            if (cm.size > 1) {
                for (int q = 0; q < cm.size; ++q)
                    if (q != cm.rank) cm.peer_gtotals[q][gi] = tot;
                __threadfence_system();
            } else {
                __threadfence();
            }
            atomicAdd(a.gdone, 1ull);
        }
    }
    if (threadIdx.x == 0) {
        int ok = 1;
        unsigned spins = 0;
        unsigned long long deadline = 0;
        while (ld_acquire_gpu_u64(a.gdone) < expect) {
            if ((++spins & 255u) == 0) {
                const unsigned long long now = kmc_globaltimer_ns();
                if (deadline == 0) deadline = now + cm.timeout_ns;
                else if (now > deadline || *(volatile int *)cm.err) { *(volatile int *)cm.err = 1; ok = 0; break; }
            }
        }
        *sm_ok = ok;
    }
    __syncthreads();
    return *sm_ok != 0;
}
// the full dot product of slot `slot`, in every thread of the CTA (fixed-order reduction over all ranks' contributions)
__device__ __forceinline__ double dot_total(const CommDev &cm, int slot, double *red, double *sm_bcast) {
    double t;
    if (cm.group_chunks) t = kmc_final_reduce(cm.gtotals + (size_t)slot * cm.ngroups_global, cm.ngroups_global, red);
    else t = kmc_final_reduce(cm.partials + (size_t)slot * cm.nchunks_global, cm.nchunks_global, red);
    if (threadIdx.x == 0) *sm_bcast = t;
    __syncthreads();
    t = *sm_bcast;
    __syncthreads();
    return t;
}

// One 256-row chunk of y = A p with the products p[row] * y[row] left in prod[] (row reduction spec of spmv_kernel).
// Kept out of line: the loop then gets the 32-register allocation of the stand-alone SpMV (8 CTAs per SM) whatever the
// persistent kernel keeps alive around it.
template <int L, int DEPTH>
__device__ __noinline__ void spmv_chunk_rows(const int *__restrict__ row_ptr, const int *__restrict__ col,
                                             const double *__restrict__ val, const double *p, double *__restrict__ Ap,
                                             int rows, int row0, int row_start, double *prod) {
    constexpr int GROUPS = CH / L;
    const int lane = threadIdx.x % L, grp = threadIdx.x / L;
#pragma unroll 1
    for (int pass = 0; pass < L; ++pass) {
        const int rl = pass * GROUPS + grp;
        const int r = row0 + rl;
        double acc = 0.0;
        if (r < rows) {
            const int s = __ldg(row_ptr + r), e = __ldg(row_ptr + r + 1);
            for (int k0 = s + lane; k0 < e; k0 += L * DEPTH) {
                double v[DEPTH], xv[DEPTH];
                int cc[DEPTH];
#pragma unroll
                for (int t = 0; t < DEPTH; ++t) {
                    const int kk = k0 + t * L;
                    const bool in = kk < e;
                    v[t] = in ? __ldcs(val + kk) : 0.0;
                    cc[t] = in ? __ldcs(col + kk) : -1;
                }
#pragma unroll
                for (int t = 0; t < DEPTH; ++t) xv[t] = (cc[t] >= 0) ? p[cc[t]] : 0.0;
#pragma unroll
                for (int t = 0; t < DEPTH; ++t)
                    if (cc[t] >= 0) acc = fma(v[t], xv[t], acc);
            }
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(KMC_FULL_MASK, acc, off);
        if (lane == 0) {
            double pv = 0.0;
            if (r < rows) {
                __stcs(Ap + r, acc);
                pv = p[row_start + r] * acc;
            }
            prod[rl] = pv;
        }
    }
}

template <int L, int DEPTH>
__global__ void __launch_bounds__(CH, 8) pcg_loop_kernel(PcgLoopArgs a, CommDev cm) {
    __shared__ double prod[CH];
    __shared__ double red[8];
    __shared__ double bcast;
    __shared__ int okf, s_chunk;
    long long nxt = 0;
    unsigned long long tbase = 0;
    CgState *st = a.st;
    if (st->done) return;  // converged at setup (uniform: every CTA reads the same word)
    const int local_groups = cm.group_chunks ? (a.nchunks + KMCB200_DOT_GROUP - 1) / KMCB200_DOT_GROUP : 0;
    const double bb = st->bb, tol2 = st->tol2;
    const int max_it = st->max_it;
    double rz = st->rz, rz_old = st->rz_old, pAp = 0.0;
    int k = st->k;
    unsigned long long target = 0, gsteps = 0, hs = a.halo_seq0, ds = a.dot_seq0;
    const unsigned all_peers = (cm.size >= 32 ? 0xffffffffu : ((1u << cm.size) - 1u));
    bool ok = true;
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long tp = prof ? kmc_globaltimer_ns() : 0;
#define PCG_TICK(q) do { if (prof) { const unsigned long long n_ = kmc_globaltimer_ns(); a.prof[q] += n_ - tp; tp = n_; } } while (0)
    while (true) {
        // ---------------- p = z (k == 1)  or  p = z + (rz / rz_old) p ; halo push -----------------------------------
        ++hs;
        const int nb = (int)(hs & 1);
        const double *p_old = cm.p_full[nb ^ 1];
        double *p_new = cm.p_full[nb];
        const bool first = (k == 1);
        const double beta = first ? 0.0 : rz / rz_old;
        for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
            const int i = c * CH + threadIdx.x;
            if (i < a.rows) {
                const int g = cm.row_start + i;
                double v;
                // streaming (evict-first) accesses for everything but p: only the new p must stay L2 resident for the gathers
                if (first) {
                    v = __ldcs(a.z + i);
                } else {
                    const double t = beta * __ldcs(p_old + g);
                    v = __ldcs(a.z + i) + t;
                }
                p_new[g] = v;
                if (cm.size > 1) {
                    unsigned m = cm.send_mask[i];
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        cm.peer_p_full[q][nb][g] = v;
                    }
                }
            }
        }
        if (cm.size > 1) __threadfence_system();
        PCG_TICK(0);
        ok = grid_barrier(a.bar, target, cm, &okf);
        if (ok && cm.size > 1) ok = peer_sync(cm, cm.peer_flag_halo, cm.flag_halo, cm.recv_mask, hs, &okf);
        if (!ok) break;
        PCG_TICK(1);
        // ---------------- Ap = A p ; p.Ap ---------------------------------------------------------------------------
        ++ds;
        // chunks are handed out by ticket (the hardware scheduler's balancing, kept): each CTA takes tickets until it draws
        // one past the end, so an iteration consumes exactly nchunks + gridDim.x of them; the next ticket is drawn while
        // the current chunk is processed
        if (threadIdx.x == 0) nxt = (long long)atomicAdd(a.work, 1ull) - (long long)tbase;
        while (true) {
            if (threadIdx.x == 0) s_chunk = (int)(nxt < (long long)a.nchunks ? nxt : (long long)a.nchunks);
            __syncthreads();
            const int c = s_chunk;
            if (c >= a.nchunks) break;
            if (threadIdx.x == 0) nxt = (long long)atomicAdd(a.work, 1ull) - (long long)tbase;
            spmv_chunk_rows<L, DEPTH>(a.row_ptr, a.col, a.val, p_new, a.Ap, a.rows, c * CH, cm.row_start, prod);
            __syncthreads();
            const double cpart = kmc_chunk_reduce_256(prod[threadIdx.x], red);
            publish_chunk(cm, 0, c, cpart);
        }
        tbase += (unsigned long long)a.nchunks + gridDim.x;
        PCG_TICK(2);
        ok = grid_barrier(a.bar, target, cm, &okf);
        if (ok && local_groups) ok = group_stage(cm, a, 0, local_groups, (gsteps += (unsigned long long)local_groups), &okf);
        if (ok && cm.size > 1) ok = peer_sync(cm, cm.peer_flag_dot, cm.flag_dot, all_peers, ds, &okf);
        if (!ok) break;
        PCG_TICK(3);
        pAp = dot_total(cm, 0, red, &bcast);
        PCG_TICK(4);
        // ---------------- x += a p ; r -= a Ap ; z = M^-1 r ; r.z ---------------------------------------------------
        ++ds;
        const double al = rz / pAp, nal = -al;
        for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
            const int i = c * CH + threadIdx.x;
            double v = 0.0;
            if (i < a.rows) {
                const double pi = p_new[cm.row_start + i];
                const double xi = fma(al, pi, __ldcs(a.x + i));
                const double ri = fma(nal, __ldcs(a.Ap + i), __ldcs(a.r + i));
                const double zi = ri * __ldcs(a.dinv + i);
                __stcs(a.x + i, xi);
                __stcs(a.r + i, ri);
                __stcs(a.z + i, zi);
                v = ri * zi;
            }
            const double cpart = kmc_chunk_reduce_256(v, red);
            publish_chunk(cm, 1, c, cpart);
        }
        PCG_TICK(5);
        ok = grid_barrier(a.bar, target, cm, &okf);
        if (ok && local_groups) ok = group_stage(cm, a, 1, local_groups, (gsteps += (unsigned long long)local_groups), &okf);
        if (ok && cm.size > 1) ok = peer_sync(cm, cm.peer_flag_dot, cm.flag_dot, all_peers, ds, &okf);
        if (!ok) break;
        PCG_TICK(6);
        rz_old = rz;
        rz = dot_total(cm, 1, red, &bcast);
        PCG_TICK(7);
        ++k;
        if (!(rz / bb > tol2 && k <= max_it)) break;
    }
#undef PCG_TICK
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->rz = rz;
        st->rz_old = rz_old;
        st->pAp = pAp;
        st->k = k;
        st->iters = k - 1;
        st->done = 1;
    }
}

// ---- split-sparse extension (reference dist_iterative/dist_spmv_split_sparse.cpp:5-79, spmm_split_sparse1):
// Ap += scatter(T_tunnel * gather(p)).  One warp per local tunnel row; row reduction spec with 32 lanes (lane l takes
// entries l, l+32, ... with fma in increasing k, then the xor butterfly 16..1); the gather reads p through the
// global-indexed p vector, so no packed copy of the tunnel entries is needed.
__global__ void __launch_bounds__(256) tunnel_spmv_kernel(TunnelDev t, const double *__restrict__ pg, double *__restrict__ Ap,
                                                         const CgState *__restrict__ st, int check_done) {
    if (check_done && st->done) return;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= t.n_local) return;
    const int s = t.row_ptr[r], e = t.row_ptr[r + 1];
    double acc = 0.0;
    for (int k0 = s + lane; k0 < e; k0 += 128) {  // 4 (val, col, x) triples in flight per lane; FMA order unchanged
        double v[4], xv[4];
        int c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u;
            const bool ok = k < e;
            v[u] = ok ? __ldcs(t.val + k) : 0.0;
            c[u] = ok ? __ldcs(t.col + k) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = (c[u] >= 0) ? __ldg(pg + __ldg(t.rows_global + c[u])) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c[u] >= 0) acc = fma(v[u], xv[u], acc);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(KMC_FULL_MASK, acc, off);
    if (lane == 0) {
        const int lr = t.rows_local[r];
        Ap[lr] = Ap[lr] + acc;  // unpack_add (utils_cg.cu)
    }
}

// p.Ap when the SpMV could not fuse it (the tunnel part is added by a second kernel)
template <bool FUSE>
__global__ void __launch_bounds__(CH) cg_pap_kernel(int rows, int nchunks, const double *__restrict__ p_full,
                                                   const double *__restrict__ Ap, CommDev cm, unsigned long long dot_seq,
                                                   CgState *__restrict__ st) {
    if (st->done) return;
    __shared__ double red[8];
    __shared__ int flag;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int i = c * CH + threadIdx.x;
        double v = 0.0;
        if (i < rows) v = p_full[cm.row_start + i] * Ap[i];
        const double cv = kmc_chunk_reduce_256(v, red);
        if (threadIdx.x == 0) publish_partial(cm, 0, c, cv);
    }
    if (FUSE) {
        if (last_cta(&st->cnt[5], &flag, false)) {
            finish_pap(cm, nchunks, dot_seq, st, red);
            if (threadIdx.x == 0) st->cnt[5] = 0;
        }
    }
}

__global__ void __launch_bounds__(CH) dot_kernel(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                double *__restrict__ partials, CgState *__restrict__ st) {
    __shared__ double red[8];
    __shared__ int flag;
    long long i = (long long)blockIdx.x * CH + threadIdx.x;
    double p = (i < n) ? u[i] * v[i] : 0.0;
    double c = kmc_chunk_reduce_256(p, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = c;
    if (last_cta(&st->cnt[3], &flag, false)) {
        const int nchunks = (int)gridDim.x;
        double tot;
        if (nchunks <= 256) {
            tot = kmc_final_reduce(partials, nchunks, red);
        } else {  // two-level combine of the summation spec: group totals behind the partials
            double *gt = partials + nchunks;
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            const int ngroups = (nchunks + KMCB200_DOT_GROUP - 1) / KMCB200_DOT_GROUP;
            for (int g = w; g < ngroups; g += CH / 32) {
                const int c0 = g * KMCB200_DOT_GROUP + lane;
                const double v0 = (c0 < nchunks) ? __ldcg(partials + c0) : 0.0;
                const double v1 = (c0 + 32 < nchunks) ? __ldcg(partials + c0 + 32) : 0.0;
                double t = kmc_warp_xor_sum(v0);
                t = t + kmc_warp_xor_sum(v1);
#pragma unroll
                for (int q = 0; q < 6; ++q) t = t + 0.0;
                if (lane == 0) gt[g] = t;
            }
            __threadfence();
            __syncthreads();
            tot = kmc_final_reduce(gt, ngroups, red);
        }
        if (threadIdx.x == 0) {
            st->scalar_out = tot;
            st->cnt[3] = 0;
        }
    }
}

int ensure_cg_workspace(kmcb200_ctx *ctx, long long nchunks) {
    if (!ctx->cg_state) {
        KMC_CUDA(cudaMalloc(&ctx->cg_state, sizeof(CgState)));
        KMC_CUDA(cudaMemsetAsync(ctx->cg_state, 0, sizeof(CgState), ctx->stream));
    }
    size_t need = (size_t)(2 * nchunks + 16);
    if (ctx->partials_cap < need) {
        if (ctx->partials) {
            KMC_CUDA(cudaStreamSynchronize(ctx->stream));
            KMC_CUDA(cudaFree(ctx->partials));
        }
        KMC_CUDA(cudaMalloc(&ctx->partials, need * sizeof(double)));
        ctx->partials_cap = need;
    }
    return 0;
}

}  // namespace

// y = A xg for this rank's rows; xg is indexed by global row and must already hold the halo entries (or be awaited
// through halo_seq).  with_dot: fused p.Ap into CgState::pAp.
// fuse_final: the SpMV's last CTA completes the dot product itself (small problems); otherwise the caller launches
// dot_finalize_kernel afterwards.
static int spmv_launch(kmcb200_ctx *ctx, kmcb200_kmat *K, const double *xg, double *y, bool with_dot, int fuse_final,
                       unsigned long long dot_seq, bool pdl = false) {
    constexpr int L = KMCB200_SPMV_LANES;
    const CommDev &cm = K->comm->dev;
    unsigned blocks = (unsigned)((K->rows + CH - 1) / CH);
    kmc_count_launch();
    const size_t dyn = (size_t)K->plan_max_unique * sizeof(double);
    // The staged kernel only pays off when the rows of a chunk share columns (reuse >> 1); for the DeviceKMC lattices
    // the measured reuse is 1.3-1.7 and it is SLOWER (371 us vs 197 us at 62 M nnz), so it is opt-in.
    if (K->plan_max_unique > 0 && dyn <= 200 * 1024 && cm.size == 1) {
        if (dyn > ctx->smem_cfg_staged) {
            KMC_CUDA(cudaFuncSetAttribute(spmv_staged_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            KMC_CUDA(cudaFuncSetAttribute(spmv_staged_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            ctx->smem_cfg_staged = dyn;
        }
        if (with_dot)
            spmv_staged_kernel<L, true><<<blocks, CH, dyn, ctx->stream>>>(K->rows, K->row_ptr, K->lcol, K->val, K->u_ptr,
                                                                         K->u_col, xg, y, cm, ctx->cg_state);
        else
            spmv_staged_kernel<L, false><<<blocks, CH, dyn, ctx->stream>>>(K->rows, K->row_ptr, K->lcol, K->val, K->u_ptr,
                                                                          K->u_col, xg, y, cm, ctx->cg_state);
    } else {
        static const int depth = getenv("KMCB200_SPMV_DEPTH") ? atoi(getenv("KMCB200_SPMV_DEPTH")) : 4;
#define KMC_SPMV_GO(DOTV, D, F)                                                                                      \
    KMC_CUDA(kmc_launch_pdl(spmv_kernel<L, DOTV, D, F>, blocks, CH, 0, ctx->stream, pdl, K->rows, (const int *)K->row_ptr, \
                            (const int *)K->col, (const double *)K->val, xg, y, cm, dot_seq, ctx->cg_state))
        if (with_dot && fuse_final) {
            if (depth <= 4) KMC_SPMV_GO(true, 4, true); else KMC_SPMV_GO(true, 7, true);
        } else if (with_dot) {
            if (depth <= 4) KMC_SPMV_GO(true, 4, false); else KMC_SPMV_GO(true, 7, false);
        } else {
            if (depth <= 4) KMC_SPMV_GO(false, 4, false); else KMC_SPMV_GO(false, 7, false);
        }
#undef KMC_SPMV_GO
    }
    KMC_CUDA(cudaGetLastError());
    return 0;
}

// copies this rank's x entries into the global-indexed p vector and pushes the halo rows to the peers
static int push_vector(kmcb200_ctx *ctx, kmcb200_kmat *K, const double *x_local, int *buf_out,
                       unsigned long long *halo_seq_out) {
    kmcb200_comm *C = K->comm;
    const unsigned long long hs = ++C->halo_seq;
    const int buf = (int)(hs & 1);
    const int nch = (K->rows + CH - 1) / CH;
    const unsigned eb = (unsigned)(nch < ctx->sm_count * 8 ? (nch > 0 ? nch : 1) : ctx->sm_count * 8);
    kmc_count_launch();
    cg_pupdate_kernel<0><<<eb, CH, 0, ctx->stream>>>(K->rows, nch, x_local, nullptr, C->dev.p_full[buf], C->dev, buf, hs,
                                                    ctx->cg_state);
    KMC_CUDA(cudaGetLastError());
    *buf_out = buf;
    *halo_seq_out = hs;
    return 0;
}

extern "C" int kmcb200_spmv(kmcb200_ctx *ctx, kmcb200_kmat *K, const double *x_local, double *y_local) {
    KMC_CHECK_ARG(ctx && K && x_local && y_local, "null pointer");
    KMC_CHECK_ARG(K->comm != nullptr, "kmat has no exchange plan");
    KMC_TRY(ensure_cg_workspace(ctx, 1));
    static const bool force_dot = getenv("KMCB200_SPMV_FORCE_DOT") != nullptr;  // diagnostics: time the dot variant
    if (K->comm->size == 1) return spmv_launch(ctx, K, x_local, y_local, force_dot, 0, 0);
    if (!(K->comm->peers_open && K->comm->masks_set)) {
        kmc_set_error("row-sharded SpMV: kmcb200_comm_open_peers and kmcb200_comm_set_send_masks must be called first");
        return KMCB200_E_COMM;
    }
    int buf;
    unsigned long long hs;
    KMC_TRY(push_vector(ctx, K, x_local, &buf, &hs));
    (void)hs;
    return spmv_launch(ctx, K, K->comm->dev.p_full[buf], y_local, false, 1, 0);
}

// y = A x and x.(A x) in one pass: the fused kernel the PCG iteration uses (single rank; x_local is the full vector).
// The scalar is left on the device (CgState::pAp) unless dot_host != NULL (then: host sync).
extern "C" int kmcb200_spmv_dot(kmcb200_ctx *ctx, kmcb200_kmat *K, const double *x_local, double *y_local,
                                double *dot_host) {
    KMC_CHECK_ARG(ctx && K && x_local && y_local, "null pointer");
    KMC_CHECK_ARG(K->comm != nullptr && K->comm->size == 1, "kmcb200_spmv_dot is a single-rank call");
    const int nchunks = (K->rows + CH - 1) / CH;
    KMC_TRY(ensure_cg_workspace(ctx, nchunks));
    CgState *st = ctx->cg_state;  // `done` is 0 outside a PCG solve ...
    if (ctx->cg_done_stale) {     // ... unless one returned early with an error
        KMC_CUDA(cudaMemsetAsync(&st->done, 0, sizeof(int), ctx->stream));
        ctx->cg_done_stale = false;
    }
    const int fuse = nchunks <= 1024 ? 1 : 0;
    const unsigned long long ds = ++K->comm->dot_seq;
    KMC_TRY(spmv_launch(ctx, K, x_local, y_local, true, fuse, ds));
    if (!fuse) {
        kmc_count_launch();
        dot_finalize_kernel<<<1, K->comm->dev.group_chunks ? FIN_WIDE : CH, 0, ctx->stream>>>(K->comm->dev, 0, nchunks, ds, st);
        KMC_CUDA(cudaGetLastError());
    }
    if (dot_host) {
        KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, &st->pAp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        *dot_host = *(double *)ctx->h_mail;
    }
    return 0;
}

extern "C" int kmcb200_dot(kmcb200_ctx *ctx, const double *u, const double *v, long long n, double *result_host) {
    KMC_CHECK_ARG(ctx && u && v && result_host && n >= 0, "arguments");
    long long nchunks = (n + CH - 1) / CH;
    if (nchunks == 0) nchunks = 1;
    KMC_TRY(ensure_cg_workspace(ctx, nchunks));
    kmc_count_launch();
    dot_kernel<<<(unsigned)nchunks, CH, 0, ctx->stream>>>(n, u, v, ctx->partials, ctx->cg_state);
    KMC_CUDA(cudaGetLastError());
    KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, &ctx->cg_state->scalar_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *result_host = *(double *)ctx->h_mail;
    return 0;
}

static int tunnel_launch(kmcb200_ctx *ctx, const TunnelDev &t, const double *pg, double *Ap, int check_done) {
    if (t.n_local <= 0) return 0;
    kmc_count_launch();
    tunnel_spmv_kernel<<<(unsigned)((t.n_local + 7) / 8), 256, 0, ctx->stream>>>(t, pg, Ap, ctx->cg_state, check_done);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

// y = T_neighbor x + scatter(T_tunnel gather(x)) on one rank (spmm_split_sparse1)
int kmc_split_spmv(kmcb200_ctx *ctx, kmcb200_kmat *K, const TunnelDev *tun, const double *x_local, double *y_local) {
    KMC_CHECK_ARG(K->comm != nullptr && K->comm->size == 1, "split-sparse SpMV: single-rank call");
    KMC_TRY(ensure_cg_workspace(ctx, 1));
    KMC_TRY(spmv_launch(ctx, K, x_local, y_local, false, 0, 0));
    if (tun) KMC_TRY(tunnel_launch(ctx, *tun, x_local, y_local, 0));
    return 0;
}

extern "C" int kmcb200_pcg_jacobi(kmcb200_ctx *ctx, kmcb200_kmat *K, double *r_local, double *x_local,
                                  const double *diag_inv_local, double relative_tolerance, int max_iterations,
                                  int *iterations_host) {
    return kmc_pcg_run(ctx, K, nullptr, r_local, x_local, diag_inv_local, relative_tolerance, max_iterations,
                       iterations_host);
}

// Jacobi-PCG; tun != NULL: the operator is T_neighbor + tunnel sub-block (conjugate_gradient_jacobi_split_sparse,
// dist_iterative/dist_conjugate_gradient_split_sparse.cpp:18-166), same update order.
int kmc_pcg_run(kmcb200_ctx *ctx, kmcb200_kmat *K, const TunnelDev *tun, double *r_local, double *x_local,
                const double *diag_inv_local, double relative_tolerance, int max_iterations, int *iterations_host) {
    KMC_CHECK_ARG(ctx && K && r_local && x_local && diag_inv_local, "null pointer");
    KMC_CHECK_ARG(K->comm != nullptr, "kmat has no exchange plan");
    kmcb200_comm *C = K->comm;
    KMC_CHECK_ARG(tun == nullptr || C->size == 1, "split-sparse PCG with a tunnel block: single-rank call");
    if (C->size > 1 && !(C->peers_open && C->masks_set)) {
        kmc_set_error("row-sharded PCG: kmcb200_comm_open_peers and kmcb200_comm_set_send_masks must be called first");
        return KMCB200_E_COMM;
    }
    const int rows = K->rows;
    const unsigned nchunks = (unsigned)((rows + CH - 1) / CH);
    KMC_TRY(ensure_cg_workspace(ctx, nchunks));
    CgState *st = ctx->cg_state;
    {   // bounded peer waits (comm.cuh): error word in the solver state, timeout from the environment (default 20 s)
        static const double tmo_ms = getenv("KMCB200_COMM_TIMEOUT_MS") ? atof(getenv("KMCB200_COMM_TIMEOUT_MS")) : 20000.0;
        C->dev.err = &st->comm_error;
        C->dev.timeout_ns = (unsigned long long)(tmo_ms * 1e6);
    }
    // host-initialised part of the state (counters stay 0 between launches)
    CgState *h = (CgState *)ctx->h_mail;
    memset(h, 0, sizeof(CgState));
    h->tol2 = relative_tolerance * relative_tolerance;  // dist_conjugate_gradient.cpp:217
    h->max_it = max_iterations;
    h->done = 0;
    ctx->cg_done_stale = true;  // until the normal exit below
    KMC_CUDA(cudaMemcpyAsync(st, h, sizeof(CgState), cudaMemcpyHostToDevice, ctx->stream));
    // A*x0 (:191), residual + preconditioned residual + both setup dots (:187-213)
    int buf = 0;
    unsigned long long hs = 0;
    if (C->size == 1) {
        KMC_TRY(spmv_launch(ctx, K, x_local, K->Ap, false, 1, 0));
        if (tun) KMC_TRY(tunnel_launch(ctx, *tun, x_local, K->Ap, 0));
        hs = C->halo_seq;
        buf = (int)(hs & 1);
    } else {
        KMC_TRY(push_vector(ctx, K, x_local, &buf, &hs));
        KMC_TRY(spmv_launch(ctx, K, C->dev.p_full[buf], K->Ap, false, 1, 0));
    }
    kmc_count_launch();
    const unsigned eb = (unsigned)((int)nchunks < ctx->sm_count * 8 ? (nchunks > 0 ? nchunks : 1) : ctx->sm_count * 8);
    // Large problems: the dot products are completed by a separate 1-CTA kernel, so the thousands of producing CTAs need
    // no fence + atomic election (measured: 13-26 % of the SpMV at >= 62 M non-zeros).  Small problems keep the fused
    // last-CTA completion and save the launches.
    const int fuse = nchunks <= 1024 ? 1 : 0;
    {
        const unsigned long long ds = ++C->dot_seq;
        if (fuse) cg_init_kernel<true><<<eb, CH, 0, ctx->stream>>>(rows, (int)nchunks, r_local, K->Ap, diag_inv_local, K->z, C->dev, ds, st);
        else cg_init_kernel<false><<<eb, CH, 0, ctx->stream>>>(rows, (int)nchunks, r_local, K->Ap, diag_inv_local, K->z, C->dev, ds, st);
        if (!fuse) { kmc_count_launch(); dot_finalize_kernel<<<1, C->dev.group_chunks ? FIN_WIDE : CH, 0, ctx->stream>>>(C->dev, 2, (int)nchunks, ds, st); }
    }
    KMC_CUDA(cudaGetLastError());
    // ---- the iterations: the multi-kernel loop below (default), or ONE persistent cooperative kernel
    // (KMCB200_PCG_PERSISTENT=1; not for the tunnel operator).  Measured on one B200 at 2.3 M rows: 254 vs 226 us per
    // iteration -- see DESIGN.md 6 / profiles/r2_pcg_multigpu.md: the persistent form is slower at every size measured.
    const bool want_persist = getenv("KMCB200_PCG_PERSISTENT") != nullptr && atoi(getenv("KMCB200_PCG_PERSISTENT")) != 0 &&
                              getenv("KMCB200_PCG_PROFILE") == nullptr;  // (read per call: the tests switch it)
    const bool persistent = !tun && want_persist;
    if (persistent) {
        constexpr int L = KMCB200_SPMV_LANES;
        if (ctx->pcg_loop_occ == 0) {
            int occ = 0;
            KMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_loop_kernel<L, 4>, CH, 0));
            ctx->pcg_loop_occ = occ > 0 ? occ : -1;
        }
        if (ctx->pcg_loop_occ < 0) { kmc_set_error("pcg_loop_kernel does not fit on an SM"); return KMCB200_E_CUDA; }
        const size_t ws = 64 * sizeof(unsigned long long);  // [0] barrier top, [1] tickets, [2] groups done, [8..15] profile, [16..47] barrier sub-counters
        if (ctx->pcg_loop_ws_bytes < ws) {
            if (ctx->pcg_loop_ws) { KMC_CUDA(cudaStreamSynchronize(ctx->stream)); KMC_CUDA(cudaFree(ctx->pcg_loop_ws)); }
            KMC_CUDA(cudaMalloc(&ctx->pcg_loop_ws, ws));
            ctx->pcg_loop_ws_bytes = ws;
        }
        KMC_CUDA(cudaMemsetAsync(ctx->pcg_loop_ws, 0, ws, ctx->stream));
        // as many CTAs as are co-resident, trimmed so that every CTA walks the same number of chunks (+- 1)
        const unsigned max_grid = (unsigned)(ctx->pcg_loop_occ * ctx->sm_count);
        const unsigned per = (nchunks + max_grid - 1) / max_grid;
        const unsigned grid = per ? (nchunks + per - 1) / per : 1;
        PcgLoopArgs a;
        a.rows = rows; a.nchunks = (int)nchunks;
        a.row_ptr = K->row_ptr; a.col = K->col; a.val = K->val; a.dinv = diag_inv_local;
        a.x = x_local; a.r = r_local; a.z = K->z; a.Ap = K->Ap;
        a.st = st;
        a.bar = (unsigned long long *)ctx->pcg_loop_ws;
        a.work = (unsigned long long *)ctx->pcg_loop_ws + 1;
        a.gdone = (unsigned long long *)ctx->pcg_loop_ws + 2;
        static const bool loop_prof = getenv("KMCB200_PCG_LOOP_PROFILE") != nullptr;
        a.prof = loop_prof ? (unsigned long long *)((char *)ctx->pcg_loop_ws + 64) : nullptr;
        a.halo_seq0 = C->halo_seq; a.dot_seq0 = C->dot_seq;
        CommDev cmv = C->dev;
        void *kargs[] = {&a, &cmv};
        kmc_count_launch();
        KMC_CUDA(cudaLaunchCooperativeKernel((const void *)pcg_loop_kernel<L, 4>, dim3(grid ? grid : 1), dim3(CH), kargs, 0, ctx->stream));
    }
    int *h_flags = (int *)((char *)ctx->h_mail + 512);
    auto read_flags = [&]() -> int {
        KMC_CUDA(cudaMemcpyAsync(h_flags, &st->k, 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_flags[4]) {  // CgState::comm_error
            kmc_set_error("row-sharded PCG: rank %d waited more than %.0f ms for a peer's flag (a peer left the solve or died)",
                          C->rank, (double)C->dev.timeout_ns * 1e-6);
            return KMCB200_E_COMM;
        }
        return 0;
    };
    KMC_TRY(read_flags());
    if (persistent) {  // the kernel ran h_flags[3] iterations: one halo step and two dot steps each, on every rank
        C->halo_seq += (unsigned long long)h_flags[3];
        C->dot_seq += 2ull * (unsigned long long)h_flags[3];
        if (getenv("KMCB200_PCG_LOOP_PROFILE") && h_flags[3] > 0) {
            unsigned long long pr[8];
            cudaMemcpy(pr, (char *)ctx->pcg_loop_ws + 64, sizeof(pr), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[pcg loop] rank %d rows %d, %d iterations, us/iteration of CTA 0: p-update %.1f | barrier %.1f | spmv %.1f | barrier %.1f | "
                            "dot %.1f | update %.1f | barrier %.1f | dot %.1f\n", C->rank, rows, h_flags[3],
                    pr[0] * 1e-3 / h_flags[3], pr[1] * 1e-3 / h_flags[3], pr[2] * 1e-3 / h_flags[3], pr[3] * 1e-3 / h_flags[3],
                    pr[4] * 1e-3 / h_flags[3], pr[5] * 1e-3 / h_flags[3], pr[6] * 1e-3 / h_flags[3], pr[7] * 1e-3 / h_flags[3]);
        }
    }
    int batch = 4;
    // Programmatic dependent launch for the kernels of the iteration (KMCB200_PDL=0 switches it off): each kernel's launch
    // and CTA ramp-up overlap the tail of its predecessor; the kernels wait (griddepcontrol.wait) before their first access.
    static const bool pdl_env = !(getenv("KMCB200_PDL") && atoi(getenv("KMCB200_PDL")) == 0);
    const bool pdl = pdl_env && !tun;
    const unsigned fin_threads = C->dev.group_chunks ? FIN_WIDE : CH;
    // KMCB200_PCG_PROFILE=1: CUDA events around every kernel of the iteration (diagnostics only; serialises nothing
    // by itself, the events sit on the same stream)
    static const bool profile = getenv("KMCB200_PCG_PROFILE") != nullptr;
    static cudaEvent_t pev[4 * 32];
    static bool pev_init = false;
    cudaEvent_t pall[2] = {nullptr, nullptr};
    double pt[3] = {0, 0, 0}, pgap = 0;
    long pn = 0;
    if (profile) {
        if (!pev_init) { for (auto &e : pev) cudaEventCreate(&e); pev_init = true; }
        for (auto &e : pall) cudaEventCreate(&e);
        cudaEventRecord(pall[0], ctx->stream);
    }
    int launched_iters = 0;
    while (!h_flags[2]) {  // done
        for (int b = 0; b < batch; ++b) {
            const unsigned long long hs2 = ++C->halo_seq;
            const int nb = (int)(hs2 & 1);
            const bool rec = profile;
            cudaEvent_t *pe = pev + 4 * b;
            if (rec) cudaEventRecord(pe[0], ctx->stream);
            kmc_count_launch();
            KMC_CUDA(kmc_launch_pdl(cg_pupdate_kernel<1>, eb, CH, 0, ctx->stream, pdl, rows, (int)nchunks, K->z,
                                    C->dev.p_full[nb ^ 1], C->dev.p_full[nb], C->dev, nb, hs2, st));
            if (rec) cudaEventRecord(pe[1], ctx->stream);
            if (!tun) {
                const unsigned long long ds = ++C->dot_seq;
                KMC_TRY(spmv_launch(ctx, K, C->dev.p_full[nb], K->Ap, true, fuse, ds, pdl));
                if (!fuse) { kmc_count_launch(); KMC_CUDA(kmc_launch_pdl(dot_finalize_kernel, 1u, fin_threads, 0, ctx->stream, pdl, C->dev, 0, (int)nchunks, ds, st)); }
            } else {  // neighbour part, tunnel part, then p.Ap over the sum
                const unsigned long long ds = ++C->dot_seq;
                KMC_TRY(spmv_launch(ctx, K, C->dev.p_full[nb], K->Ap, false, 0, 0));
                KMC_TRY(tunnel_launch(ctx, *tun, C->dev.p_full[nb], K->Ap, 1));
                kmc_count_launch();
                if (fuse) cg_pap_kernel<true><<<eb, CH, 0, ctx->stream>>>(rows, (int)nchunks, C->dev.p_full[nb], K->Ap, C->dev, ds, st);
                else cg_pap_kernel<false><<<eb, CH, 0, ctx->stream>>>(rows, (int)nchunks, C->dev.p_full[nb], K->Ap, C->dev, ds, st);
                if (!fuse) { kmc_count_launch(); dot_finalize_kernel<<<1, C->dev.group_chunks ? FIN_WIDE : CH, 0, ctx->stream>>>(C->dev, 0, (int)nchunks, ds, st); }
            }
            if (rec) cudaEventRecord(pe[2], ctx->stream);
            kmc_count_launch();
            {
                const unsigned long long ds = ++C->dot_seq;
                if (fuse)
                    KMC_CUDA(kmc_launch_pdl(cg_update_kernel<true>, eb, CH, 0, ctx->stream, pdl, rows, (int)nchunks, C->dev.p_full[nb],
                                            K->Ap, diag_inv_local, x_local, r_local, K->z, C->dev, ds, st));
                else
                    KMC_CUDA(kmc_launch_pdl(cg_update_kernel<false>, eb, CH, 0, ctx->stream, pdl, rows, (int)nchunks, C->dev.p_full[nb],
                                            K->Ap, diag_inv_local, x_local, r_local, K->z, C->dev, ds, st));
                if (!fuse) { kmc_count_launch(); KMC_CUDA(kmc_launch_pdl(dot_finalize_kernel, 1u, fin_threads, 0, ctx->stream, pdl, C->dev, 1, (int)nchunks, ds, st)); }
            }
            if (rec) cudaEventRecord(pe[3], ctx->stream);
            launched_iters++;
        }
        KMC_CUDA(cudaGetLastError());
        KMC_TRY(read_flags());
        if (profile) {
            float ms;
            for (int b = 0; b < batch; ++b) {
                for (int q = 0; q < 3; ++q) { cudaEventElapsedTime(&ms, pev[4 * b + q], pev[4 * b + q + 1]); pt[q] += ms; if (q == 1 && getenv("KMCB200_PCG_PROFILE_VERBOSE")) fprintf(stderr, "%.0f ", 1e3 * ms); }
                if (b + 1 < batch) { cudaEventElapsedTime(&ms, pev[4 * b + 3], pev[4 * (b + 1)]); pgap += ms; }
                pn++;
            }
        }
        if (batch < 32) batch *= 2;
    }
    if (profile) {
        float tot = 0;
        cudaEventRecord(pall[1], ctx->stream);
        cudaEventSynchronize(pall[1]);
        cudaEventElapsedTime(&tot, pall[0], pall[1]);
        fprintf(stderr, "[pcg profile] loop total %.3f ms for %d launched iterations (%d converged)\n", tot, launched_iters, h_flags[3]);
        if (pn > 0)
            fprintf(stderr, "[pcg profile] rank %d rows %d: pupdate %.1f us, spmv+dot %.1f us, update+dot %.1f us, gap %.1f us (avg of %ld launched iterations; sums %.2f %.2f %.2f ms)\n",
                    C->rank, rows, 1e3 * pt[0] / pn, 1e3 * pt[1] / pn, 1e3 * pt[2] / pn, 1e3 * pgap / pn, pn, pt[0], pt[1], pt[2]);
        for (auto &e : pall) cudaEventDestroy(e);
    }
    // leave the state ready for kmcb200_spmv_dot / kmcb200_dot (their kernels early-exit while `done` is set)
    KMC_CUDA(cudaMemsetAsync(&st->done, 0, sizeof(int), ctx->stream));
    ctx->cg_done_stale = false;
    if (iterations_host) *iterations_host = h_flags[3];
    return 0;
}

extern "C" int kmcb200_background_potential(kmcb200_ctx *ctx, kmcb200_kmat *K, int N, int N_left, int N_right,
                                            const int *element, const int *charge, const int *metals_host,
                                            int num_metals, double Vd, double high_G, double low_G,
                                            double *site_potential_boundary, int *iterations_host) {
    KMC_CHECK_ARG(site_potential_boundary != nullptr, "site_potential_boundary");
    KMC_TRY(kmcb200_assemble_K(ctx, K, N, N_left, N_right, element, charge, metals_host, num_metals, Vd, high_G, low_G));
    int N_interface = N - (N_left + N_right);
    double relative_tolerance = 1e-14 * N_interface;  // src/potential_solver_gpu.cu:885
    int max_iterations = 10000;                       // :886
    double *v_soln = site_potential_boundary + N_left + K->row_start;  // :861
    return kmcb200_pcg_jacobi(ctx, K, K->rhs, v_soln, K->inv_diag, relative_tolerance, max_iterations, iterations_host);
}
