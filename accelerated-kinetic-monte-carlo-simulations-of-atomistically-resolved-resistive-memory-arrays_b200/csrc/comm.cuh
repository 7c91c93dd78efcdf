// comm.cuh -- row-sharded solver exchange over NVLink peer memory (one process per GPU).
//
// Replaces the reference's GPU-aware MPI traffic inside the PCG: the pack -> MPI_Isend/Irecv -> unpack halo exchange of
// dspmv::gpu_packing_cam (dist_iterative/dist_spmv_gpu_packing.cpp:106-228) and the MPI_Allreduce after every hipblasDdot
// (dist_iterative/dist_conjugate_gradient.cpp:188,213,241,265).  Here the kernel that PRODUCES a value also delivers it.
// Inside the PCG loop (CommDev::ll_mode, the default; see the comment at that field):
//   * the kernels that compute z store the z entries a peer's matrix block references into that peer's z_full; every
//     rank forms the halo entries of p itself, so p is not exchanged;
//   * dot contributions (group totals / chunk partials) travel as self-validating 16-byte cells; every rank then reduces
//     ALL contributions in the fixed order of the summation spec, so every rank holds the bit-identical scalar without a
//     collective call or a host round trip.
// One-off sharded SpMVs (kmcb200_spmv, A x0 of a solve) and KMCB200_COMM_LL=0 use the flag protocol:
//   * the p-update kernel stores each new p entry into its own p vector and, for rows a peer's matrix block
//     references, directly into that peer's p vector (remote NVLink stores), then raises a flag;
//   * the dot-completing CTA pushes its contributions to all peers' arrays, fences, raises a flag and waits for theirs.
// With size == 1 the same kernels run with empty peer loops (this is the single-GPU path).
#pragma once
#include "common.cuh"

#define KMC_MAX_RANKS 8
#define KMC_DOT_SLOTS 5  // 0: p.Ap, 1: r.z, 2: b.b, 3: r.z (setup), 4: kmcb200_dot

// by-value kernel argument
struct CommDev {
    int rank, size;
    int row_start;            // first owned global row (multiple of 256 when size > 1)
    int n_global;             // global number of rows
    int nchunks_global;       // ceil(n_global / 256)
    int chunk_start;          // row_start / 256
    unsigned recv_mask;       // peers this rank's matrix rows read p entries from
    double *p_full[2];        // local, double buffered, indexed by global row
    double *partials;         // local, KMC_DOT_SLOTS * nchunks_global (chunk partials; exchanged only when group_chunks == 0)
    int group_chunks;         // 0: single-level dot combine (<= 256 chunks); else KMCB200_DOT_GROUP
    int ngroups_global;       // ceil(nchunks_global / group_chunks) (0 when single-level)
    int group_start;          // chunk_start / group_chunks
    double *gtotals;          // local, KMC_DOT_SLOTS * ngroups_global: the group totals of ALL ranks
    double *peer_gtotals[KMC_MAX_RANKS];
    unsigned long long *flag_dot;   // local, [KMC_MAX_RANKS] written by peers
    unsigned long long *flag_halo;  // local, [KMC_MAX_RANKS] written by peers
    double *peer_p_full[KMC_MAX_RANKS][2];
    double *peer_partials[KMC_MAX_RANKS];
    unsigned long long *peer_flag_dot[KMC_MAX_RANKS];   // peer's flag_dot array (we write entry [rank])
    unsigned long long *peer_flag_halo[KMC_MAX_RANKS];
    const unsigned char *send_mask;  // local rows: bit q set -> peer q needs this row's p entry
    // all-gather of row slices (potentials): the owner stages its slice in its own arena, peers pull it over NVLink
    double *gather;                                   // local staging region (gather_cap doubles)
    const double *peer_gather[KMC_MAX_RANKS];
    unsigned long long *flag_gather, *flag_ack;       // local, [KMC_MAX_RANKS] each, written by peers
    unsigned long long *peer_flag_gather[KMC_MAX_RANKS], *peer_flag_ack[KMC_MAX_RANKS];
    unsigned long long timeout_ns;   // bound of every peer-flag wait
    int *err;                        // device word raised when a wait timed out (CgState::comm_error of the context)
    // ---- PCG loop at size > 1 (ll_mode, the default): two exchanges per iteration instead of three, none with a fence
    // on its critical path.  (1) The kernels that compute z (setup, x/r/z update) store the z entries a peer's matrix
    // block references straight into that peer's z_full and fence them before they end; every rank then forms the halo
    // entries of p = z + beta p itself (same operands, same operations => same bits), so p needs no exchange.  (2) Dot
    // contributions travel as 16-byte {lo, seq, hi, seq} cells (each 8-byte half carries its own sequence number, so the
    // receiver needs no flag and neither side a fence); cells are double buffered by the parity of the sequence number.
    // A peer's cells of dot s arrive after its z stores of the same step were fenced, so "all cells of r.z present"
    // implies "all z halo entries present".
    int ll_mode;
    int ll_vals;                          // cells per dot: nchunks_global (single-level combine) or ngroups_global
    uint4 *ll;                            // local, [2][2][ll_vals]
    uint4 *peer_ll[KMC_MAX_RANKS];
    double *z_full;                       // local, indexed by global row; only halo entries are ever written (by peers)
    double *peer_z_full[KMC_MAX_RANKS];
    const int *halo_rows;                 // global rows this rank's matrix block reads from peers, ascending
    int nhalo;
};

struct kmcb200_comm {
    kmcb200_ctx *ctx = nullptr;
    int rank = 0, size = 1;
    int n_global = 0, nchunks_global = 0;
    std::vector<int> counts, displs;
    char *arena = nullptr;  // local allocation shared with the peers through CUDA IPC
    size_t arena_bytes = 0;
    size_t off_p[2] = {0, 0}, off_partials = 0, off_gtotals = 0, off_flag_dot = 0, off_flag_halo = 0;
    size_t off_gather = 0, off_flag_gather = 0, off_flag_ack = 0, off_ll = 0, off_z = 0;
    int *halo_rows = nullptr;  // device, nhalo entries (built by kmcb200_comm_set_send_masks)
    int nhalo = 0;
    long long gather_cap = 0;          // doubles in the all-gather staging region
    unsigned long long gather_seq = 0;  // host-side call counter (identical on every rank)
    int *err_word = nullptr;            // device int raised by a timed-out wait outside a PCG solve
    int group_chunks = 0, ngroups_global = 0;
    bool masks_set = false;  // kmcb200_comm_set_send_masks was called (required before any sharded SpMV / PCG)
    char *peer_arena[KMC_MAX_RANKS] = {nullptr};
    bool peers_open = false;
    unsigned char *send_mask = nullptr;  // counts[rank] bytes
    unsigned recv_mask = 0;
    unsigned long long dot_seq = 0, halo_seq = 0;  // host-side launch counters (identical on every rank)
    CommDev dev;
};

int kmc_comm_create_local(kmcb200_ctx *ctx, int n_rows, kmcb200_comm **out);  // size == 1
void kmc_comm_fill_dev(kmcb200_comm *c);

// ---- device side --------------------------------------------------------------------------------------------------
// Flag protocol: writer = data stores, ONE __threadfence_system(), then relaxed flag stores to all peers (a release
// store per peer would pay one system fence each: measured ~3 us x 7 peers per exchange); reader = relaxed polling, then
// ONE __threadfence_system() before touching the data.
__device__ __forceinline__ void kmc_store_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long kmc_load_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte cell of the fence-free dot exchange (the LL idea of NCCL: data and sequence number share an 8-byte store).
// The cell carries the low 32 bits of the 64-bit dot sequence number.  A stale cell could only be mistaken for a fresh
// one if it had last been written exactly 2^32 dots earlier; every cell of a parity is rewritten at least once per solve
// (the two setup dots use both cell rows), i.e. every few hundred dots, so that cannot happen.
__device__ __forceinline__ void kmc_ll_store(uint4 *cell, double v, unsigned seq32) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned)b), "r"(seq32),
                 "r"((unsigned)(b >> 32)), "r"(seq32)
                 : "memory");
}
__device__ __forceinline__ bool kmc_ll_try_load(const uint4 *cell, unsigned seq32, double *v) {
    unsigned a, fa, b, fb;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(cell) : "memory");
    if (fa != seq32 || fb != seq32) return false;
    *v = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
    return true;
}
__device__ __forceinline__ unsigned long long kmc_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// One thread waits until every peer in `mask` has published sequence number >= seq (caller fences afterwards).
// The wait is BOUNDED: if a peer never arrives (it left the solve early, its process died) the waiter gives up after
// timeout_ns, raises *err and returns false; the kernels of the solve then wind down (CgState::done) and the host call
// returns KMCB200_E_COMM instead of leaving every GPU of the job spinning.
__device__ __forceinline__ bool kmc_wait_flags(const unsigned long long *flags, unsigned mask, int self,
                                               unsigned long long seq, unsigned long long timeout_ns, int *err) {
    unsigned long long deadline = 0;
    for (int q = 0; q < KMC_MAX_RANKS; ++q) {
        if (q == self || !((mask >> q) & 1u)) continue;
        unsigned spins = 0;
        while (kmc_load_relaxed_sys(flags + q) < seq) {
            __nanosleep(20);
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = kmc_globaltimer_ns();
                if (deadline == 0) deadline = now + timeout_ns;
                else if (now > deadline || (err && *(volatile int *)err)) {
                    if (err) *(volatile int *)err = 1;
                    return false;
                }
            }
        }
    }
    return true;
}
