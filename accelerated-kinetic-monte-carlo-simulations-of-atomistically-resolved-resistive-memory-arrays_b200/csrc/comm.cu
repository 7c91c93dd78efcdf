// comm.cu -- exchange plan of the row-sharded solver (see comm.cuh): arena allocation, CUDA-IPC peer mapping, halo
// need map / send masks.  Bootstrap data (64-byte IPC handles, need maps) is moved between the processes by the
// caller (torch.distributed in this repo's Python mirror, any out-of-band channel in a C++ host).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "comm.cuh"
#include "kmat.cuh"

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__global__ void need_map_kernel(int rows, const int *__restrict__ row_ptr, const int *__restrict__ col, int own_lo,
                                int own_hi, unsigned char *__restrict__ need) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
        int j = col[k];
        if (j < own_lo || j >= own_hi) need[j] = 1;
    }
}

// all_need: size x n_global bytes (row q = need map of rank q).  send_mask[i] bit q <=> rank q needs my row i.
__global__ void send_mask_kernel(int rows, int row_start, int n_global, int size, int rank,
                                 const unsigned char *__restrict__ all_need, unsigned char *__restrict__ send_mask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    unsigned m = 0;
    for (int q = 0; q < size; ++q)
        if (q != rank && all_need[(size_t)q * n_global + row_start + i]) m |= 1u << q;
    send_mask[i] = (unsigned char)m;
}

__global__ void recv_mask_kernel(int n_global, const unsigned char *__restrict__ my_need, const int *__restrict__ displs,
                                 int size, unsigned *__restrict__ mask_out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_global || !my_need[j]) return;
    int q = 0;
    while (q + 1 < size && j >= displs[q + 1]) ++q;
    atomicOr(mask_out, 1u << q);
}

// appends every global row j with my_need[j] to the list (order fixed afterwards by a host-side sort: setup only)
__global__ void halo_list_kernel(int n_global, const unsigned char *__restrict__ my_need, int *__restrict__ list,
                                 int *__restrict__ count) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_global || !my_need[j]) return;
    list[atomicAdd(count, 1)] = j;
}

int comm_alloc(kmcb200_comm *c) {
    const size_t n = (size_t)c->n_global;
    size_t off = 0;
    c->off_p[0] = off; off = align_up(off + n * sizeof(double), 256);
    c->off_p[1] = off; off = align_up(off + n * sizeof(double), 256);
    c->off_partials = off; off = align_up(off + (size_t)KMC_DOT_SLOTS * c->nchunks_global * sizeof(double), 256);
    c->off_gtotals = off; off = align_up(off + (size_t)KMC_DOT_SLOTS * (c->ngroups_global + 1) * sizeof(double), 256);
    c->off_flag_dot = off; off += 256;
    c->off_flag_halo = off; off += 256;
    c->off_flag_gather = off; off += 256;
    c->off_flag_ack = off; off += 256;
    c->off_gather = off; off = align_up(off + (size_t)(c->gather_cap + 1) * sizeof(double), 256);
    if (c->size > 1) {  // fence-free dot cells + z halo landing zone (comm.cuh)
        const size_t ll_vals = c->group_chunks ? c->ngroups_global : c->nchunks_global;
        c->off_ll = off; off = align_up(off + 4 * (ll_vals + 1) * sizeof(uint4), 256);
        c->off_z = off; off = align_up(off + n * sizeof(double), 256);
    }
    c->arena_bytes = align_up(off, 2 << 20);
    KMC_CUDA(cudaMalloc((void **)&c->arena, c->arena_bytes));
    KMC_CUDA(cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->ctx->stream));
    KMC_CUDA(cudaMalloc((void **)&c->err_word, 64));
    KMC_CUDA(cudaMemsetAsync(c->err_word, 0, 64, c->ctx->stream));
    KMC_CUDA(cudaMalloc((void **)&c->send_mask, (size_t)c->counts[c->rank] + 1));
    KMC_CUDA(cudaMemsetAsync(c->send_mask, 0, (size_t)c->counts[c->rank] + 1, c->ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(c->ctx->stream));
    c->peer_arena[c->rank] = c->arena;
    return 0;
}

}  // namespace

void kmc_comm_fill_dev(kmcb200_comm *c) {
    CommDev &d = c->dev;
    d.rank = c->rank;
    d.size = c->size;
    d.row_start = c->displs[c->rank];
    d.n_global = c->n_global;
    d.nchunks_global = c->nchunks_global;
    d.chunk_start = d.row_start / KMCB200_CHUNK;
    d.recv_mask = c->recv_mask;
    d.p_full[0] = (double *)(c->arena + c->off_p[0]);
    d.p_full[1] = (double *)(c->arena + c->off_p[1]);
    d.partials = (double *)(c->arena + c->off_partials);
    d.group_chunks = c->group_chunks;
    d.ngroups_global = c->ngroups_global;
    d.group_start = c->group_chunks ? d.chunk_start / c->group_chunks : 0;
    d.gtotals = (double *)(c->arena + c->off_gtotals);
    d.flag_dot = (unsigned long long *)(c->arena + c->off_flag_dot);
    d.flag_halo = (unsigned long long *)(c->arena + c->off_flag_halo);
    d.gather = (double *)(c->arena + c->off_gather);
    {
        static const bool ll_env = !(getenv("KMCB200_COMM_LL") && atoi(getenv("KMCB200_COMM_LL")) == 0);
        d.ll_mode = (c->size > 1 && ll_env) ? 1 : 0;
    }
    d.ll_vals = c->group_chunks ? c->ngroups_global : c->nchunks_global;
    d.ll = c->size > 1 ? (uint4 *)(c->arena + c->off_ll) : nullptr;
    d.z_full = c->size > 1 ? (double *)(c->arena + c->off_z) : nullptr;
    d.halo_rows = c->halo_rows;
    d.nhalo = c->nhalo;
    d.flag_gather = (unsigned long long *)(c->arena + c->off_flag_gather);
    d.flag_ack = (unsigned long long *)(c->arena + c->off_flag_ack);
    for (int q = 0; q < KMC_MAX_RANKS; ++q) {
        char *base = (q < c->size) ? c->peer_arena[q] : nullptr;
        d.peer_p_full[q][0] = base ? (double *)(base + c->off_p[0]) : nullptr;
        d.peer_p_full[q][1] = base ? (double *)(base + c->off_p[1]) : nullptr;
        d.peer_partials[q] = base ? (double *)(base + c->off_partials) : nullptr;
        d.peer_gtotals[q] = base ? (double *)(base + c->off_gtotals) : nullptr;
        d.peer_flag_dot[q] = base ? (unsigned long long *)(base + c->off_flag_dot) : nullptr;
        d.peer_flag_halo[q] = base ? (unsigned long long *)(base + c->off_flag_halo) : nullptr;
        d.peer_gather[q] = base ? (const double *)(base + c->off_gather) : nullptr;
        d.peer_ll[q] = (base && c->size > 1) ? (uint4 *)(base + c->off_ll) : nullptr;
        d.peer_z_full[q] = (base && c->size > 1) ? (double *)(base + c->off_z) : nullptr;
        d.peer_flag_gather[q] = base ? (unsigned long long *)(base + c->off_flag_gather) : nullptr;
        d.peer_flag_ack[q] = base ? (unsigned long long *)(base + c->off_flag_ack) : nullptr;
    }
    d.send_mask = c->send_mask;
    if (d.timeout_ns == 0) d.timeout_ns = 20000000000ull;  // 20 s until kmcb200_pcg_jacobi sets the configured bound
}

int kmc_comm_create_local(kmcb200_ctx *ctx, int n_rows, kmcb200_comm **out) {
    int counts[1] = {n_rows}, displs[1] = {0};
    return kmcb200_comm_create(ctx, 0, 1, n_rows, counts, displs, out);
}

extern "C" int kmcb200_comm_create(kmcb200_ctx *ctx, int rank, int size, int n_global_rows, const int *counts,
                                   const int *displs, kmcb200_comm **comm_out) {
    return kmcb200_comm_create_ex(ctx, rank, size, n_global_rows, counts, displs, 0, comm_out);
}

extern "C" int kmcb200_comm_create_ex(kmcb200_ctx *ctx, int rank, int size, int n_global_rows, const int *counts,
                                      const int *displs, long long gather_capacity, kmcb200_comm **comm_out) {
    KMC_CHECK_ARG(ctx && counts && displs && comm_out && gather_capacity >= 0, "null pointer / capacity");
    KMC_CHECK_ARG(size >= 1 && size <= KMC_MAX_RANKS && rank >= 0 && rank < size, "rank/size (<= 8 ranks)");
    long long sum = 0;
    const int nchunks_g = (n_global_rows + KMCB200_CHUNK - 1) / KMCB200_CHUNK;
    const int gran = (nchunks_g > 256 ? KMCB200_DOT_GROUP : 1) * KMCB200_CHUNK;
    for (int q = 0; q < size; ++q) {
        KMC_CHECK_ARG(displs[q] == sum, "displs must be the prefix sums of counts");
        KMC_CHECK_ARG(size == 1 || displs[q] % gran == 0 || counts[q] == 0,
                      "rank boundaries must be multiples of 256 rows (16384 rows for systems of more than 65536 rows): "
                      "kmcb200_partition_aligned");
        sum += counts[q];
    }
    KMC_CHECK_ARG(sum == n_global_rows, "counts do not add up to n_global_rows");
    kmcb200_comm *c = new kmcb200_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->size = size;
    c->n_global = n_global_rows;
    c->nchunks_global = (n_global_rows + KMCB200_CHUNK - 1) / KMCB200_CHUNK;
    c->group_chunks = c->nchunks_global > 256 ? KMCB200_DOT_GROUP : 0;
    c->ngroups_global = c->group_chunks ? (c->nchunks_global + c->group_chunks - 1) / c->group_chunks : 0;
    c->counts.assign(counts, counts + size);
    c->displs.assign(displs, displs + size);
    c->gather_cap = size > 1 ? gather_capacity : 0;
    int rc = comm_alloc(c);
    if (rc) { delete c; return rc; }
    c->peers_open = (size == 1);
    kmc_comm_fill_dev(c);
    *comm_out = c;
    return 0;
}

extern "C" int kmcb200_comm_destroy(kmcb200_comm *c) {
    if (!c) return 0;
    cudaStreamSynchronize(c->ctx->stream);
    for (int q = 0; q < c->size; ++q)
        if (q != c->rank && c->peer_arena[q]) cudaIpcCloseMemHandle(c->peer_arena[q]);
    cudaFree(c->arena);
    cudaFree(c->send_mask);
    cudaFree(c->halo_rows);
    cudaFree(c->err_word);
    delete c;
    return 0;
}

extern "C" int kmcb200_comm_ipc_handle(kmcb200_comm *c, void *handle64_host) {
    KMC_CHECK_ARG(c && handle64_host, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    KMC_CUDA(cudaIpcGetMemHandle(&h, c->arena));
    memcpy(handle64_host, &h, 64);
    return 0;
}

extern "C" int kmcb200_comm_open_peers(kmcb200_comm *c, const void *handles_host) {
    KMC_CHECK_ARG(c && handles_host, "null pointer");
    for (int q = 0; q < c->size; ++q) {
        if (q == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles_host + 64 * q, 64);
        void *p = nullptr;
        KMC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_arena[q] = (char *)p;
    }
    c->peers_open = true;
    kmc_comm_fill_dev(c);
    return 0;
}

extern "C" int kmcb200_kmat_attach_comm(kmcb200_kmat *K, kmcb200_comm *c) {
    KMC_CHECK_ARG(K && c, "null pointer");
    KMC_CHECK_ARG(K->rows == c->counts[c->rank] && K->row_start == c->displs[c->rank] && K->cols_global == c->n_global,
                  "matrix rows do not match the communicator's partition");
    if (K->comm && K->owns_comm) kmcb200_comm_destroy(K->comm);
    K->comm = c;
    K->owns_comm = false;
    return 0;
}

extern "C" int kmcb200_kmat_need_map(kmcb200_kmat *K, unsigned char *need_dev) {
    KMC_CHECK_ARG(K && need_dev, "null pointer");
    kmcb200_ctx *ctx = K->ctx;
    KMC_CUDA(cudaMemsetAsync(need_dev, 0, (size_t)K->cols_global, ctx->stream));
    kmc_count_launch();
    need_map_kernel<<<(K->rows + 127) / 128, 128, 0, ctx->stream>>>(K->rows, K->row_ptr, K->col, K->row_start,
                                                                   K->row_start + K->rows, need_dev);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_comm_set_send_masks(kmcb200_comm *c, const unsigned char *all_need_dev) {
    KMC_CHECK_ARG(c && all_need_dev, "null pointer");
    kmcb200_ctx *ctx = c->ctx;
    const int rows = c->counts[c->rank];
    unsigned *d_mask = nullptr;
    int *d_displs = nullptr;
    KMC_TRY(kmc_scratch(ctx, 5, 256, (void **)&d_mask));
    d_displs = (int *)(d_mask + 8);
    KMC_CUDA(cudaMemsetAsync(d_mask, 0, 32, ctx->stream));
    KMC_CUDA(cudaMemcpyAsync(d_displs, c->displs.data(), c->size * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (rows > 0) {
        kmc_count_launch();
        send_mask_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(rows, c->displs[c->rank], c->n_global, c->size, c->rank,
                                                                    all_need_dev, c->send_mask);
    }
    kmc_count_launch();
    recv_mask_kernel<<<(c->n_global + 255) / 256, 256, 0, ctx->stream>>>(c->n_global, all_need_dev + (size_t)c->rank * c->n_global,
                                                                        d_displs, c->size, d_mask);
    KMC_CUDA(cudaGetLastError());
    unsigned m = 0;
    KMC_CUDA(cudaMemcpyAsync(&m, d_mask, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    c->recv_mask = m & ~(1u << c->rank);
    {   // the halo rows of this rank (global ids, ascending): the p-update forms their p entries locally (comm.cuh)
        int *d_list = nullptr, *d_count = (int *)(d_mask + 32);
        KMC_TRY(kmc_scratch(ctx, 6, (size_t)c->n_global * sizeof(int) + 16, (void **)&d_list));
        KMC_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
        kmc_count_launch();
        halo_list_kernel<<<(c->n_global + 255) / 256, 256, 0, ctx->stream>>>(c->n_global, all_need_dev + (size_t)c->rank * c->n_global,
                                                                            d_list, d_count);
        KMC_CUDA(cudaGetLastError());
        int cnt = 0;
        KMC_CUDA(cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        std::vector<int> h((size_t)cnt);
        if (cnt > 0) KMC_CUDA(cudaMemcpy(h.data(), d_list, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        if (c->halo_rows) { KMC_CUDA(cudaFree(c->halo_rows)); c->halo_rows = nullptr; }
        KMC_CUDA(cudaMalloc((void **)&c->halo_rows, ((size_t)cnt + 1) * sizeof(int)));
        if (cnt > 0) KMC_CUDA(cudaMemcpy(c->halo_rows, h.data(), (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice));
        c->nhalo = cnt;
    }
    c->masks_set = true;
    kmc_comm_fill_dev(c);
    return 0;
}

extern "C" int kmcb200_comm_info(kmcb200_comm *c, int *rank, int *size, unsigned *recv_mask, long long *arena_bytes) {
    KMC_CHECK_ARG(c != nullptr, "comm");
    if (rank) *rank = c->rank;
    if (size) *size = c->size;
    if (recv_mask) *recv_mask = c->recv_mask;
    if (arena_bytes) *arena_bytes = (long long)c->arena_bytes;
    return 0;
}


// ---- all-gather of row slices over NVLink peer memory ------------------------------------------------------------------
// Replaces the MPI_Gatherv + MPI_Bcast of the potentials (reference src/kmc_main.cpp:367-384,411-427,
// src/potential_solver_gpu.cu:1133-1142).  Every rank stages its slice in ITS OWN arena and raises a flag at the peers; the
// peers pull the slice straight out of the owner's memory.  An acknowledgement flag keeps the owner from overwriting the
// staging region before every peer has pulled the previous round.  All waits are bounded (comm.cuh).
namespace {
__global__ void ag_wait_acks_kernel(CommDev cm, unsigned long long prev_seq, int *err) {
    if (threadIdx.x == 0 && prev_seq > 0) {
        kmc_wait_flags(cm.flag_ack, (1u << cm.size) - 1u, cm.rank, prev_seq, cm.timeout_ns, err);
        __threadfence_system();
    }
}
__global__ void ag_stage_kernel(const double *__restrict__ src, long long count, double *__restrict__ stage) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        stage[i] = src[i];
}
__global__ void ag_raise_kernel(CommDev cm, unsigned long long seq, int which) {  // which: 0 = data ready, 1 = pulled
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int q = 0; q < cm.size; ++q)
            if (q != cm.rank) kmc_store_relaxed_sys((which ? cm.peer_flag_ack[q] : cm.peer_flag_gather[q]) + cm.rank, seq);
    }
}
struct AgPlan {
    int counts[KMC_MAX_RANKS], displs[KMC_MAX_RANKS];
};
__global__ void __launch_bounds__(256) ag_pull_kernel(CommDev cm, AgPlan plan, unsigned long long seq, double *__restrict__ dst,
                                                     int blocks_per_peer, int *err) {
    const int q = blockIdx.x / blocks_per_peer, b = blockIdx.x % blocks_per_peer;
    if (q == cm.rank) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = kmc_wait_flags(cm.flag_gather, 1u << q, cm.rank, seq, cm.timeout_ns, err) ? 1 : 0;
        __threadfence_system();
    }
    __syncthreads();
    if (!ok) return;
    const double *src = cm.peer_gather[q];
    double *out = dst + plan.displs[q];
    const long long n = plan.counts[q];
    for (long long i = (long long)b * 256 + threadIdx.x; i < n; i += (long long)blocks_per_peer * 256) out[i] = src[i];
}
}  // namespace

extern "C" int kmcb200_comm_allgather(kmcb200_comm *c, double *vec_dev, const int *counts_host, const int *displs_host) {
    KMC_CHECK_ARG(c && vec_dev && counts_host && displs_host, "null pointer");
    if (c->size == 1) return 0;
    KMC_CHECK_ARG(c->peers_open, "kmcb200_comm_open_peers was not called");
    KMC_CHECK_ARG(counts_host[c->rank] <= c->gather_cap, "slice larger than the communicator's gather capacity");
    kmcb200_ctx *ctx = c->ctx;
    AgPlan plan;
    for (int q = 0; q < KMC_MAX_RANKS; ++q) {
        plan.counts[q] = q < c->size ? counts_host[q] : 0;
        plan.displs[q] = q < c->size ? displs_host[q] : 0;
    }
    const unsigned long long seq = ++c->gather_seq;
    const long long mine = counts_host[c->rank];
    kmc_count_launch();
    ag_wait_acks_kernel<<<1, 32, 0, ctx->stream>>>(c->dev, seq - 1, c->err_word);
    if (mine > 0) {
        unsigned blocks = (unsigned)((mine + 255) / 256);
        if (blocks > (unsigned)ctx->sm_count * 4) blocks = ctx->sm_count * 4;
        kmc_count_launch();
        ag_stage_kernel<<<blocks, 256, 0, ctx->stream>>>(vec_dev + displs_host[c->rank], mine, c->dev.gather);
    }
    kmc_count_launch();
    ag_raise_kernel<<<1, 32, 0, ctx->stream>>>(c->dev, seq, 0);
    const int bpp = 32;
    kmc_count_launch();
    ag_pull_kernel<<<c->size * bpp, 256, 0, ctx->stream>>>(c->dev, plan, seq, vec_dev, bpp, c->err_word);
    kmc_count_launch();
    ag_raise_kernel<<<1, 32, 0, ctx->stream>>>(c->dev, seq, 1);
    KMC_CUDA(cudaGetLastError());
    int e = 0;
    KMC_CUDA(cudaMemcpyAsync(&e, c->err_word, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e) {
        kmc_set_error("kmcb200_comm_allgather: rank %d timed out waiting for a peer (a peer left or died)", c->rank);
        return KMCB200_E_COMM;
    }
    return 0;
}

// ---- out-of-band rendezvous through a shared directory (bootstrap only: 64-byte IPC handles, need maps, barriers) -------
// Any channel works for these few exchanges (the Python driver uses torch.distributed); this one needs nothing but a
// directory every rank of the node can see (e.g. under /dev/shm) and is what the C++ host uses (no MPI in this image).
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>

struct kmcb200_rdv {
    std::string dir;
    int rank = 0, size = 1;
    unsigned long long round = 0;
    double timeout_s = 120.0;
};

extern "C" int kmcb200_rdv_open(const char *dir, int rank, int size, kmcb200_rdv **out) {
    KMC_CHECK_ARG(dir && out && size >= 1 && rank >= 0 && rank < size, "arguments");
    mkdir(dir, 0777);  // may already exist
    struct stat st;
    if (stat(dir, &st) != 0 || !S_ISDIR(st.st_mode)) {
        kmc_set_error("rendezvous directory %s is not usable", dir);
        return KMCB200_E_IO;
    }
    kmcb200_rdv *r = new kmcb200_rdv();
    r->dir = dir; r->rank = rank; r->size = size;
    if (getenv("KMCB200_RDV_TIMEOUT_S")) r->timeout_s = atof(getenv("KMCB200_RDV_TIMEOUT_S"));
    *out = r;
    return 0;
}
extern "C" int kmcb200_rdv_close(kmcb200_rdv *r) {
    if (!r) return 0;
    // the files of the last two rounds stay: a slower peer may not have read them yet (the directory is the launcher's
    // to remove)
    delete r;
    return 0;
}
extern "C" int kmcb200_rdv_allgather(kmcb200_rdv *r, const void *mine, size_t bytes, void *all) {
    KMC_CHECK_ARG(r && mine && all, "null pointer");
    const unsigned long long k = ++r->round;
    auto name = [&](unsigned long long round, int q) { return r->dir + "/ag" + std::to_string(round) + "_" + std::to_string(q); };
    {   // publish: write to a temporary name, then rename (readers never see a partial file)
        const std::string tmp = name(k, r->rank) + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(mine, 1, bytes, f) != bytes) { if (f) fclose(f); kmc_set_error("rendezvous: cannot write %s", tmp.c_str()); return KMCB200_E_IO; }
        fclose(f);
        if (rename(tmp.c_str(), name(k, r->rank).c_str()) != 0) { kmc_set_error("rendezvous: rename failed"); return KMCB200_E_IO; }
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int q = 0; q < r->size; ++q) {
        char *dst = (char *)all + (size_t)q * bytes;
        if (q == r->rank) { memcpy(dst, mine, bytes); continue; }
        for (;;) {
            struct stat st;
            if (stat(name(k, q).c_str(), &st) == 0 && (size_t)st.st_size == bytes) {
                FILE *f = fopen(name(k, q).c_str(), "rb");
                if (f) {
                    const size_t got = fread(dst, 1, bytes, f);
                    fclose(f);
                    if (got == bytes) break;
                }
            }
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > r->timeout_s) {
                kmc_set_error("rendezvous: rank %d waited %.0f s for rank %d (round %llu)", r->rank, r->timeout_s, q, k);
                return KMCB200_E_COMM;
            }
            usleep(200);
        }
    }
    // every rank has passed round k-1 (it published round k only after reading all of k-1): files of k-2 are garbage
    if (k > 2) unlink(name(k - 2, r->rank).c_str());
    return 0;
}
extern "C" int kmcb200_rdv_barrier(kmcb200_rdv *r) {
    char mine = 1, all[KMC_MAX_RANKS * 4];
    KMC_CHECK_ARG(r && r->size <= KMC_MAX_RANKS * 4, "rendezvous");
    return kmcb200_rdv_allgather(r, &mine, 1, all);
}
