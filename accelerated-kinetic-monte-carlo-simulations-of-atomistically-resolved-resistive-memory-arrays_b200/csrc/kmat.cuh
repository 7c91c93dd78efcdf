// kmat.cuh -- the K matrix object (reference: Distributed_matrix + Distributed_vector + contact CSR blocks,
// dist_iterative/dist_objects.h:10-232, src/gpu_buffers.h K_distributed/K_p_distributed/left_*/right_*).
#pragma once
#include "common.cuh"

struct kmcb200_comm;  // multi-GPU exchange plan (comm.cu)

struct kmcb200_kmat {
    kmcb200_ctx *ctx = nullptr;
    int rows = 0;         // rows owned by this rank
    int row_start = 0;    // first owned row (interior numbering)
    int cols_global = 0;  // interior size (N_interface)
    long long nnz = 0, left_nnz = 0, right_nnz = 0;
    bool owns_csr = false;
    // CSR with GLOBAL interior column ids, ascending inside a row
    int *row_ptr = nullptr, *col = nullptr;
    double *val = nullptr;
    // contact blocks (columns: left = site id, right = site id - (N_left + N_interface))
    int *left_row_ptr = nullptr, *left_col = nullptr, *right_row_ptr = nullptr, *right_col = nullptr;
    double *inv_diag = nullptr, *rhs = nullptr;
    // PCG workspace (persistent): p is indexed by GLOBAL interior row (halo entries land in place)
    double *Ap = nullptr, *z = nullptr;  // (the global-indexed p vectors live in the exchange arena, comm.cuh)
    unsigned char *site_class = nullptr;  // N bytes, (re)built by assemble
    size_t site_class_cap = 0;
    kmcb200_comm *comm = nullptr;  // exchange plan; a private size-1 plan unless kmcb200_kmat_attach_comm was called
    bool owns_comm = false;
    // SpMV plan (spmv_plan.cu): per 256-row chunk the sorted unique columns it touches (u_col, CSR-like u_ptr) and a
    // 16-bit chunk-local column id per non-zero (lcol).  The SpMV stages x[u_col] in shared memory once per chunk and
    // gathers from there; HBM traffic per non-zero drops from 12 B (val + int32 col) to 10 B (val + uint16 lcol).
    int *u_ptr = nullptr, *u_col = nullptr;
    unsigned short *lcol = nullptr;
    int plan_max_unique = 0;  // 0: no plan (fall back to the direct-gather kernel)
};

int kmc_kmat_finalize(kmcb200_kmat *K);  // allocates the PCG workspace + builds the SpMV plan
int kmc_build_spmv_plan(kmcb200_kmat *K);
