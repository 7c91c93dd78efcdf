// pcg.cuh -- internal interface of the PCG / SpMV layer (pcg.cu) used by the Kirchhoff chain (kirchhoff.cu)
#pragma once
#include "kmat.cuh"

// The tunnel sub-block of the split-sparse format (reference Distributed_subblock_sparse, dist_iterative/dist_objects.h:51-64):
// CSR over the GLOBAL tunnel-point columns for this rank's tunnel rows.
struct TunnelDev {
    int n_local = 0;                 // tunnel rows owned by this rank
    const int *row_ptr = nullptr;    // n_local + 1
    const int *col = nullptr;        // global tunnel-point ids, ascending inside a row
    const double *val = nullptr;
    const int *rows_global = nullptr;  // matrix row (global) of every tunnel point: the gather index into p
    const int *rows_local = nullptr;   // local matrix row of each local tunnel row: the scatter index into Ap
};

int kmc_pcg_run(kmcb200_ctx *ctx, kmcb200_kmat *K, const TunnelDev *tun, double *r_local, double *x_local,
                const double *diag_inv_local, double relative_tolerance, int max_iterations, int *iterations_host);
int kmc_split_spmv(kmcb200_ctx *ctx, kmcb200_kmat *K, const TunnelDev *tun, const double *x_local, double *y_local);
