// cellgrid.cuh -- uniform cell grid used as a CANDIDATE ENUMERATOR for the distance predicates of the
// neighbour table, the cutoff counts and the K sparsity.  The predicate itself is always re-evaluated with
// the reference's arithmetic (common.cuh), so results are identical to the reference's O(N^2) kernels.
#pragma once
#include "common.cuh"

struct CellGridDev {
    double x0, y0, z0, hx, hy, hz;
    int nx, ny, nz;
    int wrap_yz;
    double latty, lattz;
    const int *cell_start;  // ncell+1
    const int *items;       // site ids grouped by cell (unordered inside a cell)

    __device__ __forceinline__ int clampi(int c, int n) const { return c < 0 ? 0 : (c >= n ? n - 1 : c); }
    __device__ __forceinline__ int cx(double x) const { return clampi((int)floor((x - x0) / hx), nx); }
    __device__ __forceinline__ int cy(double y) const {
        if (wrap_yz) {
            double f = y / latty;
            f -= floor(f);
            return clampi((int)(f * ny), ny);
        }
        return clampi((int)floor((y - y0) / hy), ny);
    }
    __device__ __forceinline__ int cz(double z) const {
        if (wrap_yz) {
            double f = z / lattz;
            f -= floor(f);
            return clampi((int)(f * nz), nz);
        }
        return clampi((int)floor((z - z0) / hz), nz);
    }
    __device__ __forceinline__ int cell(int a, int b, int c) const { return (a * ny + b) * nz + c; }

    // calls f(j) for every binned site j in the 27 (deduplicated under wrap) cells around p
    template <class F>
    __device__ __forceinline__ void for_each_candidate(double px, double py, double pz, F &&f) const {
        int a = cx(px), b = cy(py), c = cz(pz);
        int ys[3], zs[3], nys = 0, nzs = 0;
        for (int d = -1; d <= 1; ++d) {
            int v = b + d;
            if (wrap_yz) v = ((v % ny) + ny) % ny;
            else if (v < 0 || v >= ny) continue;
            bool dup = false;
            for (int q = 0; q < nys; ++q) dup |= (ys[q] == v);
            if (!dup) ys[nys++] = v;
        }
        for (int d = -1; d <= 1; ++d) {
            int v = c + d;
            if (wrap_yz) v = ((v % nz) + nz) % nz;
            else if (v < 0 || v >= nz) continue;
            bool dup = false;
            for (int q = 0; q < nzs; ++q) dup |= (zs[q] == v);
            if (!dup) zs[nzs++] = v;
        }
        for (int da = -1; da <= 1; ++da) {
            int aa = a + da;
            if (aa < 0 || aa >= nx) continue;
            for (int q = 0; q < nys; ++q)
                for (int r = 0; r < nzs; ++r) {
                    int cc = cell(aa, ys[q], zs[r]);
                    int e = cell_start[cc + 1];
                    for (int s = cell_start[cc]; s < e; ++s) f(items[s]);
                }
        }
    }
};

// Builds a grid over sites [first, first+count) with cell edge >= cutoff*1.0001.  Arrays live in ctx scratch
// slots 0..2 (valid until the next kmc_build_cellgrid call).  Host-synchronising (bounding box).
int kmc_build_cellgrid(kmcb200_ctx *ctx, const double *x, const double *y, const double *z, int first, int count,
                       double cutoff, int pbc, const double *lattice_host, CellGridDev *grid_out);
