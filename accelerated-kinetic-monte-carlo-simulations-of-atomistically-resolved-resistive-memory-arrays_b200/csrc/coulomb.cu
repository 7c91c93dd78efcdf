// coulomb.cu -- a5 site charges, a8 screened-Coulomb potential of the charged defects, a9 potential sum.
// Reference: src/potential_solver_gpu.cu:12-85 (update_charge), :1525-1564 + :1620-1655
// (calculate_pairwise_interaction_indexed / poisson_gridless_gpu), :832-843 + :1130-1151 (sum).
//
// The reference streams a materialised N x N_cutoff int32 list (635 MB at 5 nm) per step with one thread per
// site to find the O(10^2) charged sources of that site.  Here the charged sites (Q << N) are compacted in
// ascending site order each step and an N x Q kernel tiles them through shared memory; the membership test of
// the cutoff list (element class, r < cutoff, i != j) is evaluated inline, and every site sums its sources in
// ascending j like the reference thread does.
#include "cellgrid.cuh"

namespace {

// potential_solver_gpu.cu:12-63.  One thread per site; only V / Od sites touch their neighbour row.
__global__ void __launch_bounds__(256) update_charge_kernel(const int *__restrict__ element, int *__restrict__ charge,
                                                           const int *__restrict__ neigh, int nn, unsigned metal_mask,
                                                           int row_start, int row_count) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_count) return;
    int i = idx + row_start;
    int el = element[i];
    if (el == KMCB200_VACANCY) {
        int c = 2, Vnn = 0;
        const int *row = neigh + (size_t)idx * nn;
        for (int n = 0; n < nn; ++n) {
            int j = row[n];
            if (j >= 0) {
                int ej = element[j];
                if (ej == KMCB200_VACANCY) Vnn++;
                if ((metal_mask >> ej) & 1u) c = 0;
                if (Vnn >= 2) c = 0;
            }
        }
        charge[i] = c;
    } else if (el == KMCB200_OXYGEN_DEFECT) {
        int c = -2;
        const int *row = neigh + (size_t)idx * nn;
        for (int n = 0; n < nn; ++n) {
            int j = row[n];
            if (j >= 0 && ((metal_mask >> element[j]) & 1u)) c = 0;
        }
        charge[i] = c;
    }
}

__global__ void charged_flag_kernel(const int *__restrict__ element, const int *__restrict__ charge, int N,
                                    int *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    flag[i] = (i < N && charge[i] != 0 && kmc_possibly_charged(element[i])) ? 1 : 0;
}

struct __align__(16) Source {
    double x, y, z;
    int charge, idx;
};

__global__ void charged_scatter_kernel(const int *__restrict__ charge, const int *__restrict__ element,
                                       const double *__restrict__ x, const double *__restrict__ y,
                                       const double *__restrict__ z, int N, const int *__restrict__ offs,
                                       const int *__restrict__ site_cell, Source *__restrict__ src,
                                       int *__restrict__ src_cell, int *__restrict__ src_cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (charge[i] != 0 && kmc_possibly_charged(element[i])) {
        Source s;
        s.x = x[i]; s.y = y[i]; s.z = z[i]; s.charge = charge[i]; s.idx = i;
        int o = offs[i];
        src[o] = s;
        const int sc = site_cell[i];
        src_cell[o] = sc;
        atomicAdd(src_cell_count + sc, 1);
    }
}

// order-sensitive 64-bit checksum of the coordinates: guards the cached target plan against a caller that reuses
// the same device address for a different structure
__global__ void pos_checksum_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                    const double *__restrict__ z, int N, unsigned long long *__restrict__ out) {
    unsigned long long h = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        unsigned long long a = (unsigned long long)__double_as_longlong(x[i]);
        unsigned long long b = (unsigned long long)__double_as_longlong(y[i]);
        unsigned long long c = (unsigned long long)__double_as_longlong(z[i]);
        h += (a * 0x9E3779B97F4A7C15ull + (b ^ (c << 1))) * (2ull * (unsigned long long)i + 1ull);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) h += __shfl_xor_sync(KMC_FULL_MASK, h, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}

// ---- static target plan: sites grouped by 20 A cell -------------------------------------------------------
__global__ void site_cell_kernel(CellGridDev g, const double *__restrict__ x, const double *__restrict__ y,
                                 const double *__restrict__ z, int N, int *__restrict__ site_cell) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) site_cell[i] = g.cell(g.cx(x[i]), g.cy(y[i]), g.cz(z[i]));
}
__global__ void cell_blocks_kernel(const int *__restrict__ cell_start, int ncell, int ct, int *__restrict__ nblk) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= ncell) nblk[c] = (c < ncell) ? (cell_start[c + 1] - cell_start[c] + ct - 1) / ct : 0;
}
__global__ void block_map_kernel(const int *__restrict__ blk_start, int ncell, int *__restrict__ blk_cell) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    for (int b = blk_start[c]; b < blk_start[c + 1]; ++b) blk_cell[b] = c;
}

// ---- per step: ascending-j list of the charged sources in the 27-cell neighbourhood of every cell ----------
// Sources are binned by cell (count / scan / atomic fill: unordered inside a cell); then one CTA per target cell
// gathers the source ids of its <= 27 neighbour cells into shared memory, sorts them (bitonic) and writes the list,
// so every target still sums its sources in ascending j.  O(Q) binning + O(m log^2 m) per cell instead of scanning all
// Q sources for every cell.
__global__ void src_bin_fill_kernel(int Q, const int *__restrict__ src_cell, const int *__restrict__ cstart,
                                    int *__restrict__ fill, int *__restrict__ src_by_cell) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int c = src_cell[q];
    src_by_cell[cstart[c] + atomicAdd(fill + c, 1)] = q;
}

// neighbourhood size of every cell that holds targets of the requested row range
__global__ void nbr_count_kernel(int nx, int ny, int nz, const int *__restrict__ cstart,
                                 const int *__restrict__ cell_min, const int *__restrict__ cell_max, int row_lo,
                                 int row_hi, int *__restrict__ cnt, int *__restrict__ max_cnt) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int ncell = nx * ny * nz;
    if (c > ncell) return;
    int n = 0;
    if (c < ncell && cell_max[c] >= row_lo && cell_min[c] < row_hi) {
        int a = c / (ny * nz), b = (c / nz) % ny, d = c % nz;
        for (int da = -1; da <= 1; ++da)
            for (int db = -1; db <= 1; ++db)
                for (int dd = -1; dd <= 1; ++dd) {
                    int aa = a + da, bb = b + db, ee = d + dd;
                    if (aa < 0 || aa >= nx || bb < 0 || bb >= ny || ee < 0 || ee >= nz) continue;
                    int cc = (aa * ny + bb) * nz + ee;
                    n += cstart[cc + 1] - cstart[cc];
                }
    }
    cnt[c] = n;
    if (n > 0) atomicMax(max_cnt, n);
}

__global__ void __launch_bounds__(128) nbr_fill_sort_kernel(int nx, int ny, int nz, const int *__restrict__ cstart,
                                                           const int *__restrict__ src_by_cell,
                                                           const int *__restrict__ list_start,
                                                           int *__restrict__ lists) {
    extern __shared__ int keys[];
    __shared__ int seg_start[28];
    const int c = blockIdx.x;
    const int m = list_start[c + 1] - list_start[c];
    if (m == 0) return;
    const int a = c / (ny * nz), b = (c / nz) % ny, d = c % nz;
    if (threadIdx.x == 0) {
        int acc = 0, s = 0;
        for (int da = -1; da <= 1; ++da)
            for (int db = -1; db <= 1; ++db)
                for (int dd = -1; dd <= 1; ++dd) {
                    int aa = a + da, bb = b + db, ee = d + dd;
                    seg_start[s++] = acc;
                    if (aa < 0 || aa >= nx || bb < 0 || bb >= ny || ee < 0 || ee >= nz) continue;
                    int cc = (aa * ny + bb) * nz + ee;
                    acc += cstart[cc + 1] - cstart[cc];
                }
        seg_start[27] = acc;
    }
    __syncthreads();
    {
        int s = 0;
        for (int da = -1; da <= 1; ++da)
            for (int db = -1; db <= 1; ++db)
                for (int dd = -1; dd <= 1; ++dd, ++s) {
                    int len = seg_start[s + 1] - seg_start[s];
                    if (len == 0) continue;
                    int cc = ((a + da) * ny + (b + db)) * nz + (d + dd);
                    const int *from = src_by_cell + cstart[cc];
                    for (int i = threadIdx.x; i < len; i += blockDim.x) keys[seg_start[s] + i] = from[i];
                }
    }
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = m + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int va = keys[i], vb = keys[ixj];
                    bool up = ((i & k) == 0);
                    if ((va > vb) == up) { keys[i] = vb; keys[ixj] = va; }
                }
            }
            __syncthreads();
        }
    }
    int *out = lists + list_start[c];
    for (int i = threadIdx.x; i < m; i += blockDim.x) out[i] = keys[i];
}

__global__ void cell_minmax_kernel(const int *__restrict__ site_cell, int N, int *__restrict__ cell_min,
                                   int *__restrict__ cell_max) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = site_cell[i];
    atomicMin(cell_min + c, i);
    atomicMax(cell_max + c, i);
}

constexpr int CT = 128;    // target sites per CTA
constexpr int TILE = 128;  // sources per shared-memory tile

// potential_solver_gpu.cu:1541-1562: V_i = sum_j v_solve(1e-10*|r_ij|, q_j), ascending j, overwrite.
// One CTA = up to CT targets of one cell; sources = that cell's neighbourhood list, tiled through shared memory.
// Two phases per tile, because only ~15 % of the tested pairs are inside the cutoff and a thread-per-target loop that
// evaluates erfc / div under the test runs that path with ~5 of 32 lanes active:
//   1. every thread tests its target against the tile's sources on the SQUARED distance (d2 <= d2max is exactly
//      sqrt(d2) < cutoff: d2max is the largest double whose correctly rounded square root is below the cutoff, found on
//      the host) and queues the tile positions of its hits (one byte each, shared memory, [slot][thread]);
//   2. the threads evaluate their queued hits in lock-step, in queue (= ascending source) order, so a target's sum keeps
//      the reference's order and bits while the expensive path runs with most lanes active.
__global__ void __launch_bounds__(CT) coulomb_cell_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                         const double *__restrict__ z, const Source *__restrict__ src,
                                                         const int *__restrict__ blk_cell,
                                                         const int *__restrict__ blk_start,
                                                         const int *__restrict__ cell_tstart,
                                                         const int *__restrict__ titems,
                                                         const int *__restrict__ list_start,
                                                         const int *__restrict__ lists, double sigma, double k,
                                                         double d2max, int row_start, int row_count,
                                                         double *__restrict__ pot,
                                                         unsigned long long *__restrict__ pair_counter) {
    __shared__ Source tile[TILE];
    __shared__ unsigned char hitq[TILE][CT];
    const int c = blk_cell[blockIdx.x];
    const int tpos = cell_tstart[c] + (blockIdx.x - blk_start[c]) * CT + threadIdx.x;
    const bool in_cell = tpos < cell_tstart[c + 1];
    const int i = in_cell ? titems[tpos] : -1;
    const bool active = in_cell && i >= row_start && i < row_start + row_count;
    if (!__syncthreads_or((int)active)) return;  // row-sharded call: none of this CTA's targets belongs to this rank
    double xi = 0, yi = 0, zi = 0;
    if (active) { xi = x[i]; yi = y[i]; zi = z[i]; }
    const int ls = list_start[c], le = list_start[c + 1];
    double local = 0.0;
    int hits = 0;  // pairs inside the cutoff (the ones that pay erfc / div)
    for (int base = ls; base < le; base += TILE) {
        int nt = min(TILE, le - base);
        __syncthreads();
        if (threadIdx.x < nt) tile[threadIdx.x] = src[lists[base + threadIdx.x]];
        __syncthreads();
        int nq = 0;
        if (active) {
#pragma unroll 4
            for (int t = 0; t < nt; ++t) {
                const double dx = tile[t].x - xi, dy = tile[t].y - yi, dz = tile[t].z - zi;
                const double d2 = dx * dx + dy * dy + dz * dz;  // the argument of kmc_dist_nopbc's sqrt, same roundings
                if (d2 <= d2max && tile[t].idx != i) hitq[nq++][threadIdx.x] = (unsigned char)t;
            }
        }
        const int nq_warp = __reduce_max_sync(KMC_FULL_MASK, nq);
        for (int q = 0; q < nq_warp; ++q) {
            if (q < nq) {
                const Source s = tile[hitq[q][threadIdx.x]];
                const double dist = 1e-10 * kmc_dist_nopbc(xi, yi, zi, s.x, s.y, s.z);
                local += kmc_v_solve(dist, s.charge, sigma, k);
            }
        }
        hits += nq;
    }
    if (active) pot[i] = local;
    if (pair_counter) {
        if (threadIdx.x == 0) atomicAdd(pair_counter, (unsigned long long)(le - ls) * CT);
        const int wh = __reduce_add_sync(KMC_FULL_MASK, hits);
        if ((threadIdx.x & 31) == 0 && wh) atomicAdd(pair_counter + 3, (unsigned long long)wh);
    }
}

__global__ void sum_kernel(double *__restrict__ a, const double *__restrict__ b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

}  // namespace

// largest double whose (correctly rounded) square root is still below the cutoff: d2 <= d2max <=> sqrt(d2) < cutoff
extern "C" double kmcb200_cutoff_d2max(double cutoff) {
    if (!(cutoff > 0.0)) return -1.0;  // no pair is in range (d2 >= 0)
    double d2max = cutoff * cutoff;
    while (sqrt(d2max) >= cutoff) d2max = nextafter(d2max, 0.0);
    while (sqrt(nextafter(d2max, INFINITY)) < cutoff) d2max = nextafter(d2max, INFINITY);
    return d2max;
}

extern "C" int kmcb200_update_charge(kmcb200_ctx *ctx, const int *element, int *charge, const int *neigh, int N, int nn,
                                     const int *metals_host, int num_metals, int row_start, int row_count) {
    KMC_CHECK_ARG(ctx && element && charge && (neigh || row_count == 0), "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    KMC_CHECK_ARG(num_metals >= 0 && num_metals <= KMCB200_MAX_METALS && (num_metals == 0 || metals_host), "metals");
    unsigned metal_mask = 0;
    for (int m = 0; m < num_metals; ++m) metal_mask |= 1u << metals_host[m];
    if (row_count == 0) return 0;
    kmc_count_launch();
    update_charge_kernel<<<(row_count + 255) / 256, 256, 0, ctx->stream>>>(element, charge, neigh, nn, metal_mask,
                                                                          row_start, row_count);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

// Static plan for one (positions, N, cutoff): 20 A cell of every site, sites grouped by cell, CTA -> cell map.
struct CoulombPlan {
    const double *x = nullptr;
    int N = 0;
    double cutoff = 0;
    unsigned long long checksum = 0;
    CellGridDev g;
    int ncell = 0, nblocks = 0;
    int *site_cell = nullptr, *cell_tstart = nullptr, *titems = nullptr, *blk_start = nullptr, *blk_cell = nullptr;
    int *list_start = nullptr;  // ncell+1 (per step)
    int *cell_min = nullptr, *cell_max = nullptr;  // smallest / largest site id of a cell (static)
    int *src_cstart = nullptr, *src_fill = nullptr;  // ncell+1 each (per step)
};
// the plan is owned by the context (kmcb200_ctx::coulomb_plan): two contexts on one device do not share or free each
// other's arrays, and kmcb200_destroy releases it
static void free_plan_arrays(CoulombPlan &P) {
    cudaFree(P.site_cell); cudaFree(P.cell_tstart); cudaFree(P.titems); cudaFree(P.blk_start); cudaFree(P.blk_cell);
    cudaFree(P.list_start); cudaFree(P.cell_min); cudaFree(P.cell_max); cudaFree(P.src_cstart); cudaFree(P.src_fill);
    P = CoulombPlan();
}
void kmc_coulomb_plan_free(kmcb200_ctx *ctx) {
    if (!ctx->coulomb_plan) return;
    CoulombPlan *P = (CoulombPlan *)ctx->coulomb_plan;
    free_plan_arrays(*P);
    delete P;
    ctx->coulomb_plan = nullptr;
}

static int build_plan(kmcb200_ctx *ctx, CoulombPlan &P, int N, const double *x, const double *y, const double *z,
                      double cutoff, unsigned long long checksum) {
    if (P.x == x && P.N == N && P.cutoff == cutoff && P.checksum == checksum) return 0;
    if (P.site_cell) {
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        free_plan_arrays(P);
    }
    CellGridDev g;
    KMC_TRY(kmc_build_cellgrid(ctx, x, y, z, 0, N, cutoff, 0, nullptr, &g));  // scratch slots 0,1 hold start/items
    int ncell = g.nx * g.ny * g.nz;
    if (g.nx >= 1024 || g.ny >= 1024 || g.nz >= 1024) {
        kmc_set_error("Coulomb cell grid %dx%dx%d exceeds 1023 cells per axis", g.nx, g.ny, g.nz);
        return KMCB200_E_CAPACITY;
    }
    KMC_CUDA(cudaMalloc(&P.site_cell, (size_t)N * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.cell_tstart, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.titems, (size_t)N * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.blk_start, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.list_start, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.cell_min, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.cell_max, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.src_cstart, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMalloc(&P.src_fill, (size_t)(ncell + 1) * sizeof(int)));
    KMC_CUDA(cudaMemsetAsync(P.cell_min, 0x7f, (size_t)(ncell + 1) * sizeof(int), ctx->stream));
    KMC_CUDA(cudaMemsetAsync(P.cell_max, 0xff, (size_t)(ncell + 1) * sizeof(int), ctx->stream));
    KMC_CUDA(cudaMemcpyAsync(P.cell_tstart, g.cell_start, (size_t)(ncell + 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    KMC_CUDA(cudaMemcpyAsync(P.titems, g.items, (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    kmc_count_launch();
    site_cell_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(g, x, y, z, N, P.site_cell);
    kmc_count_launch();
    cell_minmax_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(P.site_cell, N, P.cell_min, P.cell_max);
    kmc_count_launch();
    cell_blocks_kernel<<<(ncell + 1 + 255) / 256, 256, 0, ctx->stream>>>(P.cell_tstart, ncell, CT, P.blk_start);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(kmc_exclusive_scan_i32(ctx, P.blk_start, P.blk_start, (long long)ncell + 1, 4));
    int nblocks = 0;
    KMC_CUDA(cudaMemcpyAsync(&nblocks, P.blk_start + ncell, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    KMC_CUDA(cudaMalloc(&P.blk_cell, (size_t)(nblocks + 1) * sizeof(int)));
    kmc_count_launch();
    block_map_kernel<<<(ncell + 255) / 256, 256, 0, ctx->stream>>>(P.blk_start, ncell, P.blk_cell);
    KMC_CUDA(cudaGetLastError());
    P.g = g;
    P.g.cell_start = P.cell_tstart;
    P.g.items = P.titems;
    P.ncell = ncell;
    P.nblocks = nblocks;
    P.x = x; P.N = N; P.cutoff = cutoff; P.checksum = checksum;
    return 0;
}

extern "C" int kmcb200_poisson_gridless(kmcb200_ctx *ctx, int N, const double *x, const double *y, const double *z,
                                        const int *element, const int *charge, double sigma, double k,
                                        double cutoff_radius, int row_start, int row_count,
                                        double *site_potential_charge) {
    KMC_CHECK_ARG(ctx && x && y && z && element && charge && site_potential_charge, "null pointer");
    KMC_CHECK_ARG(row_start >= 0 && row_count >= 0 && row_start + row_count <= N, "row range");
    if (row_count == 0) return 0;
    if (!ctx->coulomb_plan) ctx->coulomb_plan = new CoulombPlan();
    CoulombPlan &P = *(CoulombPlan *)ctx->coulomb_plan;
    // 1. charged sites, ascending site order (+ coordinate checksum for the cached plan)
    int *offs = nullptr, *src_cell = nullptr, *lists = nullptr;
    Source *src = nullptr;
    unsigned long long *csum = nullptr;
    KMC_TRY(kmc_scratch(ctx, 6, (size_t)(N + 1) * sizeof(int), (void **)&offs));
    KMC_TRY(kmc_scratch(ctx, 10, 64, (void **)&csum));
    KMC_CUDA(cudaMemsetAsync(csum, 0, 32, ctx->stream));  // [0] pair tests, [1] checksum, [2] d_max, [3] pairs in range
    kmc_count_launch();
    pos_checksum_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(x, y, z, N, csum + 1);
    kmc_count_launch();
    charged_flag_kernel<<<(N + 1 + 255) / 256, 256, 0, ctx->stream>>>(element, charge, N, offs);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(kmc_exclusive_scan_i32(ctx, offs, offs, (long long)N + 1, 4));  // offs[N] = Q
    KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, offs + N, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaMemcpyAsync((char *)ctx->h_mail + 8, csum + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    int Q = *(int *)ctx->h_mail;
    unsigned long long checksum = *(unsigned long long *)((char *)ctx->h_mail + 8);
    KMC_TRY(build_plan(ctx, P, N, x, y, z, cutoff_radius, checksum));
    KMC_TRY(kmc_scratch(ctx, 7, (size_t)(Q + 1) * sizeof(Source), (void **)&src));
    KMC_TRY(kmc_scratch(ctx, 8, (size_t)(2 * Q + 2) * sizeof(int), (void **)&src_cell));
    int *src_by_cell = src_cell + Q + 1;
    KMC_CUDA(cudaMemsetAsync(P.src_cstart, 0, (size_t)(P.ncell + 1) * sizeof(int), ctx->stream));
    KMC_CUDA(cudaMemsetAsync(P.src_fill, 0, (size_t)(P.ncell + 1) * sizeof(int), ctx->stream));
    kmc_count_launch();
    charged_scatter_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(charge, element, x, y, z, N, offs, P.site_cell, src,
                                                                    src_cell, P.src_cstart);
    KMC_CUDA(cudaGetLastError());
    // 2. bin the sources by cell, then the sorted 27-cell neighbourhood list of every cell that holds requested targets
    KMC_TRY(kmc_exclusive_scan_i32(ctx, P.src_cstart, P.src_cstart, (long long)P.ncell + 1, 4));
    if (Q > 0) {
        kmc_count_launch();
        src_bin_fill_kernel<<<(Q + 255) / 256, 256, 0, ctx->stream>>>(Q, src_cell, P.src_cstart, P.src_fill, src_by_cell);
    }
    int *d_max = (int *)(csum + 2);
    KMC_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    kmc_count_launch();
    nbr_count_kernel<<<(P.ncell + 1 + 255) / 256, 256, 0, ctx->stream>>>(P.g.nx, P.g.ny, P.g.nz, P.src_cstart, P.cell_min,
                                                                        P.cell_max, row_start, row_start + row_count,
                                                                        P.list_start, d_max);
    KMC_CUDA(cudaGetLastError());
    KMC_TRY(kmc_exclusive_scan_i32(ctx, P.list_start, P.list_start, (long long)P.ncell + 1, 4));
    KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, P.list_start + P.ncell, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaMemcpyAsync((char *)ctx->h_mail + 8, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    int total = *(int *)ctx->h_mail;
    int max_nbr = *(int *)((char *)ctx->h_mail + 8);
    KMC_TRY(kmc_scratch(ctx, 9, (size_t)(total + 1) * sizeof(int), (void **)&lists));
    if (total > 0) {
        int n2 = 1;
        while (n2 < max_nbr) n2 <<= 1;
        size_t dyn = (size_t)n2 * sizeof(int);
        if (dyn > 200 * 1024) {
            kmc_set_error("Coulomb sum: %d charged sources around one 20 A cell exceed the shared-memory sort capacity", max_nbr);
            return KMCB200_E_CAPACITY;
        }
        if (dyn > ctx->smem_cfg_coulomb) {
            KMC_CUDA(cudaFuncSetAttribute(nbr_fill_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            ctx->smem_cfg_coulomb = dyn;
        }
        kmc_count_launch();
        nbr_fill_sort_kernel<<<P.ncell, 128, dyn, ctx->stream>>>(P.g.nx, P.g.ny, P.g.nz, P.src_cstart, src_by_cell, P.list_start,
                                                               lists);
        KMC_CUDA(cudaGetLastError());
    }
    // 3. the pair sum
    unsigned long long *pairs = csum;  // csum[0] (tests) and csum[3] (in range, via pairs_in) were zeroed above
    if (P.nblocks > 0) {
        const double d2max = kmcb200_cutoff_d2max(cutoff_radius);
        kmc_count_launch();
        coulomb_cell_kernel<<<P.nblocks, CT, 0, ctx->stream>>>(x, y, z, src, P.blk_cell, P.blk_start, P.cell_tstart,
                                                              P.titems, P.list_start, lists, sigma, k, d2max,
                                                              row_start, row_count, site_potential_charge, pairs);
        KMC_CUDA(cudaGetLastError());
    }
    ctx->last_num_charged = Q;
    ctx->pair_counter_dev = pairs;
    return 0;
}

extern "C" int kmcb200_poisson_stats(kmcb200_ctx *ctx, long long *num_charged, long long *pair_tests,
                                     long long *pairs_in_range) {
    KMC_CHECK_ARG(ctx != nullptr, "ctx");
    if (num_charged) *num_charged = ctx->last_num_charged;
    if (pair_tests) *pair_tests = 0;
    if (pairs_in_range) *pairs_in_range = 0;
    if ((pair_tests || pairs_in_range) && ctx->pair_counter_dev) {
        KMC_CUDA(cudaMemcpyAsync(ctx->h_mail, ctx->pair_counter_dev, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        KMC_CUDA(cudaStreamSynchronize(ctx->stream));
        const unsigned long long *h = (const unsigned long long *)ctx->h_mail;
        if (pair_tests) *pair_tests = (long long)h[0];
        if (pairs_in_range) *pairs_in_range = (long long)h[3];
    }
    return 0;
}

extern "C" int kmcb200_sum_potential(kmcb200_ctx *ctx, int N, double *site_potential_charge,
                                     const double *site_potential_boundary) {
    KMC_CHECK_ARG(ctx && site_potential_charge && site_potential_boundary && N >= 0, "arguments");
    if (N == 0) return 0;
    kmc_count_launch();
    sum_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(site_potential_charge, site_potential_boundary, N);
    KMC_CUDA(cudaGetLastError());
    return 0;
}
