// kmc_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see kmc_oracle.h header comment).
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC   (oracle/Makefile)
// -ffp-contract=off is REQUIRED: every fused multiply-add in the spec is an explicit
// std::fma() call; everything else is separately rounded, exactly like the CUDA side
// (compiled with --fmad=false + explicit fma()).
//
// Citations are relative to /root/reference.

#include "kmc_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- distances: gpu_solvers.h:280-319 --------------------------------------------
// pow(a,2) is a*a exactly; each product/sum separately rounded (no contraction).
inline double dist_nopbc(double x1, double y1, double z1, double x2, double y2, double z2) {
    double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
inline double dist_pbc(double x1, double y1, double z1, double x2, double y2, double z2,
                       double latty, double lattz, int pbc) {
    if (pbc == 1) {
        double dist_x = x1 - x2;
        double fy = (y1 - y2) / latty;
        fy -= std::round(fy);
        double fz = (z1 - z2) / lattz;
        fz -= std::round(fz);
        double dy = fy * latty, dz = fz * lattz;
        return std::sqrt(dist_x * dist_x + dy * dy + dz * dz);
    }
    return dist_nopbc(x1, y1, z1, x2, y2, z2);
}

// gpu_solvers.h:321-328 (same left-to-right evaluation order)
inline double v_solve(double r_dist, int charge, double sigma, double k) {
    const double q = 1.60217663e-19;
    return (double)charge * std::erfc(r_dist / (sigma * std::sqrt(2.0))) * k * q / r_dist;
}

inline bool is_metal(int el, const int *metals, int num_metals) {
    for (int m = 0; m < num_metals; ++m)
        if (metals[m] == el) return true;
    return false;
}
inline bool possibly_charged(int el) {  // neighbor_lists_gpu.cu:96,123
    return el == ORC_OXYGEN_DEFECT || el == ORC_O || el == ORC_VACANCY || el == ORC_DEFECT;
}

// ---- cell grid used only as a candidate enumerator (same predicate afterwards) ----
struct CellGrid {
    double x0 = 0, y0 = 0, z0 = 0, hx = 1, hy = 1, hz = 1;
    int nx = 1, ny = 1, nz = 1;
    bool wrap_yz = false;
    double latty = 1, lattz = 1;
    std::vector<int> start, items;

    inline int cx(double x) const { int c = (int)std::floor((x - x0) / hx); return std::min(std::max(c, 0), nx - 1); }
    inline int cy(double y) const {
        if (wrap_yz) { double f = y / latty; f -= std::floor(f); int c = (int)(f * ny); return std::min(std::max(c, 0), ny - 1); }
        int c = (int)std::floor((y - y0) / hy); return std::min(std::max(c, 0), ny - 1);
    }
    inline int cz(double z) const {
        if (wrap_yz) { double f = z / lattz; f -= std::floor(f); int c = (int)(f * nz); return std::min(std::max(c, 0), nz - 1); }
        int c = (int)std::floor((z - z0) / hz); return std::min(std::max(c, 0), nz - 1);
    }
    inline int cell(int a, int b, int c) const { return (a * ny + b) * nz + c; }

    // sites [first, first+count) are binned; item ids are global site ids
    void build(const double *x, const double *y, const double *z, int first, int count, double cutoff,
               bool pbc, const double *lattice) {
        double h = cutoff * 1.0001;
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300, zmin = 1e300, zmax = -1e300;
        for (int i = first; i < first + count; ++i) {
            xmin = std::min(xmin, x[i]); xmax = std::max(xmax, x[i]);
            ymin = std::min(ymin, y[i]); ymax = std::max(ymax, y[i]);
            zmin = std::min(zmin, z[i]); zmax = std::max(zmax, z[i]);
        }
        if (count == 0) { xmin = xmax = ymin = ymax = zmin = zmax = 0; }
        x0 = xmin; y0 = ymin; z0 = zmin;
        nx = std::max(1, (int)std::floor((xmax - xmin) / h) + 1); hx = h;
        wrap_yz = pbc;
        if (pbc) {
            latty = lattice[1]; lattz = lattice[2];
            ny = std::max(1, (int)std::floor(latty / h));
            nz = std::max(1, (int)std::floor(lattz / h));
        } else {
            ny = std::max(1, (int)std::floor((ymax - ymin) / h) + 1); hy = h;
            nz = std::max(1, (int)std::floor((zmax - zmin) / h) + 1); hz = h;
        }
        size_t ncell = (size_t)nx * ny * nz;
        start.assign(ncell + 1, 0);
        std::vector<int> cid(count);
        for (int i = 0; i < count; ++i) {
            int g = first + i;
            cid[i] = cell(cx(x[g]), cy(y[g]), cz(z[g]));
            start[cid[i] + 1]++;
        }
        for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
        items.resize(count);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int i = 0; i < count; ++i) items[fill[cid[i]]++] = first + i;  // ascending id inside a cell
    }

    // collect candidate ids around (x,y,z) into out (unsorted)
    template <class F>
    void for_each_candidate(double px, double py, double pz, F &&f) const {
        int a = cx(px), b = cy(py), c = cz(pz);
        int ys[3], zs[3], nys = 0, nzs = 0;
        for (int d = -1; d <= 1; ++d) {
            int v = b + d;
            if (wrap_yz) v = ((v % ny) + ny) % ny; else if (v < 0 || v >= ny) continue;
            bool dup = false; for (int q = 0; q < nys; ++q) dup |= (ys[q] == v);
            if (!dup) ys[nys++] = v;
        }
        for (int d = -1; d <= 1; ++d) {
            int v = c + d;
            if (wrap_yz) v = ((v % nz) + nz) % nz; else if (v < 0 || v >= nz) continue;
            bool dup = false; for (int q = 0; q < nzs; ++q) dup |= (zs[q] == v);
            if (!dup) zs[nzs++] = v;
        }
        for (int da = -1; da <= 1; ++da) {
            int aa = a + da;
            if (aa < 0 || aa >= nx) continue;
            for (int q = 0; q < nys; ++q)
                for (int r = 0; r < nzs; ++r) {
                    int cc = cell(aa, ys[q], zs[r]);
                    for (int s = start[cc]; s < start[cc + 1]; ++s) f(items[s]);
                }
        }
    }
};

// ---- summation spec --------------------------------------------------
// 32-lane butterfly: v += shfl_xor(v, off) for off = 16,8,4,2,1; result of lane 0.
inline double warp_xor_reduce32(const double *v) {
    double a[32], b[32];
    for (int l = 0; l < 32; ++l) a[l] = v[l];
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) b[l] = a[l] + a[l ^ off];
        for (int l = 0; l < 32; ++l) a[l] = b[l];
    }
    return a[0];
}
// 256 values (thread t holds v[t]): 8 warp butterflies, then warp sums added in warp order.
inline double chunk_reduce_256(const double *v) {
    double s = warp_xor_reduce32(v);
    for (int w = 1; w < 8; ++w) s = s + warp_xor_reduce32(v + 32 * w);
    return s;
}
// cross-chunk combine: thread t sums partials t, t+256, ... sequentially from 0.0, then chunk_reduce_256.
inline double final_reduce(const double *partials, long n) {
    double s[256];
    for (int t = 0; t < 256; ++t) {
        double acc = 0.0;
        for (long k = t; k < n; k += 256) acc = acc + partials[k];
        s[t] = acc;
    }
    return chunk_reduce_256(s);
}

// Cross-chunk combine of the dot products (summation spec).  Up to 256 chunks (65 536 rows): the
// single-level final_reduce.  Above: two levels -- groups of ORC_DOT_GROUP = 64 consecutive chunks are reduced first (a
// group = chunk_reduce_256 of its 64 partials padded with zeros), then final_reduce runs over the group totals.  On
// several GPUs only the group totals cross NVLink (64x fewer values than the chunk partials); the result is independent
// of the GPU count when rank boundaries are multiples of 64 chunks.
double combine_partials(const double *partials, long nchunks) {
    if (nchunks <= 256) return final_reduce(partials, nchunks);
    long ngroups = (nchunks + ORC_DOT_GROUP - 1) / ORC_DOT_GROUP;
    std::vector<double> gt((size_t)ngroups);
    for (long g = 0; g < ngroups; ++g) {
        double v[256];
        for (int t = 0; t < 256; ++t) {
            long c = g * ORC_DOT_GROUP + t;
            v[t] = (t < ORC_DOT_GROUP && c < nchunks) ? partials[c] : 0.0;
        }
        gt[g] = chunk_reduce_256(v);
    }
    return final_reduce(gt.data(), ngroups);
}

double dot_spec(const double *u, const double *v, long n) {
    long nchunks = (n + ORC_CHUNK - 1) / ORC_CHUNK;
    std::vector<double> partials((size_t)std::max<long>(nchunks, 1), 0.0);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < nchunks; ++c) {
        double vals[ORC_CHUNK];
        for (int t = 0; t < ORC_CHUNK; ++t) {
            long i = c * ORC_CHUNK + t;
            vals[t] = (i < n) ? u[i] * v[i] : 0.0;
        }
        partials[c] = chunk_reduce_256(vals);
    }
    return combine_partials(partials.data(), nchunks);
}

void spmv_spec(long n, const int *row_ptr, const int *col, const double *data, const double *x, double *y,
               int lanes) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double acc[32];
        int s = row_ptr[i], e = row_ptr[i + 1];
        for (int l = 0; l < lanes; ++l) {
            double a = 0.0;
            for (int k = s + l; k < e; k += lanes) a = std::fma(data[k], x[col[k]], a);
            acc[l] = a;
        }
        double tmp[32];
        for (int off = lanes / 2; off >= 1; off >>= 1) {
            for (int l = 0; l < lanes; ++l) tmp[l] = acc[l] + acc[l ^ off];
            for (int l = 0; l < lanes; ++l) acc[l] = tmp[l];
        }
        y[i] = acc[0];
    }
}

// ---- hierarchical event sums -------------------------------------
// scan_256: inclusive scan of 256 values as ONE warp computes it on the GPU.  Lane l (0..31) owns the 8
// consecutive elements 8l..8l+7 and sums them sequentially (a[l][k]); the 32 lane totals are scanned with
// Kogge-Stone (d = 1,2,4,8,16: if lane >= d: S += S[lane-d]); incl[8l+k] = S[l-1] + a[l][k] (a[0][k] for l = 0).
// Returns the group total S[31] (the Kogge-Stone value; it may differ from incl[255] in the last bit).
double scan_256(const double *v, double *incl) {
    double a[32][8], S[32], nS[32];
    for (int l = 0; l < 32; ++l) {
        a[l][0] = v[8 * l];
        for (int k = 1; k < 8; ++k) a[l][k] = a[l][k - 1] + v[8 * l + k];
        S[l] = a[l][7];
    }
    for (int d = 1; d <= 16; d <<= 1) {
        for (int l = 0; l < 32; ++l) nS[l] = (l >= d) ? (S[l - d] + S[l]) : S[l];
        for (int l = 0; l < 32; ++l) S[l] = nS[l];
    }
    if (incl)
        for (int l = 0; l < 32; ++l)
            for (int k = 0; k < 8; ++k) incl[8 * l + k] = (l > 0) ? (S[l - 1] + a[l][k]) : a[l][k];
    return S[31];
}
void block_scan_256(const double *v, double *incl) { scan_256(v, incl); }

struct EventSums {
    int N, nn;
    long nchunk, nsuper;
    std::vector<double> rowsum, chunksum, supersum, topcum;
    double total = 0.0;

    // row sum: lane l holds p[l] + p[l+32] (0 beyond nn), then the 32-lane butterfly (xor 16,8,4,2,1)
    inline double row_sum(const double *prob, long r) const {
        const double *p = prob + r * (long)nn;
        double v[32];
        for (int l = 0; l < 32; ++l) {
            double p0 = (l < nn) ? p[l] : 0.0;
            double p1 = (l + 32 < nn) ? p[l + 32] : 0.0;
            v[l] = p0 + p1;
        }
        return warp_xor_reduce32(v);
    }
    void chunk_vals(long c, double *v) const {
        for (int t = 0; t < 256; ++t) {
            long r = c * 256 + t;
            v[t] = (r < N) ? rowsum[r] : 0.0;
        }
    }
    void super_vals(long s, double *v) const {
        for (int t = 0; t < 256; ++t) {
            long c = s * 256 + t;
            v[t] = (c < nchunk) ? chunksum[c] : 0.0;
        }
    }
    void recompute_chunk(long c) {
        double v[256];
        chunk_vals(c, v);
        chunksum[c] = scan_256(v, nullptr);
    }
    void recompute_super(long s) {
        double v[256];
        super_vals(s, v);
        supersum[s] = scan_256(v, nullptr);
    }
    // top level: one more block_scan_256 over the (<= 256) super sums; topcum = its inclusive prefixes
    void recompute_top() {
        double v[256];
        for (int t = 0; t < 256; ++t) v[t] = (t < nsuper) ? supersum[t] : 0.0;
        topcum.resize(256);
        total = scan_256(v, topcum.data());
    }
    void build(int N_, int nn_, const double *prob) {
        N = N_; nn = nn_;
        nchunk = (N + 255) / 256;
        nsuper = (nchunk + 255) / 256;
        if (nsuper > 256) { std::fprintf(stderr, "oracle: event hierarchy supports <= 16.7M sites\n"); std::abort(); }
        rowsum.resize(N); chunksum.resize(nchunk); supersum.resize(nsuper); topcum.resize(nsuper);
#pragma omp parallel for schedule(static)
        for (long r = 0; r < N; ++r) rowsum[r] = row_sum(prob, r);
#pragma omp parallel for schedule(static)
        for (long c = 0; c < nchunk; ++c) recompute_chunk(c);
        for (long s = 0; s < nsuper; ++s) recompute_super(s);
        recompute_top();
    }
    double psum() const { return total; }

    // first index t in incl[0..255] with incl[t] > number, else last t with v[t] > 0 (clamp)
    static int pick(const double *v, const double *incl, double number) {
        for (int t = 0; t < 256; ++t)
            if (incl[t] > number) return t;
        for (int t = 255; t >= 0; --t)
            if (v[t] > 0.0) return t;
        return -1;
    }
    long select(const double *prob, double number) const {
        if (!(psum() > 0.0)) return -1;
        double tv[256];
        for (int t = 0; t < 256; ++t) tv[t] = (t < nsuper) ? supersum[t] : 0.0;
        long s = pick(tv, topcum.data(), number);
        if (s < 0) return -1;
        if (s > 0) number = number - topcum[s - 1];
        double v[256], incl[256];
        super_vals(s, v);
        block_scan_256(v, incl);
        int tc = pick(v, incl, number);
        if (tc < 0) return -1;
        if (tc > 0) number = number - incl[tc - 1];
        long c = s * 256 + tc;
        chunk_vals(c, v);
        block_scan_256(v, incl);
        int tr = pick(v, incl, number);
        if (tr < 0) return -1;
        if (tr > 0) number = number - incl[tr - 1];
        long r = c * 256 + tr;
        const double *p = prob + r * (long)nn;
        double acc = p[0];
        int sel = -1;
        for (int n = 0; n < nn; ++n) {
            if (n > 0) acc = acc + p[n];
            if (acc > number) { sel = n; break; }
        }
        if (sel < 0) for (int n = nn - 1; n >= 0; --n) if (p[n] > 0.0) { sel = n; break; }
        if (sel < 0) return -1;
        return r * (long)nn + sel;
    }
};

struct Rng {  // random_num.h:4-26
    std::mt19937 gen;
    std::uniform_real_distribution<double> dist{0.0, 1.0};
    explicit Rng(unsigned seed) : gen(0) { gen.seed(seed); }
    double next() { return dist(gen); }
};

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// (torchrun exports OMP_NUM_THREADS=1 before the process starts; the CPU arm asks for all the cores it may use)
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------
// a1: neighbor_lists_gpu.cu:55-78 (kernel), :257-290 (nn_dist = 3.5 and nn = 52 are passed by
// the caller here; the reference hard-codes them at :265-266).
void orc_neighbor_list(int N, const double *x, const double *y, const double *z, double nn_dist, int nn,
                       int row_start, int row_count, int use_cells, int *out) {
    CellGrid grid;
    if (use_cells) grid.build(x, y, z, 0, N, nn_dist, false, nullptr);
#pragma omp parallel
    {
        std::vector<int> cand;
#pragma omp for schedule(dynamic, 64)
        for (int idx = 0; idx < row_count; ++idx) {
            int i = idx + row_start;
            int *row = out + (size_t)idx * nn;
            for (int n = 0; n < nn; ++n) row[n] = -1;
            int counter = 0;
            if (!use_cells) {
                for (int j = 0; j < N; ++j) {
                    double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
                    bool neighbor = (d < nn_dist && i != j);
                    if (neighbor && counter < nn) row[counter++] = j;
                }
            } else {
                cand.clear();
                grid.for_each_candidate(x[i], y[i], z[i], [&](int j) {
                    double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
                    if (d < nn_dist && i != j) cand.push_back(j);
                });
                std::sort(cand.begin(), cand.end());
                for (int j : cand)
                    if (counter < nn) row[counter++] = j;
            }
        }
    }
}

// a2: neighbor_lists_gpu.cu:80-104
void orc_cutoff_count(int N, const int *element, const double *x, const double *y, const double *z,
                      double cutoff, int row_start, int row_count, int *count_out) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int idx = 0; idx < row_count; ++idx) {
        int i = idx + row_start;
        int c = 0;
        for (int j = 0; j < N; ++j) {
            double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
            if (d < cutoff && i != j && possibly_charged(element[j])) c++;
        }
        count_out[idx] = c;
    }
}
// a2: neighbor_lists_gpu.cu:107-136
void orc_cutoff_list(int N, const int *element, const double *x, const double *y, const double *z,
                     double cutoff, int max_num_cutoff, int row_start, int row_count, int *idx_out) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int idx = 0; idx < row_count; ++idx) {
        int i = idx + row_start;
        int *row = idx_out + (size_t)idx * max_num_cutoff;
        for (int n = 0; n < max_num_cutoff; ++n) row[n] = -1;
        int counter = 0;
        for (int j = 0; j < N; ++j) {
            double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
            if (d < cutoff && possibly_charged(element[j]) && counter < max_num_cutoff && i != j) row[counter++] = j;
        }
    }
}

// a3: iterative_solvers_gpu.cu:96-124 (count), :126-157 (fill), :159-218 (driver)
long orc_block_sparsity(const double *x, const double *y, const double *z, const double *lattice, int pbc,
                        double cutoff, int size_i, int size_j, int start_i, int start_j, int use_cells,
                        int N_total, int *row_ptr, int *col_out) {
    (void)N_total;
    CellGrid grid;
    if (use_cells) grid.build(x, y, z, start_j, size_j, cutoff, pbc == 1, lattice);
    std::vector<int> nnz_row(size_i, 0);
    auto row_cols = [&](int row, std::vector<int> &cols) {
        cols.clear();
        int i = start_i + row;
        if (!use_cells) {
            for (int c = 0; c < size_j; ++c) {
                int j = start_j + c;
                double d = dist_pbc(x[i], y[i], z[i], x[j], y[j], z[j], lattice[1], lattice[2], pbc);
                if (d < cutoff) cols.push_back(c);
            }
        } else {
            grid.for_each_candidate(x[i], y[i], z[i], [&](int j) {
                double d = dist_pbc(x[i], y[i], z[i], x[j], y[j], z[j], lattice[1], lattice[2], pbc);
                if (d < cutoff) cols.push_back(j - start_j);
            });
            std::sort(cols.begin(), cols.end());
        }
    };
#pragma omp parallel
    {
        std::vector<int> cols;
#pragma omp for schedule(dynamic, 64)
        for (int row = 0; row < size_i; ++row) {
            row_cols(row, cols);
            nnz_row[row] = (int)cols.size();
        }
    }
    row_ptr[0] = 0;
    long acc = 0;
    for (int row = 0; row < size_i; ++row) { acc += nnz_row[row]; row_ptr[row + 1] = (int)acc; }
    if (col_out) {
#pragma omp parallel
        {
            std::vector<int> cols;
#pragma omp for schedule(dynamic, 64)
            for (int row = 0; row < size_i; ++row) {
                row_cols(row, cols);
                std::copy(cols.begin(), cols.end(), col_out + row_ptr[row]);
            }
        }
    }
    return acc;
}

// a5: potential_solver_gpu.cu:12-63.  (Vnn is per site: the reference launches >= one thread per
// site, :75-79, so its thread-level accumulator never spans two sites.)
void orc_update_charge(int N, int nn, const int *element, int *charge, const int *neigh, const int *metals,
                       int num_metals, int row_start, int row_end) {
    (void)N;
#pragma omp parallel for schedule(static)
    for (int i = row_start; i < row_end; ++i) {
        int idx = i - row_start;
        if (element[i] == ORC_VACANCY) {
            int c = 2, Vnn = 0;
            for (int s = idx * nn; s < (idx + 1) * nn; ++s) {
                int j = neigh[s];
                if (j >= 0) {
                    if (element[j] == ORC_VACANCY) Vnn++;
                    if (is_metal(element[j], metals, num_metals)) c = 0;
                    if (Vnn >= 2) c = 0;
                }
            }
            charge[i] = c;
        }
        if (element[i] == ORC_OXYGEN_DEFECT) {
            int c = -2;
            for (int s = idx * nn; s < (idx + 1) * nn; ++s) {
                int j = neigh[s];
                if (j >= 0 && is_metal(element[j], metals, num_metals)) c = 0;
            }
            charge[i] = c;
        }
    }
}

// a6.  Off-diagonals: calc_off_diagonal_dist (potential_solver_gpu.cu:246-285).
// Diagonal: reduce_rows_into_diag (:774-794) sums the row's stored values sequentially in column
// order (the diagonal slot holds 0 from the memset at :920) and does diag -= tmp; contact terms:
// reduce_contact_into_diag (:323-367), sequential in column order; insert_into_diag (:795-814)
// stores (diag + left) + right; inverse_diag (:817-830) 1/((diag+left)+right); rhs (:438-454)
// left*VL + right*VR with VL=-Vd/2, VR=+Vd/2 (:866-867).
void orc_assemble_K(int N, int N_left, int N_right, const int *element, const int *charge, const int *metals,
                    int num_metals, const int *row_ptr, const int *col, const int *left_row_ptr,
                    const int *left_col, const int *right_row_ptr, const int *right_col, double Vd, double high_G,
                    double low_G, double *data, double *inv_diag, double *rhs) {
    int n = N - N_left - N_right;
    double VL = -Vd / 2, VR = Vd / 2;
    auto cond = [&](int i, int j) -> double {
        bool metal1 = is_metal(element[i], metals, num_metals);
        bool metal2 = is_metal(element[j], metals, num_metals);
        bool cv1 = (element[i] == ORC_VACANCY) && (charge[i] == 0);
        bool cv2 = (element[j] == ORC_VACANCY) && (charge[j] == 0);
        return ((metal1 && metal2) || (cv1 && cv2)) ? high_G : low_G;
    };
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        int i = N_left + r;
        double tmp = 0.0;
        int diag_slot = -1;
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            int j = N_left + col[k];
            if (i != j) {
                data[k] = -cond(i, j);
            } else {
                data[k] = 0.0;
                diag_slot = k;
            }
            tmp += data[k];
        }
        double diag = 0.0;
        diag -= tmp;
        double left = 0.0;
        for (int k = left_row_ptr[r]; k < left_row_ptr[r + 1]; ++k) left += cond(i, 0 + left_col[k]);
        double right = 0.0;
        for (int k = right_row_ptr[r]; k < right_row_ptr[r + 1]; ++k) right += cond(i, N_left + n + right_col[k]);
        double d = diag + left + right;
        if (diag_slot >= 0) data[diag_slot] = d;
        inv_diag[r] = 1.0 / d;
        rhs[r] = left * VL + right * VR;
    }
}

double orc_dot(const double *u, const double *v, long n) { return dot_spec(u, v, n); }
void orc_spmv(long n, const int *row_ptr, const int *col, const double *data, const double *x, double *y,
              int lanes) {
    spmv_spec(n, row_ptr, col, data, x, y, lanes);
}

// a7: dist_conjugate_gradient.cpp:149-276 (update order preserved exactly; the vendor BLAS/SpMV
// summation orders are replaced by the summation spec).
int orc_pcg_jacobi(long n, const int *row_ptr, const int *col, const double *data, const double *inv_diag,
                   double *r, double *x, double tol, int max_it, int lanes, double *stats) {
    std::vector<double> p((size_t)n), Ap((size_t)n), z((size_t)n);
    double bb = dot_spec(r, r, n);                           // :187
    spmv_spec(n, row_ptr, col, data, x, Ap.data(), lanes);  // :191 (p <- x0 at :178)
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        r[i] = r[i] - Ap[i];           // :201 daxpy(-1)
        z[i] = r[i] * inv_diag[i];     // :204
    }
    double rz = dot_spec(r, z.data(), n);  // :212
    double r0 = 0.0;
    int k = 1;
    while (rz / bb > tol * tol && k <= max_it) {  // :217
        if (k > 1) {
            double b = rz / r0;  // :220
#pragma omp parallel for schedule(static)
            for (long i = 0; i < n; ++i) {
                double t = b * p[i];  // dscal :221
                p[i] = z[i] + t;      // daxpy alpha=1 :222
            }
        } else {
#pragma omp parallel for schedule(static)
            for (long i = 0; i < n; ++i) p[i] = z[i];  // :226
        }
        spmv_spec(n, row_ptr, col, data, p.data(), Ap.data(), lanes);  // :232
        double pAp = dot_spec(p.data(), Ap.data(), n);                 // :240
        double a = rz / pAp;                                           // :243
        double na = -a;                                                // :249
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            x[i] = std::fma(a, p[i], x[i]);    // :246
            r[i] = std::fma(na, Ap[i], r[i]);  // :250
            z[i] = r[i] * inv_diag[i];         // :254
        }
        r0 = rz;                          // :251
        rz = dot_spec(r, z.data(), n);    // :264
        k++;
    }
    if (stats) { stats[0] = rz; stats[1] = bb; }
    return k - 1;
}

// a8: potential_solver_gpu.cu:1541-1562.  The cutoff list (built once from the initial elements,
// neighbor_lists_gpu.cu:107-136) is not materialised: its membership test is evaluated inline.  The
// class {d,Od,V,O} is closed under all four events (kmc_events.cu:305-328), so testing the current
// element is equivalent to testing the initial one.
void orc_coulomb(int N, const double *x, const double *y, const double *z, const int *element, const int *charge,
                 double sigma, double k, double cutoff, int row_start, int row_count, double *pot) {
    std::vector<int> q;  // ascending j
    for (int j = 0; j < N; ++j)
        if (charge[j] != 0 && possibly_charged(element[j])) q.push_back(j);
#pragma omp parallel for schedule(dynamic, 256)
    for (int idx = 0; idx < row_count; ++idx) {
        int i = idx + row_start;
        double local = 0.0;
        for (int j : q) {
            if (i == j) continue;
            double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
            if (d < cutoff) {
                double dist = 1e-10 * d;
                local += v_solve(dist, charge[j], sigma, k);
            }
        }
        pot[i] = local;
    }
}

// a8 with a 20 A cell grid as the candidate enumerator (same predicate, same ascending-j summation order as
// orc_coulomb, hence bit-identical results; tests/test_oracle.py checks that).  This is the variant orc_superstep and the
// CPU baseline use: the all-sources loop above is O(N*Q).
void orc_coulomb_cells(int N, const double *x, const double *y, const double *z, const int *element,
                       const int *charge, double sigma, double k, double cutoff, int row_start, int row_count,
                       double *pot) {
    std::vector<int> q;  // ascending j
    for (int j = 0; j < N; ++j)
        if (charge[j] != 0 && possibly_charged(element[j])) q.push_back(j);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < N; ++i) {
        lo[0] = std::min(lo[0], x[i]); hi[0] = std::max(hi[0], x[i]);
        lo[1] = std::min(lo[1], y[i]); hi[1] = std::max(hi[1], y[i]);
        lo[2] = std::min(lo[2], z[i]); hi[2] = std::max(hi[2], z[i]);
    }
    if (N == 0) return;
    const double h = cutoff * 1.0001;
    int nc[3];
    for (int d = 0; d < 3; ++d) nc[d] = std::max(1, (int)std::floor((hi[d] - lo[d]) / h) + 1);
    auto cidx = [&](double v, int d) { int c = (int)std::floor((v - lo[d]) / h); return std::min(std::max(c, 0), nc[d] - 1); };
    size_t ncell = (size_t)nc[0] * nc[1] * nc[2];
    std::vector<int> start(ncell + 1, 0), items(q.size());
    std::vector<int> cid(q.size());
    for (size_t s = 0; s < q.size(); ++s) {
        int j = q[s];
        cid[s] = (cidx(x[j], 0) * nc[1] + cidx(y[j], 1)) * nc[2] + cidx(z[j], 2);
        start[cid[s] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (size_t s = 0; s < q.size(); ++s) items[fill[cid[s]]++] = q[s];  // ascending j inside a cell
    }
#pragma omp parallel
    {
        std::vector<int> cand;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < row_count; ++idx) {
            int i = idx + row_start;
            cand.clear();
            int a = cidx(x[i], 0), b = cidx(y[i], 1), c = cidx(z[i], 2);
            for (int da = -1; da <= 1; ++da) for (int db = -1; db <= 1; ++db) for (int dc = -1; dc <= 1; ++dc) {
                int aa = a + da, bb = b + db, cc = c + dc;
                if (aa < 0 || aa >= nc[0] || bb < 0 || bb >= nc[1] || cc < 0 || cc >= nc[2]) continue;
                size_t cell = ((size_t)aa * nc[1] + bb) * nc[2] + cc;
                for (int s = start[cell]; s < start[cell + 1]; ++s) cand.push_back(items[s]);
            }
            std::sort(cand.begin(), cand.end());
            double local = 0.0;
            for (int j : cand) {
                if (i == j) continue;
                double d = dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
                if (d < cutoff) {
                    double dist = 1e-10 * d;
                    local += v_solve(dist, charge[j], sigma, k);
                }
            }
            pot[i] = local;
        }
    }
}

// a10: kmc_events.cu:130-229
void orc_build_events(int N, int nn, const int *neigh, const int *layer, double T_bg, double freq, double sigma,
                      double k, const double *x, const double *y, const double *z, const double *pot,
                      const int *element, const int *charge, const double *E_gen, const double *E_rec,
                      const double *E_Vdiff, const double *E_Odiff, int row_start, int row_count, int *type_out,
                      double *prob_out) {
    const double kB = 8.617333262e-5;  // kmc_events.cu:5
    const double epsilon = 1e-200;     // :150
    long total = (long)row_count * nn;
#pragma omp parallel for schedule(static)
    for (long id = 0; id < total; ++id) {
        int ev = ORC_NULL_EVENT;
        double P = 0.0;
        int i = (int)(id / nn) + row_start;
        int j = neigh[id];
        if (j >= 0 && j < N) {
            double dist = 1e-10 * dist_nopbc(x[i], y[i], z[i], x[j], y[j], z[j]);
            if (element[i] == ORC_DEFECT && element[j] == ORC_O) {  // :158
                double E = 2 * (pot[i] - pot[j]);
                double EA = E_gen[layer[j]] - E - 0.0;
                ev = ORC_VACANCY_GENERATION;
                P = freq * (1 / (std::exp(EA / (kB * T_bg)) + epsilon));
            }
            if (element[i] == ORC_OXYGEN_DEFECT && element[j] == ORC_VACANCY) {  // :171
                double self_int_V = v_solve(dist, 2, sigma, k);
                int charge_state = charge[i] - charge[j];
                double E = charge_state * ((pot[i] - pot[j]) + (charge_state / 2) * self_int_V);
                double EA = E_rec[layer[j]] - E - 0.0;
                ev = ORC_VACANCY_RECOMBINATION;
                P = freq * (1 / (std::exp(EA / (kB * T_bg)) + epsilon));
            }
            if (element[i] == ORC_VACANCY && element[j] == ORC_O) {  // :188
                double self_int_V = 0.0;
                if (charge[i] != 0) self_int_V = v_solve(dist, charge[i], sigma, k);
                double E = (charge[i] - charge[j]) * ((pot[i] - pot[j]) + self_int_V);
                double EA = E_Vdiff[layer[j]] - E - 0.0;
                ev = ORC_VACANCY_DIFFUSION;
                P = freq * (1 / (std::exp(EA / (kB * T_bg)) + epsilon));
            }
            if (element[i] == ORC_OXYGEN_DEFECT && element[j] == ORC_DEFECT) {  // :207
                double self_int_V = 0.0;
                if (charge[i] != 0) self_int_V = v_solve(dist, 2, sigma, k);
                double E = (charge[i] - charge[j]) * ((pot[i] - pot[j]) - self_int_V);
                double EA = E_Odiff[layer[j]] - E - 0.0;
                ev = ORC_ION_DIFFUSION;
                P = freq * (1 / (std::exp(EA / (kB * T_bg)) + epsilon));
            }
        }
        type_out[id] = ev;
        prob_out[id] = P;
    }
}

// Device::makeSubstoichiometric (Device.cpp:180-211): n_add = int(concentration * #O); draw loc = int(u * N_atom) from
// mt19937(seed) until n_add oxygen ATOMS (sites that are neither d nor Od, in site order) have become vacancies.
int orc_make_substoichiometric(int N, int *element, double vacancy_concentration, unsigned rnd_seed) {
    std::vector<int> atom_ind;
    int num_O = 0;
    for (int i = 0; i < N; ++i) {
        if (element[i] == ORC_O) num_O++;
        if (element[i] != ORC_DEFECT && element[i] != ORC_OXYGEN_DEFECT) atom_ind.push_back(i);  // Device.cpp:116-143
    }
    const int N_atom = (int)atom_ind.size();
    int num_V_add = (int)(vacancy_concentration * num_O);
    const int added = num_V_add;
    Rng rng(rnd_seed);
    while (num_V_add > 0) {
        double random_num = rng.next();
        int loc = (int)(random_num * N_atom);
        if (element[atom_ind[loc]] == ORC_O) {
            element[atom_ind[loc]] = ORC_VACANCY;
            num_V_add--;
        }
    }
    return added;
}

void *orc_rng_create(unsigned seed) { return new Rng(seed); }
void orc_rng_destroy(void *rng) { delete (Rng *)rng; }
double orc_rng_next(void *rng) { return ((Rng *)rng)->next(); }
void orc_rng_get_state(void *rng, unsigned *mt624, int *pos) {
    // libstdc++ operator<< prints the 624 state words followed by the position
    std::stringstream ss;
    ss << ((Rng *)rng)->gen;
    for (int i = 0; i < 624; ++i) { unsigned long v; ss >> v; mt624[i] = (unsigned)v; }
    unsigned long p; ss >> p; *pos = (int)p;
}

void orc_block_scan_256(const double *v, double *incl) { block_scan_256(v, incl); }

long orc_select_event(int N, int nn, const double *prob, double number, double *psum_out) {
    EventSums es;
    es.build(N, nn, prob);
    if (psum_out) *psum_out = es.psum();
    return es.select(prob, number);
}

// a10: kmc_events.cu:448-516.  thrust::inclusive_scan + upper_bound are replaced by the
// hierarchical sums of the summation spec (same selection rule: first slot whose inclusive
// cumulative rate exceeds u*Psum).  Event application: execute_event :292-331; zero-out rule:
// zero_out_events_split :247-266.  Two RNG draws per event (:469, :515); the returned time is the
// LAST drawn residence time (:515, :562).
int orc_event_loop(int N, int nn, const int *neigh, int *type, double *prob, int *element, int *charge,
                   double freq, void *rng_, int max_events, int max_log, int *log_out, double *psum_out,
                   double *event_time_out) {
    Rng *rng = (Rng *)rng_;
    EventSums es;
    es.build(N, nn, prob);
    // reverse adjacency: slots whose neighbour is s (static; equivalent to the reference's full scan)
    std::vector<int> rev_ptr((size_t)N + 1, 0), rev_slot;
    {
        long total = (long)N * nn;
        for (long s = 0; s < total; ++s) if (neigh[s] >= 0) rev_ptr[neigh[s] + 1]++;
        for (int i = 0; i < N; ++i) rev_ptr[i + 1] += rev_ptr[i];
        rev_slot.resize(rev_ptr[N]);
        std::vector<int> fill(rev_ptr.begin(), rev_ptr.end() - 1);
        for (long s = 0; s < total; ++s) if (neigh[s] >= 0) rev_slot[fill[neigh[s]]++] = (int)s;
    }
    double event_time = 0.0;
    int n_events = 0;
    std::vector<long> dirty_rows;
    while (event_time < 1 / freq && (max_events <= 0 || n_events < max_events)) {
        double Psum = es.psum();
        double number = rng->next() * Psum;
        long slot = es.select(prob, number);
        if (slot >= 0) {
            int i = (int)(slot / nn);
            int j = neigh[slot];
            int t = type[slot];
            if (n_events < max_log && log_out) {
                log_out[4 * n_events + 0] = i; log_out[4 * n_events + 1] = j;
                log_out[4 * n_events + 2] = t; log_out[4 * n_events + 3] = (int)slot;
            }
            if (n_events < max_log && psum_out) psum_out[n_events] = Psum;
            if (t == ORC_VACANCY_GENERATION) {
                element[i] = ORC_OXYGEN_DEFECT; element[j] = ORC_VACANCY; charge[i] = -2; charge[j] = 2;
            } else if (t == ORC_VACANCY_RECOMBINATION) {
                element[i] = ORC_DEFECT; element[j] = ORC_O; charge[i] = 0; charge[j] = 0;
            } else if (t == ORC_VACANCY_DIFFUSION || t == ORC_ION_DIFFUSION) {
                std::swap(element[i], element[j]);
                std::swap(charge[i], charge[j]);
            }
            // zero out: every slot (r,n) with neigh>=0 and (r==i || nb==j || r==j || nb==i)
            dirty_rows.clear();
            for (int s : {i, j}) {
                for (int n = 0; n < nn; ++n) {
                    long sl = (long)s * nn + n;
                    if (neigh[sl] >= 0) { type[sl] = ORC_NULL_EVENT; prob[sl] = 0.0; }
                }
                dirty_rows.push_back(s);
                for (int q = rev_ptr[s]; q < rev_ptr[s + 1]; ++q) {
                    int sl = rev_slot[q];
                    type[sl] = ORC_NULL_EVENT; prob[sl] = 0.0;
                    dirty_rows.push_back(sl / nn);
                }
            }
            std::sort(dirty_rows.begin(), dirty_rows.end());
            dirty_rows.erase(std::unique(dirty_rows.begin(), dirty_rows.end()), dirty_rows.end());
            long last_c = -1, last_s = -1;
            std::vector<long> dc, ds;
            for (long r : dirty_rows) {
                es.rowsum[r] = es.row_sum(prob, r);
                if (r / 256 != last_c) { last_c = r / 256; dc.push_back(last_c); }
            }
            for (long c : dc) {
                es.recompute_chunk(c);
                if (c / 256 != last_s) { last_s = c / 256; ds.push_back(last_s); }
            }
            std::sort(ds.begin(), ds.end());
            ds.erase(std::unique(ds.begin(), ds.end()), ds.end());
            for (long s : ds) es.recompute_super(s);
            es.recompute_top();
        }
        event_time = -std::log(rng->next()) / Psum;
        n_events++;
    }
    *event_time_out = event_time;
    return n_events;
}

// kmc_main.cpp:328-540 on one rank: update_charge_gpu (:342) -> background_potential_gpu_sparse (:364)
// -> poisson_gridless_gpu (:405) -> sum_and_gather_potential (:479) -> execute_kmc_step_mpi (:491).
void orc_superstep(const orc_params *p, const double *x, const double *y, const double *z, const int *layer,
                   const int *neigh, const int *row_ptr, const int *col, const int *left_row_ptr,
                   const int *left_col, const int *right_row_ptr, const int *right_col, int *element, int *charge,
                   double *pot_boundary, double *pot_total, void *rng, int max_log, int *log_out,
                   orc_step_info *info) {
    int N = p->N, nn = p->nn, NL = p->N_left, NR = p->N_right;
    int n = N - NL - NR;
    double t0 = now_s();
    orc_update_charge(N, nn, element, charge, neigh, p->metals, p->num_metals, 0, N);
    double t1 = now_s();
    long nnz = row_ptr[n];
    std::vector<double> data((size_t)nnz), inv_diag((size_t)n), rhs((size_t)n);
    orc_assemble_K(N, NL, NR, element, charge, p->metals, p->num_metals, row_ptr, col, left_row_ptr, left_col,
                   right_row_ptr, right_col, p->Vd, p->high_G, p->low_G, data.data(), inv_diag.data(), rhs.data());
    double tol = p->cg_tol_per_row * n;  // potential_solver_gpu.cu:885
    info->cg_iterations = orc_pcg_jacobi(n, row_ptr, col, data.data(), inv_diag.data(), rhs.data(),
                                         pot_boundary + NL, tol, p->cg_max_it, p->spmv_lanes, nullptr);
    double t2 = now_s();
    orc_coulomb_cells(N, x, y, z, element, charge, p->sigma, p->k, p->cutoff_radius, 0, N, pot_total);
    double t3 = now_s();
    for (int i = 0; i < N; ++i) pot_total[i] += pot_boundary[i];  // potential_solver_gpu.cu:832-843,1147
    std::vector<int> type((size_t)N * nn);
    std::vector<double> prob((size_t)N * nn);
    orc_build_events(N, nn, neigh, layer, p->T_bg, p->freq, p->sigma, p->k, x, y, z, pot_total, element, charge,
                     p->E_gen, p->E_rec, p->E_Vdiff, p->E_Odiff, 0, N, type.data(), prob.data());
    double et = 0.0;
    info->n_events = orc_event_loop(N, nn, neigh, type.data(), prob.data(), element, charge, p->freq, rng, 0,
                                    max_log, log_out, nullptr, &et);
    info->event_time = et;
    double t4 = now_s();
    info->t_charge = t1 - t0; info->t_boundary = t2 - t1; info->t_coulomb = t3 - t2; info->t_events = t4 - t3;
}

}  // extern "C"
