// events.cu -- a10/a11: KMC event rates, hierarchical rate sums, device-resident residence-time loop.
// Reference: src/kmc_events.cu:130-229 (build_event_list_split), :247-266 (zero_out_events_split),
// :268-331 (read_out_event / execute_event), :333-563 (execute_kmc_step_mpi), :566-572 (copytoConstMemory);
// src/random_num.h:4-26 (host mt19937 + uniform_real_distribution<double>).
//
// The reference does, PER EVENT: thrust::inclusive_scan over all N*nn rates, a D2H of the total, a host RNG
// draw, thrust::upper_bound, two <<<1,64>>> kernels, a full zero-out pass over N*nn slots and three host syncs.
// Here the rates are reduced once per superstep into a 3-level hierarchy (row of nn slots -> chunk of 256 rows
// -> super of 256 chunks); ONE persistent CTA then runs the whole `while (event_time < 1/freq)` loop on the
// device: MT19937 draw, top-down selection (one inclusive prefix per level + first-greater search), event
// application, zero-out through a static reverse-neighbour index, and repair of only the touched partial sums.
// The selection rule is the reference's (first slot whose inclusive cumulative rate exceeds u*Psum); the
// association of the partial sums is the summation spec shared with the CPU oracle.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {
constexpr int EV_THREADS = 512;
constexpr int MAX_SUPER = 256;
}  // namespace

struct EvEnergies {
    double E_gen[KMCB200_MAX_LAYERS], E_rec[KMCB200_MAX_LAYERS], E_Vdiff[KMCB200_MAX_LAYERS], E_Odiff[KMCB200_MAX_LAYERS];
};

struct EvResult {
    double event_time;
    double psum_last;
    int n_events;
    int error;  // reserved (always 0: the per-event work lists are bounded by 2 * REV_STRIDE + 2 chunks)
};

struct kmcb200_events {
    kmcb200_ctx *ctx = nullptr;
    int N = 0, nn = 0;
    long long nchunk = 0, nsuper = 0;
    double *prob = nullptr;
    unsigned char *type = nullptr;
    double *rowsum = nullptr, *chunksum = nullptr, *supersum = nullptr;
    // inclusive scan_256 prefixes kept next to the sums, so the selector compares instead of re-scanning:
    // rowincl[256 c + t] over the rows of chunk c, chunkincl[256 s + t] over the chunks of super s (both padded)
    // (rowsum and rowincl share ONE allocation: rowincl = rowsum + nchunk*256, so a single L2 access-policy window covers both)
    double *rowincl = nullptr, *chunkincl = nullptr;
    size_t l2_window_bytes = 0;  // 0: no persisting-L2 window (not supported / disabled)
    float l2_hit_ratio = 1.0f;
    int *rev = nullptr;  // N * REV_STRIDE
    // revpos[r*nn + n] = position of slot (r, n) inside the reverse-index row of its neighbour (static);
    // nzflag[s*64 + q] == gen  <=>  the slot named by rev[s*64 + q] may hold a non-zero rate in the current event list
    // (set by the rate build, cleared when the slot is zeroed through that reverse-index entry; the slots of an event's own
    // two rows keep their flags -- a stale "may" only costs a redundant, bit-identical row sum).  gen is
    // 1..255 and advances with every build, so the flags of older lists expire without a pass over the array.
    unsigned char *revpos = nullptr, *nzflag = nullptr;
    int gen = 0;
    unsigned char *row_active = nullptr;  // N bytes: the site could start an event at the last rate build (build_rates_kernel)
    bool list_built = false;              // false until the first build has written every row of prob / type
    unsigned *mt = nullptr;  // 624 words + pos
    int *log = nullptr;
    double *log_psum = nullptr;
    int log_cap = 0;
    EvResult *result = nullptr;
    EvEnergies energies;
    bool energies_set = false;
    int last_n_events = 0;
};

namespace {

constexpr int REV_STRIDE = 64;  // max number of rows that list a given site (in-degree of the neighbour graph)
// reverse neighbour index, fixed stride: rev[s*64 + q] = (r << 6) | n for the slots (r, n) with neigh[r][n] == s,
// -1 padded (r < 2^24, n < 64: no division in the event loop)
__global__ void rev_fill_kernel(const int *__restrict__ neigh, long long total, int nn, int *__restrict__ fill,
                                int *__restrict__ rev, unsigned char *__restrict__ revpos, int *__restrict__ overflow) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    int j = neigh[s];
    if (j >= 0) {
        int pos = atomicAdd(fill + j, 1);
        if (pos < REV_STRIDE - 1) {  // last entry stays free (-1)
            rev[(size_t)j * REV_STRIDE + pos] = (int)((s / nn) << 6) | (int)(s % nn);
            revpos[s] = (unsigned char)pos;
        } else {
            atomicExch(overflow, 1);
        }
    }
}

// ---- summation spec scan_256, evaluated by ONE warp: lane l owns the 8 consecutive elements 8l..8l+7 ----------
// a[k] = sequential prefix inside the lane; lane totals are Kogge-Stone scanned across the warp;
// incl[8l+k] = S[l-1] + a[k]; group total = S[31].
struct Scan256 {
    double a[8];   // in: values, out: lane-local inclusive prefixes
    double excl;   // S[l-1] (unused for lane 0)
    double total;  // S[31]
};
__device__ __forceinline__ void warp_scan_256(Scan256 &r) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 1; k < 8; ++k) r.a[k] = r.a[k - 1] + r.a[k];
    double S = r.a[7];
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) {
        double o = __shfl_up_sync(KMC_FULL_MASK, S, d);
        if (lane >= d) S = o + S;
    }
    r.excl = __shfl_up_sync(KMC_FULL_MASK, S, 1);
    r.total = __shfl_sync(KMC_FULL_MASK, S, 31);
}
// two independent scans with their dependent chains interleaved (the latency of one)
__device__ __forceinline__ void warp_scan_256x2(Scan256 &r, Scan256 &q) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        r.a[k] = r.a[k - 1] + r.a[k];
        q.a[k] = q.a[k - 1] + q.a[k];
    }
    double S = r.a[7], T = q.a[7];
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) {
        const double o = __shfl_up_sync(KMC_FULL_MASK, S, d);
        const double u = __shfl_up_sync(KMC_FULL_MASK, T, d);
        if (lane >= d) { S = o + S; T = u + T; }
    }
    r.excl = __shfl_up_sync(KMC_FULL_MASK, S, 1);
    q.excl = __shfl_up_sync(KMC_FULL_MASK, T, 1);
    r.total = __shfl_sync(KMC_FULL_MASK, S, 31);
    q.total = __shfl_sync(KMC_FULL_MASK, T, 31);
}
__device__ __forceinline__ double scan_incl(const Scan256 &r, int k) {
    return ((threadIdx.x & 31) > 0) ? (r.excl + r.a[k]) : r.a[k];
}
// first t (0..255) with incl[t] > number, else the last t with v[t] > 0, else -1; *prev = incl[t-1] (0 for t == 0).
// v: the original values (before the scan).
__device__ __forceinline__ int warp_pick_256(const Scan256 &sc, const double v[8], double number, double *prev) {
    double inc[8];
    int kfirst = 8, klast = -1;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        inc[k] = scan_incl(sc, k);
        if (inc[k] > number) kfirst = k;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (v[k] > 0.0) klast = k;
    int tsel = -1;
    unsigned bf = __ballot_sync(KMC_FULL_MASK, kfirst < 8);
    if (bf) {
        int l = __ffs(bf) - 1;
        tsel = l * 8 + __shfl_sync(KMC_FULL_MASK, kfirst, l);
    } else {
        unsigned bl = __ballot_sync(KMC_FULL_MASK, klast >= 0);
        if (bl) {
            int l = 31 - __clz(bl);
            tsel = l * 8 + __shfl_sync(KMC_FULL_MASK, klast, l);
        }
    }
    double pv = 0.0;
    if (tsel > 0) {
        int pl = (tsel - 1) >> 3, pk = (tsel - 1) & 7;
        double cand = inc[0];
#pragma unroll
        for (int k = 1; k < 8; ++k)
            if (k == pk) cand = inc[k];
        pv = __shfl_sync(KMC_FULL_MASK, cand, pl);
    }
    *prev = pv;
    return tsel;
}
// The same selection from STORED inclusive prefixes (inc = incl[8 lane .. 8 lane + 7]); the values themselves are only
// needed when no prefix exceeds number (load_v fetches them then).
template <class LoadV>
__device__ __forceinline__ int warp_pick_incl(const double inc[8], double number, double *prev, double *cur, LoadV load_v) {
    int kfirst = 8;
#pragma unroll
    for (int k = 7; k >= 0; --k)
        if (inc[k] > number) kfirst = k;
    int tsel = -1;
    unsigned bf = __ballot_sync(KMC_FULL_MASK, kfirst < 8);
    if (bf) {
        int l = __ffs(bf) - 1;
        tsel = l * 8 + __shfl_sync(KMC_FULL_MASK, kfirst, l);
    } else {
        double v[8];
        load_v(v);
        int klast = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (v[k] > 0.0) klast = k;
        unsigned bl = __ballot_sync(KMC_FULL_MASK, klast >= 0);
        if (bl) {
            int l = 31 - __clz(bl);
            tsel = l * 8 + __shfl_sync(KMC_FULL_MASK, klast, l);
        }
    }
    double pv = 0.0, cv = 0.0;
    if (tsel >= 0) {  // prev = incl[tsel - 1] (0 for tsel == 0), cur = incl[tsel]
        const int pt = tsel > 0 ? tsel - 1 : 0;
        const int pl = pt >> 3, pk = pt & 7, cl = tsel >> 3, ck = tsel & 7;
        double candp = inc[0], candc = inc[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            if (k == pk) candp = inc[k];
            if (k == ck) candc = inc[k];
        }
        pv = __shfl_sync(KMC_FULL_MASK, candp, pl);
        cv = __shfl_sync(KMC_FULL_MASK, candc, cl);
        if (tsel == 0) pv = 0.0;
    }
    *prev = pv;
    *cur = cv;
    return tsel;
}
// max over t < tsel of incl[t] (incl[8 lane + k] held as inc[k]); -inf when tsel == 0.  With it the selection rule
// "first t whose inclusive prefix exceeds number" can be VALIDATED for a given t without looking at the 256 values again:
// t is the selected element  <=>  !(max_before > number) && incl[t] > number.
__device__ __forceinline__ double warp_max_before(const double inc[8], int tsel) {
    const int lane = threadIdx.x & 31;
    double m = -1.0 / 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (8 * lane + k < tsel && inc[k] > m) m = inc[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double o = __shfl_xor_sync(KMC_FULL_MASK, m, off);
        if (o > m) m = o;
    }
    return m;
}
// butterfly row sum of the summation spec: lane l holds p[l] + p[l+32]
__device__ __forceinline__ double warp_row_sum(double p0, double p1) { return kmc_warp_xor_sum(p0 + p1); }

// rate of one (i, slot) pair: kmc_events.cu:141-228
__device__ __forceinline__ double event_rate(int i, int j, int el_i, int c_i, double pot_i, double xi, double yi,
                                             double zi, const int *__restrict__ element,
                                             const int *__restrict__ charge, const double *__restrict__ pot,
                                             const int *__restrict__ layer, const double *__restrict__ x,
                                             const double *__restrict__ y, const double *__restrict__ z,
                                             const EvEnergies &E, double kT, double freq, double sigma, double k,
                                             int &ev) {
    const double epsilon = 1e-200;  // :150
    ev = KMCB200_NULL_EVENT;
    int el_j = element[j];
    double P = 0.0;
    if (el_i == KMCB200_DEFECT && el_j == KMCB200_O) {  // :158
        double Eo = 2 * (pot_i - pot[j]);
        double EA = E.E_gen[layer[j]] - Eo - 0.0;
        ev = KMCB200_VACANCY_GENERATION;
        P = freq * (1 / (exp(EA / kT) + epsilon));
    } else if (el_i == KMCB200_OXYGEN_DEFECT && el_j == KMCB200_VACANCY) {  // :171
        double dist = 1e-10 * kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
        double self_int_V = kmc_v_solve(dist, 2, sigma, k);
        int charge_state = c_i - charge[j];
        double Eo = charge_state * ((pot_i - pot[j]) + (charge_state / 2) * self_int_V);
        double EA = E.E_rec[layer[j]] - Eo - 0.0;
        ev = KMCB200_VACANCY_RECOMBINATION;
        P = freq * (1 / (exp(EA / kT) + epsilon));
    } else if (el_i == KMCB200_VACANCY && el_j == KMCB200_O) {  // :188
        double self_int_V = 0.0;
        if (c_i != 0) {
            double dist = 1e-10 * kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
            self_int_V = kmc_v_solve(dist, c_i, sigma, k);
        }
        double Eo = (c_i - charge[j]) * ((pot_i - pot[j]) + self_int_V);
        double EA = E.E_Vdiff[layer[j]] - Eo - 0.0;
        ev = KMCB200_VACANCY_DIFFUSION;
        P = freq * (1 / (exp(EA / kT) + epsilon));
    } else if (el_i == KMCB200_OXYGEN_DEFECT && el_j == KMCB200_DEFECT) {  // :207
        double self_int_V = 0.0;
        if (c_i != 0) {
            double dist = 1e-10 * kmc_dist_nopbc(xi, yi, zi, x[j], y[j], z[j]);
            self_int_V = kmc_v_solve(dist, 2, sigma, k);
        }
        double Eo = (c_i - charge[j]) * ((pot_i - pot[j]) - self_int_V);
        double EA = E.E_Odiff[layer[j]] - Eo - 0.0;
        ev = KMCB200_ION_DIFFUSION;
        P = freq * (1 / (exp(EA / kT) + epsilon));
    }
    return P;
}

// One CTA per chunk of 256 rows; each of its 8 warps walks 32 rows, lanes cover the nn (<= 64) slots of a row.
// Writes event_prob / event_type (coalesced), the row sums and the chunk sum.
// Incremental across supersteps: a row whose site can start no event (an ordinary lattice atom: ~95 % of the
// sites) holds 52 zero rates and NULL types.  Such a row is written only if it COULD act at the previous build
// (row_active, one byte per site); rows that were and are inactive still hold the zeros of an earlier build -- the event
// loop only ever clears slots -- and are skipped, which takes ~1.1 GB of stores per superstep down to the active rows.
// full != 0 (first build of a list): every row is written.
__global__ void __launch_bounds__(256) build_rates_kernel(int N, int nn, const int *__restrict__ neigh,
                                                         const int *__restrict__ layer, double kT, double freq,
                                                         double sigma, double k, const double *__restrict__ x,
                                                         const double *__restrict__ y, const double *__restrict__ z,
                                                         const double *__restrict__ pot,
                                                         const int *__restrict__ element,
                                                         const int *__restrict__ charge, EvEnergies E,
                                                         double *__restrict__ prob, unsigned char *__restrict__ type,
                                                         double *__restrict__ rowsum, double *__restrict__ chunksum,
                                                         double *__restrict__ rowincl,
                                                         const unsigned char *__restrict__ revpos,
                                                         unsigned char *__restrict__ nzflag, int gen,
                                                         unsigned char *__restrict__ row_active, int full) {
    __shared__ double rs[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = blockIdx.x * 256 + w * 32;
    double my_rowsum = 0.0;  // lane q ends up holding the sum of row row0 + q
    int my_el = -1;
    bool my_act = false, my_was = false;
    if (row0 + lane < N) {
        my_el = element[row0 + lane];
        my_act = (my_el == KMCB200_DEFECT || my_el == KMCB200_OXYGEN_DEFECT || my_el == KMCB200_VACANCY);
        my_was = full || row_active[row0 + lane] != 0;
        row_active[row0 + lane] = my_act ? 1 : 0;
    }
    unsigned todo = __ballot_sync(KMC_FULL_MASK, my_act || my_was);
    while (todo) {
        const int q = __ffs(todo) - 1;
        todo &= todo - 1;
        int i = row0 + q;
        double s = 0.0;
        {
            int el_i = __shfl_sync(KMC_FULL_MASK, my_el, q);
            bool can_act = (el_i == KMCB200_DEFECT || el_i == KMCB200_OXYGEN_DEFECT || el_i == KMCB200_VACANCY);
            double P0 = 0.0, P1 = 0.0;
            int e0 = KMCB200_NULL_EVENT, e1 = KMCB200_NULL_EVENT;
            size_t base = (size_t)i * nn;
            if (can_act) {
                int c_i = charge[i];
                double pot_i = pot[i], xi = x[i], yi = y[i], zi = z[i];
                if (lane < nn) {
                    int j = neigh[base + lane];
                    if (j >= 0 && j < N)
                        P0 = event_rate(i, j, el_i, c_i, pot_i, xi, yi, zi, element, charge, pot, layer, x, y, z, E, kT,
                                        freq, sigma, k, e0);
                    if (P0 != 0.0) nzflag[(size_t)j * REV_STRIDE + revpos[base + lane]] = (unsigned char)gen;
                }
                if (lane + 32 < nn) {
                    int j = neigh[base + lane + 32];
                    if (j >= 0 && j < N)
                        P1 = event_rate(i, j, el_i, c_i, pot_i, xi, yi, zi, element, charge, pot, layer, x, y, z, E, kT,
                                        freq, sigma, k, e1);
                    if (P1 != 0.0) nzflag[(size_t)j * REV_STRIDE + revpos[base + lane + 32]] = (unsigned char)gen;
                }
            }
            if (lane < nn) { prob[base + lane] = P0; type[base + lane] = (unsigned char)e0; }
            if (lane + 32 < nn) { prob[base + lane + 32] = P1; type[base + lane + 32] = (unsigned char)e1; }
            // row sum (summation spec): butterfly over the lanes; all-zero rows short-cut to +0.0
            if (__any_sync(KMC_FULL_MASK, (P0 != 0.0) || (P1 != 0.0))) s = warp_row_sum(P0, P1);
        }
        if (lane == q) my_rowsum = s;
    }
    if (row0 + lane < N) rowsum[row0 + lane] = my_rowsum;
    rs[w * 32 + lane] = my_rowsum;
    __syncthreads();
    // scan_256 of the 256 row sums by warp 0
    if (w == 0) {
        Scan256 sc;
#pragma unroll
        for (int k = 0; k < 8; ++k) sc.a[k] = rs[8 * lane + k];
        warp_scan_256(sc);
        if (lane == 0) chunksum[blockIdx.x] = sc.total;
        double2 *dst2 = reinterpret_cast<double2 *>(rowincl + (size_t)blockIdx.x * 256 + 8 * lane);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst2[k] = make_double2(scan_incl(sc, 2 * k), scan_incl(sc, 2 * k + 1));
    }
}

// one warp per super: scan_256 total of its 256 chunk sums
__global__ void __launch_bounds__(32) super_sums_kernel(const double *__restrict__ chunksum, long long nchunk,
                                                       double *__restrict__ supersum, double *__restrict__ chunkincl) {
    const int lane = threadIdx.x;
    Scan256 sc;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long c = (long long)blockIdx.x * 256 + 8 * lane + k;
        sc.a[k] = (c < nchunk) ? chunksum[c] : 0.0;
    }
    warp_scan_256(sc);
    if (lane == 0) supersum[blockIdx.x] = sc.total;
#pragma unroll
    for (int k = 0; k < 8; ++k) chunkincl[(size_t)blockIdx.x * 256 + 8 * lane + k] = scan_incl(sc, k);
}

// ---- MT19937 (std::mt19937) + libstdc++ generate_canonical<double,53> -----------------------------------
__device__ __forceinline__ unsigned mt_next32(unsigned *mt, int &pos) {
    if (pos >= 624) {
        for (int i = 0; i < 624; ++i) {
            unsigned yv = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
        }
        pos = 0;
    }
    unsigned yv = mt[pos++];
    yv ^= (yv >> 11);
    yv ^= (yv << 7) & 0x9d2c5680u;
    yv ^= (yv << 15) & 0xefc60000u;
    yv ^= (yv >> 18);
    return yv;
}
// uniform_real_distribution<double>(0,1)(mt19937): two 32-bit draws, (x0 + x1*2^32) / 2^64, clamped below 1
__device__ __forceinline__ double mt_next_double(unsigned *mt, int &pos) {
    double x0 = (double)mt_next32(mt, pos);
    double x1 = (double)mt_next32(mt, pos);
    double sum = x0 + x1 * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = 0.99999999999999988897769753748;  // nextafter(1.0, 0.0)
    return ret;
}

#ifdef KMC_EV_PROFILE
#define EV_TICK(k) do { if (tid == 0) { long long now_ = clock64(); ph[k] += now_ - t_last; t_last = now_; } } while (0)
#else
#define EV_TICK(k) do { } while (0)
#endif

// KMC_EV_TRACE: per-warp clock64 stamps at the phase boundaries of 64 consecutive events (diagnostics: who waits for whom)
#ifdef KMC_EV_TRACE
#define EV_TRACE_FIRST 2000
#define EV_TR(k) do { const int te_ = tr_e - EV_TRACE_FIRST + ((k) == 0 || ((k) >= 7 && (k) <= 11) ? 1 : 0); if (lane == 0 && te_ >= 0 && te_ < 64 && a.phase_cycles) a.phase_cycles[(te_ * 16 + warp) * 16 + (k)] = clock64(); } while (0)
#else
#define EV_TR(k) do { } while (0)
#endif

struct EvLoopArgs {
    int N, nn;
    long long nchunk, nsuper;
    const int *neigh;
    double *prob;
    unsigned char *type;
    double *rowsum, *chunksum, *supersum, *rowincl, *chunkincl;
    const int *rev;
    const unsigned char *revpos;  // position of each slot in its neighbour's reverse-index row
    unsigned char *nzflag;        // == gen: the slot named by the reverse-index entry may hold a non-zero rate
    int gen;
    int *element, *charge;
    unsigned *mt_state;  // 624 + pos
    double inv_freq_threshold;  // 1/freq
    int max_events;
    int *log;
    double *log_psum;
    int log_cap;
    EvResult *result;
    int use_spec;  // warp 15 predicts the next event and stages its data in shared memory (never changes a result)
    long long *phase_cycles;  // 16 counters (KMC_EV_PROFILE builds)
};

// Hand-overs between warps of the one CTA through SHARED memory (data stores, then a volatile flag store; volatile flag
// load, then data loads): a warp's shared-memory accesses are performed in program order, so only the compiler has to be
// kept from reordering them.  (A membar.cta would also wait for the warp's outstanding GLOBAL accesses -- a full round
// trip on the critical path.)  Hand-overs of global data go through the block barriers.
#define KMC_CBAR() asm volatile("" ::: "memory")
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Shared-memory layout of a 256-group that one warp scans (lane l owns elements 8l..8l+7): element t lives at
// (t & 7) * 32 + (t >> 3), so the 32 lanes of a warp reading "their k-th element" hit 32 consecutive doubles (no bank
// conflicts; the natural layout is a 16-way conflict on every one of the 8 loads).
__device__ __forceinline__ int tpos(int t) { return ((t & 7) << 5) | (t >> 3); }

// Device-resident MT19937 that may run ahead of what the loop consumes (the uniforms of event e+2 are drawn while event
// e is repaired).  block_start = number of 32-bit words produced before the current 624-word block began; the state
// before the latest twist is kept, so at loop exit the generator can be handed back exactly `consumed` words in -- in
// step with the host generator of the reference (src/kmc_events.cu:469,515: two doubles per event).
struct EvRng {
    unsigned mt[624], backup[624];
    int pos;
    long long block_start, prev_block_start;
};
__device__ __forceinline__ unsigned evrng_next32(EvRng &g) {
    if (g.pos >= 624) {
        for (int q = 0; q < 624; ++q) g.backup[q] = g.mt[q];
        g.prev_block_start = g.block_start;
        g.block_start += 624;
    }
    return mt_next32(g.mt, g.pos);
}
__device__ __forceinline__ double evrng_next_double(EvRng &g) {
    double x0 = (double)evrng_next32(g);
    double x1 = (double)evrng_next32(g);
    double sum = x0 + x1 * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = 0.99999999999999988897769753748;  // nextafter(1.0, 0.0)
    return ret;
}

// One event record, double buffered by event parity: the selector of event e+1 may run while event e is still applied
// to element/charge and logged.
struct EvRecord {
    int i, j, ty, slot;
    double psum;
    double pos, w;  // cumulative rate in front of row i and the row's total rate (for the predictor, approximate)
    int fast;       // the event is the predictor's (validated): its zero-out inputs are already staged in shared memory
};

// What the predictor publishes about the next event: its path through the hierarchy and, per level, the interval of
// `number` for which that path is what the exact selection rule picks (see warp_max_before).
struct EvPrediction {
    int ts, tc, tr, n;            // super, chunk in super, row in chunk, slot in row
    int ej, ety;                  // partner site and event type of the slot
    double lo_c, hi_c, prev_c;    // chunk level: max prefix in front, prefix at tc, prefix at tc - 1 (0 for tc == 0)
    double lo_r, hi_r, prev_r;    // row level
    double acc_b, acc_a;          // slot level: running sum before / after the slot (sequential over the non-zero slots)
};

// The persistent event loop: ONE CTA of 16 warps, three block barriers per event.
//   S   one warp (elected, see R2): top level = scan_256 of the super sums; chunk and row level = compare against the
//       STORED inclusive prefixes (chunk level in shared memory, row level from global memory); slot level = ordered walk
//       over the row's non-zero slots; publishes (i, j, type) and the residence time.  The two global round trips of this
//       phase (the chunk's 256 row prefixes, the row's rates / neighbours / types) are normally served from shared memory:
//       see P.
//   ---- barrier A
//   Z   4 warps: zero-out through the fixed-stride reverse index (2 dependent round trips) + list of rows that lost a rate
//   ---- barrier B1
//   R1  one warp per such row: butterfly row sum, dirty-chunk bitmap + list
//   ---- barrier B2
//   R2  one warp per dirty chunk: scan_256 of its 256 row sums -> chunk sum + stored row prefixes.  The warp that
//       finishes the LAST dirty chunk of a super re-scans that super; the warp that finishes the LAST dirty super is the
//       selector of the next event and goes straight on to S (counters in shared memory instead of two more barriers).
//   H   warp 14, during R: applies the event to element/charge, writes the log, draws the uniforms of the event after next.
//   P   warp 15, from barrier A on: PREDICTS the next event's chunk and row from the not-yet-repaired sums (corrected for
//       the rate the current event removes) and copies that chunk's row prefixes and that row's slots into shared memory
//       while Z / R1 / R2 run.  The selector repeats the exact selection on the repaired sums and uses the copies only if
//       it arrives at the same chunk / row and no row of that chunk changed in between -- the copies are then bit-identical
//       to what it would load, so speculation never changes a result, it only hides two memory round trips.
// SMEM: chunk sums + their stored prefixes live in dynamic shared memory (padded to whole supers with zeros).
template <bool SMEM>
__global__ void __launch_bounds__(EV_THREADS, 1) event_loop_kernel(EvLoopArgs a) {
    extern __shared__ double cs_smem[];  // [nsuper*256] chunk sums, then [nsuper*256] inclusive prefixes (tpos layout)
    __shared__ EvRng rng;
    __shared__ double ss[MAX_SUPER];           // super sums, tpos layout
    __shared__ double s_u1[4], s_lg[4];         // per event (ring of 4): selection uniform, -log(time uniform)
    __shared__ EvRecord s_rec[2];
    __shared__ unsigned chunk_bits[2048];       // dirty bitmap over <= 65536 chunks
    __shared__ int rows_list[2 * REV_STRIDE];
    __shared__ int chunk_list[2 * REV_STRIDE + 2];
    __shared__ int super_cnt[MAX_SUPER];        // dirty chunks of a super still to be re-scanned
    __shared__ int n_rows, n_chunks, n_supers;
    __shared__ double s_event_time;
    __shared__ int s_stop, s_nevents;
    // predictor -> selector
    __shared__ double s_spec_incl[256];         // row prefixes of the predicted chunk, tpos layout
    __shared__ double s_spec_p[64];             // rates of the predicted row
    __shared__ int s_spec_nb[64], s_spec_ty[64];
    __shared__ int s_spec_chunk, s_spec_r;
    __shared__ EvPrediction s_pred;
    // zero-out inputs of the predicted event, double-buffered by event parity (the predictor fills buffer (e+1)&1 while the
    // zero-out of event e may still read buffer e&1)
    __shared__ int s_pz_packed[2][2 * REV_STRIDE];   // reverse-index rows of the predicted event's two sites ...
    __shared__ double s_pz_oldp[2][2 * REV_STRIDE];  // ... and the rates those slots hold
    __shared__ int s_pz_rows[2][2 * REV_STRIDE];     // ... compacted: the rows (other than the two sites) that lose a rate
    __shared__ int s_pz_nrows[2];
    __shared__ int s_pz_event[2];                    // index of the event each buffer was staged for
    __shared__ int s_pz_pr, s_pz_pej, s_pz_req1, s_pz_req;      // predictor -> housekeeping warp: the predicted pair, = e + 1 when set
    __shared__ int s_pred_ok;                        // the prediction reached the slot level
    __shared__ int s_h_done;                    // = e + 1 once the residence time of event e is stored
    __shared__ int s_b2_event;                  // = e + 1 once event e passed barrier B2 (its dirty lists are complete)
    __shared__ int s_pred_clean;                // bit 0: predicted chunk not dirty in this event, bit 1: nor its super
    __shared__ double s_top_incl[256];          // inclusive prefixes of the super sums as the last selector scanned them
    __shared__ int s_spec_ready;                // = e once the speculation for event e is complete

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EV_THREADS / 32;
    constexpr int PW = NW - 1, HW = NW - 2;  // predictor / housekeeping warps
    constexpr int NCW = NW - 2;              // warps that take row / chunk work in R1 / R2
    const int nn = a.nn;
#ifdef KMC_EV_PROFILE
    long long ph[16] = {0};
    long long t_last = clock64();
    __shared__ unsigned long long s_prof[16];
    if (tid < 16) s_prof[tid] = 0;
#endif
    const int npad = (int)a.nsuper * 256;  // chunk arrays are padded to whole supers (tail = 0)
    double *cs, *ci;                       // chunk sums / inclusive prefixes of the chunk sums per super
    if (SMEM) {
        cs = cs_smem;
        ci = cs_smem + npad;
        for (int q = tid; q < npad; q += EV_THREADS) {
            const int d = (q & ~255) | tpos(q & 255);
            cs[d] = a.chunksum[q];
            ci[d] = a.chunkincl[q];
        }
    } else {
        cs = a.chunksum;
        ci = a.chunkincl;
    }
    // element t (0..255) of super s inside cs / ci
#define KMC_CS(s_, k_) (SMEM ? ((s_) * 256 + 32 * (k_) + lane) : ((s_) * 256 + 8 * lane + (k_)))
#define KMC_CIDX(c) (SMEM ? (((c) & ~255) | tpos((c) & 255)) : (c))
    for (int q = tid; q < 624; q += EV_THREADS) rng.mt[q] = a.mt_state[q];
    for (int q = tid; q < MAX_SUPER; q += EV_THREADS) {
        ss[tpos(q)] = (q < a.nsuper) ? a.supersum[q] : 0.0;
        super_cnt[q] = 0;
    }
    for (int q = tid; q < 2048; q += EV_THREADS) chunk_bits[q] = 0u;
    if (tid == 0) {
        rng.pos = (int)a.mt_state[624];
        rng.block_start = -(long long)rng.pos;
        rng.prev_block_start = rng.block_start;
        s_event_time = 0.0;
        s_nevents = 0;
        s_stop = 0;
        n_rows = 0;
        n_chunks = 0;
        n_supers = 0;
        s_spec_chunk = -1; s_spec_r = -1; s_spec_ready = 0; s_h_done = 0; s_b2_event = 0; s_pred_clean = 0; s_pz_req = 0; s_pz_req1 = 0; s_pz_pr = -1; s_pz_pej = -1; s_pred_ok = 0; s_pz_event[0] = s_pz_event[1] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        for (int e = 0; e < 2; ++e) {  // uniforms of events 0 and 1
            s_u1[e] = evrng_next_double(rng);
            s_lg[e] = -log(evrng_next_double(rng));
        }
    }
    __syncthreads();

    bool i_select = (warp == 0);
#ifdef KMC_EV_TRACE
    __shared__ long long s_tr_cnt[8];
    if (tid < 8) s_tr_cnt[tid] = 0;
    int tr_e = -1;  // index of the event whose phases are being stamped (set after barrier A)
#endif
    while (true) {
        // =============================== S: selector (one warp, warp-synchronous) =============================
        EV_TR(0);
        if (i_select) {
            const int e = s_nevents;  // index of the event selected now
            while (*(volatile int *)&s_h_done < e) { }  // residence time of event e - 1
            KMC_CBAR();
            const bool go = (s_event_time < a.inv_freq_threshold) && (a.max_events <= 0 || e < a.max_events);
            int ei = -1, ej = -1, ety = KMCB200_NULL_EVENT, eslot = -1;
            double Psum = 0.0, pos = 0.0, wrow = 0.0;
            int fast_rec = 0;
            if (go) {
                // ---- top level: scan_256 over the super sums (shared memory) ------------------------------
                Scan256 sc;
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { v[k] = ss[32 * k + lane]; sc.a[k] = v[k]; }
                warp_scan_256(sc);
                Psum = sc.total;
                // u1 (selection, kmc_events.cu:469) and -log(u2) (residence time, :515) were drawn two events ago
                const double u1 = s_u1[e & 3];
                double number = u1 * Psum;
                double prev, cur;
                int ts = (Psum > 0.0) ? warp_pick_256(sc, v, number, &prev) : -1;
                if (a.use_spec) {  // for the predictor of this event (the super sums do not change before its repair)
#pragma unroll
                    for (int k = 0; k < 8; ++k) s_top_incl[32 * k + lane] = scan_incl(sc, k);
                }
                EV_TR(8);
                int r = -1;  // all slot indices fit 32 bits: N * nn <= 16.7 M * 64 < 2^31
                bool spec = false;
                // ---- fast path: the predictor's path, VALIDATED level by level against the exact `number` (the intervals
                // come from arrays no event has touched since the predictor read them: chunk prefixes of a super that was
                // not re-scanned, row prefixes / rates of a chunk that is not dirty).  Every check is exactly the selection
                // rule applied to that element, so an accepted path is the path the full selection below would take.
                bool fast = false, chunk_clean = false;
                int pc = -1;
                if (ts >= 0 && a.use_spec) {
                    while (*(volatile int *)&s_spec_ready < e) { }
                    KMC_CBAR();
                    EV_TR(9);
                    pc = *(volatile int *)&s_spec_chunk;
                    const int clean = *(volatile int *)&s_pred_clean;  // the predictor's verdict on the dirty lists
                    chunk_clean = (clean & 1) != 0;
                    const bool super_clean = (clean & 2) != 0;
                    if (*(volatile int *)&s_pred_ok && super_clean && s_pred.ts == ts) {
                        const double n1 = number - prev;
                        const double n2 = n1 - s_pred.prev_c;
                        const double n3 = n2 - s_pred.prev_r;
                        if (!(s_pred.lo_c > n1) && s_pred.hi_c > n1 && !(s_pred.lo_r > n2) && s_pred.hi_r > n2 &&
                            !(s_pred.acc_b > n3) && s_pred.acc_a > n3) {
                            fast = true;
                            fast_rec = 1;
                            pos = (prev + s_pred.prev_c) + s_pred.prev_r;
                            wrow = s_pred.hi_r - s_pred.prev_r;
                            ei = (ts * 256 + s_pred.tc) * 256 + s_pred.tr;
                            ej = s_pred.ej; ety = s_pred.ety;
                            eslot = ei * nn + s_pred.n;
#ifdef KMC_EV_TRACE
                            if (lane == 0) s_tr_cnt[0]++;
#endif
                        }
                    }
#ifdef KMC_EV_TRACE
                    if (lane == 0) {
                        s_tr_cnt[1] += (s_pred_ok != 0);
                        s_tr_cnt[2] += super_clean;
                        s_tr_cnt[3] += chunk_clean;
                        s_tr_cnt[4] += (s_pred.ts == ts);
                    }
#endif
                }
                EV_TR(10);
                if (ts >= 0 && !fast) {
                    number = number - prev;
                    pos = prev;
                    // ---- chunk level: stored prefixes of super ts ---------------------------------------------
                    double inc[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) inc[k] = ci[KMC_CS(ts, k)];
                    int tc = warp_pick_incl(inc, number, &prev, &cur, [&](double *vv) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) vv[k] = cs[KMC_CS(ts, k)];
                    });
                    if (tc >= 0) {
                        number = number - prev;
                        pos = pos + prev;
                        const int chunk = ts * 256 + tc;
                        // ---- row level: stored prefixes of the chunk: the predictor's copy, else one round trip ------
                        spec = chunk_clean && (pc == chunk);  // the predictor's copy of this chunk's row prefixes
                        const int rbase = chunk * 256 + 8 * lane;
                        if (spec) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) inc[k] = s_spec_incl[32 * k + lane];
                        } else {
                            const double2 *src2 = reinterpret_cast<const double2 *>(a.rowincl + rbase);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                double2 t2 = src2[k];
                                inc[2 * k] = t2.x; inc[2 * k + 1] = t2.y;
                            }
                        }
                        int tr = warp_pick_incl(inc, number, &prev, &cur, [&](double *vv) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) vv[k] = (rbase + k < a.N) ? a.rowsum[rbase + k] : 0.0;
                        });
                        if (tr >= 0) {
                            number = number - prev;
                            pos = pos + prev;
                            wrow = cur - prev;
                            r = chunk * 256 + tr;
                        }
                    }
                }
                if (r >= 0) {
                    // ---- slot level: lanes hold slots lane and lane+32; walk the non-zero slots in order --------
                    const int base = r * nn;
                    double p0 = 0.0, p1 = 0.0;
                    int nb0 = -1, nb1 = -1, ty0 = KMCB200_NULL_EVENT, ty1 = KMCB200_NULL_EVENT;
                    if (spec && *(volatile int *)&s_spec_r == r) {
                        p0 = s_spec_p[lane]; p1 = s_spec_p[lane + 32];
                        nb0 = s_spec_nb[lane]; nb1 = s_spec_nb[lane + 32];
                        ty0 = s_spec_ty[lane]; ty1 = s_spec_ty[lane + 32];
#ifdef KMC_EV_PROFILE
                        if (lane == 0) atomicAdd(&s_prof[15], 1ull);
#endif
                    } else {
                        if (lane < 2) prefetch_l2(a.rev + r * REV_STRIDE + 32 * lane);  // for the zero-out phase
                        if (lane < nn) { p0 = a.prob[base + lane]; nb0 = a.neigh[base + lane]; ty0 = a.type[base + lane]; }
                        if (lane + 32 < nn) { p1 = a.prob[base + lane + 32]; nb1 = a.neigh[base + lane + 32]; ty1 = a.type[base + lane + 32]; }
                        // every candidate partner's reverse-index row starts its trip from DRAM now (the zero-out needs one)
                        if (p0 > 0.0) { prefetch_l2(a.rev + nb0 * REV_STRIDE); prefetch_l2(a.rev + nb0 * REV_STRIDE + 32); }
                        if (p1 > 0.0) { prefetch_l2(a.rev + nb1 * REV_STRIDE); prefetch_l2(a.rev + nb1 * REV_STRIDE + 32); }
                    }
                    unsigned m0 = __ballot_sync(KMC_FULL_MASK, p0 > 0.0), m1 = __ballot_sync(KMC_FULL_MASK, p1 > 0.0);
                    int seln = -1, lastn = -1;
                    double acc = 0.0;
                    // adding an exact zero never changes acc, so only non-zero slots can make acc exceed number
                    unsigned long long mm = ((unsigned long long)m1 << 32) | m0;
                    while (mm) {
                        int n = __ffsll((long long)mm) - 1;
                        mm &= mm - 1;
                        double pv = __shfl_sync(KMC_FULL_MASK, (n < 32) ? p0 : p1, n & 31);
                        acc = acc + pv;  // first term: 0.0 + pv == pv
                        lastn = n;
                        if (acc > number) { seln = n; break; }
                    }
                    if (seln < 0) seln = lastn;
                    if (seln >= 0) {
                        ej = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? nb0 : nb1, seln & 31);
                        ety = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? ty0 : ty1, seln & 31);
                        ei = r;
                        eslot = base + seln;
                    }
                }
            }
            EV_TR(11);
            if (lane == 0) {
                s_stop = !go;
                if (go) {
                    EvRecord &rec = s_rec[e & 1];
                    rec.i = ei; rec.j = ej; rec.ty = ety; rec.slot = eslot; rec.psum = Psum;
                    rec.pos = pos; rec.w = wrow; rec.fast = fast_rec;
                    s_nevents = e + 1;
                }
                n_chunks = 0;
                n_rows = 0;
            }
        }
        EV_TR(7);
        __syncthreads();  // ---- barrier A
#ifdef KMC_EV_TRACE
        tr_e = s_nevents - 1;
#endif
        EV_TR(1);
        EV_TICK(0);
        if (s_stop) break;
        const int ev_idx = s_nevents - 1;
        const int ei = s_rec[ev_idx & 1].i, ej = s_rec[ev_idx & 1].j;
        i_select = false;
        if (warp == PW) {
            // =========================== P: the predictor -- runs beside Z / R1 / R2 with no barrier of its own ==========
            // It reads the sums and rates while the other warps repair them.  Nothing it publishes is trusted as such: the
            // selector accepts a prediction only if (a) the predicted chunk / super is not in this event's dirty list --
            // then every array the prediction was derived from is untouched by this event, whatever the interleaving --
            // and (b) the exact `number` of the next event falls in the published intervals.
            if (a.use_spec) {
                const EvRecord rec = s_rec[ev_idx & 1];
                int pchunk = -1, pr = -1, pej = -1, ok = 0;
                if (rec.i >= 0) {
                    // the point the next draw will land on, in the coordinates of the sums BEFORE this event's repair: the
                    // event removes (to first order) row i's total rate w, which sits at cumulative position pos
                    double t = s_u1[(ev_idx + 1) & 3] * (rec.psum - rec.w);
                    if (t >= rec.pos) t = t + rec.w;
                    double prev, cur;
                    double tinc[8];  // the selector of this event left its scan of the (still unrepaired) super sums
#pragma unroll
                    for (int k = 0; k < 8; ++k) tinc[k] = s_top_incl[32 * k + lane];
                    const int ts = warp_pick_incl(tinc, t, &prev, &cur, [&](double *vv) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) vv[k] = 0.0;  // no prefix exceeds t: no prediction
                    });
                    int tc = -1;
                    double inc[8];
                    if (ts >= 0) {
                        t = t - prev;
#pragma unroll
                        for (int k = 0; k < 8; ++k) inc[k] = ci[KMC_CS(ts, k)];
                        tc = warp_pick_incl(inc, t, &prev, &cur, [&](double *vv) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) vv[k] = cs[KMC_CS(ts, k)];
                        });
                    }
                    if (tc >= 0) {
                        double pinc[8];
                        pchunk = ts * 256 + tc;
                        const double2 *src2 = reinterpret_cast<const double2 *>(a.rowincl + pchunk * 256 + 8 * lane);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double2 t2 = src2[k];
                            pinc[2 * k] = t2.x; pinc[2 * k + 1] = t2.y;
                        }
                        EV_TR(3);
                        // (while the row prefixes are in flight)
                        const double lo = warp_max_before(inc, tc);
                        if (lane == 0) {
                            s_pred.ts = ts; s_pred.tc = tc;
                            s_pred.lo_c = lo; s_pred.hi_c = cur; s_pred.prev_c = prev;
                        }
                        const double pnum = t - prev;
#pragma unroll
                        for (int k = 0; k < 8; ++k) s_spec_incl[32 * k + lane] = pinc[k];
                        const int rbase = pchunk * 256 + 8 * lane;
                        const int tr = warp_pick_incl(pinc, pnum, &prev, &cur, [&](double *vv) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) vv[k] = (rbase + k < a.N) ? a.rowsum[rbase + k] : 0.0;
                        });
                        if (tr >= 0) {
                            pr = pchunk * 256 + tr;
                            const int base = pr * nn;
                            double p0 = 0.0, p1 = 0.0;
                            int nb0 = -1, nb1 = -1, ty0 = KMCB200_NULL_EVENT, ty1 = KMCB200_NULL_EVENT;
                            if (lane < nn) { p0 = a.prob[base + lane]; nb0 = a.neigh[base + lane]; ty0 = a.type[base + lane]; }
                            if (lane + 32 < nn) { p1 = a.prob[base + lane + 32]; nb1 = a.neigh[base + lane + 32]; ty1 = a.type[base + lane + 32]; }
                            if (lane == 0) {  // the housekeeping warp starts staging this site's half now
                                s_pz_pr = pr;
                                KMC_CBAR();
                                *(volatile int *)&s_pz_req1 = ev_idx + 1;
                            }
                            EV_TR(4);
                            // (while the row is in flight)
                            const double lo_r = warp_max_before(pinc, tr);
                            const double pnum3 = pnum - prev;
                            if (lane == 0) { s_pred.tr = tr; s_pred.lo_r = lo_r; s_pred.hi_r = cur; s_pred.prev_r = prev; }
                            if (p0 > 0.0) { prefetch_l2(a.rev + nb0 * REV_STRIDE); prefetch_l2(a.rev + nb0 * REV_STRIDE + 32); }
                            if (p1 > 0.0) { prefetch_l2(a.rev + nb1 * REV_STRIDE); prefetch_l2(a.rev + nb1 * REV_STRIDE + 32); }
                            s_spec_p[lane] = p0; s_spec_p[lane + 32] = p1;
                            s_spec_nb[lane] = nb0; s_spec_nb[lane + 32] = nb1;
                            s_spec_ty[lane] = ty0; s_spec_ty[lane + 32] = ty1;
                            // the slot the exact walk would take for pnum3, with the interval of `number` that leads to it
                            const unsigned m0 = __ballot_sync(KMC_FULL_MASK, p0 > 0.0), m1 = __ballot_sync(KMC_FULL_MASK, p1 > 0.0);
                            unsigned long long mm = ((unsigned long long)m1 << 32) | m0;
                            int seln = -1;
                            double acc = 0.0, accb = 0.0;
                            while (mm) {
                                const int n = __ffsll((long long)mm) - 1;
                                mm &= mm - 1;
                                const double pv = __shfl_sync(KMC_FULL_MASK, (n < 32) ? p0 : p1, n & 31);
                                accb = acc;
                                acc = acc + pv;
                                if (acc > pnum3) { seln = n; break; }
                            }
                            if (seln >= 0) {
                                pej = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? nb0 : nb1, seln & 31);
                                const int pty = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? ty0 : ty1, seln & 31);
                                if (lane == 0) {
                                    s_pred.n = seln; s_pred.ej = pej; s_pred.ety = pty;
                                    s_pred.acc_b = accb; s_pred.acc_a = acc;
                                }
                                ok = 1;
                            }
                        }
                    }
                }
                // ---- hand the predicted pair to the housekeeping warp, which stages that event's zero-out inputs
                if (lane == 0) {
                    s_pz_pr = pr;
                    s_pz_pej = ok ? pej : -1;
                    KMC_CBAR();
                    *(volatile int *)&s_pz_req1 = ev_idx + 1;
                    *(volatile int *)&s_pz_req = ev_idx + 1;
                }
                EV_TR(5);
                // what was read above is untouched by this event iff the predicted chunk (row prefixes, rates) / super (chunk
                // prefixes) is not in this event's dirty list, which is complete once barrier B2 has been passed
                int clean = 0;
                if (pchunk >= 0) {
                    while (*(volatile int *)&s_b2_event < ev_idx + 1) { }
                    KMC_CBAR();
                    const int ndc = *(volatile int *)&n_chunks;
                    bool dc = false, dsup = false;
                    for (int q = lane; q < ndc; q += 32) {
                        const int c = chunk_list[q];
                        dc |= (c == pchunk);
                        dsup |= ((c >> 8) == (pchunk >> 8));
                    }
                    clean = (__any_sync(KMC_FULL_MASK, dc) ? 0 : 1) | (__any_sync(KMC_FULL_MASK, dsup) ? 0 : 2);
                }
                __syncwarp();
                if (lane == 0) {
                    s_spec_chunk = pchunk;
                    s_spec_r = pr;
                    s_pred_ok = ok;
                    s_pred_clean = clean;
                    KMC_CBAR();
                    *(volatile int *)&s_spec_ready = ev_idx + 1;
                }
                EV_TR(2);
            }
        } else if (warp == HW) {
            // =========================== H: housekeeping -- beside Z / R1 / R2, no barrier of its own ======================
            // residence time of this event (kmc_events.cu:515; the next selector's stop test waits for it), event
            // application, log, uniforms of event ev_idx + 2
            if (lane == 0) {
                const EvRecord rec = s_rec[ev_idx & 1];
                s_event_time = s_lg[ev_idx & 3] / rec.psum;
                KMC_CBAR();
                *(volatile int *)&s_h_done = ev_idx + 1;
                if (rec.i >= 0) {  // execute_event: kmc_events.cu:305-328
                    const int i = rec.i, j = rec.j, ty = rec.ty;
                    if (ty == KMCB200_VACANCY_GENERATION) {
                        a.element[i] = KMCB200_OXYGEN_DEFECT; a.element[j] = KMCB200_VACANCY;
                        a.charge[i] = -2; a.charge[j] = 2;
                    } else if (ty == KMCB200_VACANCY_RECOMBINATION) {
                        a.element[i] = KMCB200_DEFECT; a.element[j] = KMCB200_O;
                        a.charge[i] = 0; a.charge[j] = 0;
                    } else if (ty == KMCB200_VACANCY_DIFFUSION || ty == KMCB200_ION_DIFFUSION) {
                        int e_i = a.element[i], e_j = a.element[j], q_i = a.charge[i], q_j = a.charge[j];
                        a.element[i] = e_j; a.element[j] = e_i;
                        a.charge[i] = q_j; a.charge[j] = q_i;
                    }
                    if (ev_idx < a.log_cap) {  // event log (i, j, type, slot) + Psum before the event
                        a.log[4 * ev_idx + 0] = rec.i; a.log[4 * ev_idx + 1] = rec.j;
                        a.log[4 * ev_idx + 2] = rec.ty; a.log[4 * ev_idx + 3] = rec.slot;
                        a.log_psum[ev_idx] = rec.psum;
                    }
                }
                const int e2 = ev_idx + 2;
                s_u1[e2 & 3] = evrng_next_double(rng);
                s_lg[e2 & 3] = -log(evrng_next_double(rng));
            }
            __syncwarp();
            if (a.use_spec) {
                // ---- stage the zero-out inputs of the PREDICTED next event: the reverse-index rows of its two sites, the
                // "may be non-zero" flags of the slots they name (one round trip together with the rows -- reading the
                // rates themselves would be a second, dependent one), and the compacted list of the rows that lose a rate.
                // A flag can only err towards "non-zero" (a redundant, bit-identical row sum), never towards "zero".
                const int ri = s_rec[ev_idx & 1].i, rj = s_rec[ev_idx & 1].j;
                const int nb = (ev_idx + 1) & 1;
                int pk[4] = {-1, -1, -1, -1};
                int fl[4] = {0, 0, 0, 0};
                // first half: the predicted row's site, known one round trip before its partner
                while (*(volatile int *)&s_pz_req1 < ev_idx + 1) { }
                KMC_CBAR();
                const int pr = *(volatile int *)&s_pz_pr;
                EV_TR(3);
                if (pr >= 0) {
                    if (lane < 16) prefetch_l2(a.rowsum + (pr >> 8) * 256 + lane * 16);  // its chunk's row sums, for R2
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        pk[u] = a.rev[pr * REV_STRIDE + lane + 32 * u];
                        fl[u] = a.nzflag[pr * REV_STRIDE + lane + 32 * u];
                    }
                }
                // second half: the partner
                while (*(volatile int *)&s_pz_req < ev_idx + 1) { }
                KMC_CBAR();
                const int pej = *(volatile int *)&s_pz_pej;
                EV_TR(4);
                if (pr >= 0 && pej >= 0) {
                    if (lane < 16) prefetch_l2(a.rowsum + (pej >> 8) * 256 + lane * 16);
#pragma unroll
                    for (int u = 2; u < 4; ++u) {
                        pk[u] = a.rev[pej * REV_STRIDE + lane + 32 * (u - 2)];
                        fl[u] = a.nzflag[pej * REV_STRIDE + lane + 32 * (u - 2)];
                    }
                    int nrows = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int rr = pk[u] >> 6;
                        // (rows of this event's two sites are cleared entirely by this event's zero-out)
                        const bool nz = pk[u] >= 0 && fl[u] == a.gen && rr != ri && rr != rj;
                        s_pz_packed[nb][lane + 32 * u] = pk[u];
                        s_pz_oldp[nb][lane + 32 * u] = nz ? 1.0 : 0.0;
                        const bool loses = nz && rr != pr && rr != pej;
                        const unsigned m = __ballot_sync(KMC_FULL_MASK, loses);
                        if (loses) {
                            s_pz_rows[nb][nrows + __popc(m & ((1u << lane) - 1u))] = rr;
                            const double *rowp = a.prob + rr * nn;  // towards L2 for the next event's row sums
                            prefetch_l2(rowp); prefetch_l2(rowp + 16); prefetch_l2(rowp + 32); prefetch_l2(rowp + nn - 1);
                        }
                        nrows += __popc(m);
                    }
                    EV_TR(5);
                    if (lane == 0) { s_pz_nrows[nb] = nrows; s_pz_event[nb] = ev_idx + 1; }
                }
            }
        } else {
            // =============================== Z: zero-out ================================================================
            // zero_out_events_split (kmc_events.cu:247-266): every slot whose row or neighbour is i or j.  Padded slots
            // already hold rate 0 / NULL_EVENT, so rows i and j are cleared entirely; slots of other rows pointing at i / j
            // come from the reverse index.  4 warps (one per scheduler): the phase is two dependent round trips, not work.
            // the housekeeping warp staged this event's zero-out inputs during the previous event, if it is the predicted one
            const int zb = ev_idx & 1;
            const bool staged = s_rec[zb].fast && s_pz_event[zb] == ev_idx;
            if (staged) {
                // ---- Z and R1 in one phase: the slots to clear and the rows that lose a rate are known, so the row sums do
                // not wait for the zero-out's stores -- each warp clears "its" row's slots in registers
                if (tid < 2 * REV_STRIDE) {
                    const int s_site = (tid < REV_STRIDE) ? ei : ej;
                    const int q = tid & (REV_STRIDE - 1);
                    const int packed = s_pz_packed[zb][tid];
                    if (q < nn) {  // the event's own rows
                        const int sl = s_site * nn + q;
                        a.prob[sl] = 0.0;
                        a.type[sl] = KMCB200_NULL_EVENT;
                    }
                    if (q == 0) {  // ... whose sums become +0.0
                        a.rowsum[s_site] = 0.0;
                        const int c = s_site >> 8;
                        const unsigned bit = 1u << (c & 31);
                        if (!(atomicOr(&chunk_bits[c >> 5], bit) & bit)) {
                            chunk_list[atomicAdd(&n_chunks, 1)] = c;
                            if (atomicAdd(&super_cnt[c >> 8], 1) == 0) atomicAdd(&n_supers, 1);
                        }
                    }
                    if (packed >= 0) {
                        const int sl = (packed >> 6) * nn + (packed & 63);
                        a.type[sl] = KMCB200_NULL_EVENT;
                        if (s_pz_oldp[zb][tid] != 0.0) {
                            a.prob[sl] = 0.0;
                            a.nzflag[s_site * REV_STRIDE + q] = 0;
                        }
                    }
                }
                const int nd = s_pz_nrows[zb];
                for (int qq = warp; qq < nd; qq += NCW) {
                    const int rr = s_pz_rows[zb][qq];
                    const int pb = rr * nn;
                    double p0 = (lane < nn) ? a.prob[pb + lane] : 0.0;
                    double p1 = (lane + 32 < nn) ? a.prob[pb + lane + 32] : 0.0;
                    if (lane == 0) {  // dirty-chunk bookkeeping while the row is in flight
                        const int c = rr >> 8;
                        const unsigned bit = 1u << (c & 31);
                        if (!(atomicOr(&chunk_bits[c >> 5], bit) & bit)) {
                            chunk_list[atomicAdd(&n_chunks, 1)] = c;
                            if (atomicAdd(&super_cnt[c >> 8], 1) == 0) atomicAdd(&n_supers, 1);
                        }
                    }
                    // the slots of this row named by the two reverse-index rows (the row may be named by both)
                    unsigned mlo = 0u, mhi = 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int pk = s_pz_packed[zb][lane + 32 * u];
                        if (pk >= 0 && (pk >> 6) == rr) {
                            const int k = pk & 63;
                            if (k < 32) mlo |= 1u << k; else mhi |= 1u << (k - 32);
                        }
                    }
                    mlo = __reduce_or_sync(KMC_FULL_MASK, mlo);
                    mhi = __reduce_or_sync(KMC_FULL_MASK, mhi);
                    if ((mlo >> lane) & 1u) p0 = 0.0;
                    if ((mhi >> lane) & 1u) p1 = 0.0;
                    const double sacc = warp_row_sum(p0, p1);
                    if (lane == 0) a.rowsum[rr] = sacc;
                }
            } else {
                // (the chunks' row sums start their trip to L2 now: R2 reads them two phases later)
                if (ei >= 0 && warp == NCW - 1) prefetch_l2(a.rowsum + (((lane < 16) ? ei : ej) >> 8) * 256 + (lane & 15) * 16);
                if (ei >= 0 && tid < 2 * REV_STRIDE) {
                    const int s_site = (tid < REV_STRIDE) ? ei : ej;
                    const int q = tid & (REV_STRIDE - 1);
                    const int packed = a.rev[s_site * REV_STRIDE + q];
                    if (q < nn) {  // the event's own rows
                        const int sl = s_site * nn + q;
                        a.prob[sl] = 0.0;
                        a.type[sl] = KMCB200_NULL_EVENT;
                    }
                    if (q == 0) {  // ... whose sums become +0.0
                        a.rowsum[s_site] = 0.0;
                        const int c = s_site >> 8;
                        const unsigned bit = 1u << (c & 31);
                        if (!(atomicOr(&chunk_bits[c >> 5], bit) & bit)) {
                            chunk_list[atomicAdd(&n_chunks, 1)] = c;
                            if (atomicAdd(&super_cnt[c >> 8], 1) == 0) atomicAdd(&n_supers, 1);
                        }
                    }
                    if (packed >= 0) {
                        const int rr = packed >> 6;
                        const int sl = rr * nn + (packed & 63);
                        // a slot that already holds rate 0 does not change its row: only rows that lose a non-zero rate need
                        // their sums repaired (their recomputed sums would be bit-identical anyway)
                        const double oldp = a.prob[sl];
                        a.type[sl] = KMCB200_NULL_EVENT;
                        if (oldp != 0.0) {
                            a.prob[sl] = 0.0;
                            a.nzflag[s_site * REV_STRIDE + q] = 0;
                            if (rr != ei && rr != ej) rows_list[atomicAdd(&n_rows, 1)] = rr;
                        }
                    }
                }
                EV_TR(2);
                asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");  // ---- barrier B1 (the NCW working warps)
                EV_TR(3);
                EV_TICK(1);
                // =============================== R1: row sums, one warp per row that lost a rate ===========================
                // (a row listed twice -- it lost a rate to i and one to j -- is recomputed twice with the same result)
                if (warp < NCW) {
                    const int nd = n_rows;
                    for (int qq = warp; qq < nd; qq += NCW) {
                        const int rr = rows_list[qq];
                        const int pb = rr * nn;
                        const double p0 = (lane < nn) ? a.prob[pb + lane] : 0.0;
                        const double p1 = (lane + 32 < nn) ? a.prob[pb + lane + 32] : 0.0;
                        if (lane == 0) {  // dirty-chunk bookkeeping while the row is in flight
                            const int c = rr >> 8;
                            const unsigned bit = 1u << (c & 31);
                            if (!(atomicOr(&chunk_bits[c >> 5], bit) & bit)) {
                                chunk_list[atomicAdd(&n_chunks, 1)] = c;
                                if (atomicAdd(&super_cnt[c >> 8], 1) == 0) atomicAdd(&n_supers, 1);
                            }
                        }
                        const double sacc = warp_row_sum(p0, p1);
                        if (lane == 0) a.rowsum[rr] = sacc;
                    }
                }
            }
            EV_TR(4);
            asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");  // ---- barrier B2
            EV_TR(5);
            EV_TICK(3);
            const int nc = n_chunks;
            if (tid == 0) *(volatile int *)&s_b2_event = ev_idx + 1;  // (the barrier ordered the dirty lists before this)
            {
                if (nc == 0) i_select = (warp == 0);  // nothing changed (no selectable event): warp 0 selects again
                // ---- the common case, one or two dirty chunks in one super: ONE warp repairs them (the two chunk scans
                // interleaved), re-scans the super and goes on to select -- no hand-over, no atomics, no fences
                const int c0 = chunk_list[0], c1 = chunk_list[nc > 0 ? nc - 1 : 0];
                const bool solo = (nc == 1 || nc == 2) && ((c0 >> 8) == (c1 >> 8));
                if (solo) {
                    if (warp == 0) {
                        Scan256 sa, sb;
                        const double2 *srca = reinterpret_cast<const double2 *>(a.rowsum + c0 * 256 + 8 * lane);
                        const double2 *srcb = reinterpret_cast<const double2 *>(a.rowsum + c1 * 256 + 8 * lane);
#pragma unroll
                        for (int k = 0; k < 4; ++k) { double2 t2 = srca[k]; sa.a[2 * k] = t2.x; sa.a[2 * k + 1] = t2.y; }
#pragma unroll
                        for (int k = 0; k < 4; ++k) { double2 t2 = srcb[k]; sb.a[2 * k] = t2.x; sb.a[2 * k + 1] = t2.y; }
                        warp_scan_256x2(sa, sb);
                        EV_TR(12);
                        double2 *dsta = reinterpret_cast<double2 *>(a.rowincl + c0 * 256 + 8 * lane);
#pragma unroll
                        for (int k = 0; k < 4; ++k) dsta[k] = make_double2(scan_incl(sa, 2 * k), scan_incl(sa, 2 * k + 1));
                        if (nc == 2) {
                            double2 *dstb = reinterpret_cast<double2 *>(a.rowincl + c1 * 256 + 8 * lane);
#pragma unroll
                            for (int k = 0; k < 4; ++k) dstb[k] = make_double2(scan_incl(sb, 2 * k), scan_incl(sb, 2 * k + 1));
                        }
                        const int sidx = c0 >> 8;
                        if (lane == 0) {
                            cs[KMC_CIDX(c0)] = sa.total;
                            cs[KMC_CIDX(c1)] = sb.total;  // (c1 == c0 and sb == sa for a single chunk)
                            chunk_bits[c0 >> 5] = 0u;
                            chunk_bits[c1 >> 5] = 0u;
                            super_cnt[sidx] = 0;
                            n_supers = 0;
                        }
                        __syncwarp();
                        EV_TR(13);
                        Scan256 su;
#pragma unroll
                        for (int k = 0; k < 8; ++k) su.a[k] = cs[KMC_CS(sidx, k)];
                        warp_scan_256(su);
#pragma unroll
                        for (int k = 0; k < 8; ++k) ci[KMC_CS(sidx, k)] = scan_incl(su, k);
                        if (lane == 0) ss[tpos(sidx)] = su.total;
                        __syncwarp();
                        EV_TR(14);
                        i_select = true;
                    }
                } else
                for (int qq = warp; qq < nc; qq += NCW) {
                    const int c = chunk_list[qq];
                    Scan256 sc;
                    const double2 *src2 = reinterpret_cast<const double2 *>(a.rowsum + c * 256 + 8 * lane);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { double2 t2 = src2[k]; sc.a[2 * k] = t2.x; sc.a[2 * k + 1] = t2.y; }
                    warp_scan_256(sc);
                    if (qq == warp) EV_TR(12);
                    double2 *dst2 = reinterpret_cast<double2 *>(a.rowincl + c * 256 + 8 * lane);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst2[k] = make_double2(scan_incl(sc, 2 * k), scan_incl(sc, 2 * k + 1));
                    const int sidx = c >> 8;
                    int left = 0;
                    if (lane == 0) {
                        cs[KMC_CIDX(c)] = sc.total;
                        chunk_bits[c >> 5] = 0u;  // every set bit of this word belongs to a dirty chunk handled in this phase
                        __threadfence_block();
                        left = atomicSub(&super_cnt[sidx], 1) - 1;
                    }
                    left = __shfl_sync(KMC_FULL_MASK, left, 0);
                    if (qq == warp) EV_TR(13);
                    if (left == 0) {
                        // ---- this warp finished the last dirty chunk of super sidx: re-scan the super -------------
                        __threadfence_block();
                        Scan256 su;
#pragma unroll
                        for (int k = 0; k < 8; ++k) su.a[k] = cs[KMC_CS(sidx, k)];
                        warp_scan_256(su);
#pragma unroll
                        for (int k = 0; k < 8; ++k) ci[KMC_CS(sidx, k)] = scan_incl(su, k);
                        int sleft = 0;
                        if (lane == 0) {
                            ss[tpos(sidx)] = su.total;
                            __threadfence_block();
                            sleft = atomicSub(&n_supers, 1) - 1;
                        }
                        sleft = __shfl_sync(KMC_FULL_MASK, sleft, 0);
                        EV_TR(14);
                        if (sleft == 0) {  // ... and the last dirty super: this warp selects the next event
                            __threadfence_block();
                            i_select = true;
                        }
                    }
                }
            }
        }
        EV_TR(6);
        EV_TICK(2);
    }
#undef KMC_CIDX
#undef KMC_CS
    // hand the generator back exactly 4 words per executed event in (the uniforms drawn ahead are returned)
    {
        const long long consumed = 4ll * s_nevents;
        const bool curblk = consumed >= rng.block_start;
        const unsigned *src = curblk ? rng.mt : rng.backup;
        for (int q = tid; q < 624; q += EV_THREADS) a.mt_state[q] = src[q];
        if (tid == 0) a.mt_state[624] = (unsigned)(consumed - (curblk ? rng.block_start : rng.prev_block_start));
    }
    if (tid == 0) {
        a.result->event_time = s_event_time;
        a.result->psum_last = s_nevents > 0 ? s_rec[(s_nevents - 1) & 1].psum : 0.0;
        a.result->n_events = s_nevents;
        a.result->error = 0;
#ifdef KMC_EV_PROFILE
        if (a.phase_cycles) for (int q = 0; q < 16; ++q) a.phase_cycles[q] = (q >= 4) ? (long long)s_prof[q] : ph[q];
#endif
#ifdef KMC_EV_TRACE
        if (a.phase_cycles) for (int q = 0; q < 8; ++q) a.phase_cycles[64 * 16 * 16 + q] = s_tr_cnt[q];
#endif
    }
}

__global__ void rng_draw_kernel(unsigned *mt_state, int n, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int pos = (int)mt_state[624];
    for (int q = 0; q < n; ++q) out[q] = mt_next_double(mt_state, pos);
    mt_state[624] = (unsigned)pos;
}

}  // namespace

extern "C" int kmcb200_events_create(kmcb200_ctx *ctx, int N, int nn, const int *neigh, kmcb200_events **ev_out) {
    KMC_CHECK_ARG(ctx && neigh && ev_out, "null pointer");
    KMC_CHECK_ARG(N > 0 && nn > 0 && nn <= 64, "N > 0, 0 < nn <= 64");
    kmcb200_events *ev = new kmcb200_events();
    ev->ctx = ctx;
    ev->N = N;
    ev->nn = nn;
    ev->nchunk = (N + 255) / 256;
    ev->nsuper = (ev->nchunk + 255) / 256;
    if (ev->nsuper > MAX_SUPER) {
        delete ev;
        kmc_set_error("event hierarchy supports up to %d sites", MAX_SUPER * 65536);
        return KMCB200_E_CAPACITY;
    }
    long long total = (long long)N * nn;
    ev->log_cap = 1 << 16;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&ev->prob, (size_t)total * sizeof(double));
    A((void **)&ev->type, (size_t)total);
    A((void **)&ev->row_active, (size_t)N);
    A((void **)&ev->rowsum, (size_t)ev->nchunk * 256 * 2 * sizeof(double));  // rowsum | rowincl, padded to whole chunks (tail stays 0)
    A((void **)&ev->chunksum, (size_t)ev->nsuper * 256 * sizeof(double));  // padded to whole supers (tail stays 0)
    A((void **)&ev->supersum, (size_t)MAX_SUPER * sizeof(double));
    A((void **)&ev->chunkincl, (size_t)ev->nsuper * 256 * sizeof(double));
    A((void **)&ev->mt, 640 * sizeof(unsigned));
    A((void **)&ev->log, (size_t)ev->log_cap * 4 * sizeof(int));
    A((void **)&ev->log_psum, (size_t)ev->log_cap * sizeof(double));
    A((void **)&ev->result, sizeof(EvResult));
    if (e != cudaSuccess) {
        kmc_set_error("cudaMalloc failed in events_create: %s", cudaGetErrorString(e));
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    ev->rowincl = ev->rowsum + (size_t)ev->nchunk * 256;
    {   // persisting-L2 window over rowsum | rowincl (see kmcb200_execute_kmc_step)
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
            prop.accessPolicyMaxWindowSize > 0) {
            size_t want = (size_t)ev->nchunk * 256 * 2 * sizeof(double);
            size_t setaside = want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) == cudaSuccess) {
                ev->l2_window_bytes = want < (size_t)prop.accessPolicyMaxWindowSize ? want : (size_t)prop.accessPolicyMaxWindowSize;
                ev->l2_hit_ratio = (float)((double)setaside / (double)ev->l2_window_bytes);
                if (ev->l2_hit_ratio > 1.0f) ev->l2_hit_ratio = 1.0f;
            } else {
                cudaGetLastError();
            }
        }
    }
    // reverse neighbour index: for each site s the slots (r,n) with neigh[r][n] == s (static)
    int *fill = nullptr;
    int rc = kmc_scratch(ctx, 8, (size_t)(N + 2) * sizeof(int), (void **)&fill);
    if (rc) { kmcb200_events_destroy(ev); return rc; }
    if (cudaMalloc((void **)&ev->rev, (size_t)N * REV_STRIDE * sizeof(int)) != cudaSuccess) {
        kmc_set_error("cudaMalloc(reverse index) failed");
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    if (cudaMalloc((void **)&ev->revpos, (size_t)N * nn) != cudaSuccess ||
        cudaMalloc((void **)&ev->nzflag, (size_t)N * REV_STRIDE) != cudaSuccess) {
        kmc_set_error("cudaMalloc(reverse index flags) failed");
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    cudaMemsetAsync(ev->revpos, 0xff, (size_t)N * nn, ctx->stream);
    cudaMemsetAsync(ev->nzflag, 0, (size_t)N * REV_STRIDE, ctx->stream);
    ev->gen = 0;
    cudaMemsetAsync(ev->rev, 0xff, (size_t)N * REV_STRIDE * sizeof(int), ctx->stream);
    cudaMemsetAsync(ev->rowsum, 0, (size_t)ev->nchunk * 256 * 2 * sizeof(double), ctx->stream);
    cudaMemsetAsync(ev->chunksum, 0, (size_t)ev->nsuper * 256 * sizeof(double), ctx->stream);
    cudaMemsetAsync(fill, 0, (size_t)(N + 2) * sizeof(int), ctx->stream);
    unsigned blocks = (unsigned)((total + 255) / 256);
    kmc_count_launch();
    rev_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(neigh, total, nn, fill, ev->rev, ev->revpos, fill + N + 1);
    int ovf = 0;
    cudaMemcpyAsync(&ovf, fill + N + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        kmc_set_error("reverse index build failed: %s", cudaGetErrorString(cudaGetLastError()));
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    if (ovf) {
        kmc_set_error("a site is listed as neighbour by more than %d rows", REV_STRIDE);
        kmcb200_events_destroy(ev);
        return KMCB200_E_CAPACITY;
    }
    *ev_out = ev;
    return kmcb200_rng_seed(ev, 1u);  // rnd_seed_kmc = 1, src/structure_input.h:8
}

extern "C" int kmcb200_events_destroy(kmcb200_events *ev) {
    if (!ev) return 0;
    if (ev->ctx) cudaStreamSynchronize(ev->ctx->stream);
    cudaFree(ev->prob); cudaFree(ev->type); cudaFree(ev->row_active); cudaFree(ev->rowsum); cudaFree(ev->chunksum); cudaFree(ev->supersum);
    cudaFree(ev->chunkincl);
    cudaFree(ev->rev); cudaFree(ev->revpos); cudaFree(ev->nzflag); cudaFree(ev->mt); cudaFree(ev->log); cudaFree(ev->log_psum);
    cudaFree(ev->result);
    delete ev;
    return 0;
}

extern "C" int kmcb200_set_activation_energies(kmcb200_events *ev, int num_layers, const double *E_gen,
                                               const double *E_rec, const double *E_Vdiff, const double *E_Odiff) {
    KMC_CHECK_ARG(ev && E_gen && E_rec && E_Vdiff && E_Odiff, "null pointer");
    KMC_CHECK_ARG(num_layers > 0 && num_layers <= KMCB200_MAX_LAYERS, "num_layers");
    memset(&ev->energies, 0, sizeof(ev->energies));
    for (int l = 0; l < num_layers; ++l) {
        ev->energies.E_gen[l] = E_gen[l];
        ev->energies.E_rec[l] = E_rec[l];
        ev->energies.E_Vdiff[l] = E_Vdiff[l];
        ev->energies.E_Odiff[l] = E_Odiff[l];
    }
    ev->energies_set = true;
    return 0;
}

extern "C" int kmcb200_rng_set_state(kmcb200_events *ev, const unsigned *mt624_host, int pos) {
    KMC_CHECK_ARG(ev && mt624_host && pos >= 0 && pos <= 624, "arguments");
    unsigned h[625];
    memcpy(h, mt624_host, 624 * sizeof(unsigned));
    h[624] = (unsigned)pos;
    KMC_CUDA(cudaMemcpyAsync(ev->mt, h, sizeof(h), cudaMemcpyHostToDevice, ev->ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ev->ctx->stream));
    return 0;
}

extern "C" int kmcb200_rng_seed(kmcb200_events *ev, unsigned seed) {
    KMC_CHECK_ARG(ev != nullptr, "ev");
    unsigned mt[624];
    mt[0] = seed;  // std::mt19937::seed
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (unsigned)i;
    return kmcb200_rng_set_state(ev, mt, 624);
}

extern "C" int kmcb200_rng_get_state(kmcb200_events *ev, unsigned *mt624_host, int *pos_host) {
    KMC_CHECK_ARG(ev && mt624_host && pos_host, "arguments");
    unsigned h[625];
    KMC_CUDA(cudaMemcpyAsync(h, ev->mt, sizeof(h), cudaMemcpyDeviceToHost, ev->ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ev->ctx->stream));
    memcpy(mt624_host, h, 624 * sizeof(unsigned));
    *pos_host = (int)h[624];
    return 0;
}

extern "C" int kmcb200_rng_draw(kmcb200_events *ev, int n, double *out_host) {
    KMC_CHECK_ARG(ev && out_host && n >= 0, "arguments");
    if (n == 0) return 0;
    double *d = nullptr;
    KMC_TRY(kmc_scratch(ev->ctx, 9, (size_t)n * sizeof(double), (void **)&d));
    kmc_count_launch();
    rng_draw_kernel<<<1, 32, 0, ev->ctx->stream>>>(ev->mt, n, d);
    KMC_CUDA(cudaGetLastError());
    KMC_CUDA(cudaMemcpyAsync(out_host, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ev->ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ev->ctx->stream));
    return 0;
}

extern "C" int kmcb200_build_event_list(kmcb200_ctx *ctx, kmcb200_events *ev, int N, int nn, const int *neigh,
                                        const int *site_layer, double T_bg, double freq, double sigma, double k,
                                        const double *x, const double *y, const double *z,
                                        const double *site_potential_charge, const int *site_element,
                                        const int *site_charge) {
    KMC_CHECK_ARG(ctx && ev && neigh && site_layer && x && y && z && site_potential_charge && site_element && site_charge,
                  "null pointer");
    KMC_CHECK_ARG(N == ev->N && nn == ev->nn, "N/nn differ from events_create");
    KMC_CHECK_ARG(ev->energies_set, "kmcb200_set_activation_energies was not called");
    const double kB = 8.617333262e-5;  // src/kmc_events.cu:5
    double kT = kB * T_bg;
    if (++ev->gen > 255) {  // the generation stamps wrapped: expire every flag explicitly
        KMC_CUDA(cudaMemsetAsync(ev->nzflag, 0, (size_t)N * REV_STRIDE, ctx->stream));
        ev->gen = 1;
    }
    kmc_count_launch();
    build_rates_kernel<<<(unsigned)ev->nchunk, 256, 0, ctx->stream>>>(N, nn, neigh, site_layer, kT, freq, sigma, k, x, y,
                                                                     z, site_potential_charge, site_element,
                                                                     site_charge, ev->energies, ev->prob, ev->type,
                                                                     ev->rowsum, ev->chunksum, ev->rowincl, ev->revpos,
                                                                     ev->nzflag, ev->gen, ev->row_active,
                                                                     ev->list_built ? 0 : 1);
    KMC_CUDA(cudaGetLastError());
    ev->list_built = true;
    kmc_count_launch();
    super_sums_kernel<<<(unsigned)ev->nsuper, 32, 0, ctx->stream>>>(ev->chunksum, ev->nchunk, ev->supersum,
                                                                     ev->chunkincl);
    KMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int kmcb200_execute_kmc_step(kmcb200_ctx *ctx, kmcb200_events *ev, int N, int nn, const int *neigh,
                                        const int *site_layer, double T_bg, double freq, double sigma, double k,
                                        const double *x, const double *y, const double *z,
                                        const double *site_potential_charge, int *site_element, int *site_charge,
                                        int max_events, double *event_time_host, int *n_events_host) {
    KMC_TRY(kmcb200_build_event_list(ctx, ev, N, nn, neigh, site_layer, T_bg, freq, sigma, k, x, y, z,
                                     site_potential_charge, site_element, site_charge));
    EvLoopArgs a;
    a.N = N; a.nn = nn; a.nchunk = ev->nchunk; a.nsuper = ev->nsuper;
    a.neigh = neigh; a.prob = ev->prob; a.type = ev->type;
    a.rowsum = ev->rowsum; a.chunksum = ev->chunksum; a.supersum = ev->supersum;
    a.rowincl = ev->rowincl; a.chunkincl = ev->chunkincl;
    a.rev = ev->rev; a.revpos = ev->revpos; a.nzflag = ev->nzflag; a.gen = ev->gen;
    a.element = site_element; a.charge = site_charge;
    a.mt_state = ev->mt;
    a.inv_freq_threshold = 1 / freq;  // kmc_events.cu:448
    a.max_events = max_events;
    a.log = ev->log; a.log_psum = ev->log_psum; a.log_cap = ev->log_cap;
    a.result = ev->result;
    a.phase_cycles = nullptr;
    a.use_spec = getenv("KMCB200_EV_NO_SPEC") ? 0 : 1;  // tests run both: results must not depend on it
#ifdef KMC_EV_PROFILE
    KMC_TRY(kmc_scratch(ctx, 5, 16 * sizeof(long long), (void **)&a.phase_cycles));
#endif
#ifdef KMC_EV_TRACE
    KMC_TRY(kmc_scratch(ctx, 5, (64 * 16 * 16 + 8) * sizeof(long long), (void **)&a.phase_cycles));
    KMC_CUDA(cudaMemsetAsync(a.phase_cycles, 0, (64 * 16 * 16 + 8) * sizeof(long long), ctx->stream));
#endif
    // chunk sums + their stored prefixes live in shared memory when they fit (up to ~3.2 M sites); larger devices
    // read them from L2
    size_t dyn = (size_t)(2 * ev->nsuper * 256) * sizeof(double);
    bool chunks_in_smem = dyn <= 190 * 1024;
    if (getenv("KMCB200_EV_NO_SMEM")) chunks_in_smem = false;  // tests: exercise the large-device (> 3.2 M sites) path
    if (chunks_in_smem && ctx->smem_cfg_events == 0) {
        KMC_CUDA(cudaFuncSetAttribute(event_loop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(190 * 1024)));
        ctx->smem_cfg_events = 190 * 1024;
    }
    // The row sums and their stored prefixes (one allocation, 16 B per site) are what the loop re-reads at random for
    // every event; pin them in the persisting part of L2 for the duration of the loop so that the selector's row level
    // and the chunk re-scan are L2 hits instead of DRAM round trips.
    const bool persist = ev->l2_window_bytes > 0 && !getenv("KMCB200_EV_NO_L2_PERSIST");
    if (persist) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = ev->rowsum;
        attr.accessPolicyWindow.num_bytes = ev->l2_window_bytes;
        attr.accessPolicyWindow.hitRatio = ev->l2_hit_ratio;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        KMC_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    kmc_count_launch();
    if (chunks_in_smem) event_loop_kernel<true><<<1, EV_THREADS, dyn, ctx->stream>>>(a);
    else event_loop_kernel<false><<<1, EV_THREADS, 0, ctx->stream>>>(a);
    KMC_CUDA(cudaGetLastError());
    if (persist) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.num_bytes = 0;  // no window for whatever is launched on this stream next
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        KMC_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    EvResult *h = (EvResult *)ctx->h_mail;
    KMC_CUDA(cudaMemcpyAsync(h, ev->result, sizeof(EvResult), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (persist) cudaCtxResetPersistingL2Cache();  // the pinned lines go back to normal replacement
#ifdef KMC_EV_PROFILE
    {
        long long ph[16];
        cudaMemcpy(ph, a.phase_cycles, sizeof(ph), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ev profile] events=%d cycles/event:", h->n_events);
        for (int q = 0; q < 16; ++q) fprintf(stderr, " p%d=%.0f", q, (double)ph[q] / (h->n_events > 0 ? h->n_events : 1));
        fprintf(stderr, "\n");
    }
#endif
#ifdef KMC_EV_TRACE
    if (getenv("KMCB200_EV_TRACE_FILE") && h->n_events > EV_TRACE_FIRST + 64) {
        static int dumped = 0;
        if (dumped++ == 1) {   // second superstep
            std::vector<long long> t(64 * 16 * 16 + 8);
            cudaMemcpy(t.data(), a.phase_cycles, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[ev trace] events %lld: fast-path %lld, pred_ok %lld, super_ok %lld, chunk_valid %lld, super match %lld\n",
                    (long long)h->n_events, t[64 * 16 * 16], t[64 * 16 * 16 + 1], t[64 * 16 * 16 + 2], t[64 * 16 * 16 + 3], t[64 * 16 * 16 + 4]);
            FILE *f = fopen(getenv("KMCB200_EV_TRACE_FILE"), "w");
            for (int e = 0; e < 64; ++e)
                for (int w = 0; w < 16; ++w) {
                    fprintf(f, "%d %d", e, w);
                    for (int k = 0; k < 16; ++k) fprintf(f, " %lld", t[(e * 16 + w) * 16 + k]);
                    fprintf(f, "\n");
                }
            fclose(f);
        }
    }
#endif
    ev->last_n_events = h->n_events;
    if (event_time_host) *event_time_host = h->event_time;
    if (n_events_host) *n_events_host = h->n_events;
    return 0;
}

extern "C" int kmcb200_events_pointers(kmcb200_events *ev, double **event_prob, unsigned char **event_type) {
    KMC_CHECK_ARG(ev != nullptr, "ev");
    if (event_prob) *event_prob = ev->prob;
    if (event_type) *event_type = ev->type;
    return 0;
}

extern "C" int kmcb200_events_log(kmcb200_events *ev, int max_rows, int *log_host, double *psum_host, int *rows_host) {
    KMC_CHECK_ARG(ev && rows_host, "arguments");
    int rows = ev->last_n_events;
    if (rows > ev->log_cap) rows = ev->log_cap;
    if (rows > max_rows) rows = max_rows;
    if (rows > 0 && log_host)
        KMC_CUDA(cudaMemcpyAsync(log_host, ev->log, (size_t)rows * 4 * sizeof(int), cudaMemcpyDeviceToHost, ev->ctx->stream));
    if (rows > 0 && psum_host)
        KMC_CUDA(cudaMemcpyAsync(psum_host, ev->log_psum, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, ev->ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ev->ctx->stream));
    *rows_host = rows;
    return 0;
}
