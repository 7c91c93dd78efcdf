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
constexpr int MAX_DIRTY = 1024;
constexpr int MAX_SUPER = 256;
}  // namespace

struct EvEnergies {
    double E_gen[KMCB200_MAX_LAYERS], E_rec[KMCB200_MAX_LAYERS], E_Vdiff[KMCB200_MAX_LAYERS], E_Odiff[KMCB200_MAX_LAYERS];
};

struct EvResult {
    double event_time;
    double psum_last;
    int n_events;
    int error;  // 1: dirty list overflow
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
    double *rowincl = nullptr, *chunkincl = nullptr;
    int *rev = nullptr;  // N * REV_STRIDE
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
                                int *__restrict__ rev, int *__restrict__ overflow) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    int j = neigh[s];
    if (j >= 0) {
        int pos = atomicAdd(fill + j, 1);
        if (pos < REV_STRIDE - 1) rev[(size_t)j * REV_STRIDE + pos] = (int)((s / nn) << 6) | (int)(s % nn);  // last entry stays free (-1)
        else atomicExch(overflow, 1);
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
__device__ __forceinline__ int warp_pick_incl(const double inc[8], double number, double *prev, LoadV load_v) {
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
__global__ void __launch_bounds__(256) build_rates_kernel(int N, int nn, const int *__restrict__ neigh,
                                                         const int *__restrict__ layer, double kT, double freq,
                                                         double sigma, double k, const double *__restrict__ x,
                                                         const double *__restrict__ y, const double *__restrict__ z,
                                                         const double *__restrict__ pot,
                                                         const int *__restrict__ element,
                                                         const int *__restrict__ charge, EvEnergies E,
                                                         double *__restrict__ prob, unsigned char *__restrict__ type,
                                                         double *__restrict__ rowsum, double *__restrict__ chunksum,
                                                         double *__restrict__ rowincl) {
    __shared__ double rs[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = blockIdx.x * 256 + w * 32;
    double my_rowsum = 0.0;  // lane q ends up holding the sum of row row0 + q
    for (int q = 0; q < 32; ++q) {
        int i = row0 + q;
        double s = 0.0;
        if (i < N) {
            int el_i = element[i];
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
                }
                if (lane + 32 < nn) {
                    int j = neigh[base + lane + 32];
                    if (j >= 0 && j < N)
                        P1 = event_rate(i, j, el_i, c_i, pot_i, xi, yi, zi, element, charge, pot, layer, x, y, z, E, kT,
                                        freq, sigma, k, e1);
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

struct EvLoopArgs {
    int N, nn;
    long long nchunk, nsuper;
    const int *neigh;
    double *prob;
    unsigned char *type;
    double *rowsum, *chunksum, *supersum, *rowincl, *chunkincl;
    const int *rev;
    int *element, *charge;
    unsigned *mt_state;  // 624 + pos
    double inv_freq_threshold;  // 1/freq
    int max_events;
    int *log;
    double *log_psum;
    int log_cap;
    EvResult *result;
    int chunks_in_smem;  // chunk sums cached in dynamic shared memory for the whole loop
    long long *phase_cycles;  // 16 counters (KMC_EV_PROFILE builds)
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ bool smem_set_insert(int *table, int mask, int key) {
    // open addressing; returns true if key was not present.  table entries are -1 when empty.
    unsigned h = ((unsigned)key * 2654435761u) >> 7;
    for (int probe = 0; probe <= mask; ++probe) {
        int slot = (int)((h + probe) & (unsigned)mask);
        int old = atomicCAS(table + slot, -1, key);
        if (old == -1) return true;
        if (old == key) return false;
    }
    return true;  // table full: treat as new (duplicates only cost redundant work)
}

// Draws the (u1, u2) pair of one event ahead of time.  Remembers how to undo it (position + pre-twist state), because
// the pair drawn for the event that ends the loop is never consumed and the generator must stay in step with the host's.
__device__ __forceinline__ void rng_prefetch_pair(unsigned *mt, unsigned *mt_backup, int *mtpos, int *pos_before,
                                                  int *backup_valid, double *pair) {
    int pos = *mtpos;
    *pos_before = pos;
    if (pos + 4 > 624) {
        for (int q = 0; q < 624; ++q) mt_backup[q] = mt[q];
        *backup_valid = 1;
    } else {
        *backup_valid = 0;
    }
    pair[0] = mt_next_double(mt, pos);
    pair[1] = mt_next_double(mt, pos);
    *mtpos = pos;
}

// The persistent event loop: ONE CTA.  Every phase issues all its global loads together, so a phase costs about
// one L2/DRAM round trip; 5 block barriers per event:
//   S   warp 0 (selector, warp-synchronous): top level = scan_256 of the super sums; chunk and row level = compare
//       against the STORED inclusive prefixes (chunkincl in shared memory, rowincl: 1 RT); slot level = ordered walk
//       over the row's non-zero slots (1 RT); publishes (i, j, type); prefetches the reverse-index rows of i and of
//       every candidate j into L2.  Warp 1 meanwhile draws the uniforms of the NEXT event.
//   Z   all warps: zero-out through the fixed-stride reverse index (1 RT + 1 RT) + list of touched rows / unique
//       chunks / supers; one spare thread applies the event to element/charge, another draws the residence time
//   R1  one warp per touched row: butterfly row sum (1 RT)
//   R2  one warp per touched chunk: scan_256 of its 256 row sums (1 RT) -> chunk sum + the chunk's stored prefixes
//   U   one warp per touched super: scan_256 of its chunk sums (shared memory) -> super sum + stored prefixes
// Measured budget and the building-block latencies: DESIGN.md section 3 ("Events").
// SMEM: chunk sums + their stored prefixes live in dynamic shared memory (both padded to whole supers with zeros).
template <bool SMEM>
__global__ void __launch_bounds__(EV_THREADS, 1) event_loop_kernel(EvLoopArgs a) {
    extern __shared__ double cs_smem[];  // [nsuper*256] chunk sums, then [nsuper*256] inclusive prefixes
    __shared__ unsigned mt[624];
    __shared__ double ss[MAX_SUPER];
    __shared__ int rows_list[MAX_DIRTY], chunk_list[MAX_DIRTY], super_list[MAX_SUPER];
    __shared__ int chunk_set[1024];
    __shared__ int super_flag[MAX_SUPER];
    __shared__ int n_rows, n_chunks, n_supers;
    __shared__ double s_event_time, s_u2, s_psum;
    __shared__ double s_upair[2][2];  // uniforms of event e live in s_upair[e & 1] (drawn one event ahead by warp 1)
    __shared__ unsigned mt_backup[624];
    __shared__ int s_pos_before, s_backup_valid;
    __shared__ int s_i, s_j, s_ty, s_slot, s_stop, s_nevents, s_error, s_mtpos;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EV_THREADS / 32;
    const int nn = a.nn;
#ifdef KMC_EV_PROFILE
    long long ph[16] = {0};
    long long t_last = clock64();
#endif
    const int npad = (int)a.nsuper * 256;  // chunk arrays are padded to whole supers (tail = 0)
    double *cs, *ci;                       // chunk sums / inclusive prefixes of the chunk sums per super
    if (SMEM) {
        cs = cs_smem;
        ci = cs_smem + npad;
        for (int q = tid; q < npad; q += EV_THREADS) { cs[q] = a.chunksum[q]; ci[q] = a.chunkincl[q]; }
    } else {
        cs = a.chunksum;
        ci = a.chunkincl;
    }
    for (int q = tid; q < 624; q += EV_THREADS) mt[q] = a.mt_state[q];
    for (int q = tid; q < MAX_SUPER; q += EV_THREADS) {
        ss[q] = (q < a.nsuper) ? a.supersum[q] : 0.0;
        super_flag[q] = 0;
    }
    for (int q = tid; q < 1024; q += EV_THREADS) chunk_set[q] = -1;
    if (tid == 0) {
        s_mtpos = (int)a.mt_state[624];
        s_event_time = 0.0;
        s_nevents = 0;
        s_error = 0;
        s_stop = 0;
        n_rows = 0; n_chunks = 0; n_supers = 0;
    }
    __syncthreads();
    if (tid == 32) rng_prefetch_pair(mt, mt_backup, &s_mtpos, &s_pos_before, &s_backup_valid, s_upair[0]);
    __syncthreads();

    while (true) {
        // =============================== S: selector, warp 0 ===========================================
        if (warp == 1 && lane == 0) {  // draw the uniforms of the NEXT event while warp 0 selects the current one
            bool go1 = (s_event_time < a.inv_freq_threshold) && (a.max_events <= 0 || s_nevents < a.max_events) && !s_error;
            if (go1) rng_prefetch_pair(mt, mt_backup, &s_mtpos, &s_pos_before, &s_backup_valid, s_upair[(s_nevents + 1) & 1]);
        }
        if (warp == 0) {
            bool go = (s_event_time < a.inv_freq_threshold) && (a.max_events <= 0 || s_nevents < a.max_events) && !s_error;
            int ei = -1, ej = -1, ety = KMCB200_NULL_EVENT, eslot = -1;
            if (go) {
                // ---- top level: scan_256 over the super sums (shared memory) ------------------------------
                Scan256 sc;
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { v[k] = ss[8 * lane + k]; sc.a[k] = v[k]; }
                warp_scan_256(sc);
                const double Psum = sc.total;
                // u1 (selection, kmc_events.cu:469) and u2 (residence time, :515) were drawn by warp 1 during the
                // previous event's S phase (or at start-up): the generator is off the critical path
                const double u1 = s_upair[s_nevents & 1][0];
                if (lane == 0) {
                    s_u2 = s_upair[s_nevents & 1][1];
                    s_psum = Psum;
                }
                double number = u1 * Psum;
                double prev;
                int ts = (Psum > 0.0) ? warp_pick_256(sc, v, number, &prev) : -1;
                EV_TICK(8);
                int r = -1;  // all slot indices fit 32 bits: N * nn <= 16.7 M * 64 < 2^31
                if (ts >= 0) {
                    number = number - prev;
                    // ---- chunk level: stored prefixes of super ts ---------------------------------------------
                    double inc[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) inc[k] = ci[ts * 256 + 8 * lane + k];
                    int tc = warp_pick_incl(inc, number, &prev, [&](double *vv) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) vv[k] = cs[ts * 256 + 8 * lane + k];
                    });
                    EV_TICK(9);
                    if (tc >= 0) {
                        number = number - prev;
                        const int chunk = ts * 256 + tc;
                        // ---- row level: stored prefixes of the chunk (one L2 round trip) ------------------------
                        const int rbase = chunk * 256 + 8 * lane;
                        {
                            const double2 *src2 = reinterpret_cast<const double2 *>(a.rowincl + rbase);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                double2 t2 = src2[k];
                                inc[2 * k] = t2.x; inc[2 * k + 1] = t2.y;
                            }
                        }
                        int tr = warp_pick_incl(inc, number, &prev, [&](double *vv) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) vv[k] = (rbase + k < a.N) ? a.rowsum[rbase + k] : 0.0;
                        });
                        EV_TICK(10);
                        if (tr >= 0) {
                            number = number - prev;
                            r = chunk * 256 + tr;
                        }
                    }
                }
                if (r >= 0) {
                    // ---- slot level: lanes hold slots lane and lane+32; walk the non-zero slots in order --------
                    const int base = r * nn;
                    if (lane < 2) prefetch_l2(a.rev + r * REV_STRIDE + 32 * lane);  // for the zero-out phase
                    double p0 = 0.0, p1 = 0.0;
                    int nb0 = -1, nb1 = -1, ty0 = KMCB200_NULL_EVENT, ty1 = KMCB200_NULL_EVENT;
                    if (lane < nn) { p0 = a.prob[base + lane]; nb0 = a.neigh[base + lane]; ty0 = a.type[base + lane]; }
                    if (lane + 32 < nn) { p1 = a.prob[base + lane + 32]; nb1 = a.neigh[base + lane + 32]; ty1 = a.type[base + lane + 32]; }
                    // every candidate partner's reverse-index row starts its trip from DRAM now (the zero-out needs one)
                    if (p0 > 0.0) { prefetch_l2(a.rev + nb0 * REV_STRIDE); prefetch_l2(a.rev + nb0 * REV_STRIDE + 32); }
                    if (p1 > 0.0) { prefetch_l2(a.rev + nb1 * REV_STRIDE); prefetch_l2(a.rev + nb1 * REV_STRIDE + 32); }
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
                    EV_TICK(11);
                    if (seln >= 0) {
                        ej = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? nb0 : nb1, seln & 31);
                        ety = __shfl_sync(KMC_FULL_MASK, (seln < 32) ? ty0 : ty1, seln & 31);
                        ei = r;
                        eslot = base + seln;
                    }
                }
            }
            if (lane == 0) {
                s_stop = !go;
                s_i = ei; s_j = ej; s_ty = ety; s_slot = eslot;
                n_rows = 0;
                n_chunks = 0;
                n_supers = 0;
            }
        }
        __syncthreads();
        EV_TICK(0);
        if (s_stop) break;
        const int ei = s_i, ej = s_j;
        // =============================== Z: zero-out + apply + residence time ================================
        if (tid == EV_THREADS - 1) {  // residence time (kmc_events.cu:515)
            s_event_time = -log(s_u2) / s_psum;
            s_nevents = s_nevents + 1;
        }
        if (ei >= 0) {
            if (tid == EV_THREADS - 32) {  // execute_event: kmc_events.cu:305-328
                const int i = ei, j = ej, ty = s_ty;
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
            }
            // zero-out (zero_out_events_split, kmc_events.cu:247-266).  Padded slots already hold rate 0 / NULL_EVENT,
            // so rows ei and ej are cleared entirely; slots of other rows pointing at ei / ej come from the reverse index.
            if (tid < 2 * REV_STRIDE) {
                const int which = tid / REV_STRIDE, q = tid % REV_STRIDE;
                const int s_site = which ? ej : ei;
                if (q < nn) {
                    const int sl = s_site * nn + q;
                    a.prob[sl] = 0.0;
                    a.type[sl] = KMCB200_NULL_EVENT;
                }
                const int packed = a.rev[s_site * REV_STRIDE + q];
                int rr = -1;
                if (packed >= 0) {
                    const int sl = (packed >> 6) * nn + (packed & 63);
                    // a slot that already holds rate 0 does not change its row: only rows that lose a non-zero rate need
                    // their sums repaired (their recomputed sums would be bit-identical anyway)
                    double oldp = a.prob[sl];
                    a.type[sl] = KMCB200_NULL_EVENT;
                    if (oldp != 0.0) {
                        a.prob[sl] = 0.0;
                        rr = packed >> 6;
                    }
                } else if (q == REV_STRIDE - 1) {
                    rr = s_site;  // the event's own rows (their sums become 0)
                }
                if (rr >= 0) {
                    int pos = atomicAdd(&n_rows, 1);
                    rows_list[pos] = rr;
                    if (smem_set_insert(chunk_set, 1023, rr >> 8)) {
                        chunk_list[atomicAdd(&n_chunks, 1)] = rr >> 8;
                        if (atomicExch(&super_flag[rr >> 16], 1) == 0) super_list[atomicAdd(&n_supers, 1)] = rr >> 16;
                    }
                }
            }
        }
        __syncthreads();
        EV_TICK(1);
        if (ei >= 0) {
            if (tid == 0) {  // event log (i, j, type, slot) + Psum before the event
                int ne = s_nevents - 1;
                if (ne < a.log_cap) {
                    a.log[4 * ne + 0] = ei; a.log[4 * ne + 1] = ej; a.log[4 * ne + 2] = s_ty; a.log[4 * ne + 3] = s_slot;
                    a.log_psum[ne] = s_psum;
                }
            }
            const int nd = n_rows, nc = n_chunks;
            for (int q = tid; q < 1024; q += EV_THREADS) chunk_set[q] = -1;  // next use: the next event's Z phase
#ifdef KMC_EV_PROFILE
            if (tid == 0) { ph[14] += nd; ph[15] += nc; }
#endif
            // =============================== R1: row sums, one warp per touched row ==================================
            // (duplicates in rows_list recompute the same value)
            for (int q = warp; q < nd; q += NW) {
                const int rr = rows_list[q];
                const int pb = rr * nn;
                double p0 = (lane < nn) ? a.prob[pb + lane] : 0.0;
                double p1 = (lane + 32 < nn) ? a.prob[pb + lane + 32] : 0.0;
                double sacc = warp_row_sum(p0, p1);
                if (lane == 0) a.rowsum[rr] = sacc;
            }
            __syncthreads();
            EV_TICK(2);
            // =============================== R2: one warp per touched chunk: scan_256 of its 256 row sums; the total goes
            // to the chunk sums, the inclusive prefixes to rowincl (what the selector's row level compares against) ======
            for (int q = warp; q < nc; q += NW) {
                const int c = chunk_list[q];
                Scan256 sc;
                const double2 *src2 = reinterpret_cast<const double2 *>(a.rowsum + c * 256 + 8 * lane);
#pragma unroll
                for (int k = 0; k < 4; ++k) { double2 t2 = src2[k]; sc.a[2 * k] = t2.x; sc.a[2 * k + 1] = t2.y; }
                warp_scan_256(sc);
                if (lane == 0) {
                    cs[c] = sc.total;
                    if (SMEM) a.chunksum[c] = sc.total;
                }
                double2 *dst2 = reinterpret_cast<double2 *>(a.rowincl + c * 256 + 8 * lane);
#pragma unroll
                for (int k = 0; k < 4; ++k) dst2[k] = make_double2(scan_incl(sc, 2 * k), scan_incl(sc, 2 * k + 1));
            }
            __syncthreads();
            EV_TICK(4);
            // =============================== U: one warp per touched super: scan_256 of its chunk sums ===============
            const int nsd = n_supers;
            for (int q = warp; q < nsd; q += NW) {
                const int sidx = super_list[q];
                Scan256 su;
#pragma unroll
                for (int k = 0; k < 8; ++k) su.a[k] = cs[sidx * 256 + 8 * lane + k];
                warp_scan_256(su);
#pragma unroll
                for (int k = 0; k < 8; ++k) ci[sidx * 256 + 8 * lane + k] = scan_incl(su, k);
                if (lane == 0) {
                    ss[sidx] = su.total;
                    a.supersum[sidx] = su.total;
                    super_flag[sidx] = 0;
                }
            }
            __syncthreads();
            EV_TICK(3);
        }
    }
    // the pair drawn for the event that did not happen is handed back to the generator
    for (int q = tid; q < 624; q += EV_THREADS) a.mt_state[q] = s_backup_valid ? mt_backup[q] : mt[q];
    if (tid == 0) {
        a.mt_state[624] = (unsigned)s_pos_before;
        a.result->event_time = s_event_time;
        a.result->psum_last = s_psum;
        a.result->n_events = s_nevents;
        a.result->error = s_error;
#ifdef KMC_EV_PROFILE
        if (a.phase_cycles) for (int q = 0; q < 16; ++q) a.phase_cycles[q] = ph[q];
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
    A((void **)&ev->rowsum, (size_t)ev->nchunk * 256 * sizeof(double));  // padded to whole chunks (tail stays 0)
    A((void **)&ev->chunksum, (size_t)ev->nsuper * 256 * sizeof(double));  // padded to whole supers (tail stays 0)
    A((void **)&ev->supersum, (size_t)MAX_SUPER * sizeof(double));
    A((void **)&ev->rowincl, (size_t)ev->nchunk * 256 * sizeof(double));
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
    // reverse neighbour index: for each site s the slots (r,n) with neigh[r][n] == s (static)
    int *fill = nullptr;
    int rc = kmc_scratch(ctx, 8, (size_t)(N + 2) * sizeof(int), (void **)&fill);
    if (rc) { kmcb200_events_destroy(ev); return rc; }
    if (cudaMalloc((void **)&ev->rev, (size_t)N * REV_STRIDE * sizeof(int)) != cudaSuccess) {
        kmc_set_error("cudaMalloc(reverse index) failed");
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    cudaMemsetAsync(ev->rev, 0xff, (size_t)N * REV_STRIDE * sizeof(int), ctx->stream);
    cudaMemsetAsync(ev->rowsum, 0, (size_t)ev->nchunk * 256 * sizeof(double), ctx->stream);
    cudaMemsetAsync(ev->chunksum, 0, (size_t)ev->nsuper * 256 * sizeof(double), ctx->stream);
    cudaMemsetAsync(fill, 0, (size_t)(N + 2) * sizeof(int), ctx->stream);
    unsigned blocks = (unsigned)((total + 255) / 256);
    kmc_count_launch();
    rev_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(neigh, total, nn, fill, ev->rev, fill + N + 1);
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
    cudaFree(ev->prob); cudaFree(ev->type); cudaFree(ev->rowsum); cudaFree(ev->chunksum); cudaFree(ev->supersum);
    cudaFree(ev->rowincl); cudaFree(ev->chunkincl);
    cudaFree(ev->rev); cudaFree(ev->mt); cudaFree(ev->log); cudaFree(ev->log_psum);
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
    kmc_count_launch();
    build_rates_kernel<<<(unsigned)ev->nchunk, 256, 0, ctx->stream>>>(N, nn, neigh, site_layer, kT, freq, sigma, k, x, y,
                                                                     z, site_potential_charge, site_element,
                                                                     site_charge, ev->energies, ev->prob, ev->type,
                                                                     ev->rowsum, ev->chunksum, ev->rowincl);
    KMC_CUDA(cudaGetLastError());
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
    a.rev = ev->rev;
    a.element = site_element; a.charge = site_charge;
    a.mt_state = ev->mt;
    a.inv_freq_threshold = 1 / freq;  // kmc_events.cu:448
    a.max_events = max_events;
    a.log = ev->log; a.log_psum = ev->log_psum; a.log_cap = ev->log_cap;
    a.result = ev->result;
    a.phase_cycles = nullptr;
#ifdef KMC_EV_PROFILE
    KMC_TRY(kmc_scratch(ctx, 5, 16 * sizeof(long long), (void **)&a.phase_cycles));
#endif
    // chunk sums + their stored prefixes live in shared memory when they fit (up to ~3.2 M sites); larger devices
    // read them from L2
    size_t dyn = (size_t)(2 * ev->nsuper * 256) * sizeof(double);
    a.chunks_in_smem = dyn <= 190 * 1024 ? 1 : 0;
    if (getenv("KMCB200_EV_NO_SMEM")) a.chunks_in_smem = 0;  // tests: exercise the large-device (> 3.2 M sites) path
    if (a.chunks_in_smem && ctx->smem_cfg_events == 0) {
        KMC_CUDA(cudaFuncSetAttribute(event_loop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(190 * 1024)));
        ctx->smem_cfg_events = 190 * 1024;
    }
    kmc_count_launch();
    if (a.chunks_in_smem) event_loop_kernel<true><<<1, EV_THREADS, dyn, ctx->stream>>>(a);
    else event_loop_kernel<false><<<1, EV_THREADS, 0, ctx->stream>>>(a);
    KMC_CUDA(cudaGetLastError());
    EvResult *h = (EvResult *)ctx->h_mail;
    KMC_CUDA(cudaMemcpyAsync(h, ev->result, sizeof(EvResult), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h->error) {
        kmc_set_error("event loop: more than %d touched rows in one event", MAX_DIRTY);
        return KMCB200_E_CAPACITY;
    }
#ifdef KMC_EV_PROFILE
    {
        long long ph[16];
        cudaMemcpy(ph, a.phase_cycles, sizeof(ph), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ev profile] events=%d cycles/event:", h->n_events);
        for (int q = 0; q < 16; ++q) fprintf(stderr, " p%d=%.0f", q, (double)ph[q] / (h->n_events > 0 ? h->n_events : 1));
        fprintf(stderr, "\n");
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
