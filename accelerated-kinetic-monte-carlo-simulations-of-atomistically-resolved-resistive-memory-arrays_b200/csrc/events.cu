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
constexpr int EV_THREADS = 1024;
constexpr int EV_GROUPS = EV_THREADS / 256;
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
    int *rev_ptr = nullptr, *rev_slot = nullptr;
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

__global__ void rev_count_kernel(const int *__restrict__ neigh, long long total, int *__restrict__ cnt) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    int j = neigh[s];
    if (j >= 0) atomicAdd(cnt + j, 1);
}
__global__ void rev_fill_kernel(const int *__restrict__ neigh, long long total, const int *__restrict__ rev_ptr,
                                int *__restrict__ fill, int *__restrict__ rev_slot) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    int j = neigh[s];
    if (j >= 0) rev_slot[rev_ptr[j] + atomicAdd(fill + j, 1)] = (int)s;
}

// summation spec block_scan_256 for one 256-thread group; every thread of the CTA must call it.
__device__ __forceinline__ double block_scan_256_dev(double v, int t, double *sm8) {
    double x = kmc_warp_inclusive_scan(v);
    int w = t >> 5;
    if ((t & 31) == 31) sm8[w] = x;
    __syncthreads();
    double incl = x;
    if (w > 0) {
        double pre = sm8[0];
        for (int q = 1; q < w; ++q) pre = pre + sm8[q];
        incl = pre + x;
    }
    __syncthreads();
    return incl;
}

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
                                                         double *__restrict__ rowsum, double *__restrict__ chunksum) {
    __shared__ double rs[256];
    __shared__ double sm8[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = blockIdx.x * 256 + w * 32;
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
            // row sum: sequential over the nn slots (summation spec); all-zero rows short-cut to +0.0
            if (__any_sync(KMC_FULL_MASK, (P0 != 0.0) || (P1 != 0.0))) {
                s = __shfl_sync(KMC_FULL_MASK, P0, 0);
                for (int n = 1; n < nn; ++n) {
                    double v = __shfl_sync(KMC_FULL_MASK, (n < 32) ? P0 : P1, n & 31);
                    s = s + v;
                }
            }
            if (lane == 0) rowsum[i] = s;
        }
        if (lane == 0) rs[w * 32 + q] = s;
    }
    __syncthreads();
    double incl = block_scan_256_dev(rs[threadIdx.x], threadIdx.x, sm8);
    if (threadIdx.x == 255) chunksum[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(256) super_sums_kernel(const double *__restrict__ chunksum, long long nchunk,
                                                        double *__restrict__ supersum) {
    __shared__ double sm8[8];
    long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    double v = (c < nchunk) ? chunksum[c] : 0.0;
    double incl = block_scan_256_dev(v, threadIdx.x, sm8);
    if (threadIdx.x == 255) supersum[blockIdx.x] = incl;
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

struct EvLoopArgs {
    int N, nn;
    long long nchunk, nsuper;
    const int *neigh;
    double *prob;
    unsigned char *type;
    double *rowsum, *chunksum, *supersum;
    const int *rev_ptr, *rev_slot;
    int *element, *charge;
    unsigned *mt_state;  // 624 + pos
    double inv_freq_threshold;  // 1/freq
    int max_events;
    int *log;
    double *log_psum;
    int log_cap;
    EvResult *result;
};

// first t with incl > number, else last t with v > 0, else -1.  Group-parallel; result broadcast through smem.
__device__ __forceinline__ void group_pick(double incl, double v, double number, int t, int *sm_first, int *sm_last) {
    if (incl > number) atomicMin(sm_first, t);
    if (v > 0.0) atomicMax(sm_last, t);
}

__global__ void __launch_bounds__(EV_THREADS, 1) event_loop_kernel(EvLoopArgs a) {
    __shared__ unsigned mt[624];
    __shared__ int mt_pos;
    __shared__ double ss[MAX_SUPER];
    __shared__ double incl_sm[256];
    __shared__ double sm8[EV_GROUPS][8];
    __shared__ int dirty[MAX_DIRTY];
    __shared__ int ndirty;
    __shared__ unsigned char super_dirty[MAX_SUPER];
    __shared__ int pick_first, pick_last;
    __shared__ double s_number, s_psum, s_event_time;
    __shared__ int s_sel, s_i, s_j, s_stop, s_nevents, s_error;

    const int tid = threadIdx.x;
    const int grp = tid >> 8, t = tid & 255;
    const int nn = a.nn;
    for (int q = tid; q < 624; q += EV_THREADS) mt[q] = a.mt_state[q];
    for (int q = tid; q < MAX_SUPER; q += EV_THREADS) {
        ss[q] = (q < a.nsuper) ? a.supersum[q] : 0.0;
        super_dirty[q] = 0;
    }
    if (tid == 0) {
        mt_pos = (int)a.mt_state[624];
        s_event_time = 0.0;
        s_nevents = 0;
        s_error = 0;
        s_stop = 0;
    }
    __syncthreads();

    while (true) {
        // ---- loop condition (kmc_events.cu:448) + top level ---------------------------------------------
        if (tid == 0) {
            bool go = (s_event_time < a.inv_freq_threshold) && (a.max_events <= 0 || s_nevents < a.max_events) && !s_error;
            s_stop = !go;
            if (go) {
                double acc = ss[0];
                for (int s = 1; s < a.nsuper; ++s) acc = acc + ss[s];
                double Psum = acc;
                s_psum = Psum;
                double number = mt_next_double(mt, mt_pos) * Psum;  // :469
                int sel = -1;
                if (Psum > 0.0) {
                    double c = ss[0], prev = 0.0;
                    for (int s = 0; s < a.nsuper; ++s) {
                        if (s > 0) { prev = c; c = c + ss[s]; }
                        if (c > number) { sel = s; if (s > 0) number = number - prev; break; }
                    }
                    if (sel < 0) {  // rounding pushed number past the total: clamp to the last non-empty super
                        double c2 = ss[0];
                        double prev2 = 0.0;
                        int last = -1; double last_prev = 0.0;
                        for (int s = 0; s < a.nsuper; ++s) {
                            if (s > 0) { prev2 = c2; c2 = c2 + ss[s]; }
                            if (ss[s] > 0.0) { last = s; last_prev = prev2; }
                        }
                        sel = last;
                        if (sel > 0) number = number - last_prev;
                    }
                }
                s_sel = sel;
                s_number = number;
                pick_first = 256;
                pick_last = -1;
            }
        }
        __syncthreads();
        if (s_stop) break;
        int sel_super = s_sel;
        if (sel_super >= 0) {
            // ---- chunk level (all groups compute redundantly; group 0 publishes) ---------------------------
            long long c_idx = (long long)sel_super * 256 + t;
            double v = (c_idx < a.nchunk) ? a.chunksum[c_idx] : 0.0;
            double incl = block_scan_256_dev(v, t, sm8[grp]);
            if (grp == 0) { incl_sm[t] = incl; group_pick(incl, v, s_number, t, &pick_first, &pick_last); }
            __syncthreads();
            if (tid == 0) {
                int tc = (pick_first < 256) ? pick_first : pick_last;
                if (tc > 0) s_number = s_number - incl_sm[tc - 1];
                s_sel = tc;
                pick_first = 256;
                pick_last = -1;
            }
            __syncthreads();
            int tc = s_sel;
            long long chunk = (long long)sel_super * 256 + tc;
            // ---- row level --------------------------------------------------------------------------------
            long long r_idx = chunk * 256 + t;
            v = (tc >= 0 && r_idx < a.N) ? a.rowsum[r_idx] : 0.0;
            incl = block_scan_256_dev(v, t, sm8[grp]);
            if (grp == 0) { incl_sm[t] = incl; group_pick(incl, v, s_number, t, &pick_first, &pick_last); }
            __syncthreads();
            // ---- slot level + apply (thread 0) --------------------------------------------------------------
            if (tid == 0) {
                int tr = (tc >= 0) ? ((pick_first < 256) ? pick_first : pick_last) : -1;
                long long slot = -1;
                if (tr >= 0) {
                    double number = s_number;
                    if (tr > 0) number = number - incl_sm[tr - 1];
                    long long r = chunk * 256 + tr;
                    const double *p = a.prob + r * (long long)nn;
                    double acc = p[0];
                    int seln = -1;
                    for (int n = 0; n < nn; ++n) {
                        if (n > 0) acc = acc + p[n];
                        if (acc > number) { seln = n; break; }
                    }
                    if (seln < 0)
                        for (int n = nn - 1; n >= 0; --n)
                            if (p[n] > 0.0) { seln = n; break; }
                    if (seln >= 0) slot = r * (long long)nn + seln;
                }
                s_i = -1;
                s_j = -1;
                if (slot >= 0) {
                    int i = (int)(slot / nn);
                    int j = a.neigh[slot];
                    int ty = a.type[slot];
                    int ne = s_nevents;
                    if (ne < a.log_cap) {
                        a.log[4 * ne + 0] = i; a.log[4 * ne + 1] = j; a.log[4 * ne + 2] = ty; a.log[4 * ne + 3] = (int)slot;
                        a.log_psum[ne] = s_psum;
                    }
                    // execute_event: kmc_events.cu:305-328
                    if (ty == KMCB200_VACANCY_GENERATION) {
                        a.element[i] = KMCB200_OXYGEN_DEFECT; a.element[j] = KMCB200_VACANCY;
                        a.charge[i] = -2; a.charge[j] = 2;
                    } else if (ty == KMCB200_VACANCY_RECOMBINATION) {
                        a.element[i] = KMCB200_DEFECT; a.element[j] = KMCB200_O;
                        a.charge[i] = 0; a.charge[j] = 0;
                    } else if (ty == KMCB200_VACANCY_DIFFUSION || ty == KMCB200_ION_DIFFUSION) {
                        int te = a.element[i]; a.element[i] = a.element[j]; a.element[j] = te;
                        int tq = a.charge[i]; a.charge[i] = a.charge[j]; a.charge[j] = tq;
                    }
                    s_i = i;
                    s_j = j;
                    dirty[0] = i;
                    dirty[1] = j;
                    ndirty = 2;
                } else {
                    ndirty = 0;
                }
            }
            __syncthreads();
            const int ei = s_i, ej = s_j;
            if (ei >= 0) {
                // ---- zero-out (zero_out_events_split, kmc_events.cu:247-266) through the reverse index -------
                for (int q = tid; q < 2 * nn; q += EV_THREADS) {
                    long long sl = (long long)((q < nn) ? ei : ej) * nn + (q % nn);
                    if (a.neigh[sl] >= 0) { a.prob[sl] = 0.0; a.type[sl] = KMCB200_NULL_EVENT; }
                }
                for (int which = 0; which < 2; ++which) {
                    int s = which ? ej : ei;
                    int b = a.rev_ptr[s], e = a.rev_ptr[s + 1];
                    for (int q = b + tid; q < e; q += EV_THREADS) {
                        int sl = a.rev_slot[q];
                        a.prob[sl] = 0.0;
                        a.type[sl] = KMCB200_NULL_EVENT;
                        int pos = atomicAdd(&ndirty, 1);
                        if (pos < MAX_DIRTY) dirty[pos] = sl / nn;
                        else s_error = 1;
                    }
                }
                __syncthreads();
                int nd = min(ndirty, MAX_DIRTY);
                // ---- repair row sums ---------------------------------------------------------------------
                for (int q = tid; q < nd; q += EV_THREADS) {
                    int r = dirty[q];
                    const double *p = a.prob + (long long)r * nn;
                    double s = p[0];
                    for (int n = 1; n < nn; ++n) s = s + p[n];
                    a.rowsum[r] = s;
                    super_dirty[r >> 16] = 1;
                }
                __syncthreads();
                // ---- repair chunk sums: EV_GROUPS chunks per round ------------------------------------------
                for (int base = 0; base < nd; base += EV_GROUPS) {
                    int q = base + grp;
                    long long c = -1;
                    if (q < nd) {
                        c = dirty[q] >> 8;
                        if (q > 0 && (dirty[q - 1] >> 8) == c) c = -1;  // same chunk as the previous entry
                    }
                    long long r_i = c * 256 + t;
                    double vv = (c >= 0 && r_i < a.N) ? a.rowsum[r_i] : 0.0;
                    double inc = block_scan_256_dev(vv, t, sm8[grp]);
                    if (c >= 0 && t == 255) a.chunksum[c] = inc;
                }
                __syncthreads();
                // ---- repair super sums ---------------------------------------------------------------------
                for (int s = 0; s < a.nsuper; ++s) {
                    if (super_dirty[s]) {  // uniform (shared memory)
                        long long c_i = (long long)s * 256 + t;
                        double vv = (c_i < a.nchunk) ? a.chunksum[c_i] : 0.0;
                        double inc = block_scan_256_dev(vv, t, sm8[grp]);
                        if (tid == 255) { ss[s] = inc; a.supersum[s] = inc; }
                        __syncthreads();
                        if (tid == 0) super_dirty[s] = 0;
                    }
                }
                __syncthreads();
            }
        }
        // ---- residence time (kmc_events.cu:515) -----------------------------------------------------------
        if (tid == 0) {
            double u2 = mt_next_double(mt, mt_pos);
            s_event_time = -log(u2) / s_psum;
            s_nevents = s_nevents + 1;
        }
        __syncthreads();
    }
    for (int q = tid; q < 624; q += EV_THREADS) a.mt_state[q] = mt[q];
    if (tid == 0) {
        a.mt_state[624] = (unsigned)mt_pos;
        a.result->event_time = s_event_time;
        a.result->psum_last = s_psum;
        a.result->n_events = s_nevents;
        a.result->error = s_error;
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
    A((void **)&ev->rowsum, (size_t)N * sizeof(double));
    A((void **)&ev->chunksum, (size_t)ev->nchunk * sizeof(double));
    A((void **)&ev->supersum, (size_t)MAX_SUPER * sizeof(double));
    A((void **)&ev->rev_ptr, (size_t)(N + 1) * sizeof(int));
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
    int rc = kmc_scratch(ctx, 8, (size_t)(N + 1) * sizeof(int), (void **)&fill);
    if (rc) { kmcb200_events_destroy(ev); return rc; }
    cudaMemsetAsync(ev->rev_ptr, 0, (size_t)(N + 1) * sizeof(int), ctx->stream);
    cudaMemsetAsync(fill, 0, (size_t)(N + 1) * sizeof(int), ctx->stream);
    unsigned blocks = (unsigned)((total + 255) / 256);
    kmc_count_launch();
    rev_count_kernel<<<blocks, 256, 0, ctx->stream>>>(neigh, total, ev->rev_ptr);
    rc = kmc_exclusive_scan_i32(ctx, ev->rev_ptr, ev->rev_ptr, (long long)N + 1, 4);
    if (rc) { kmcb200_events_destroy(ev); return rc; }
    int M = 0;
    cudaMemcpyAsync(&M, ev->rev_ptr + N, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaMalloc((void **)&ev->rev_slot, (size_t)(M + 1) * sizeof(int)) != cudaSuccess) {
        kmc_set_error("reverse index build failed: %s", cudaGetErrorString(cudaGetLastError()));
        kmcb200_events_destroy(ev);
        return KMCB200_E_CUDA;
    }
    kmc_count_launch();
    rev_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(neigh, total, ev->rev_ptr, fill, ev->rev_slot);
    if (cudaGetLastError() != cudaSuccess) { kmc_set_error("rev_fill launch failed"); kmcb200_events_destroy(ev); return KMCB200_E_CUDA; }
    *ev_out = ev;
    return kmcb200_rng_seed(ev, 1u);  // rnd_seed_kmc = 1, src/structure_input.h:8
}

extern "C" int kmcb200_events_destroy(kmcb200_events *ev) {
    if (!ev) return 0;
    if (ev->ctx) cudaStreamSynchronize(ev->ctx->stream);
    cudaFree(ev->prob); cudaFree(ev->type); cudaFree(ev->rowsum); cudaFree(ev->chunksum); cudaFree(ev->supersum);
    cudaFree(ev->rev_ptr); cudaFree(ev->rev_slot); cudaFree(ev->mt); cudaFree(ev->log); cudaFree(ev->log_psum);
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
                                                                     ev->rowsum, ev->chunksum);
    KMC_CUDA(cudaGetLastError());
    kmc_count_launch();
    super_sums_kernel<<<(unsigned)ev->nsuper, 256, 0, ctx->stream>>>(ev->chunksum, ev->nchunk, ev->supersum);
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
    a.rev_ptr = ev->rev_ptr; a.rev_slot = ev->rev_slot;
    a.element = site_element; a.charge = site_charge;
    a.mt_state = ev->mt;
    a.inv_freq_threshold = 1 / freq;  // kmc_events.cu:448
    a.max_events = max_events;
    a.log = ev->log; a.log_psum = ev->log_psum; a.log_cap = ev->log_cap;
    a.result = ev->result;
    kmc_count_launch();
    event_loop_kernel<<<1, EV_THREADS, 0, ctx->stream>>>(a);
    KMC_CUDA(cudaGetLastError());
    EvResult *h = (EvResult *)ctx->h_mail;
    KMC_CUDA(cudaMemcpyAsync(h, ev->result, sizeof(EvResult), cudaMemcpyDeviceToHost, ctx->stream));
    KMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h->error) {
        kmc_set_error("event loop: more than %d touched rows in one event", MAX_DIRTY);
        return KMCB200_E_CAPACITY;
    }
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
