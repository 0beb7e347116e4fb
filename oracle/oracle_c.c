/* TEST INFRASTRUCTURE (oracle) -- not product code.
 *
 * Plain-C restatement of the reference sampler, one function per reference file (paths relative to
 * /root/reference/code/matlab/), following the .m sources statement by statement in fp64:
 *
 *   select_random.m:14-20, asub2ind.m:13-14, bn_sample.m:39-58, dbn_sample.m:36-166 (both
 *   branches), dbn_hierarchical_sample.m:9-37, resample_events.m:11-37, dediscretize.m:7-41,
 *   events2samples.m:9-27, UncorEncounterModel.m:244-307 (driver loop incl. rejection).
 *
 * It exists (a) to check the CUDA path at sizes the Python oracle cannot reach and (b) as the CPU
 * baseline ("port") timed by bench.py on the GPU host's cores (OpenMP over samples).  It keeps the
 * reference's data flow -- event lists, resample pass, de-discretisation pass, dense expansion --
 * and uses none of the product's tricks (no word-space thresholds: every draw is a cumsum + fp64
 * compare exactly like select_random.m).  The model arrays come from the oracle's own reader
 * (oracle/em_read.py); nothing is shared with libemb200.so.  Uniforms: keyed Philox, stream spec v5
 * (oracle/philox.py).
 *
 * PARITY UNPINNED: validated only against the Python oracle (tests/test_oracle_c.py), which in turn
 * is pinned only by SURVEY.md A.8 known answers -- the reference has no tests and MATLAB is absent.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXV 32
#define MAXR 128

typedef struct {
    int32_t n_initial, n_transition, n_dyn, n_gated;
    int32_t is_dynvar_depend;
    const uint8_t* G_initial;     /* n_initial^2, [parent*n + child] */
    const uint8_t* G_transition;  /* n_transition^2 */
    const int32_t* r;             /* r_transition (length n_transition; first n_initial = r_initial) */
    const double* W_initial;      /* N + alpha, concatenated column-major tables */
    const int64_t* off_initial;   /* n_initial offsets into W_initial */
    const double* W_transition;
    const int64_t* off_transition; /* n_transition offsets (-1 for non-dynamic) */
    const int32_t* order_initial;  /* 1-based */
    const int32_t* order_transition;
    const int32_t* temporal_map;   /* n_dyn x 2, 1-based */
    const double* boundaries;      /* concatenated */
    const int32_t* boundaries_off; /* n_initial + 1 */
    const int32_t* zero_bins;      /* 0 = none */
    const double* rates;           /* n_initial */
    const int32_t* gated;          /* n_gated, 1-based ids */
    const uint64_t* gate_G;
    /* driver options */
    const int32_t* start;          /* n_initial, 0 = free */
    int32_t reject_uncor, idx_v, idx_dh, idx_L; /* 1-based */
    int32_t is_quantize500, n_layers;
    const double* layers;          /* n_layers x 2 */
    int32_t max_attempts;
} oc_model;

/* ---- keyed Philox4x32-10 (stream spec v5: counter = (sample_hi, sample_lo, attempt<<16 | purpose<<8 | sub, index); the step stream, purpose 2, carries no attempt) ----------------------------------------------------- */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* o) {
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
/* one cached block per purpose so that the port does not pay 4x for its uniforms (the reference's
 * MT19937 draw is cheap; the CPU baseline should not be inflated by the keyed generator) */
typedef struct { uint64_t seed, sample; uint32_t attempt; int nd, nw, n_initial;
                 uint32_t c_idx[8], c_w3[8], c_o[8][4]; int c_ok[8]; } ukey;
static uint32_t word(ukey* k, uint32_t purpose, uint32_t index, uint32_t sub, uint32_t lane) {
    const uint32_t w3 = ((purpose == 2 ? 0u : k->attempt) << 16) | (purpose << 8) | sub;
    const int q = (int)(purpose & 7);
    if (!k->c_ok[q] || k->c_idx[q] != index || k->c_w3[q] != w3) {
        philox((uint32_t)(k->sample >> 32), (uint32_t)k->sample, w3, index, (uint32_t)k->seed, (uint32_t)(k->seed >> 32), k->c_o[q]);
        k->c_idx[q] = index; k->c_w3[q] = w3; k->c_ok[q] = 1;
    }
    return k->c_o[q][lane];
}
static uint32_t word_at(ukey* k, uint32_t purpose, uint64_t p) { return word(k, purpose, (uint32_t)(p >> 2), 0, (uint32_t)(p & 3)); }
/* select word of initial variable i (0-based) (stream spec v5): four consecutive samples share a call (lane = sample & 3);
 * the cache slot of purpose 1 holds one call, so a variable's word is recomputed when the index changes */
static uint32_t init_word(ukey* k, int i) {
    const uint64_t s = k->sample;
    k->sample = s >> 2;
    const uint32_t w = word(k, 1, (uint32_t)i, 0, (uint32_t)(s & 3));
    k->sample = s;
    k->c_ok[1] = 0;   /* the cached call belongs to the shifted sample: never reuse it under the real one */
    return w;
}
static double init_dd_u(ukey* k, int i, int n) {
    const uint32_t a = init_word(k, i), b = init_word(k, (i + 1) % (n > 2 ? n : 2));
    return ((double)((uint32_t)(a * 0x85EBCA6Bu + b) >> 9) + 0.5) * 1.1920928955078125e-07; /* 23 bits, like the step values */
}
/* step word of second e, gated ordinal g (stream spec v5): one call = four consecutive seconds of one variable */
static uint32_t step_word(ukey* k, int e, int g) { return word(k, 2, (uint32_t)((e >> 2) * k->nw + g), 0, (uint32_t)(e & 3)); }
static double u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }

/* ---- select_random.m:17-20 --------------------------------------------------------------------- */
static int select_random(const double* weights, int r, double rnd) {
    double s[MAXR];
    double acc = 0.0;
    for (int m = 0; m < r; ++m) { acc += weights[m]; s[m] = acc; }   /* cumsum */
    volatile double sthres = s[r - 1] * rnd;
    for (int m = 0; m < r; ++m) if (s[m] >= sthres) return m + 1;     /* find(x, 1, 'first') */
    return r; /* unreachable for finite weights */
}

/* ---- asub2ind.m:13-14 over the parents of column `child` of G ---------------------------------- */
static int64_t parent_index(const uint8_t* G, int n, int child, const int32_t* r, const double* x) {
    double k = 1.0, ndx = 1.0;
    for (int p = 0; p < n; ++p)
        if (G[(size_t)p * n + child]) { ndx += k * (x[p] - 1.0); k *= (double)r[p]; }
    return (int64_t)ndx;
}

typedef struct { double dt, var, val; int kind; int second; } event; /* kind 0 trans, 1 gate, 2 end */

/* ---- dediscretize.m:7-41 ------------------------------------------------------------------------ */
static int dd_needs_u(const oc_model* M, int var, double d) {
    int nb = M->boundaries_off[var] - M->boundaries_off[var - 1];
    return nb > 0 && !(M->zero_bins[var - 1] != 0 && (double)M->zero_bins[var - 1] == d);
}
static double dediscretize(const oc_model* M, int var, double d, double rnd) {
    const double* prm = M->boundaries + M->boundaries_off[var - 1];
    int nb = M->boundaries_off[var] - M->boundaries_off[var - 1];
    if (nb == 0) return d;
    if (M->zero_bins[var - 1] != 0 && (double)M->zero_bins[var - 1] == d) return 0.0;
    int dd = (int)d;
    double a = prm[dd - 1], b = prm[dd];
    volatile double w = b - a;
    volatile double t = w * rnd;
    return a + t;
}

/* one sample: returns attempts used (>=1) or -1 */
static int sample_one(const oc_model* M, uint64_t seed, uint64_t sample, int T, event* ev, event* ev2,
                      double* init_bins, double* initial, double* dense_vals, double* dense_bins, int* n_events_out) {
    const int n = M->n_initial, nt = M->n_transition, nd = M->n_dyn;
    ukey K;
    memset(&K, 0, sizeof(K));
    K.seed = seed; K.sample = sample; K.nd = nd; K.nw = M->n_gated; K.n_initial = n;
    int gate_of_var[MAXV + 1];
    for (int i = 0; i <= n; ++i) gate_of_var[i] = -1;
    for (int g = 0; g < M->n_gated; ++g) gate_of_var[M->gated[g]] = g;
    int gate_of_dyn[8];   /* stream spec v5: the word of (second, variable) also selects the variable's transition */
    for (int d = 0; d < nd; ++d) gate_of_dyn[d] = gate_of_var[M->temporal_map[2 * d]];
    for (int attempt = 0; attempt <= M->max_attempts; ++attempt) {
        K.attempt = (uint32_t)attempt;
        double x[MAXV + 8];
        memset(x, 0, sizeof(x));
        /* bn_sample.m:41-56 */
        for (int oi = 0; oi < n; ++oi) {
            int i = M->order_initial[oi];
            if (M->start && M->start[i - 1]) { x[i - 1] = (double)M->start[i - 1]; continue; }
            int64_t j = parent_index(M->G_initial, n, i - 1, M->r, x);
            const double* w = M->W_initial + M->off_initial[i - 1] + (j - 1) * M->r[i - 1];
            x[i - 1] = (double)select_random(w, M->r[i - 1], u01(init_word(&K, i - 1)));
        }
        for (int i = 0; i < n; ++i) init_bins[i] = x[i];
        /* dbn_sample.m:38-166 */
        int ne = 0;
        if (nd > 0 && T >= 1) {
            double delta_t = 0;
            double s[8][MAXR];
            int rdyn[8];
            if (!M->is_dynvar_depend) { /* :110-135 frozen parent configuration */
                for (int oi = 0; oi < nt; ++oi) {
                    int ii = M->order_transition[oi];
                    for (int d = 0; d < nd; ++d) if (M->temporal_map[2 * d + 1] == ii) {
                        int64_t j = parent_index(M->G_transition, nt, ii - 1, M->r, x);
                        const double* w = M->W_transition + M->off_transition[ii - 1] + (j - 1) * M->r[ii - 1];
                        double acc = 0;
                        rdyn[d] = M->r[ii - 1];
                        for (int m = 0; m < rdyn[d]; ++m) { acc += w[m]; s[d][m] = acc; }
                    }
                }
            }
            for (int t = 2; t <= T; ++t) {
                delta_t += 1;
                double x_old[MAXV + 8];
                memcpy(x_old, x, sizeof(x));
                if (M->is_dynvar_depend) { /* :69-79 */
                    for (int oi = 0; oi < nt; ++oi) {
                        int i = M->order_transition[oi];
                        for (int d = 0; d < nd; ++d) if (M->temporal_map[2 * d + 1] == i) {
                            int64_t j = parent_index(M->G_transition, nt, i - 1, M->r, x);
                            const double* w = M->W_transition + M->off_transition[i - 1] + (j - 1) * M->r[i - 1];
                            double rnd = u01(step_word(&K, t - 1, gate_of_dyn[d]));
                            x[i - 1] = (double)select_random(w, M->r[i - 1], rnd);
                        }
                    }
                } else { /* :143-146 */
                    for (int d = 0; d < nd; ++d) {
                        int ii = M->temporal_map[2 * d + 1];
                        volatile double sthres = s[d][rdyn[d] - 1] * u01(step_word(&K, t - 1, gate_of_dyn[d]));
                        int m = 0;
                        while (!(s[d][m] >= sthres)) ++m;
                        x[ii - 1] = (double)(m + 1);
                    }
                }
                for (int d = 0; d < nd; ++d) x[M->temporal_map[2 * d] - 1] = x[M->temporal_map[2 * d + 1] - 1]; /* :82 */
                for (int i = 1; i <= n; ++i)
                    if (x[i - 1] != x_old[i - 1]) { /* :84-92 */
                        ev[ne].dt = delta_t; ev[ne].var = i; ev[ne].val = x[i - 1]; ev[ne].kind = 0; ev[ne].second = t - 1;
                        ++ne; delta_t = 0;
                    }
            }
        }
        /* dbn_hierarchical_sample.m:15-19 */
        double total = 0;
        for (int e = 0; e < ne; ++e) total += ev[e].dt;
        ev[ne].dt = (double)T - total; ev[ne].var = 0; ev[ne].val = 0; ev[ne].kind = 2; ev[ne].second = T; ++ne;
        /* resample_events.m:11-37 */
        int n2 = 0, second = 0;
        double xr[MAXV];
        for (int i = 0; i < n; ++i) xr[i] = init_bins[i];
        for (int e = 0; e < ne; ++e) {
            int hold = (int)ev[e].dt;
            if (hold == 0) { ev2[n2++] = ev[e]; }
            else {
                double delta_t = 0;
                for (int jj = 0; jj < hold; ++jj) {
                    ++second;
                    int first = 1;
                    delta_t += 1;
                    for (int i = 1; i <= n; ++i) { /* changes = find(rand(size(rates)) < rates) */
                        double u = 0.5;
                        if (gate_of_var[i] >= 0) u = u01(step_word(&K, second, gate_of_var[i]) * 0x9E3779B1u);
                        if (u < M->rates[i - 1]) {
                            ev2[n2].dt = first ? delta_t : 0; ev2[n2].var = i; ev2[n2].val = xr[i - 1];
                            ev2[n2].kind = 1; ev2[n2].second = second; ++n2;
                            first = 0;
                        }
                    }
                    if (!first) delta_t = 0;
                }
                ev2[n2] = ev[e]; ev2[n2].dt = delta_t; ++n2;
            }
            if (ev[e].var > 0) xr[(int)ev[e].var - 1] = ev[e].val;
        }
        /* dbn_hierarchical_sample.m:25-37 */
        for (int i = 1; i <= n; ++i) {
            double rnd = 0.5;
            if (dd_needs_u(M, i, x[i - 1] * 0 + init_bins[i - 1])) rnd = init_dd_u(&K, i - 1, n);
            initial[i - 1] = dediscretize(M, i, init_bins[i - 1], rnd);
        }
        for (int e = 0; e + 1 < n2; ++e) {
            int var = (int)ev2[e].var;
            double rnd = 0.5;
            if (dd_needs_u(M, var, ev2[e].val)) {
                /* stream spec v5: fired-gate and transition values both read the variable's word of that second and, as an
                 * independent partner, the same variable's word of the cyclically next second of the same call */
                int g = 0;
                while (M->gated[g] != var) ++g;
                const int sec = ev2[e].second;
                uint32_t k = step_word(&K, sec, g);
                uint32_t kn = step_word(&K, (sec & ~3) | ((sec + 1) & 3), g);
                uint32_t h = k * 0x85EBCA6Bu + kn;
                rnd = ((double)(h >> 9) + 0.5) * 1.1920928955078125e-07; /* 2^-23 */
            }
            /* keep the bin for the dense bin expansion in .second's place holder */
            double bin = ev2[e].val;
            ev2[e].val = dediscretize(M, var, bin, rnd);
            ev2[e].dt = ev2[e].dt; ev2[e].second = (int)bin;
        }
        /* driver: UncorEncounterModel.m:259-279 */
        int good = 1;
        if (M->reject_uncor) {
            if (M->n_layers > 0 || M->is_quantize500) {
                double h = initial[M->idx_L - 1];
                if (M->n_layers > 0) {
                    int L = (int)initial[M->idx_L - 1];
                    volatile double dlt = M->layers[2 * (L - 1) + 1] - M->layers[2 * (L - 1)];
                    volatile double tt = u01(word(&K, 4, 0, 0, 0)) * dlt;
                    h = M->layers[2 * (L - 1)] + tt;
                }
                if (initial[M->idx_dh - 1] == 0 && M->is_quantize500) {
                    double m = fmod(h, 500.0);
                    if (m < 0) m += 500.0;
                    h = 500.0 * (floor(h / 500.0) + (m > 250.0 ? 1.0 : 0.0));
                }
                initial[M->idx_L - 1] = h;
            }
            volatile double lhs = initial[M->idx_v - 1] * 1.68781;
            good = lhs > fabs(initial[M->idx_dh - 1]) / 60.0;
        }
        if (!good) continue;
        /* events2samples.m:9-27 (values and, for the checker, bins) */
        if (dense_vals) {
            double xv[MAXV], xb[MAXV];
            for (int i = 0; i < n; ++i) { xv[i] = initial[i]; xb[i] = init_bins[i]; }
            int t = 0;
            for (int e = 0; e < n2; ++e) {
                int dt = (int)ev2[e].dt;
                if (ev2[e].var == 0) {
                    t = t + 1;
                    for (int c = t; c <= t + dt - 1; ++c)
                        for (int i = 0; i < n; ++i) { dense_vals[(size_t)i * T + c - 1] = xv[i]; dense_bins[(size_t)i * T + c - 1] = xb[i]; }
                } else {
                    if (dt > 0) {
                        for (int c = t + 1; c <= t + dt; ++c)
                            for (int i = 0; i < n; ++i) { dense_vals[(size_t)i * T + c - 1] = xv[i]; dense_bins[(size_t)i * T + c - 1] = xb[i]; }
                        t = t + dt;
                    }
                    xv[(int)ev2[e].var - 1] = ev2[e].val;
                    xb[(int)ev2[e].var - 1] = (double)ev2[e].second;
                }
            }
        }
        if (n_events_out) *n_events_out = n2;
        return attempt + 1;
    }
    return -1;
}

/* Batch driver (UncorEncounterModel.m:244 loop / em_sample.m:78 loop), OpenMP over samples.
 * Outputs (any may be NULL):
 *   init_bins  [n][n_initial] int8,  init_values [n][n_initial] double, attempts [n] int32,
 *   samples    [n][n_initial][T] double  (out_samples),  sample_bins [n][n_initial][T] int8,
 *   n_events   [n] int32.
 * Returns 0, or -1 if a sample exhausted max_attempts. */
int oc_sample_tracks(const oc_model* M, uint64_t seed, uint64_t first_sample, int64_t n, int32_t T, int32_t n_threads,
                     int8_t* init_bins, double* init_values, int32_t* attempts, double* samples, int8_t* sample_bins,
                     int32_t* n_events) {
    int status = 0;
    const int ni = M->n_initial;
    const size_t cap = (size_t)(M->n_dyn + M->n_gated + 1) * (size_t)(T + 1) + 8;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        event* ev = (event*)malloc(cap * sizeof(event));
        event* ev2 = (event*)malloc(cap * sizeof(event));
        double* dv = (double*)malloc((size_t)ni * T * sizeof(double));
        double* db = (double*)malloc((size_t)ni * T * sizeof(double));
        double ib[MAXV], iv[MAXV];
#pragma omp for schedule(dynamic, 64)
        for (int64_t s = 0; s < n; ++s) {
            int ne = 0;
            int a = sample_one(M, seed, first_sample + (uint64_t)s, T, ev, ev2, ib, iv, dv, db, &ne);
            if (a < 0) {
#pragma omp atomic write
                status = -1;
                a = M->max_attempts + 1;
            }
            if (attempts) attempts[s] = a;
            if (n_events) n_events[s] = ne;
            for (int i = 0; i < ni; ++i) {
                if (init_bins) init_bins[s * ni + i] = (int8_t)ib[i];
                if (init_values) init_values[s * ni + i] = iv[i];
            }
            if (samples) memcpy(samples + (size_t)s * ni * T, dv, (size_t)ni * T * sizeof(double));
            if (sample_bins)
                for (size_t q = 0; q < (size_t)ni * T; ++q) sample_bins[(size_t)s * ni * T + q] = (int8_t)db[q];
        }
        free(ev); free(ev2); free(dv); free(db);
    }
    return status;
}

/* bn_sample.m:39-58 over num_samples + dediscretize of the initial vector (config 2; also
 * @CorTerminalModel/sample.m:29-77 with box rejection when box_lo/box_hi are given). */
int oc_sample_initial(const oc_model* M, uint64_t seed, uint64_t first_sample, int64_t n, int32_t n_threads,
                      const double* box_lo, const double* box_hi, int8_t* bins, double* values, int32_t* attempts) {
    int status = 0;
    const int ni = M->n_initial;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t s = 0; s < n; ++s) {
        ukey K;
        memset(&K, 0, sizeof(K));
        K.seed = seed; K.sample = first_sample + (uint64_t)s; K.n_initial = ni;
        double x[MAXV], v[MAXV];
        int a;
        for (a = 0; a <= M->max_attempts; ++a) {
            K.attempt = (uint32_t)a;
            memset(x, 0, sizeof(x));
            for (int oi = 0; oi < ni; ++oi) {
                int i = M->order_initial[oi];
                if (M->start && M->start[i - 1]) { x[i - 1] = (double)M->start[i - 1]; continue; }
                int64_t j = parent_index(M->G_initial, ni, i - 1, M->r, x);
                const double* w = M->W_initial + M->off_initial[i - 1] + (j - 1) * M->r[i - 1];
                x[i - 1] = (double)select_random(w, M->r[i - 1], u01(init_word(&K, i - 1)));
            }
            int good = 1;
            for (int i = 1; i <= ni; ++i) {
                double rnd = 0.5;
                if (dd_needs_u(M, i, x[i - 1])) rnd = init_dd_u(&K, i - 1, ni);
                v[i - 1] = dediscretize(M, i, x[i - 1], rnd);
                if (box_lo && !(v[i - 1] >= box_lo[i - 1] && v[i - 1] <= box_hi[i - 1])) good = 0;
            }
            if (good) break;
        }
        if (a > M->max_attempts) {
#pragma omp atomic write
            status = -1;
            a = M->max_attempts;
        }
        if (attempts) attempts[s] = a + 1;
        for (int i = 0; i < ni; ++i) {
            if (bins) bins[s * ni + i] = (int8_t)x[i];
            if (values) values[s * ni + i] = v[i];
        }
    }
    return status;
}

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
