// emb_fast.cuh -- register-resident track sampler for the common model shapes.
//
// Same semantics and same keyed stream as track_generic (emb_device.cuh); what changes is where the
// state lives.  Template parameters fix the number of bins of every dynamic variable (RS packs up to
// four of them, one per byte, in temporal_map order) and the number of gated variables NG, so that
//   * the frozen inverse-CDF thresholds of the fast branch (dbn_sample.m:110-135) sit in registers,
//   * the nw = ND + NG words of four consecutive seconds come from exactly nw Philox calls whose
//     outputs are indexed statically (stream spec v1: p = e*nw + slot, so 4 seconds = nw blocks),
//   * bins are packed 16 per uint4 store and values 4 per float4 store ([var][t/16][sample][16]).
// Requirements checked on the host (emb_kernels.cu: pick_fast): every dynamic variable is gated and
// the gated list ends with the dynamic variables in temporal_map order (true for every shipped
// model); all gate thresholds G are in [1, 2^32 - 1].
#pragma once
#include "emb_device.cuh"

namespace emb {

constexpr int FAST_MAX_EDGES = 160;  // sum of bins over the gated variables (shared-memory edge table)

template <uint32_t RS>
struct DynShape {
    static constexpr int R(int d) { return (int)((RS >> (8 * d)) & 0xFFu); }
    static constexpr int ND = (R(0) > 0) + (R(1) > 0) + (R(2) > 0) + (R(3) > 0);
    static constexpr int RP(int d) { return (R(d) + 3) & ~3; }
    static constexpr int RMAX = R(0) > R(1) ? (R(0) > R(2) ? (R(0) > R(3) ? R(0) : R(3)) : (R(2) > R(3) ? R(2) : R(3)))
                                            : (R(1) > R(2) ? (R(1) > R(3) ? R(1) : R(3)) : (R(2) > R(3) ? R(2) : R(3)));
    static constexpr int RPMAX = (RMAX + 3) & ~3;
};

// per-block constants of the fast kernel (shared memory on the device)
struct FastShared {
    double edges[2 * FAST_MAX_EDGES];  // {a, w} per (gated ordinal, bin)
};

// fill FastShared (called by all threads of a block with their index, or by the host with tid=0,nthreads=1)
EMB_HD void fast_fill_shared(const DevModel& M, FastShared& S, int tid, int nthreads) {
    int base = 0;
    for (int g = 0; g < M.n_gated; ++g) {
        const int v = M.gated_var[g];
        const int r = M.init[v].r;
        if (M.edge_off[v] >= 0)
            for (int q = tid; q < 2 * r; q += nthreads) S.edges[2 * base + q] = ldg64(M.edges + M.edge_off[v] + q);
        base += r;
    }
}

template <uint32_t RS, int NG, bool FAST, class HistInc>
EMB_HD void track_fast(const DevModel& M, const SampleParams& P, const TrackOut& O, int64_t s, const FastShared& S,
                       HistInc hist_inc) {
    using SH = DynShape<RS>;
    constexpr int ND = SH::ND;
    constexpr int NW = ND + NG;
    constexpr int RPM = SH::RPMAX;
    const uint64_t sample = P.first_sample + (uint64_t)s;
    const int T = P.T;
    const int64_t N = P.n;

    // ---- initial network (once per track; generic code, cost amortised over T steps) ------------
    uint32_t bin[ND];        // current 0-based bins of the dynamic variables
    uint32_t gbin[NG];       // current 0-based bins of the gated variables (last ND mirror bin[])
    float val[NG];           // current continuous values of the gated (= time-varying) variables
    uint32_t thr[ND][RPM];   // fast branch: frozen thresholds, [RP-1] = lead
    uint32_t cbase[ND];      // slow branch: column offset from the parents that never change
    int attempt;
    {
        uint8_t x[MAXX];
        double vals[MAXV];
        for (int i = 0; i < MAXX; ++i) x[i] = 0;
        attempt = sample_initial(M, P, sample, x, vals);
        if (attempt < 0) {
            if (O.status) *O.status = 1;
            attempt = P.max_attempts;
        }
        if (O.attempts) O.attempts[s] = (uint16_t)(attempt + 1);
        for (int i = 0; i < M.n_initial; ++i) {
            if (O.init_bins) O.init_bins[(int64_t)i * N + s] = (int8_t)(x[i] + 1);
            if (O.init_values) O.init_values[(int64_t)i * N + s] = vals[i];
            if (O.hist_initial) hist_inc(0, i, x[i]);
        }
        if (T <= 0 || (!O.bins && !O.values && !O.hist_transition)) return;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            bin[d] = x[M.dyn_t[d]];
            x[M.dyn_t1[d]] = x[M.dyn_t[d]];
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            gbin[g] = x[M.gated_var[g]];
            val[g] = (float)vals[M.gated_var[g]];
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            if (FAST) {
                const uint32_t* col = node_column(M.dyn[d], M.thr_trans, x);
#pragma unroll
                for (int m = 0; m < SH::RP(d); ++m) thr[d][m] = ldg32(col + m);
            } else {
                // column offset contributed by parents that are neither X(t) nor X(t+1) of a dynamic variable
                const Node& nd = M.dyn[d];
                uint32_t o = nd.off;
                for (int p = 0; p < nd.np; ++p) {
                    bool isdyn = false;
                    for (int e = 0; e < ND; ++e) isdyn = isdyn || nd.par[p] == M.dyn_t[e] || nd.par[p] == M.dyn_t1[e];
                    if (!isdyn) o += nd.stride_rp[p] * (uint32_t)x[nd.par[p]];
                }
                cbase[d] = o;
            }
        }
    }
    // slow branch: coefficients of the dynamic parents (uniform across threads)
    uint32_t ct[ND][ND], c1[ND][ND];
    if (!FAST) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
            for (int e = 0; e < ND; ++e) {
                ct[d][e] = 0;
                c1[d][e] = 0;
                for (int p = 0; p < M.dyn[d].np; ++p) {
                    if (M.dyn[d].par[p] == M.dyn_t[e]) ct[d][e] = M.dyn[d].stride_rp[p];
                    if (M.dyn[d].par[p] == M.dyn_t1[e]) c1[d][e] = M.dyn[d].stride_rp[p];
                }
            }
    }
    // edge-table bases of the gated variables (uniform)
    int ebase[NG];
    {
        int b = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            ebase[g] = b;
            b += M.init[M.gated_var[g]].r;
        }
    }

    const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
    const uint32_t c0 = (uint32_t)sample, c1w = (uint32_t)(sample >> 32);
    const uint32_t w3 = ((uint32_t)attempt << 16) | (P_STEP << 8);
    const uint32_t w3dd = ((uint32_t)attempt << 16) | (P_STEP_DD << 8);
    const int nch16 = (T + 15) >> 4, nch4 = nch16 * 4;

    uint32_t bpack[ND][4];
#pragma unroll
    for (int d = 0; d < ND; ++d) bpack[d][0] = bpack[d][1] = bpack[d][2] = bpack[d][3] = 0;

    for (int grp = 0; grp < nch4; ++grp) {
        // ---- the 4*NW words of seconds e = 4*grp .. 4*grp+3 : NW Philox calls ------------------
        uint32_t W[4 * NW];
#pragma unroll
        for (int c = 0; c < NW; ++c)
            philox4x32_10(c0, c1w, (uint32_t)(grp * NW + c), w3, k0, k1, W[4 * c], W[4 * c + 1], W[4 * c + 2], W[4 * c + 3]);
        float vb[NG][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = 4 * grp + j;
            if (e > 0 && e < T) {
                // ---- resample gates on the pre-transition bins (resample_events.m:23-29) ---------
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t k = W[j * NW + ND + g];
                    if (k < (uint32_t)M.gate_G[g]) {
                        const int v = M.gated_var[g];
                        double x;
                        if (M.edge_off[v] < 0) x = (double)(gbin[g] + 1);
                        else if ((uint32_t)M.zero_bin[v] == gbin[g] + 1) x = 0.0;
                        else {
                            const double u = dmul(dadd((double)k, 0.5), M.gate_inv[g]);
                            const double* ed = S.edges + 2 * (ebase[g] + (int)gbin[g]);
                            x = dadd(ed[0], dmul(ed[1], u));
                        }
                        val[g] = (float)x;
                    }
                }
                // ---- transitions ------------------------------------------------------------------
                uint32_t nb[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) nb[d] = 0;
                if (FAST) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const uint32_t k = W[j * NW + d];
                        uint32_t b = thr[d][SH::RP(d) - 1];
#pragma unroll
                        for (int m = 0; m < SH::R(d) - 1; ++m) b += (k > thr[d][m]) ? 1u : 0u;
                        nb[d] = b;
                    }
                } else {
#pragma unroll
                    for (int od = 0; od < ND; ++od) {
                        const int dsel = M.order_dyn[od];   // uniform: which variable is sampled od-th
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            if (dsel == d) {
                                uint32_t o = cbase[d];
#pragma unroll
                                for (int e2 = 0; e2 < ND; ++e2) o += ct[d][e2] * bin[e2] + c1[d][e2] * nb[e2];
                                const uint32_t* col = M.thr_trans + o;
                                uint32_t t[RPM];
#pragma unroll
                                for (int q = 0; q < SH::RP(d); q += 4) {
#if defined(__CUDA_ARCH__)
                                    const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(col + q));
                                    t[q] = v4.x; t[q + 1] = v4.y; t[q + 2] = v4.z; t[q + 3] = v4.w;
#else
                                    t[q] = col[q]; t[q + 1] = col[q + 1]; t[q + 2] = col[q + 2]; t[q + 3] = col[q + 3];
#endif
                                }
                                const uint32_t k = W[j * NW + d];
                                uint32_t b = t[SH::RP(d) - 1];
#pragma unroll
                                for (int m = 0; m < SH::R(d) - 1; ++m) b += (k > t[m]) ? 1u : 0u;
                                nb[d] = b;
                            }
                        }
                    }
                }
                // ---- map back + change events (dbn_sample.m:82-92) --------------------------------
                bool any = false;
#pragma unroll
                for (int d = 0; d < ND; ++d) any = any || (nb[d] != bin[d]);
                if (any) {
                    uint32_t dd0, dd1, dd2, dd3;
                    philox4x32_10(c0, c1w, (uint32_t)e, w3dd, k0, k1, dd0, dd1, dd2, dd3);
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (nb[d] != bin[d]) {
                            const int g = NG - ND + d;
                            const int v = M.gated_var[g];
                            bin[d] = nb[d];
                            gbin[g] = nb[d];
                            double x;
                            if (M.edge_off[v] < 0) x = (double)(nb[d] + 1);
                            else if ((uint32_t)M.zero_bin[v] == nb[d] + 1) x = 0.0;
                            else {
                                const uint32_t kk = d == 0 ? dd0 : d == 1 ? dd1 : d == 2 ? dd2 : dd3;
                                const double* ed = S.edges + 2 * (ebase[g] + (int)nb[d]);
                                x = dadd(ed[0], dmul(ed[1], u01(kk)));
                            }
                            val[g] = (float)x;
                        }
                    }
                }
            }
            const bool live = e < T;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                bpack[d][grp & 3] |= (live ? bin[d] + 1u : 0u) << (8 * j);
                if (live && e > 0 && O.hist_transition) hist_inc(1, d, (int)bin[d]);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) vb[g][j] = live ? val[g] : 0.0f;
        }
        if (O.values) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float* dst = O.values + (((int64_t)g * nch4 + grp) * N + s) * 4;
#if defined(__CUDA_ARCH__)
                __stcs(reinterpret_cast<float4*>(dst), make_float4(vb[g][0], vb[g][1], vb[g][2], vb[g][3]));
#else
                dst[0] = vb[g][0]; dst[1] = vb[g][1]; dst[2] = vb[g][2]; dst[3] = vb[g][3];
#endif
            }
        }
        if ((grp & 3) == 3) {
            if (O.bins) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    int8_t* dst = O.bins + (((int64_t)d * nch16 + (grp >> 2)) * N + s) * 16;
#if defined(__CUDA_ARCH__)
                    __stcs(reinterpret_cast<uint4*>(dst), make_uint4(bpack[d][0], bpack[d][1], bpack[d][2], bpack[d][3]));
#else
                    for (int q = 0; q < 4; ++q)
                        for (int b = 0; b < 4; ++b) dst[q * 4 + b] = (int8_t)((bpack[d][q] >> (8 * b)) & 0xFF);
#endif
                }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) bpack[d][0] = bpack[d][1] = bpack[d][2] = bpack[d][3] = 0;
        }
    }
}

// ---- shape selection (host) --------------------------------------------------------------------
// Returns the RS code (bins of the dynamic variables, one per byte) if the model satisfies the fast
// kernel's structural requirements, else 0.
inline uint32_t fast_shape_of(const DevModel& M) {
    const int nd = M.n_dyn, ng = M.n_gated;
    if (nd < 1 || nd > 4 || ng < nd) return 0;
    int edges = 0;
    for (int g = 0; g < ng; ++g) {
        if (M.gate_G[g] < 1 || M.gate_G[g] > 0xFFFFFFFFull) return 0;
        edges += M.init[M.gated_var[g]].r;
    }
    if (edges > FAST_MAX_EDGES) return 0;
    for (int d = 0; d < nd; ++d)
        if (M.gated_var[ng - nd + d] != M.dyn_t[d]) return 0;   // gated list must end with the dynamic variables
    uint32_t rs = 0;
    for (int d = 0; d < nd; ++d) rs |= (uint32_t)M.dyn[d].r << (8 * d);
    return rs;
}

// Shapes compiled into libemb200.so: X(RS, NG, FAST).  Anything else runs on k_tracks_generic.
//   0x070705 = bins (5,7,7): all 7-variable uncor models, uncor v1, littoral, glider/paraglider/fai/paramotor/
//              skydiving/blimp;  0x070905: dueregard;  0x050707: haa;  0x09090909: cor_v1 / littoral_cor;  0x07: balloons
#define EMB_FAST_SHAPES(X)                                                                         \
    X(0x070705u, 3, true) X(0x070705u, 4, true) X(0x070705u, 5, true)                              \
    X(0x070705u, 3, false) X(0x070705u, 4, false) X(0x070705u, 5, false)                           \
    X(0x070905u, 5, true) X(0x050707u, 7, true) X(0x09090909u, 4, false) X(0x09090909u, 4, true)   \
    X(0x07u, 1, true)

}  // namespace emb
