// emb_fast.cuh -- register-resident, branch-free track sampler for the common model shapes.
//
// Same semantics and same keyed stream (spec v5, oracle/philox.py) as track_generic (emb_device.cuh);
// what changes is where the state lives and that nothing in the per-second loop diverges.  Template
// parameters fix the number of bins of every dynamic variable (RS packs up to four of them, one per
// byte, in temporal_map order) and the number of gated variables NG, so that
//   * the frozen inverse-CDF thresholds of the fast branch (dbn_sample.m:110-135) sit in registers,
//   * the nw = NG words of four consecutive seconds (one per gated variable and second) come from exactly nw
//     Philox calls whose outputs are indexed statically (p = e*nw + g, so 4 seconds = nw blocks),
//   * a second costs, per gated variable, one compare (gate), one compare (bin changed), and one
//     predicated fp32 de-discretisation  value = fma(slope, fma(f, s, c), base)  whose operands come
//     from a 16-byte shared-memory entry indexed by the bin (no fp64, no int->float conversion: the
//     23-bit value-word fraction is placed in the mantissa of f in [1,2)),
//   * bins and values leave as 4-second tiles: one 4-byte store per dynamic variable and one 16-byte
//     store per gated variable per thread, contiguous across the warp, at compile-time offsets from one
//     running pointer (layout [grp][tile of 128 tracks][var][128][4], emb_device.cuh: tile_offset).
// Requirements checked on the host (fast_shape_of): the gated list ends with the dynamic variables in
// temporal_map order (true for every shipped model); all gate thresholds G < 2^32; no gated bin
// straddles zero without being the zero bin (DevModel::fast32_ok).
#pragma once
#include "emb_device.cuh"
#include "emb_terminal.cuh"   // sincosd (fused integration)

namespace emb {

constexpr int FAST_MAX_EDGES = 160;  // sum of bins over the gated variables (shared-memory entry table)

template <uint32_t RS>
struct DynShape {
    static constexpr int R(int d) { return (int)((RS >> (8 * d)) & 0xFFu); }
    static constexpr int ND = (R(0) > 0) + (R(1) > 0) + (R(2) > 0) + (R(3) > 0);
    static constexpr int RP(int d) { return (R(d) + 3) & ~3; }
    static constexpr int RMAX = R(0) > R(1) ? (R(0) > R(2) ? (R(0) > R(3) ? R(0) : R(3)) : (R(2) > R(3) ? R(2) : R(3)))
                                            : (R(1) > R(2) ? (R(1) > R(3) ? R(1) : R(3)) : (R(2) > R(3) ? R(2) : R(3)));
    static constexpr int RPMAX = (RMAX + 3) & ~3;
};

struct DdEntry {
    float slope, base, s, c;
};

// groups of four seconds whose shared Philox words (philox_call) sit in shared memory at a time
constexpr int UT_GROUPS = 160;

// per-block constants of the fast kernel (shared memory on the device)
struct alignas(16) FastShared {
    DdEntry ent[FAST_MAX_EDGES];  // per (gated ordinal, bin), DevModel::dd32
};
// the index-only part of the Philox calls of UT_GROUPS consecutive groups (NW calls per group)
template <int NW>
struct alignas(16) CallTable {
    uint4 e[UT_GROUPS * NW];
};
template <int NW>
EMB_HD void fill_call_table(CallTable<NW>& U, const SampleParams& P, uint32_t c0, uint32_t c2, int grp0, int ngroups, int tid,
                            int nthreads) {
    for (int q = tid; q < ngroups * NW; q += nthreads) U.e[q] = philox_call(c0, c2, (uint32_t)(grp0 * NW + q), P.rk);
}
// 16-byte load from the call table: on the device through its 32-bit shared-window address (one LDS with an immediate offset;
// a generic pointer makes ptxas rebuild the shared window base in every group)
EMB_HD uint4 lds_call(const uint4* base, uint32_t saddr, int c) {
#if defined(__CUDA_ARCH__)
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr + 16u * (uint32_t)c));
    return v;
#else
    (void)saddr;
    return base[c];
#endif
}
EMB_HD void block_sync() {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#endif
}

// fill FastShared (called by all threads of a block with their index, or by the host with tid=0,nthreads=1)
EMB_HD void fast_fill_shared(const DevModel& M, FastShared& S, int tid, int nthreads) {
    int total = 0;
    for (int g = 0; g < M.n_gated; ++g) total += M.init[M.gated_var[g]].r;
    float* dst = reinterpret_cast<float*>(S.ent);
    for (int q = tid; q < 4 * total; q += nthreads) dst[q] = M.dd32[q];
}

// the 23 fraction bits of the value word of a step word k (partner kn) as a float in [1,2)
// (stream spec v5: u_dd = (f - 1) + 2^-24): one IMAD (hash) + one funnel shift that drops the exponent of 1.0f on top
EMB_HD float dd_fraction(uint32_t k, uint32_t kn) {
    const uint32_t h = k * DD_MULT + kn;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__funnelshift_r(h, 0x7Fu, 9));
#else
    const uint32_t b = (h >> 9) | 0x3F800000u;
    float f;
    __builtin_memcpy(&f, &b, 4);
    return f;
#endif
}

// ---- counting compare:  acc + [k > t]  with nt = ~t ---------------------------------------------------
// k > t  <=>  k + ~t carries out of 32 bits.  Written as add.cc / addc so that SASS is one IADD3 that
// only produces the carry predicate (alu pipe) plus one carry-consuming add that ptxas is free to emit
// as IMAD.X on the fma pipe -- the plain C form (ISETP + increment + predicated move) costs three
// instructions, two of them on the alu pipe that the Philox LOP3s already saturate.
EMB_HD uint32_t add_gt(uint32_t acc, uint32_t k, uint32_t nt) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(acc) : "r"(k), "r"(nt));
    return acc;
#else
    return acc + (k > ~nt ? 1u : 0u);
#endif
}

// The same count on the otherwise idle fp64 pipe: with kb = 2^52 + k and td = 2^52 + t (both exact: the word sits in
// the low half of the mantissa under the high word 0x43300000) the double compare kb > td *is* the integer compare
// k > t.  SASS: one DSETP (fp64 pipe) + one predicated add that ptxas places on whichever integer pipe is free.
EMB_HD double biased_double(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)w);
#else
    return 4503599627370496.0 + (double)w;
#endif
}
EMB_HD uint32_t add_gtd(uint32_t acc, double kb, double td) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(acc) : "d"(kb), "d"(td));
    return acc;
#else
    return acc + (kb > td ? 1u : 0u);
#endif
}

// bit d set: the thresholds of dynamic variable d are compared on the fp64 pipe (fast branch only).  6 = the two 7-bin
// variables of the (5,7,7) shapes; the 5-bin variable stays on the carry chain, which balances the alu, fma and fp64 pipes
// (tools/ubench/ubench3.cu and profiles/r1_ubench_fp64_compare.txt hold the measurements, including an all-fp64
// DADD + DFMA.RZ accumulator that needs no integer instruction but was slower)
#ifndef EMB_SLOW_INLINE_PHILOX   // 1: the per-step branch computes whole Philox calls in the thread (no call table): it is bound by
                                 // the latency of its column gathers, and the extra shared-memory round trip per call cost 5-11 %
                                 // (glider_v1 2.47 -> 2.22 ms, cor_v1 10.0 -> 9.6 ms) where the fast branch gains 8 % from the table
#define EMB_SLOW_INLINE_PHILOX 1
#endif
#ifndef EMB_F64CMP
#define EMB_F64CMP 6
#endif

EMB_HD uint32_t f2u_(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    __builtin_memcpy(&u, &f, 4);
    return u;
#endif
}
EMB_HD float fmaf_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// EV: 0 = dense outputs only, 1 = also count the rows of the event list, 2 = also write them (see emb200.h: emb_event),
//     3 = dense outputs (if present) + the fused Euler loop of sample2track.m:199-244 (TrackOut::x)
constexpr bool ev_list(int EV) { return EV == 1 || EV == 2; }
// ORD (slow branch only): the order in which dbn_sample.m:66-79 samples the dynamic variables, two bits per position
// (order_code below); a compile-time order keeps the per-second code straight-line (no jump table, a third of the code)
template <uint32_t RS, int NG, bool FAST, bool HIST, int EV, uint32_t ORD, class HistInc>
struct FastTrack {
    using SH = DynShape<RS>;
    static constexpr int ND = SH::ND;
    static constexpr int NS = NG - ND;   // gated variables that are not dynamic (their bin never changes)
    static constexpr int NW = NG;        // stream spec v5: one word per (second, gated variable)
    static constexpr int RPM = SH::RPMAX;
    static constexpr bool F64(int d) { return FAST && ((EMB_F64CMP >> d) & 1) != 0; }

    const DevModel& M;
    const SampleParams& P;
    const FastShared& S;
    HistInc hist_inc;
    uint32_t bin[ND];         // current bins of the dynamic variables as entry-table indices: ebase + 0-based bin
    float val[NG];            // current continuous values of the gated variables
    DdEntry sent[NS > 0 ? NS : 1];   // entries of the static gated variables
    uint32_t thr[ND][RPM];    // fast branch: frozen thresholds, complemented (~t); [RP-1] = lead + ebase
                              // slow branch: the column at coff[d] exactly as stored
    double thrd[ND][RPM];     // fast branch, F64(d): 2^52 + t (see add_gtd); the lead stays in thr[d][RP-1]
    uint32_t cbase[ND];       // slow branch: column offset from the parents that never change
    uint32_t coff[ND];        // slow branch: offset of the column held in thr[d][] (0xFFFFFFFF: none yet)
    uint32_t ct[ND][ND], c1[ND][ND];   // slow branch: strides of the dynamic parents (uniform)
    int ebase[NG];            // entry-table bases (uniform)
    uint32_t G[NG];           // gate thresholds (uniform)
    PhiloxTrack pt;           // track-invariant part of the step stream's Philox calls
    uint32_t c0w, c1w;        // EMB_SLOW_INLINE_PHILOX: counter words 0, 1 (sample_hi, sample_lo)
    uint32_t sbin1[NS > 0 ? NS : 1];   // EV: 1-based bins of the static gated variables
    uint32_t ev_last, ev_n;            // EV: second of the last row, rows so far
    long long ev_i;                    // EV == 2: next row of this track
    double ix, iy, iz, ispeed, ihead;  // EV == 3: state of the Euler loop (emb_integrate.cuh: integrate_track, same arithmetic)
    bool icfit, ibad;
    uint32_t* O_words;                 // EV == 2: TrackOut::ev_words / ev_dts
    uint8_t* O_dts;
    EventFormat O_fmt;

    EMB_HD FastTrack(const DevModel& M_, const SampleParams& P_, const FastShared& S_, HistInc h)
        : M(M_), P(P_), S(S_), hist_inc(h) {}

    // one row [dt, var, value] of out_events (dbn_hierarchical_sample.m:9-37 after resample_events.m:11-37):
    // dt = seconds since the previous row, 0 for further rows of the same second
    // gord: 1-based gated ordinal, 0 in the closing row; frac: the 23 value bits (emb_device.cuh: pack_event_word)
    EMB_HD void emit(bool on, uint32_t e, uint32_t gord, uint32_t bin1, uint32_t frac) {
        if (!on) return;
        if (EV == 2) {
            store_event(O_words, O_dts, ev_i, e - ev_last, gord, bin1, frac, O_fmt);
            ++ev_i;
        }
        ev_last = e;
        ++ev_n;
    }

    // one group of four seconds e = 4*grp .. 4*grp+3; CHECK = the group may contain e == 0 or e >= T
    template <bool CHECK>
    EMB_HD void group(int grp, int T, uint32_t (&bout)[ND], float (&vout)[NG][4], const uint4* ut, uint32_t ut_s) {
        uint32_t W[4 * NW];
#pragma unroll
        for (int c = 0; c < NW; ++c) {
#if defined(EMB_MEASURE_3CALLS)
            if (FAST && NS == 1 && ND == 3 && c == 0) continue;
#endif
#if EMB_SLOW_INLINE_PHILOX
            if (!FAST) philox4x32_10_rk(c0w, c1w, P_STEP << 8, (uint32_t)(grp * NW + c), P.rk, W[4 * c], W[4 * c + 1], W[4 * c + 2], W[4 * c + 3]);
            else
#endif
            philox_finish(pt, lds_call(ut, ut_s, c), P.rk, W[4 * c], W[4 * c + 1], W[4 * c + 2], W[4 * c + 3]);
        }
#if defined(EMB_MEASURE_3CALLS)   // measurement only: the static variable's words from the dynamic variables' calls
        if (FAST && NS == 1 && ND == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) W[j] = W[4 + j] ^ ((W[8 + j] << 11) | (W[8 + j] >> 21)) ^ ((W[12 + j] << 22) | (W[12 + j] >> 10));
        }
#endif
#pragma unroll
        for (int d = 0; d < ND; ++d) bout[d] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = 4 * grp + j;
            const bool act = !CHECK || (e > 0 && e < T);
            const bool live = !CHECK || e < T;
            const bool act_gate = !CHECK || (e > 0 && e <= T);   // EV: second T still draws its gates (resample_events.m:23)
            // ---- transitions: new bins from the words of this second --------------------------------
            uint32_t nb[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) nb[d] = bin[d];
            if (FAST) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const uint32_t k = W[4 * (NS + d) + j];
                    uint32_t b = thr[d][SH::RP(d) - 1];
                    if (F64(d)) {
                        const double kb = biased_double(k);
#pragma unroll
                        for (int m = 0; m < SH::R(d) - 1; ++m) b = add_gtd(b, kb, thrd[d][m]);
                    } else {
#pragma unroll
                        for (int m = 0; m < SH::R(d) - 1; ++m) b = add_gt(b, k, thr[d][m]);
                    }
                    nb[d] = act ? b : bin[d];
                }
            } else if (act) {
#pragma unroll
                for (int d = 0; d < ND; ++d) nb[d] = 0;
#pragma unroll
                for (int od = 0; od < ND; ++od) {
                    const int dsel = (int)((ORD >> (2 * od)) & 3u);   // which variable is sampled od-th
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (dsel == d) {
                            uint32_t o = cbase[d];
#pragma unroll
                            for (int e2 = 0; e2 < ND; ++e2) o += ct[d][e2] * bin[e2] + c1[d][e2] * nb[e2];   // cbase absorbs ebase
                            // the column of the previous second stays in registers; a lane gathers only when its parent
                            // configuration changed (a few percent of the track-seconds), so a warp touches a handful of
                            // L1 lines per second instead of 32 per load
                            if (__builtin_expect(o != coff[d], 0)) {
                                coff[d] = o;
                                const uint32_t* col = M.thr_trans + o;
#pragma unroll
                                for (int q = 0; q < SH::RP(d); q += 4) {
#if defined(__CUDA_ARCH__)
                                    const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(col + q));
                                    thr[d][q] = v4.x; thr[d][q + 1] = v4.y; thr[d][q + 2] = v4.z; thr[d][q + 3] = v4.w;
#else
                                    thr[d][q] = col[q]; thr[d][q + 1] = col[q + 1]; thr[d][q + 2] = col[q + 2]; thr[d][q + 3] = col[q + 3];
#endif
                                }
                            }
                            // thresholds as stored (not complemented): the borrow chain counts downwards; last slot = lead
                            const uint32_t k = W[4 * (NS + d) + j];
                            uint32_t c = 0;
#pragma unroll
                            for (int m = 0; m < SH::R(d) - 1; ++m) c = sub_gt(c, k, thr[d][m]);
                            const uint32_t b = thr[d][SH::RP(d) - 1] + (uint32_t)ebase[NS + d] - c;
                            nb[d] = b;
                        }
                    }
                }
            }
            // ---- values: gate (resample_events.m:23-29) and/or bin change (dbn_sample.m:82-92) ------
            // bin[] / nb[] hold entry-table indices (ebase + bin), see track_fast
            bool ev_chg[ND];
            uint32_t ev_frac[ND];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t k = W[4 * g + j];
                const uint32_t kn = W[4 * g + ((j + 1) & 3)];   // partner word of the value (spec v5)
                const int d = g - NS;
                DdEntry en;
                if (g >= NS) en = S.ent[nb[d >= 0 ? d : 0]];
                else en = sent[g < NS ? g : 0];
                const float gp = fmaf_rn(dd_fraction(k, kn), en.s, en.c);
                const float cand = fmaf_rn(en.slope, gp, en.base);
                const bool fired = (k * GATE_MULT) < G[g];
                const bool changed = g >= NS && nb[d >= 0 ? d : 0] != bin[d >= 0 ? d : 0];
                if ((fired || changed) && act) val[g] = cand;
                if (ev_list(EV)) {
                    // gate row: the re-emitted *current* bin (resample_events.m:26-29); when the variable also changes
                    // in this second the row is hidden in the dense output but present in the list.  A row carries the 23
                    // value bits, not the value (pack_event_word): gate and transition row of a second share them.
                    const uint32_t fr = f2u_(dd_fraction(k, kn)) & 0x7FFFFFu;
                    uint32_t b1 = g >= NS ? bin[d >= 0 ? d : 0] - (uint32_t)ebase[g] + 1u : sbin1[g < NS ? g : 0];
                    emit(fired && act_gate, (uint32_t)e, (uint32_t)g + 1u, b1, fr);
                    if (g >= NS) {
                        ev_chg[d >= 0 ? d : 0] = changed && act;
                        ev_frac[d >= 0 ? d : 0] = fr;
                    }
                }
                if (g >= NS) bin[d >= 0 ? d : 0] = nb[d >= 0 ? d : 0];
                vout[g][j] = live ? val[g] : 0.0f;
            }
            if (ev_list(EV)) {   // transition rows follow the gate rows of the same second, variables ascending (dbn_sample.m:84-92)
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    emit(ev_chg[d], (uint32_t)e, (uint32_t)(NS + d) + 1u, bin[d] - (uint32_t)ebase[NS + d] + 1u, ev_frac[d]);
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (CHECK) bout[d] |= (live ? bin[d] - (uint32_t)ebase[NS + d] + 1u : 0u) << (8 * j);
                else bout[d] += bin[d] << (8 * j);     // (1 - ebase) of all four seconds is added once below
                if (HIST && act) hist_inc(1, d, (int)bin[d] - ebase[NS + d]);
            }
        }
        if (!CHECK) {
#pragma unroll
            for (int d = 0; d < ND; ++d) bout[d] += 0x01010101u * (1u - (uint32_t)ebase[NS + d]);
        }
    }

    // EV == 3: the four seconds of a group through the Euler loop of sample2track.m:199-218 (integrate_track statement for
    // statement, on the same fp32 values the dense output holds, so both routes give identical points).  The caller runs it one
    // group BEHIND the sampling (fast_groups): the fp64 chain of four dependent sincosd evaluations and the integer work of the
    // next group's Philox calls and selects are independent and interleave.
    EMB_HD void pick_rates(const float (&vout)[NG][4], const XyzOut& X, float (&a)[4], float (&v)[4], float (&t)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] = 0.0f; v[j] = 0.0f; t[j] = 0.0f;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g == X.g_acc) a[j] = vout[g][j];
                if (g == X.g_vr) v[j] = vout[g][j];
                if (g == X.g_turn) t[j] = vout[g][j];
            }
        }
    }
    EMB_HD void integrate4(int grp, int T, const float (&a)[4], const float (&v)[4], const float (&t)[4], const XyzOut& X, float* out,
                           int64_t N) {
        const int64_t fs = (int64_t)(T + 1) * N;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * grp + j;
            if (c >= T) break;
            double sh, ch;
            sincosd(ihead, sh, ch);
            const double nx = dadd(ix, dmul(ispeed, ch)), ny = dadd(iy, dmul(ispeed, sh));   // :212-213 (old speed, old heading)
            iz = dadd(iz, dmul((double)v[j], X.ur_vertrate));                               // :208 (+ unit conversion :134)
            ispeed = dadd(ispeed, dmul((double)a[j], X.ur_speed));                          // :209 (:135)
            ihead = dadd(ihead, dmul((double)t[j], X.ur_heading));                          // :210 (:136)
            ix = nx;
            iy = ny;
            icfit = icfit || iz < 0.0;
            ibad = ibad || ispeed <= X.min_speed || ispeed >= X.max_speed;
            if (out) {
                float* o = out + (int64_t)(c + 1) * N;
                EMB_STREAM_F32(o, (float)ix);
                EMB_STREAM_F32(o + fs, (float)iy);
                EMB_STREAM_F32(o + 2 * fs, (float)iz);
            }
        }
    }
};

// the loop over the four-second groups of one track; BOTH = both dense outputs are present (no per-group pointer tests)
template <uint32_t RS, int NG, bool FAST, bool HIST, int EV, uint32_t ORD, class HistInc, bool BOTH>
EMB_HD void fast_groups(FastTrack<RS, NG, FAST, HIST, EV, ORD, HistInc>& ft, const SampleParams& P, const TrackOut& O,
                        CallTable<NG>& U, int64_t s, bool valid, int tid, int nthreads, uint32_t c0, uint32_t c2) {
    constexpr int ND = DynShape<RS>::ND;
    const int T = P.T;
    const int64_t N = P.n;
    const int nch4 = (T + 3) >> 2;
    const int nfull = T >> 2;       // groups 1 .. nfull-1 contain only seconds 1 <= e < T
    const int ngrp = ev_list(EV) ? (T + 4) >> 2 : nch4;   // the event list also needs the gates of second T
    const int64_t ntile = num_tiles(N);
    // (a null output only predicates the stores off: its pointer is advanced but never dereferenced)
    const bool wb = BOTH || O.bins != nullptr, wv = BOTH || O.values != nullptr;
    int8_t* pb = O.bins + tile_offset(ND, ntile, 0, 0, s);
    float* pv = O.values + tile_offset(NG, ntile, 0, 0, s);
    const int64_t bstep = ntile * (ND * TRACK_TILE * 4), vstep = ntile * (NG * TRACK_TILE * 4);
    float xa[4], xv[4], xt[4];      // EV == 3: rates of the group the Euler loop has not taken yet
    int xgrp = -1;
    for (int grp0 = 0; grp0 < ngrp; grp0 += UT_GROUPS) {
        const int gcount = ngrp - grp0 < UT_GROUPS ? ngrp - grp0 : UT_GROUPS;
        if (FAST || !EMB_SLOW_INLINE_PHILOX) {
            if (grp0 > 0) block_sync();
            fill_call_table<NG>(U, P, c0, c2, grp0, gcount, tid, nthreads);
            block_sync();
        }
        if (!valid) continue;
#if defined(__CUDA_ARCH__)
        uint32_t ut_s = (uint32_t)__cvta_generic_to_shared(U.e);
#else
        uint32_t ut_s = 0;
#endif
        const uint4* ut = U.e;
        for (int grp = grp0; grp < grp0 + gcount; ++grp, ut += NG, ut_s += 16u * NG) {
            uint32_t bout[ND];
            float vout[NG][4];
            if (grp > 0 && grp < nfull) ft.template group<false>(grp, T, bout, vout, ut, ut_s);
            else ft.template group<true>(grp, T, bout, vout, ut, ut_s);
            if (ev_list(EV) && grp >= nch4) break;
            if (EV == 3) {   // the Euler loop runs one group behind (see integrate4)
                if (xgrp >= 0) ft.integrate4(xgrp, T, xa, xv, xt, O.x, O.x.xyz ? O.x.xyz + s : nullptr, N);
                ft.pick_rates(vout, O.x, xa, xv, xt);
                xgrp = grp;
            }
            if (wv) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float* dst = pv + g * (TRACK_TILE * 4);
#if defined(__CUDA_ARCH__)
                    __stcs(reinterpret_cast<float4*>(dst), make_float4(vout[g][0], vout[g][1], vout[g][2], vout[g][3]));
#else
                    dst[0] = vout[g][0]; dst[1] = vout[g][1]; dst[2] = vout[g][2]; dst[3] = vout[g][3];
#endif
                }
            }
            pv += vstep;
            if (wb) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    int8_t* dst = pb + d * (TRACK_TILE * 4);
#if defined(__CUDA_ARCH__)
                    __stcs(reinterpret_cast<uint32_t*>(dst), bout[d]);
#else
                    for (int b = 0; b < 4; ++b) dst[b] = (int8_t)((bout[d] >> (8 * b)) & 0xFF);
#endif
                }
            }
            pb += bstep;
        }
    }
    if (EV == 3 && valid && xgrp >= 0) ft.integrate4(xgrp, T, xa, xv, xt, O.x, O.x.xyz ? O.x.xyz + s : nullptr, N);
}

// `valid` = this thread owns track s (s < P.n); the other threads of a block only help filling the call table U.
// (tid, nthreads) = the caller's index among the threads that share U (the host emulation calls with 0, 1).
template <uint32_t RS, int NG, bool FAST, bool HIST, int EV, uint32_t ORD, class HistInc>
EMB_HD void track_fast(const DevModel& M, const SampleParams& P, const TrackOut& O, int64_t s, bool valid, const FastShared& S,
                       CallTable<NG>& U, int tid, int nthreads, HistInc hist_inc) {
    using FT = FastTrack<RS, NG, FAST, HIST, EV, ORD, HistInc>;
    using SH = DynShape<RS>;
    constexpr int ND = FT::ND, NS = FT::NS;
    const uint64_t sample = P.first_sample + (uint64_t)s;
    const int T = P.T;
    const int64_t N = P.n;
    FT ft(M, P, S, hist_inc);
    // step stream (spec v5): every track of a launch shares c0 = first_sample >> 32 (the host splits a launch at multiples
    // of 2^32: SampleParams::s_begin) and c2 = P_STEP << 8
    const uint32_t c0 = (uint32_t)((P.first_sample + (uint64_t)P.s_begin) >> 32), c2 = P_STEP << 8;
    const int nch4 = (T + 3) >> 2;
    const int nfull = T >> 2;       // groups 1 .. nfull-1 contain only seconds 1 <= e < T
    const int ngrp = ev_list(EV) ? (T + 4) >> 2 : nch4;   // the event list also needs the gates of second T
    const bool steps = T > 0 && (EV || O.bins || O.values || O.hist_transition);   // uniform

    // ---- initial network (once per track; generic code, cost amortised over T seconds) -----------
    if (valid) {
        uint8_t x[MAXX];
        double vals[MAXV];
        for (int i = 0; i < MAXX; ++i) x[i] = 0;
        int attempt = sample_initial(M, P, sample, x, vals);
        if (attempt < 0) {
            if (O.status) *O.status = attempt == -2 ? 2 : 1;
            attempt = P.max_attempts;
        }
        if (O.attempts) O.attempts[s] = (uint16_t)(attempt + 1);
        for (int i = 0; i < M.n_initial; ++i) {
            if (O.init_bins) O.init_bins[(int64_t)i * (O.init_stride ? O.init_stride : N) + s] = (int8_t)(x[i] + 1);
            if (O.init_values) O.init_values[(int64_t)i * (O.init_stride ? O.init_stride : N) + s] = vals[i];
            if (O.hist_initial) hist_inc(0, i, x[i]);
        }
        if (steps) {
            int b = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                ft.ebase[g] = b;
                b += M.init[M.gated_var[g]].r;
                ft.G[g] = (uint32_t)M.gate_G[g];
                ft.val[g] = (float)vals[M.gated_var[g]];
                if (g < NS) {
                    ft.sent[g] = S.ent[ft.ebase[g] + (int)x[M.gated_var[g]]];
                    ft.sbin1[g] = (uint32_t)x[M.gated_var[g]] + 1u;
                }
            }
            if (EV == 3) {                                                               // sample2track.m:190-194
                ft.iz = vals[O.x.i_alt];
                ft.ispeed = dmul(vals[O.x.i_speed], O.x.ur_speed);
                ft.ihead = 0.0;
                ft.ix = 0.0;
                ft.iy = 0.0;
                ft.icfit = ft.iz < 0.0;
                ft.ibad = ft.ispeed <= O.x.min_speed || ft.ispeed >= O.x.max_speed;
                if (O.x.xyz) {
                    float* o = O.x.xyz + s;
                    const int64_t fs = (int64_t)(T + 1) * N;
                    EMB_STREAM_F32(o, 0.0f);
                    EMB_STREAM_F32(o + fs, 0.0f);
                    EMB_STREAM_F32(o + 2 * fs, (float)ft.iz);
                }
            }
            ft.ev_last = 0;
            ft.ev_n = 0;
            ft.ev_i = EV == 2 ? O.ev_offsets[s] : 0;
            ft.O_words = O.ev_words;
            ft.O_dts = O.ev_dts;
            ft.O_fmt = EventFormat{O.ev_gord_bits, O.ev_dt_bytes};
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                ft.bin[d] = (uint32_t)ft.ebase[NS + d] + x[M.dyn_t[d]];
                x[M.dyn_t1[d]] = x[M.dyn_t[d]];
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (FAST) {
                    const uint32_t* col = node_column(M.dyn[d], M.thr_trans, x);
#pragma unroll
                    for (int m = 0; m < SH::RP(d); ++m) {
                        if (FT::F64(d) && m < SH::RP(d) - 1) ft.thrd[d][m] = biased_double(ldg32(col + m));
                        else ft.thr[d][m] = m < SH::RP(d) - 1 ? ~ldg32(col + m) : ldg32(col + m) + (uint32_t)ft.ebase[NS + d];
                    }
                } else {
                    // column offset contributed by parents that are neither X(t) nor X(t+1) of a dynamic variable
                    const Node& nd = M.dyn[d];
                    uint32_t o = nd.off;
                    for (int p = 0; p < nd.np; ++p) {
                        bool isdyn = false;
                        for (int e = 0; e < ND; ++e) isdyn = isdyn || nd.par[p] == M.dyn_t[e] || nd.par[p] == M.dyn_t1[e];
                        if (!isdyn) o += nd.stride_rp[p] * (uint32_t)x[nd.par[p]];
                    }
                    ft.cbase[d] = o;
                    ft.coff[d] = 0xFFFFFFFFu;
                }
            }
            if (!FAST) {   // strides of the dynamic parents (uniform across threads)
#pragma unroll
                for (int d = 0; d < ND; ++d)
#pragma unroll
                    for (int e = 0; e < ND; ++e) {
                        ft.ct[d][e] = 0;
                        ft.c1[d][e] = 0;
                        for (int p = 0; p < M.dyn[d].np; ++p) {
                            if (M.dyn[d].par[p] == M.dyn_t[e]) ft.ct[d][e] = M.dyn[d].stride_rp[p];
                            if (M.dyn[d].par[p] == M.dyn_t1[e]) ft.c1[d][e] = M.dyn[d].stride_rp[p];
                        }
                        // bin[] / nb[] carry ebase: take it out of the column offset once
                        ft.cbase[d] -= (ft.ct[d][e] + ft.c1[d][e]) * (uint32_t)ft.ebase[NS + e];
                    }
            }
            ft.pt = philox_track(c0, (uint32_t)sample, c2, P.rk);
            ft.c0w = c0;
            ft.c1w = (uint32_t)sample;
        }
    }
    if (!steps) return;

    // [grp][tile][var][128][4]: one running pointer per output, variables at compile-time offsets, one uniform stride per group
    // (the loop is instantiated for "both dense outputs present", the bench case, and for the general case)
    if (O.bins && O.values) fast_groups<RS, NG, FAST, HIST, EV, ORD, HistInc, true>(ft, P, O, U, s, valid, tid, nthreads, c0, c2);
    else fast_groups<RS, NG, FAST, HIST, EV, ORD, HistInc, false>(ft, P, O, U, s, valid, tid, nthreads, c0, c2);
    if (EV == 3 && valid && O.x.is_good) O.x.is_good[s] = (uint8_t)(!ft.icfit && !ft.ibad);   // sample2track.m:234-244
    if (ev_list(EV) && valid) {   // closing row [T - sum(dt), 0, 0] (dbn_hierarchical_sample.m:15-19)
        ft.emit(true, (uint32_t)T, 0u, 0u, 0u);
        if (EV == 1) O.ev_counts[s] = ft.ev_n;
    }
}

// ---- shape selection (host) --------------------------------------------------------------------
// Returns the RS code (bins of the dynamic variables, one per byte) if the model satisfies the fast
// kernel's structural requirements, else 0.
inline uint32_t fast_shape_of(const DevModel& M) {
    const int nd = M.n_dyn, ng = M.n_gated;
    if (nd < 1 || nd > 4 || ng < nd || !M.fast32_ok) return 0;
    int edges = 0;
    for (int g = 0; g < ng; ++g) {
        if (M.gate_G[g] > 0xFFFFFFFFull) return 0;
        edges += M.init[M.gated_var[g]].r;
    }
    if (edges > FAST_MAX_EDGES) return 0;
    for (int d = 0; d < nd; ++d)
        if (M.gated_var[ng - nd + d] != M.dyn_t[d]) return 0;   // gated list must end with the dynamic variables
    uint32_t rs = 0;
    for (int d = 0; d < nd; ++d) rs |= (uint32_t)M.dyn[d].r << (8 * d);
    return rs;
}

// order_transition restricted to the dynamic variables, two bits per position (fast-branch models: 0, never read)
inline uint32_t order_code(const DevModel& M) {
    uint32_t c = 0;
    if (!M.fast)
        for (int od = 0; od < M.n_dyn && od < 4; ++od) c |= ((uint32_t)M.order_dyn[od] & 3u) << (2 * od);
    return c;
}
constexpr uint32_t order_code_of(int a, int b, int c, int d) { return (uint32_t)(a | (b << 2) | (c << 4) | (d << 6)); }

// Shapes compiled into libemb200.so: X(RS, NG, FAST, ORD).  Anything else runs on k_tracks_generic.
//   0x070705 = bins (5,7,7): all 7-variable uncor models, blimp (fast branch); glider/paraglider/paramotor/skydiving/
//              littoral_uncor/uncor_1200code_v1 (slow branch, order dh', dpsi', dv') and fai1/fai5 (dh', dv', dpsi');
//   0x070905: dueregard;  0x050707: haa;  0x09090909: cor_v1 (slow, file order) / littoral_cor;  0x07: balloons
#define EMB_FAST_SHAPES(X)                                                                         \
    X(0x070705u, 3, true, 0u) X(0x070705u, 4, true, 0u) X(0x070705u, 5, true, 0u)                   \
    X(0x070705u, 3, false, order_code_of(1, 2, 0, 0)) X(0x070705u, 3, false, order_code_of(1, 0, 2, 0)) \
    X(0x070705u, 4, false, order_code_of(0, 1, 2, 0))   /* 7-variable uncor models with emb_sample_opts::correct_dbn */ \
    X(0x070905u, 5, true, 0u) X(0x050707u, 7, true, 0u)                                             \
    X(0x09090909u, 4, false, order_code_of(0, 1, 2, 3)) X(0x09090909u, 4, true, 0u)                 \
    X(0x07u, 1, true, 0u)

}  // namespace emb
