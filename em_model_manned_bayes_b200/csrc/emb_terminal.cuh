// emb_terminal.cuh -- terminal trajectory chains: @CorTerminalModel/createEncounter.m:93-329
// (PropagateTrajectory, CreateStartDistribution, CheckTrajectoryConditions) with the fwd/bck
// concatenation and time sort of :74-84 folded into the output addressing.
//
// One call of terminal_chain = one chain = one (encounter, aircraft, direction).  A chain is a
// sequential walk of at most floor(tmax_s)+1 states; every state costs one dbn_sample(t_max = 2) of the
// aircraft's trajectory DBN with all six initial variables preset (dbn_sample.m:95-166, frozen-parent
// branch: three selects from columns addressed by the re-discretised state), the dynamic-limit
// rejection loop of createEncounter.m:192-243 and the kinematic update of :171-184 / :246-256 in fp64.
//
// Uniforms (stream spec v5, terminal part -- oracle/terminal.py): Philox counter
//   (encounter_hi, encounter_lo, attempt << 16 | purpose << 8 | chain, step ii), chain = 2*aircraft + (0 fwd, 1 bck);
//   purpose TERM_SEL: lane d = row 2 of the rand(2,1) of the d-th dynamic variable (dbn_sample.m:133,144);
//   purpose TERM_DD : lane d = the rand of dediscretize for its event (createEncounter.m:203,208,216).
#pragma once
#include "emb_device.cuh"

namespace emb {

#ifndef EMB_TERM_PREFETCH
#define EMB_TERM_PREFETCH 0
#endif
constexpr uint32_t P_TERM_SEL = 5, P_TERM_DD = 6;
constexpr int TERM_NMODELS = 10;   // own {landing, takeoff} x {fwd, bck}, intruder {landing, takeoff, transit} x {fwd, bck}
constexpr int TERM_FIELDS = 5;     // x_nm, y_nm, z_ft, heading_deg, v_ft_s  (t_s is the slot index)
constexpr double TERM_FT_PER_NM = 6076.1154855643;   // createEncounter.m:172
constexpr double TERM_NM_PER_FT = 1.0 / TERM_FT_PER_NM;   // correctly rounded reciprocal (div_const)
// Cutpoint tables of one trajectory model, searched without branches (term_cell): TERM_NCUT tables of TERM_CUT_MAX slots, the
// cutpoints ascending and padded with +inf.  The chain never needs the distance or the bearing themselves, only their cells:
//   TC_DIST2: norm([x y]) >= cut (createEncounter.m:277 on d_nm)  <=>  x*x + y*y >= min{s : sqrt(s) >= cut}  (sqrt is monotone
//             and correctly rounded, so the threshold in s is exact and the fp64 square root leaves the per-second path);
//   TC_BEAR : wrapTo360(atan2d(y, x)) >= cut  <=>  pseudo-angle(x, y) >= pseudo-angle(cosd cut, sind cut), see pseudo_angle();
//   TC_HDG, TC_ALT, TC_SPD: cutpoints_initial{4..6} as they are (em_read.m:128-136).
constexpr int TERM_CUT_MAX = 64;   // at most 63 cutpoints (64 bins) per variable of a trajectory model
constexpr int TERM_NCUT = 5;
enum { TC_DIST2 = 0, TC_BEAR = 1, TC_HDG = 2, TC_ALT = 3, TC_SPD = 4 };

// what a chain needs from one trajectory model (built on the host from HostModel::dev)
struct TermModel {
    // heading', altitude', speed' (temporal_map rows 0..2): column = off + sum_i stride[i] * bin_i over the six
    // initial variables (asub2ind.m:13-14 strides times the padded column length; 0 for non-parents)
    uint32_t off[3];
    uint32_t rp[3];
    uint32_t stride[3][6];
    const uint32_t* thr;      // word-space threshold table of the transition network
    const double* edges;      // {a, b-a} pairs (HostModel::edges)
    int32_t edge_off[6];      // per initial variable: offset into edges (doubles), -1 = no boundaries
    int32_t r[6];
    int32_t i_dist, i_bear;   // 0-based positions of "distance" and "bearing" (createEncounter.m:112-113)
    int32_t alt_hi;           // discreteValidAlt = 1..alt_hi        (createEncounter.m:120), 0 = empty
    int32_t spd_lo, spd_hi;   // discreteValidV   = spd_lo..spd_hi   (:123-125), lo > hi = empty
    // tests on d_nm = norm([x y]) taken on s = x*x + y*y (TC_DIST2 above):
    double dist_max_sq;       // d_nm > bounds_initial(idx.dist, 2)  <=>  s >= dist_max_sq   (:263, :310)
    double quarter_sq;        // d_nm <= 0.25                        <=>  s <  quarter_sq    (:312)
};

struct TermLimits {           // @CorTerminalModel/getDynamicLimits.m:14-62
    double minVel, maxVel, maxTurn, maxAlt, maxVR;
};

struct TermParams {
    uint64_t seed, first_sample;
    int64_t n;
    double tmax_s;
    int32_t tmax;             // floor(tmax_s): a chain has at most tmax + 1 states
    int32_t max_attempts;
    uint32_t rk[20];          // Philox round keys of `seed` (k0 + i*W0, k1 + i*W1; term_round_keys), constant operands of the rounds
    const double* geo;        // sample_geo fields, row geo_row[k] of a [rows][geo_stride] array
    int64_t geo_stride;
    int32_t geo_row[12];      // own_{intent, distance, bearing, alt, heading, speed}, int_{...}
    TermLimits lim[2];
    const double* cuts;       // [TERM_NMODELS][TERM_NCUT][TERM_CUT_MAX] cutpoint tables (make_term_cuts), same order as m[]
    TermModel m[TERM_NMODELS];// [aircraft 0: (intent-1)*2 + dir | aircraft 1: 4 + (intent-1)*2 + dir], dir 0 fwd / 1 bck
};

struct TermOut {
    float* traj;              // [TERM_FIELDS][2][2*tmax+1][n], slot k <-> t_s = k - tmax, NaN where the aircraft has no state
    int16_t* len;             // [4][n] states per chain (numel(fwd.t_s), numel(bck.t_s))
    int32_t* status;          // |= 1: inner resample loop exhausted max_attempts; |= 2: unknown intent (createEncounter.m:22,37)
};

// ---- MATLAB built-ins as restated by the oracle (oracle/terminal.py) --------------------------------
EMB_HD double fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
// a / c for a compile-time divisor, correctly rounded like the IEEE division it replaces (rc = 1/c correctly rounded): the
// reciprocal product is corrected twice with the exact FMA residual (Markstein: with rc = RN(1/c) and q within one ulp of a/c,
// RN(q + RN(a - q*c)*rc) is the correctly rounded quotient; the first correction makes q faithful, the second applies the
// theorem).  Five fp64 instructions instead of the ~25 of a division; finite operands far from the subnormal range only.
EMB_HD double div_const(double a, double c, double rc) {
    const double q0 = dmul(a, rc);
    const double q1 = fma_rn(fma_rn(-q0, c, a), rc, q0);
    return fma_rn(fma_rn(-q1, c, a), rc, q1);
}

// sind/cosd the way MATLAB evaluates them: the argument is reduced in DEGREES (exactly) to t in [-45, 45] plus a quadrant, and
// only t is converted to radians; the two kernels are the fdlibm polynomials for |a| <= pi/4 (error < 1 ulp).  Multiples of 90
// degrees give exactly 0 and +-1.  ~45 fp64 instructions where sincos() of a radian argument costs several hundred (it cannot
// know the argument is below 2*pi and carries a Payne-Hanek path through local memory).
struct SinCos {
    double s, c;
};
EMB_HD SinCos sincosd_pair(double x) {
    // fmod(x, 360) is x itself for |x| < 360 (every angle this path produces); fmod proper is a long software loop on the GPU
    const double r = ::fabs(x) < 360.0 ? x : ::fmod(x, 360.0);
    // nearest quadrant, |qn| <= 4 (round to nearest through the 1.5 * 2^52 shift: two exact-by-construction additions)
    const double qn = dadd(dadd(dmul(r, 0.011111111111111112), 6755399441055744.0), -6755399441055744.0);
    const double t = fma_rn(-90.0, qn, r);                      // exact: a multiple of ulp(r) that is smaller than r
    const double a = dmul(t, 0.017453292519943295);             // pi/180
    const double z = dmul(a, a);
    // sin(a) = a + a*z*(S1 + z*(S2 + ... z*S6))
    double ps = fma_rn(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma_rn(z, ps, 2.75573137070700676789e-06);
    ps = fma_rn(z, ps, -1.98412698298579493134e-04);
    ps = fma_rn(z, ps, 8.33333333332248946124e-03);
    ps = fma_rn(z, ps, -1.66666666666666324348e-01);
    const double sn = fma_rn(dmul(a, z), ps, a);
    // cos(a) = w + ((1 - w) - z/2 + z*z*(C1 + z*(C2 + ... z*C6))),  w = 1 - z/2
    double pc = fma_rn(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma_rn(z, pc, -2.75573143513906633035e-07);
    pc = fma_rn(z, pc, 2.48015872894767294178e-05);
    pc = fma_rn(z, pc, -1.38888888888741095749e-03);
    pc = fma_rn(z, pc, 4.16666666666666019037e-02);
    const double hz = dmul(0.5, z), w = dadd(1.0, -hz);
    const double cs = dadd(w, fma_rn(dmul(z, z), pc, dadd(dadd(1.0, -w), -hz)));
    // sin(t + 90 q), cos(t + 90 q); "+ 0.0" turns a -0 of the negated kernel value into the +0 of the reference's table
    const int q = (int)qn & 3;                                  // two's complement: -1 -> 3, -2 -> 2, -3 -> 1
    SinCos o;
    o.s = dadd((q & 1) ? ((q & 2) ? -cs : cs) : ((q & 2) ? -sn : sn), 0.0);
    o.c = dadd((q & 1) ? ((q & 2) ? sn : -sn) : ((q & 2) ? -cs : cs), 0.0);
    return o;
}
EMB_HD void sincosd(double x, double& s, double& c) {
    const SinCos r = sincosd_pair(x);
    s = r.s;
    c = r.c;
}
EMB_HD double atan2d(double y, double x) { return dmul(::atan2(y, x), 57.29577951308232); }   // 180/pi
// wrapTo360(atan2d(y, x)): atan2d lies in [-180, 180], where mod(a, 360) is a + 360 for a < 0 and a otherwise (the
// general formula a - 360*floor(a/360) gives bit-identical results there), so no division is needed
EMB_HD double heading_of(double y, double x) {
    const double a = atan2d(y, x);
    return a < 0.0 ? dadd(a, 360.0) : a;
}
EMB_HD double wrap360(double a) {                       // wrapTo360 of any finite angle, result in [0, 360]
    if (a >= 0.0 && a < 360.0) return a;
    if (a >= 360.0 && a < 720.0) return dadd(a, -360.0);              // exact (Sterbenz), equal to fmod(a, 360)
    const double r = (a < 0.0 && a > -360.0) ? a : ::fmod(a, 360.0);  // fmod(a, 360) = a for |a| < 360
    return r < 0.0 ? dadd(r, 360.0) : r;
}
EMB_HD double pseudo_angle(double x, double y) {        // [0, 4), increasing with wrapTo360(atan2d(y, x)); (0, 0) -> 0 like atan2
    const double d = dadd(::fabs(x), ::fabs(y));
    if (!(d > 0.0)) return 0.0;
    const double q = x / d;
    return y >= 0.0 ? dadd(1.0, -q) : dadd(3.0, q);
}
EMB_HD double round2(double x) {                        // round(x, 2), half away from zero
    const double y = dmul(x, 100.0);
    const double m = ::floor(dadd(::fabs(y), 0.5));
    return div_const(y >= 0.0 ? m : -m, 100.0, 0.01);
}
EMB_HD double norm2(double a, double b) { return ::sqrt(dadd(dmul(a, a), dmul(b, b))); }

// discretize_bayes.m:14-22, 0-based: #{j : v >= cut[j]} over an ascending table of 63 slots padded with +inf (make_term_model).
// No branch, no loop, no data-dependent trip count: six steps of one load at an immediate offset, one compare and one
// predicated pointer bump.
EMB_HD int term_cell(const double* cut, double v) {
    const double* p = cut;
    if (v >= p[31]) p += 32;
    if (v >= p[15]) p += 16;
    if (v >= p[7]) p += 8;
    if (v >= p[3]) p += 4;
    if (v >= p[1]) p += 2;
    if (v >= p[0]) p += 1;
    return (int)(p - cut);
}

// What a chain reads from its model in every state, gathered in one place: in shared memory on the device (one per intent of the
// block's chain id; lanes of a warp differ in intent, and a lane-indexed read of the kernel parameters is an address computation
// plus a replayed constant load per field), a local copy in the host emulation.
struct TermLane {
    uint32_t off[3], rp[3], stride[3][6];
    int32_t edge_off[3];      // heading, altitude, speed
    int32_t alt_hi, spd_lo, spd_hi;
    const uint32_t* thr;
    const double* edges;
    double dist_max_sq, quarter_sq;
};
EMB_HD void term_lane_fill(const TermModel& M, TermLane& C) {
    for (int d = 0; d < 3; ++d) {
        C.off[d] = M.off[d];
        C.rp[d] = M.rp[d];
        for (int i = 0; i < 6; ++i) C.stride[d][i] = M.stride[d][i];
        C.edge_off[d] = M.edge_off[3 + d];
    }
    C.alt_hi = M.alt_hi;
    C.spd_lo = M.spd_lo;
    C.spd_hi = M.spd_hi;
    C.thr = M.thr;
    C.edges = M.edges;
    C.dist_max_sq = M.dist_max_sq;
    C.quarter_sq = M.quarter_sq;
}
EMB_HD double term_dedisc(const TermLane& C, int d, int b, uint32_t k) {   // dediscretize.m:39, two-argument call; d = 0..2
    if (C.edge_off[d] < 0) return (double)(b + 1);
    double a, w;
    ldg_pair(C.edges + C.edge_off[d] + 2 * b, a, w);
    return dadd(a, dmul(w, u01(k)));
}

// select_bin (emb_device.cuh) with the loads of twelve thresholds in flight at a time: the plain loop waits for one 16-byte load
// per trip (nine in a row for a 36-bin heading column, the first stall of the chain kernel).  Every slot is counted, the
// padding (0xFFFFFFFF, never exceeded) and the last one, which holds `lead`; that one is taken out again at the end.
EMB_HD int select_bin3(const uint32_t* col, uint32_t rp, uint32_t k, uint32_t& lead_out) {
    uint32_t neg = 0;             // minus the number of slots below k
    for (uint32_t q = 0; q < rp; q += 12) {
#if defined(__CUDA_ARCH__)
        const uint4 none = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(col + q));
        const uint4 b = q + 4 < rp ? __ldg(reinterpret_cast<const uint4*>(col + q + 4)) : none;
        const uint4 c = q + 8 < rp ? __ldg(reinterpret_cast<const uint4*>(col + q + 8)) : none;
        neg = sub_gt(sub_gt(sub_gt(sub_gt(neg, k, a.x), k, a.y), k, a.z), k, a.w);
        neg = sub_gt(sub_gt(sub_gt(sub_gt(neg, k, b.x), k, b.y), k, b.z), k, b.w);
        neg = sub_gt(sub_gt(sub_gt(sub_gt(neg, k, c.x), k, c.y), k, c.z), k, c.w);
#else
        for (uint32_t m = q; m < q + 12 && m < rp; ++m) neg = sub_gt(neg, k, col[m]);
#endif
    }
    const uint32_t lead = ldg32(col + rp - 1);
    lead_out = lead;
    return (int)sub_gt(lead - neg, k, lead);
}

// The select of one dynamic variable with a one-entry cache in registers.  With the stay prior most seconds keep the bin, and
// the column only changes when a parent's cell does, so a lane remembers for which (column, bin) it last selected and the
// interval of words (T'[bin-1], T'[bin]] that gives that bin again (thresholds ascend from slot `lead` on, pack_column); a word
// inside it needs no memory access at all.  What this buys is L1 tag bandwidth: every lane's column is its own cache line, so a
// 16-byte load of a warp is 32 tag lookups -- 519 per warp and state before, the busiest unit of the SM (0.75 per clock).
struct SelectCache {
    uint32_t key = 0xFFFFFFFFu, lo1 = 0, width = 0;   // key = column offset * 64 + bin (tables < 2^26 words, bins <= 64)
};
EMB_HD int select_cached(const uint32_t* thr, uint32_t co, uint32_t rp, uint32_t b, uint32_t k, SelectCache& c) {
#if !defined(EMB_TERM_NOCACHE)
    if (co * 64u + b == c.key && k - c.lo1 <= c.width) return (int)b;
#endif
    const uint32_t* col = thr + co;
    uint32_t lead;
    const int n = select_bin3(col, rp, k, lead);
#if !defined(EMB_TERM_NOCACHE)
    const uint32_t lo1 = (uint32_t)n > lead ? ldg32(col + n - 1) + 1u : 0u;
    const uint32_t hi = (uint32_t)n + 1u < rp ? ldg32(col + n) : 0xFFFFFFFFu;
    c.key = co * 64u + (uint32_t)n;
    c.lo1 = lo1;
    c.width = hi - lo1;
#endif
    return n;
}
EMB_HD void term_round_keys(uint64_t seed, uint32_t (&rk)[20]) {
    for (int i = 0; i < 10; ++i) {
        rk[2 * i] = (uint32_t)seed + (uint32_t)i * PHILOX_W0;
        rk[2 * i + 1] = (uint32_t)(seed >> 32) + (uint32_t)i * PHILOX_W1;
    }
}

// request the cache lines of a packed column (rp <= 64 words, 16-byte aligned: at most three 128-byte lines)
EMB_HD void prefetch_column(const uint32_t* col, uint32_t rp) {
#if defined(__CUDA_ARCH__) && EMB_TERM_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(col));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(col + rp - 1));
    if (rp > 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(col + (rp >> 1)));
#else
    (void)col;
    (void)rp;
#endif
}

#if defined(__CUDA_ARCH__)
#define EMB_STREAM_F32(p, v) __stcs((p), (v))
#define EMB_FLAG_OR(p, v) atomicOr((p), (v))
#else
#define EMB_STREAM_F32(p, v) (*(p) = (v))
#define EMB_FLAG_OR(p, v) (*(p) |= (v))
#endif

// One chain.  `s` = encounter index within this call, chain = 2*aircraft + direction.
// cuts_sh / lanes_sh: the cutpoint tables ([3][TERM_NCUT][TERM_CUT_MAX]) and per-state constants ([3]) of the (up to three)
// models this chain id can use, one per intent, in shared memory; nullptr in the host emulation (read from TermParams).
//
// The walk is ONE loop whose trip is a (state, attempt) pair: createEncounter.m's `while is_resample` (:192-243) inside its
// `for ii` (:160) would make a warp repeat the select for as long as ANY of its lanes is re-drawing (1.7 passes per state on
// the bench models, the extra ones with three lanes active); flattened, a lane that must re-draw simply does not advance its state
// in this trip while its neighbours move on.  Lanes of a warp then sit at most a few states apart, so their stores touch a few
// neighbouring rows of the [slot][n] output instead of one (L2 merges the sectors before they leave for HBM).
EMB_HD void terminal_chain(const TermParams& P, const TermOut& O, int64_t s, int chain, const double* cuts_sh = nullptr,
                           const TermLane* lanes_sh = nullptr) {
    const int ac = chain >> 1, dir = chain & 1;
    const double dt_s = dir ? -1.0 : 1.0;
    const int64_t N = P.n;
    const int K = P.tmax + 1;                           // maximum number of states
    const int64_t S = 2 * (int64_t)P.tmax + 1;          // slots per aircraft
    const uint64_t sample = P.first_sample + (uint64_t)s;
    const double* g = P.geo + s;
    const int32_t* row = P.geo_row + 6 * ac;
    const int intent = (int)g[(int64_t)row[0] * P.geo_stride];
    const double distance = g[(int64_t)row[1] * P.geo_stride], bearing = g[(int64_t)row[2] * P.geo_stride];
    double z_ft = g[(int64_t)row[3] * P.geo_stride];
    double heading_deg = g[(int64_t)row[4] * P.geo_stride];
    const double v0 = g[(int64_t)row[5] * P.geo_stride];
    const TermLimits& L = P.lim[ac];

    const int64_t fstride = 2 * S * N;                   // between fields
    const int64_t sstep = dir ? -N : N;                  // between consecutive states of this chain
    float* slotp = O.traj ? O.traj + ((int64_t)ac * S + P.tmax) * N + s : nullptr;   // slot of t_s = 0, field 0
    auto put = [&](int f, float v) { EMB_STREAM_F32(slotp + f * fstride, v); };
#if defined(__CUDA_ARCH__)
    const float qnan = __int_as_float(0x7FC00000);
#else
    const float qnan = __builtin_nanf("");
#endif
    const bool bad_intent = intent < 1 || intent > (ac ? 3 : 2);                 // createEncounter.m:14-38
    if (bad_intent && O.status) EMB_FLAG_OR(O.status, 2);
    const int islot = (bad_intent ? 1 : intent) - 1, mi = (ac ? 4 : 0) + islot * 2 + dir;
#if defined(__CUDA_ARCH__)
    const TermLane& C = lanes_sh[islot];
    const double* cuts = cuts_sh + islot * (TERM_NCUT * TERM_CUT_MAX);
#else
    TermLane C_;
    term_lane_fill(P.m[mi], C_);
    const TermLane& C = lanes_sh ? lanes_sh[islot] : C_;
    const double* cuts = cuts_sh ? cuts_sh + islot * (TERM_NCUT * TERM_CUT_MAX) : P.cuts + mi * (TERM_NCUT * TERM_CUT_MAX);
#endif

    double sb, cb, sh, ch;
    sincosd(bearing, sb, cb);
    double x = dmul(distance, cb), y = dmul(distance, sb);                        // :45-46
    sincosd(heading_deg, sh, ch);
    double vx = dadd(dmul(ch, v0), -dmul(sh, 0.0)), vy = dadd(dmul(sh, v0), dmul(ch, 0.0));   // :150
    double t_s = 0.0, z_prev = 0.0;
    bool go = !bad_intent;
    int len = 0;
    // Three quantities the reference re-derives every second are carried instead, because they only change at events:
    //  * curr_hdg = wrapTo360(atan2d(vy, vx)) (:176): (vx, vy) is only ever set to v*(cosd h, sind h) (:150, :228-229) or rotated by
    //    delta (:252-255), so its direction is h resp. the previous direction + delta (equal to the atan2d value to ~1e-14 deg);
    //  * speed = norm(v) (:168, :290): re-taken from (vx, vy) whenever v changed (speed event or rotation), like the reference;
    //  * the cells of heading_deg, z_ft and speed (:278-293): re-discretised when the value changed.
    // a turn toward the desired heading runs at exactly +-maxTurn for all but its last second (:246-256): its sine and cosine
    // are taken once per chain (sind(-x) = -sind(x), cosd(-x) = cosd(x) exactly)
    const SinCos turn_sc = sincosd_pair(L.maxTurn);
    double curr_hdg = (vx == 0.0 && vy == 0.0) ? 0.0 : wrap360(heading_deg);
    double speed = norm2(vx, vy);
    uint32_t b_hdg = 0, b_alt = 0, b_spd = 0;
    if (go) {
        b_hdg = (uint32_t)term_cell(cuts + TC_HDG * TERM_CUT_MAX, heading_deg);
        b_alt = (uint32_t)term_cell(cuts + TC_ALT * TERM_CUT_MAX, z_ft);
        b_spd = (uint32_t)term_cell(cuts + TC_SPD * TERM_CUT_MAX, speed);
    }

    int ii = 1;                       // state being produced (:160)
    uint32_t attempt = 0;             // pass of the resample loop of this state (:192)
    uint32_t co0 = 0, co1 = 0, co2 = 0;   // columns of heading', altitude', speed' for this state (frozen parents, dbn_sample.m:110-135)
    double d_sq = 0.0;
    bool ev_any = false, ev_v = false;
    SelectCache sc0, sc1, sc2;
    // Opening of state ii (:163-187): record it, step the position, take the cells of the new position and the three columns.
    // It runs at the END of the previous trip (and once before the loop): everything it reads is known there, so the columns'
    // cache lines are requested (prefetch into L1) a whole state opening before the select reads them.
    auto open_state = [&]() {
        const bool store = O.traj && !(dir && ii == 1);                              // [fwd, bck(2:end)] (:77)
        ++len;
        double z_rec = z_ft;
        if (ii > 1) {                                                                 // :180-184
            const double diff = dadd(z_ft, -z_prev);
            const double lim = ::fmin(L.maxVR, ::fabs(diff));
            z_rec = dadd(z_prev, diff > 0.0 ? lim : diff < 0.0 ? -lim : dmul(0.0, lim));
        }
        z_prev = z_rec;
        if (store) {
            put(0, (float)x);
            put(1, (float)y);
            put(2, (float)z_rec);
            put(3, (float)curr_hdg);
            put(4, (float)speed);
        }
        x = dadd(x, div_const(dmul(vx, dt_s), TERM_FT_PER_NM, TERM_NM_PER_FT));       // :171-173
        y = dadd(y, div_const(dmul(vy, dt_s), TERM_FT_PER_NM, TERM_NM_PER_FT));
        // CreateStartDistribution (:268-294), 0-based bins; the cells of d_nm = norm([x y]) and of the bearing
        // wrapTo360(atan2d(y, x)) (:277, :293) come from x*x + y*y and the pseudo-angle (TC_DIST2, TC_BEAR)
        d_sq = dadd(dmul(x, x), dmul(y, y));
        uint32_t st[6];
        st[0] = (uint32_t)(intent - 1);
        st[1] = (uint32_t)term_cell(cuts + TC_DIST2 * TERM_CUT_MAX, d_sq);
        st[2] = (uint32_t)term_cell(cuts + TC_BEAR * TERM_CUT_MAX, pseudo_angle(x, y));
        st[3] = b_hdg;
        st[4] = b_alt;
        st[5] = b_spd;
        co0 = C.off[0];
        co1 = C.off[1];
        co2 = C.off[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            co0 += C.stride[0][i] * st[i];
            co1 += C.stride[1][i] * st[i];
            co2 += C.stride[2][i] * st[i];
        }
        prefetch_column(C.thr + co0, C.rp[0]);
        prefetch_column(C.thr + co1, C.rp[1]);
        prefetch_column(C.thr + co2, C.rp[2]);
        ev_any = false;
        ev_v = false;
    };
    if (go) open_state();
    while (ii <= K) {
        if (!go) {                                                                    // the chain has ended: NaN in its slots
            if (O.traj && !(dir && ii == 1)) for (int f = 0; f < TERM_FIELDS; ++f) put(f, qnan);
            ++ii;
            if (slotp) slotp += sstep;
            continue;
        }

        // One pass of `while is_resample` (:192-243) and the turn (:246-256), written without branches: in a warp of 32 chains
        // some lane has an event, a rejected bin, the last second of a turn ... in practically every trip, so every branch body
        // is executed anyway -- one after the other, each with a handful of lanes and its own dependent chain of fp64 latencies.
        // As straight-line code with selects the same instructions interleave.  Values that only change at an event are
        // recomputed from unchanged inputs in the other trips, which reproduces them bit for bit.
        uint32_t w0, w1, w2, w3, d0, d1, d2, d3;
        philox4x32_10_rk((uint32_t)(sample >> 32), (uint32_t)sample, (attempt << 16) | (P_TERM_SEL << 8) | (uint32_t)chain,
                         (uint32_t)ii, P.rk, w0, w1, w2, w3);
        philox4x32_10_rk((uint32_t)(sample >> 32), (uint32_t)sample, (attempt << 16) | (P_TERM_DD << 8) | (uint32_t)chain,
                         (uint32_t)ii, P.rk, d0, d1, d2, d3);
        const int nh = select_cached(C.thr, co0, C.rp[0], b_hdg, w0, sc0);
        const int na = select_cached(C.thr, co1, C.rp[1], b_alt, w1, sc1);
        const int nv = select_cached(C.thr, co2, C.rp[2], b_spd, w2, sc2);
        const bool e_h = nh != (int)b_hdg, e_a = na != (int)b_alt, e_v = nv != (int)b_spd;   // events in variable order 4, 5, 6
        const double h_new = term_dedisc(C, 0, nh, d0), z_new = term_dedisc(C, 1, na, d1);
        double v1 = term_dedisc(C, 2, nv, d2);
        ev_any = ev_any || e_h || e_a || e_v;
        if (e_h) heading_deg = h_new;                                                 // :200-205
        const bool alt_ok = na + 1 <= C.alt_hi;                                       // :206-212
        if (e_a && alt_ok) z_ft = z_new;
        bool redo = e_a && !alt_ok;
        const bool spd_ev = !redo && e_v;                                             // :213-233
        const bool spd_ok = nv + 1 >= C.spd_lo && nv + 1 <= C.spd_hi;
        redo = redo || (spd_ev && !spd_ok);
        if (v1 < L.minVel) v1 = L.minVel;
        if (v1 > L.maxVel) v1 = L.maxVel;
        const SinCos hsc = sincosd_pair(heading_deg);
        if (spd_ev && spd_ok) {
            vx = dadd(dmul(hsc.c, v1), -dmul(hsc.s, 0.0));                            // :228-229
            vy = dadd(dmul(hsc.s, v1), dmul(hsc.c, 0.0));
            ev_v = true;
        }
        if (redo) {
            if (attempt < (uint32_t)P.max_attempts) {
                ++attempt;
                continue;
            }
            if (O.status) EMB_FLAG_OR(O.status, 1);
        }

        // turn to the desired heading at the maximum rate (:246-256); curr_hdg is still the heading at the top of this state
        const double turn1 = round2(dadd(heading_deg, -curr_hdg));
        const double mag = ::fmin(::fabs(turn1), L.maxTurn);
        const double delta = turn1 > 0.0 ? mag : turn1 < 0.0 ? -mag : dmul(mag, 0.0);
        if (ev_v) curr_hdg = wrap360(heading_deg);                                // v was re-pointed along heading_deg (:228-229)
        {
            // a turn runs at exactly +-maxTurn for all but its last second: that sine and cosine were taken once per chain
            const SinCos dsc = sincosd_pair(delta);
            const bool full = mag == L.maxTurn;
            const double sd = full ? (delta < 0.0 ? -turn_sc.s : turn_sc.s) : dsc.s, cd = full ? turn_sc.c : dsc.c;
            const double nvx = dadd(dmul(cd, vx), -dmul(sd, vy)), nvy = dadd(dmul(sd, vx), dmul(cd, vy));
            const double nh2 = wrap360(dadd(curr_hdg, delta));
            if (delta != 0.0) {                                                       // rotation by 0 degrees is the identity
                vx = nvx;
                vy = nvy;
                curr_hdg = nh2;
            }
        }
        if (vx == 0.0 && vy == 0.0) curr_hdg = 0.0;                               // atan2d(0, 0) = 0
        // The next state records norm(v) (:168) and discretises it (:290) from (vx, vy) exactly as the reference does, so a
        // speed that was clamped to minVel/maxVel (:221-226) and sits on a bin edge (the dynamic limits are round numbers) lands
        // in the same cell as in the reference given the same sind/cosd; likewise the cells of heading_deg and z_ft (:278-289).
        speed = norm2(vx, vy);
        b_spd = (uint32_t)term_cell(cuts + TC_SPD * TERM_CUT_MAX, speed);
        b_hdg = (uint32_t)term_cell(cuts + TC_HDG * TERM_CUT_MAX, heading_deg);
        b_alt = (uint32_t)term_cell(cuts + TC_ALT * TERM_CUT_MAX, z_ft);
        t_s = dadd(t_s, dt_s);
        // CheckTrajectoryConditions (:296-329); the tests on d_nm are taken on d_sq (TermModel::dist_max_sq, quarter_sq)
        const bool violate = ::fabs(t_s) > P.tmax_s || d_sq >= C.dist_max_sq || ((intent == 1 || intent == 2) && d_sq < C.quarter_sq) ||
                             (ac == 0 && y > 0.25);
        go = !violate;
        ++ii;
        attempt = 0;
        if (slotp) slotp += sstep;
        if (go && ii <= K) open_state();
    }
    if (O.len) O.len[(int64_t)chain * N + s] = (int16_t)len;
}

// ---- encounter screening on the merged trajectories -------------------------------------------------
// getGeneratedMissDistance (@CorTerminalModel/CorTerminalModel.m:117-133), the overlap length of track.m:88 and
// CheckRunwayProximity (CorTerminalModel.m:187-210) for one encounter, read from the [field][aircraft][slot][n] output.
struct ScreenParams {
    int64_t n;
    int32_t tmax;
    double thres_dist_ft, thres_altlow_ft;     // track.m thresDist_ft / thresAltLow_ft
    const float* traj;
    const int16_t* len;
    double* hmd_ft;                            // [n] nullable
    double* vmd_ft;                            // [n]
    int16_t* tcpa;                             // [3][n]: tcpa_s, tcpa_index_own, tcpa_index_int (1-based)
    int16_t* enc_time_s;                       // [n] numel(intersect(t_s1, t_s2))
    uint8_t* runway;                           // [n] bit0 is_close1, bit1 is_low1, bit2 is_close2, bit3 is_low2
};

EMB_HD void screen_encounter(const ScreenParams& P, int64_t s) {
    const int64_t N = P.n, S = 2 * (int64_t)P.tmax + 1, fs = 2 * S * N;
    int lo[2], hi[2];
    for (int a = 0; a < 2; ++a) {
        lo[a] = P.tmax - ((int)P.len[(int64_t)(2 * a + 1) * N + s] - 1);
        hi[a] = P.tmax + (int)P.len[(int64_t)(2 * a) * N + s] - 1;
    }
    const float* t0 = P.traj + s;                 // aircraft 0, field 0
    const float* t1 = P.traj + S * N + s;         // aircraft 1, field 0
    // getGeneratedMissDistance: common times, first minimum of the horizontal distance
    const int clo = lo[0] > lo[1] ? lo[0] : lo[1], chi = hi[0] < hi[1] ? hi[0] : hi[1];
    double best = 0.0, vmd = 0.0;
    int kbest = -1;
    for (int k = clo; k <= chi; ++k) {
        const int64_t o = (int64_t)k * N;
        const double dx = dadd((double)t0[o], -(double)t1[o]);                      // :120
        const double dy = dadd((double)t0[fs + o], -(double)t1[fs + o]);            // :121
        const double d = dmul(::sqrt(dadd(dmul(dx, dx), dmul(dy, dy))), TERM_FT_PER_NM);   // :122
        if (kbest < 0 || d < best) {                                                // :125 min() returns the first minimum
            best = d;
            kbest = k;
            vmd = dadd((double)t1[2 * fs + o], -(double)t0[2 * fs + o]);            // :123, :126
        }
    }
    if (P.hmd_ft) P.hmd_ft[s] = kbest < 0 ? 0.0 : best;
    if (P.vmd_ft) P.vmd_ft[s] = vmd;
    if (P.tcpa) {
        P.tcpa[s] = (int16_t)(kbest < 0 ? 0 : kbest - P.tmax);                      // :127
        P.tcpa[N + s] = (int16_t)(kbest < 0 ? 0 : kbest - lo[0] + 1);               // :128
        P.tcpa[2 * N + s] = (int16_t)(kbest < 0 ? 0 : kbest - lo[1] + 1);           // :129
    }
    if (P.enc_time_s) P.enc_time_s[s] = (int16_t)(chi >= clo ? chi - clo + 1 : 0);  // track.m:88
    if (P.runway) {                                                                 // CorTerminalModel.m:187-210
        uint32_t bits = 0;
        for (int a = 0; a < 2; ++a) {
            const float* t = a ? t1 : t0;
            bool close = false, low = false;
            for (int k = lo[a]; k <= hi[a]; ++k) {
                const int64_t o = (int64_t)k * N;
                const double x = (double)t[o], y = (double)t[fs + o];
                const double d_ft = dmul(::hypot(x, y), 1.68781);                   // :196-197 (the reference's own constant)
                if (d_ft <= P.thres_dist_ft) {
                    close = true;
                    low = low || (double)t[2 * fs + o] <= P.thres_altlow_ft;        // :203-205
                }
            }
            bits |= (close ? 1u : 0u) << (2 * a) | (low ? 1u : 0u) << (2 * a + 1);
        }
        P.runway[s] = (uint8_t)bits;
    }
}

}  // namespace emb
