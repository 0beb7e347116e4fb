// emb_initial.cuh -- register-resident sampler of the initial network (bn_sample.m:39-58 over
// num_samples, plus dbn_hierarchical_sample.m:25-31 de-discretisation) for models whose topological
// order is the identity (every shipped model except cor_v1 / paramotor / skydiving) when no rejection
// test is requested.  Same keyed stream as sample_initial (emb_device.cuh): word k_i selects variable i and
// k_i * DD_MULT + k_(i+1) de-discretises it (stream spec v5, INIT).
//
// One thread samples SPT = 4 consecutive samples so that
//   * the gathers of the four samples (L1/L2-resident threshold columns) are in flight together,
//   * bins leave as one 4-byte store per variable (a warp writes 128 contiguous bytes) and values as
//     two 16-byte stores per variable,
// and the column offset of variable i is accumulated with compile-time indices
//   off_i = off0_i + sum_{p<i} S[p][i] * x_p        (S = asub2ind strides x rp, 0 when p is not a parent)
// so the state never leaves registers.
#pragma once
#include "emb_device.cuh"

namespace emb {

constexpr int INIT_SPT = 4;

// the 23 value bits of (k, partner) as a float in [1,2)  (same construction as emb_fast.cuh: dd_fraction)
EMB_HD float dd_fraction32(uint32_t k, uint32_t kn) {
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

// upper-triangular stride matrix of the identity-order initial network (kernel parameter, constant bank)
struct InitStrides {
    uint32_t S[MAXV][MAXV];
    uint32_t off[MAXV];      // first word of variable i's columns in the table the kernel reads
    uint32_t src_off[MAXV + 1];   // ... in DevModel::thr_init (source of the shared-memory copy); [n] = table length
    uint32_t rp[MAXV];       // padded column length in DevModel::thr_init
    uint32_t pad;            // extra words between consecutive columns of the shared-memory copy (0: plain copy)
    uint32_t words;          // length of the table the kernel reads
};

inline bool initial_fast_ok(const DevModel& M, const SampleParams& P) {
    if (P.reject_mode != 0 || P.start_ps) return false;
    for (int i = 0; i < M.n_initial; ++i)
        if (M.order_initial[i] != i) return false;
    return true;
}

// pad = 0: the table as it lies in global memory.  pad = 4: the shared-memory copy leaves four words between consecutive
// columns -- columns of 8 words start in bank groups {0,2,4,6} only (16-byte accesses: 8 groups of 4 banks), so a warp's 32
// column gathers pile up 8 deep; with a 12-word pitch they spread over all 8 groups and an LDS.128 costs its minimum of 4
// wavefronts (ncu: 9.5 wavefronts per LDS.128 before, the kernel at 67 % of the shared-memory bandwidth).
inline void fill_init_strides(const DevModel& M, int table_words, uint32_t pad, InitStrides& st) {
    for (int p = 0; p < MAXV; ++p)
        for (int i = 0; i < MAXV; ++i) st.S[p][i] = 0;
    st.pad = pad;
    uint32_t o = 0;
    for (int i = 0; i < M.n_initial; ++i) {
        const uint32_t rp = (uint32_t)M.init[i].rp, pitch = rp + pad;
        const uint32_t end = i + 1 < M.n_initial ? M.init[i + 1].off : (uint32_t)table_words;
        const uint32_t ncol = (end - M.init[i].off) / rp;
        st.src_off[i] = M.init[i].off;
        st.rp[i] = rp;
        st.off[i] = pad ? o : M.init[i].off;
        o += ncol * pitch;
        for (int q = 0; q < M.init[i].np; ++q) st.S[M.init[i].par[q]][i] = M.init[i].stride_rp[q] / rp * pitch;
    }
    st.src_off[M.n_initial] = (uint32_t)table_words;
    st.words = pad ? o : (uint32_t)table_words;
}

// bins of INIT_SPT samples in the same variable: the column gathers of the four samples are issued together, 16 bytes at a
// time; thresholds are counted on the borrow chain (sub_gt: 1.5 instructions per threshold) downwards from `lead`.
// Columns of up to 8 slots (every shipped initial variable except the 36-bin terminal ones) are straight-line code.
#ifndef EMB_INIT_COUNT
#define EMB_INIT_COUNT 1
#endif
EMB_HD void count_chunk4(const uint32_t* table, const uint32_t (&off)[INIT_SPT], int q, bool last, const uint32_t (&k)[INIT_SPT],
                         uint32_t (&neg)[INIT_SPT], uint32_t (&lead)[INIT_SPT]) {
    uint32_t a[INIT_SPT], b[INIT_SPT], c[INIT_SPT], d[INIT_SPT];
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) {
        const uint32_t* col = table + off[j] + q;
#if defined(__CUDA_ARCH__)
        const uint4 v = *reinterpret_cast<const uint4*>(col);   // LDS.128 (staged table) or LDG.128
        a[j] = v.x; b[j] = v.y; c[j] = v.z; d[j] = v.w;
#else
        a[j] = col[0]; b[j] = col[1]; c[j] = col[2]; d[j] = col[3];
#endif
    }
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) {
#if EMB_INIT_COUNT == 0
        neg[j] -= (uint32_t)(k[j] > a[j]) + (uint32_t)(k[j] > b[j]) + (uint32_t)(k[j] > c[j]);
        if (last) lead[j] = d[j];
        else neg[j] -= (uint32_t)(k[j] > d[j]);
#else
        neg[j] = sub_gt(sub_gt(sub_gt(neg[j], k[j], a[j]), k[j], b[j]), k[j], c[j]);
        if (last) lead[j] = d[j];                               // last slot of the column holds `lead`
        else neg[j] = sub_gt(neg[j], k[j], d[j]);
#endif
    }
}
EMB_HD void count_gt4(const uint32_t* table, const uint32_t (&off)[INIT_SPT], int rp, const uint32_t (&k)[INIT_SPT],
                      uint32_t (&bin)[INIT_SPT]) {
    uint32_t neg[INIT_SPT], lead[INIT_SPT];
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) neg[j] = lead[j] = 0;
#if EMB_INIT_COUNT == 2
    if (rp == 4) {
        count_chunk4(table, off, 0, true, k, neg, lead);
    } else if (rp == 8) {
        count_chunk4(table, off, 0, false, k, neg, lead);
        count_chunk4(table, off, 4, true, k, neg, lead);
    } else
#endif
    {
#pragma unroll 1
        for (int q = 0; q < rp; q += 4) count_chunk4(table, off, q, q + 4 >= rp, k, neg, lead);
    }
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) bin[j] = lead[j] - neg[j];
}

// the index-only part of the NW = max(NV, 2) Philox calls of the INIT stream (philox_call: the same for every thread of a launch
// whose sample groups share their high word); kernel-parameter / constant-bank resident
struct InitCalls {
    uint4 e[MAXV + 1];
    uint32_t c0;          // (first_sample >> 2) >> 32 the entries were computed for
};
inline void fill_init_calls(const SampleParams& P, int nv, InitCalls& C) {
    C.c0 = (uint32_t)((P.first_sample >> 2) >> 32);
    for (int i = 0; i < (nv > 1 ? nv : 2); ++i) C.e[i] = philox_call(C.c0, P_INIT << 8, (uint32_t)i, P.rk);
}

EMB_HD uint32_t pick4(const uint32_t (&a)[4], uint32_t l) { return l == 0 ? a[0] : l == 1 ? a[1] : l == 2 ? a[2] : a[3]; }

// `table` is the threshold table of the initial network: DevModel::thr_init, or its copy in shared memory.
// samples s0 .. s0+3 (those < P.n); bins [NV][n] int8, values [NV][n] VT = double or float (nullable), attempts [n] (nullable)
// Stream spec v5, INIT: word k_i of a sample is lane (sample & 3) of the call (sample >> 2, index i), so the four samples of a
// thread take their words of variable i from ONE call when first_sample is a multiple of four (two calls otherwise), and the
// first three rounds of every call are shared (philox_track per sample group, InitCalls per variable): NV x 14 IMAD.WIDE per
// four samples instead of the 4 x 3 x 20 of one-call-per-four-words-per-sample.
// ALIGNED: first_sample is a multiple of four (a property of the launch), so one call per variable serves the four samples
template <int NV, bool VALUES, class VT, bool ALIGNED>
EMB_HD void initial_fast4(const DevModel& M, const SampleParams& P, const InitStrides& ST, const InitCalls& IC,
                          const uint32_t* table, const float* ent /* fp32 entries of the initial variables: dd32 + 4 * ddi_off[0], or its shared copy */,
                          int64_t s0, int8_t* bins, VT* values, uint16_t* attempts) {
    constexpr int NW = NV > 1 ? NV : 2;
    const int64_t N = P.n;
    const uint64_t sample0 = P.first_sample + (uint64_t)s0;
    const uint32_t sh = ALIGNED ? 0u : (uint32_t)(sample0 & 3u);   // uniform: s0 is a multiple of four
    const uint64_t grp = sample0 >> 2;
    uint32_t W[INIT_SPT][NW];
    {
        const uint32_t c0 = (uint32_t)(grp >> 32), c2 = P_INIT << 8;
        const PhiloxTrack t0 = philox_track(c0, (uint32_t)grp, c2, P.rk);
        const uint64_t grp1 = grp + 1;
        const uint32_t c0b = (uint32_t)(grp1 >> 32);
        PhiloxTrack t1{};
        if (sh) t1 = philox_track(c0b, (uint32_t)grp1, c2, P.rk);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint32_t a[4], b[4] = {0u, 0u, 0u, 0u};
            const uint4 e0 = c0 == IC.c0 ? IC.e[i] : philox_call(c0, c2, (uint32_t)i, P.rk);
            philox_finish(t0, e0, P.rk, a[0], a[1], a[2], a[3]);
            if (sh) {
                const uint4 e1 = c0b == IC.c0 ? IC.e[i] : philox_call(c0b, c2, (uint32_t)i, P.rk);
                philox_finish(t1, e1, P.rk, b[0], b[1], b[2], b[3]);
#pragma unroll
                for (int j = 0; j < INIT_SPT; ++j) W[j][i] = (uint32_t)j + sh < 4u ? pick4(a, (uint32_t)j + sh) : pick4(b, (uint32_t)j + sh - 4u);
            } else {
#pragma unroll
                for (int j = 0; j < INIT_SPT; ++j) W[j][i] = a[j];
            }
        }
    }
    uint32_t x[INIT_SPT][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const Node& nd = M.init[i];
        if (P.start[i]) {                                                         // bn_sample.m:49 (uniform branch)
#pragma unroll
            for (int j = 0; j < INIT_SPT; ++j) x[j][i] = (uint32_t)P.start[i] - 1u;
        } else {
            uint32_t off[INIT_SPT], k[INIT_SPT], b[INIT_SPT];
#pragma unroll
            for (int j = 0; j < INIT_SPT; ++j) {
                uint32_t o = ST.off[i];
#pragma unroll
                for (int p = 0; p < i; ++p) o += ST.S[p][i] * x[j][p];           // asub2ind.m:13-14
                off[j] = o;
                k[j] = W[j][i];
            }
            count_gt4(table, off, nd.rp, k, b);                                   // select_random.m:17-20 in word space
#pragma unroll
            for (int j = 0; j < INIT_SPT; ++j) x[j][i] = b[j];
        }
    }
    const bool full = s0 + INIT_SPT <= N;
    const bool al4 = full && (N & 3) == 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (bins) {
            int8_t* dst = bins + (int64_t)i * N + s0;
            if (al4) {
                const uint32_t v = (x[0][i] + 1u) | ((x[1][i] + 1u) << 8) | ((x[2][i] + 1u) << 16) | ((x[3][i] + 1u) << 24);
#if defined(__CUDA_ARCH__)
                __stcs(reinterpret_cast<uint32_t*>(dst), v);
#else
                for (int j = 0; j < 4; ++j) dst[j] = (int8_t)((v >> (8 * j)) & 0xFF);
#endif
            } else {
#pragma unroll
                for (int j = 0; j < INIT_SPT; ++j)
                    if (s0 + j < N) dst[j] = (int8_t)(x[j][i] + 1u);
            }
        }
        if (VALUES && values) {
            VT v[INIT_SPT];
#pragma unroll
            for (int j = 0; j < INIT_SPT; ++j) {
                const int b = (int)x[j][i];
                const uint32_t kn = W[j][i + 1 < NW ? i + 1 : 0];
                if (sizeof(VT) == 4) {                                            // fp32 entries {slope, base, s, c} (emb_model.cpp: pack)
                    const float* en = ent + 4 * (M.ddi_off[i] - M.ddi_off[0] + b);
#if defined(__CUDA_ARCH__)
                    const float4 e4 = *reinterpret_cast<const float4*>(en);
                    v[j] = (VT)__fmaf_rn(e4.x, __fmaf_rn(dd_fraction32(W[j][i], kn), e4.z, e4.w), e4.y);
#else
                    v[j] = (VT)__builtin_fmaf(en[0], __builtin_fmaf(dd_fraction32(W[j][i], kn), en[2], en[3]), en[1]);
#endif
                } else {
                    v[j] = (VT)dedisc(M, i, b, u_dd(W[j][i], kn));                // dediscretize.m:22-41
                }
            }
            VT* dst = values + (int64_t)i * N + s0;
#if defined(__CUDA_ARCH__)
            if (al4 && sizeof(VT) == 8) {
                __stcs(reinterpret_cast<double2*>(dst), make_double2((double)v[0], (double)v[1]));
                __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2((double)v[2], (double)v[3]));
            } else if (al4) {
                __stcs(reinterpret_cast<float4*>(dst), make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]));
            } else
#endif
            {
#pragma unroll
                for (int j = 0; j < INIT_SPT; ++j)
                    if (s0 + j < N) dst[j] = v[j];
            }
        }
    }
    if (attempts) {
#pragma unroll
        for (int j = 0; j < INIT_SPT; ++j)
            if (s0 + j < N) attempts[s0 + j] = 1;
    }
}

// Variable counts compiled into libemb200.so: balloons 2, glider family 5, littoral/uncor v1 6,
// 7-variable uncor 7, HAA 9, terminal encounter geometry 15
#define EMB_INIT_SHAPES(X) X(2) X(5) X(6) X(7) X(9) X(15)

}  // namespace emb
