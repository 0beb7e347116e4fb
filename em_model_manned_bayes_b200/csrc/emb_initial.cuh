// emb_initial.cuh -- register-resident sampler of the initial network (bn_sample.m:39-58 over
// num_samples, plus dbn_hierarchical_sample.m:25-31 de-discretisation) for models whose topological
// order is the identity (every shipped model except cor_v1 / paramotor / skydiving) when no rejection
// test is requested.  Same keyed stream as sample_initial (emb_device.cuh): word i selects variable
// i, word NV+i de-discretises it.
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

// upper-triangular stride matrix of the identity-order initial network (kernel parameter, constant bank)
struct InitStrides {
    uint32_t S[MAXV][MAXV];
};

inline bool initial_fast_ok(const DevModel& M, const SampleParams& P) {
    if (P.reject_mode != 0 || P.start_ps) return false;
    for (int i = 0; i < M.n_initial; ++i)
        if (M.order_initial[i] != i) return false;
    return true;
}

inline void fill_init_strides(const DevModel& M, InitStrides& st) {
    for (int p = 0; p < MAXV; ++p)
        for (int i = 0; i < MAXV; ++i) st.S[p][i] = 0;
    for (int i = 0; i < M.n_initial; ++i)
        for (int q = 0; q < M.init[i].np; ++q) st.S[M.init[i].par[q]][i] = M.init[i].stride_rp[q];
}

// bins of INIT_SPT samples in the same variable: the four column gathers of every 16-byte step are issued
// together (one loop with a uniform trip count instead of four dependent loops)
EMB_HD void count_gt4(const uint32_t* table, const uint32_t (&off)[INIT_SPT], int rp, const uint32_t (&k)[INIT_SPT],
                      uint32_t (&bin)[INIT_SPT]) {
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) bin[j] = 0;
    for (int q = 0; q < rp; q += 4) {
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
        const bool last = q + 4 >= rp;
#pragma unroll
        for (int j = 0; j < INIT_SPT; ++j) {
            bin[j] += (k[j] > a[j]) + (k[j] > b[j]) + (k[j] > c[j]);
            bin[j] += last ? d[j] : (uint32_t)(k[j] > d[j]);   // last slot holds `lead`
        }
    }
}

// `table` is the threshold table of the initial network: DevModel::thr_init, or its copy in shared memory.
// samples s0 .. s0+3 (those < P.n); bins [NV][n] int8, values [NV][n] double (nullable), attempts [n] (nullable)
template <int NV, bool VALUES>
EMB_HD void initial_fast4(const DevModel& M, const SampleParams& P, const InitStrides& ST, const uint32_t* table,
                          int64_t s0, int8_t* bins, double* values, uint16_t* attempts) {
    constexpr int NWORD = VALUES ? 2 * NV : NV;
    constexpr int NCALL = (NWORD + 3) / 4;
    const int64_t N = P.n;
    uint32_t W[INIT_SPT][4 * NCALL];
#pragma unroll
    for (int j = 0; j < INIT_SPT; ++j) {
        const uint64_t sample = P.first_sample + (uint64_t)(s0 + j);
#pragma unroll
        for (int c = 0; c < NCALL; ++c)
            philox4x32_10_rk((uint32_t)(sample >> 32), (uint32_t)sample, P_INIT << 8, (uint32_t)c, P.rk, W[j][4 * c],
                             W[j][4 * c + 1], W[j][4 * c + 2], W[j][4 * c + 3]);
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
                uint32_t o = nd.off;
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
            double v[INIT_SPT];
#pragma unroll
            for (int j = 0; j < INIT_SPT; ++j) {
                const int b = (int)x[j][i];
                double u = 0.5;
                if (needs_uniform(M, i, b)) u = u01(W[j][NV + i]);
                v[j] = dedisc(M, i, b, u);                                        // dediscretize.m:22-41
            }
            double* dst = values + (int64_t)i * N + s0;
#if defined(__CUDA_ARCH__)
            if (al4) {
                __stcs(reinterpret_cast<double2*>(dst), make_double2(v[0], v[1]));
                __stcs(reinterpret_cast<double2*>(dst) + 1, make_double2(v[2], v[3]));
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
