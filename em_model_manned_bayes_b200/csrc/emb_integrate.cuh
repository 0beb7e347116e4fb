// emb_integrate.cuh -- first-order track integration of sample2track.m:188-244: the per-track
// Euler loop (:199-218) and the CFIT / speed rejection (:234-244), run on the dense compact output of
// emb_sample_tracks while it is still in HBM.
//
// One call of integrate_track = one track: reads the fp32 tiles of the three dynamic variables
// ([ceil(T/4)][ceil(n/128)][var][128][4]: a warp reads 512 contiguous bytes per variable per tile), accumulates in fp64 like the
// reference, writes T+1 points [3][T+1][n] fp32 (a warp writes 128 contiguous bytes per field per second).
#pragma once
#include "emb_terminal.cuh"   // sincosd

namespace emb {

struct IntegrateParams {
    int64_t n;
    int32_t T;
    int32_t i_alt, i_speed;        // 0-based initial variables: altitude layer value 'L', airspeed 'v'
    int32_t g_acc, g_vr, g_turn;   // ordinals of \dot v, \dot h, \dot\psi among the time-varying variables (tile index)
    int32_t n_tv;                  // number of time-varying variables (variables per tile)
    double ur_speed, ur_vertrate, ur_heading;   // sample2track.m:108-125
    double min_speed, max_speed;   // boundaries{v}([1 end]) * ur_speed (:98-99, :140-141)
    const double* init_values;     // [n_initial][n]
    const float* values;           // [ceil(T/4)][ceil(n/128)][n_tv][128][4]
    float* xyz;                    // [3][T+1][n]   x_ft, y_ft, z_ft at time_s = 0..T
    uint8_t* is_good;              // [n]  ~is_cfit & ~is_reject_speed (:241)
};

EMB_HD void integrate_track(const IntegrateParams& P, int64_t s) {
    const int64_t N = P.n;
    const int T = P.T, nch4 = (T + 3) >> 2;
    double z = P.init_values[(int64_t)P.i_alt * N + s];                              // :192
    double speed = dmul(P.init_values[(int64_t)P.i_speed * N + s], P.ur_speed);      // :129, :193
    double heading = 0.0, x = 0.0, y = 0.0;                                          // :190-194
    bool cfit = z < 0.0;                                                             // :235-237 any(z_ft < 0)
    bool bad_speed = speed <= P.min_speed || speed >= P.max_speed;                   // :240
    const int64_t fs = (int64_t)(T + 1) * N;
    float* out = P.xyz ? P.xyz + s : nullptr;
    if (out) {
        EMB_STREAM_F32(out, 0.0f);
        EMB_STREAM_F32(out + fs, 0.0f);
        EMB_STREAM_F32(out + 2 * fs, (float)z);
    }
    const int64_t ntile = num_tiles(N), gstep = ntile * ((int64_t)P.n_tv * TRACK_TILE * 4);
    const float* pa = P.values + tile_offset(P.n_tv, ntile, P.g_acc, 0, s);
    const float* pv = P.values + tile_offset(P.n_tv, ntile, P.g_vr, 0, s);
    const float* pt = P.values + tile_offset(P.n_tv, ntile, P.g_turn, 0, s);
    for (int grp = 0; grp < nch4; ++grp) {
        float a4[4], v4[4], t4[4];
#if defined(__CUDA_ARCH__)
        const float4 A = __ldcs(reinterpret_cast<const float4*>(pa)), V = __ldcs(reinterpret_cast<const float4*>(pv)),
                     Tn = __ldcs(reinterpret_cast<const float4*>(pt));
        a4[0] = A.x; a4[1] = A.y; a4[2] = A.z; a4[3] = A.w;
        v4[0] = V.x; v4[1] = V.y; v4[2] = V.z; v4[3] = V.w;
        t4[0] = Tn.x; t4[1] = Tn.y; t4[2] = Tn.z; t4[3] = Tn.w;
#else
        for (int j = 0; j < 4; ++j) { a4[j] = pa[j]; v4[j] = pv[j]; t4[j] = pt[j]; }
#endif
        pa += gstep; pv += gstep; pt += gstep;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * grp + j;               // updates of row pInd = c+1, i.e. dense column c (:203-205)
            if (c >= T) break;
            double sh, ch;
            sincosd(heading, sh, ch);
            const double nx = dadd(x, dmul(speed, ch)), ny = dadd(y, dmul(speed, sh));   // :212-213 (old speed, old heading)
            z = dadd(z, dmul((double)v4[j], P.ur_vertrate));                              // :208 (+ unit conversion :134)
            speed = dadd(speed, dmul((double)a4[j], P.ur_speed));                         // :209 (:135)
            heading = dadd(heading, dmul((double)t4[j], P.ur_heading));                   // :210 (:136)
            x = nx;
            y = ny;
            cfit = cfit || z < 0.0;
            bad_speed = bad_speed || speed <= P.min_speed || speed >= P.max_speed;
            if (out) {
                float* o = out + (int64_t)(c + 1) * N;
                EMB_STREAM_F32(o, (float)x);
                EMB_STREAM_F32(o + fs, (float)y);
                EMB_STREAM_F32(o + 2 * fs, (float)z);
            }
        }
    }
    if (P.is_good) P.is_good[s] = (uint8_t)(!cfit && !bad_speed);
}

}  // namespace emb
