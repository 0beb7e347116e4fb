// emb_terminal.cu -- sm_100a kernels of the terminal trajectory chains (emb_terminal.cuh) and of the first-order
// track integration (emb_integrate.cuh).
//
// Thread = chain.  blockIdx.y is the chain id (aircraft, direction), so a warp holds 32 consecutive
// encounters of the same chain: its stores are 128 contiguous bytes per field per step, and the only
// divergence is between lanes whose intents pick different models and lanes whose chains ended early.
// The ten models' descriptors travel in the kernel parameter block (constant bank); their threshold
// tables (a few MB each) are gathered from L2.
#include <cuda_runtime.h>

#include "emb_launch.h"
#include "emb_integrate.cuh"
#include "emb_terminal.cuh"

namespace emb {
namespace {

#ifndef EMB_TERM_BLOCK
#define EMB_TERM_BLOCK 128
#endif
constexpr int TERM_BLOCK = EMB_TERM_BLOCK;

#ifndef EMB_TERM_MINBLOCKS
#define EMB_TERM_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(TERM_BLOCK, EMB_TERM_MINBLOCKS)
k_terminal_chains(const __grid_constant__ TermParams P, const __grid_constant__ TermOut O) {
    // the cutpoint tables of the (up to three) models this block's chain id can use: every lane searches them at its own
    // index, which shared memory serves and the constant bank would serialise
    __shared__ double cuts[3 * TERM_NCUT * TERM_CUT_MAX];
    __shared__ TermLane lanes[3];
    const int chain = (int)blockIdx.y, ac = chain >> 1, dir = chain & 1;
    constexpr int per_model = TERM_NCUT * TERM_CUT_MAX;
    for (int it = 0; it < (ac ? 3 : 2); ++it) {
        const int mi = (ac ? 4 : 0) + it * 2 + dir;
        for (int j = threadIdx.x; j < per_model; j += blockDim.x) cuts[it * per_model + j] = __ldg(P.cuts + mi * per_model + j);
        if (threadIdx.x == it) term_lane_fill(P.m[mi], lanes[it]);
    }
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n) terminal_chain(P, O, s, chain, cuts, lanes);
}

// first-order track integration (emb_integrate.cuh): thread = track, HBM-bound (12 B read + 12 B written per track-second)
__global__ void __launch_bounds__(256) k_tracks_integrate(const __grid_constant__ IntegrateParams P) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n) integrate_track(P, s);
}

// encounter screening (miss distance, overlap, runway proximity): thread = encounter, HBM-bound reads of the trajectories
__global__ void __launch_bounds__(256) k_terminal_screen(const __grid_constant__ ScreenParams P) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n) screen_encounter(P, s);
}

}  // namespace

int launch_screen(const ScreenParams& P, void* stream) {
    if (P.n <= 0) return 0;
    k_terminal_screen<<<(unsigned)((P.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    g_launch_count.fetch_add(1);
    return (int)cudaGetLastError();
}

int launch_integrate(const IntegrateParams& P, void* stream) {
    if (P.n <= 0) return 0;
    k_tracks_integrate<<<(unsigned)((P.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
    g_launch_count.fetch_add(1);
    return (int)cudaGetLastError();
}

int launch_terminal(const TermParams& P, const TermOut& O, void* stream) {
    if (P.n <= 0) return 0;
    const dim3 grid((unsigned)((P.n + TERM_BLOCK - 1) / TERM_BLOCK), 4, 1);
    k_terminal_chains<<<grid, TERM_BLOCK, 0, (cudaStream_t)stream>>>(P, O);
    g_launch_count.fetch_add(1);
    return (int)cudaGetLastError();
}

}  // namespace emb
