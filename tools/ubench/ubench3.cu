// ubench3.cu -- can the idle FP64 pipe take the threshold counting off the alu pipe?  (sm_100a scratch tool)
// One iteration = one Philox4x32-10 call + 16 counting compares  acc += [k > t_m]  done with
//   NC thresholds by the carry method (add.cc / addc: IADD3 + half an IADD3.X, alu pipe),
//   ND thresholds by the exact fp64 method  s = (2^52+k) + (~t - 2^52);  acc = fma_rz(s, 2^-32, acc)
//      (s = k + ~t exactly; s*2^-32 in [0,2) truncates to the carry; acc carries 2^52 so its low word is the count),
//   NP thresholds by DSETP + predicated add.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
__device__ __forceinline__ uint32_t add_gt(uint32_t acc, uint32_t k, uint32_t nt) {
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(acc) : "r"(k), "r"(nt));
    return acc;
}
__device__ __forceinline__ void philox(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)c0 * 0xD2511F53u, p1 = (uint64_t)c2 * 0xCD9E8D57u;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

template <int NC, int ND, int NP>
__global__ void __launch_bounds__(128) k(const uint32_t* __restrict__ thr, uint32_t* out, uint32_t seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t tc[NC > 0 ? NC : 1];
    double td[ND > 0 ? ND : 1], tp[NP > 0 ? NP : 1];
#pragma unroll
    for (int m = 0; m < NC; ++m) tc[m] = ~thr[(tid * 16 + m) & 4095];
#pragma unroll
    for (int m = 0; m < ND; ++m) td[m] = (double)(~thr[(tid * 16 + NC + m) & 4095]) - 4503599627370496.0;
#pragma unroll
    for (int m = 0; m < NP; ++m) tp[m] = __hiloint2double(0x43300000, thr[(tid * 16 + NC + ND + m) & 4095]);
    uint32_t accc = 0, accp = 0;
    double accd = 4503599627370496.0;
    for (int it = 0; it < ITER; ++it) {
        uint32_t c0 = tid, c1 = it, c2 = seed, c3 = 7;
        philox(c0, c1, c2, c3, seed, 99u);
        const uint32_t w[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int m = 0; m < NC; ++m) accc = add_gt(accc, w[m & 3], tc[m]);
#pragma unroll
        for (int m = 0; m < ND; ++m) {
            const double kb = __hiloint2double(0x43300000, w[m & 3]);
            accd = __fma_rz(__dadd_rn(kb, td[m]), 2.3283064365386963e-10, accd);
        }
#pragma unroll
        for (int m = 0; m < NP; ++m) {
            const double kb = __hiloint2double(0x43300000, w[m & 3]);
            asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(accp) : "d"(kb), "d"(tp[m]));
        }
    }
    out[tid] = accc + accp + (uint32_t)__double2loint(accd);
}

template <int NC, int ND, int NP>
void run(int sm, double mhz, const uint32_t* thr) {
    uint32_t* out; cudaMalloc(&out, sm * 16 * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NC, ND, NP><<<sm * 16, 128>>>(thr, out, 12345u); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<NC, ND, NP><<<sm * 16, 128>>>(thr, out, 12345u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // 16 blocks of 4 warps per SM = 16 warps per scheduler in 4 waves of 4 resident (regs permitting)
    const double clk = ms * 1e-3 * mhz * 1e6;
    const double warp_iters_per_sched = 16.0 * 4 / 4 * ITER;
    printf("carry %2d  dfma %2d  dsetp %2d : %7.3f ms  %6.1f cycles per warp-iteration per scheduler\n", NC, ND, NP, ms,
           clk / warp_iters_per_sched);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount; double mhz = p.clockRate / 1e3;
    printf("%s SMs=%d clock=%.0f MHz (nominal)\n", p.name, sm, mhz);
    uint32_t h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 2654435761u * (i + 1);
    uint32_t* thr; cudaMalloc(&thr, sizeof h); cudaMemcpy(thr, h, sizeof h, cudaMemcpyHostToDevice);
    run<0, 0, 0>(sm, mhz, thr);
    run<16, 0, 0>(sm, mhz, thr);
    run<12, 4, 0>(sm, mhz, thr);
    run<8, 8, 0>(sm, mhz, thr);
    run<6, 10, 0>(sm, mhz, thr);
    run<4, 12, 0>(sm, mhz, thr);
    run<0, 16, 0>(sm, mhz, thr);
    run<0, 0, 16>(sm, mhz, thr);
    run<8, 0, 8>(sm, mhz, thr);
    run<4, 6, 6>(sm, mhz, thr);
    run<6, 6, 4>(sm, mhz, thr);
    return 0;
}
