// ubench.cu -- per-SM issue rates of the instructions the sampler kernels are made of (sm_100a).
// Scratch measurement tool (not product): prints warp-instructions / clk / SM for each op mix.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i; }
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (float)a[i];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = (double)a[i];
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) {  // IMAD.WIDE.U32 : 1 per op
                uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;  // + 1 LOP3
            } else if (OP == 1) {  // LOP3 only: 2 per op
                a[i] = (a[i] ^ b[i] ^ seed); b[i] = (b[i] & a[i]) | seed;
            } else if (OP == 2) {  // philox round pair: 2 IMAD.WIDE + 2 LOP3 (a[i],b[i] as c0,c2; c1,c3 folded)
                uint64_t p0 = (uint64_t)a[i] * 0xD2511F53u, p1 = (uint64_t)b[i] * 0xCD9E8D57u;
                a[i] = (uint32_t)(p1 >> 32) ^ (uint32_t)p0 ^ seed; b[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p1 ^ seed;
            } else if (OP == 3) {  // ISETP + predicated add (count compares)
                b[i] += (a[i] > b[i]) ? 1u : 0u; a[i] += (b[i] > seed) ? 1u : 0u;
            } else if (OP == 4) {  // I2FP.F32.U32 + FFMA
                f[i] = fmaf(f[i], 0.5f, (float)a[i]); a[i] += __float_as_uint(f[i]);
            } else if (OP == 5) {  // IMAD.HI.U32
                a[i] = __umulhi(a[i], 0x00800000u + b[i]) + 0x3F800000u;
            } else if (OP == 6) {  // DFMA
                d[i] = fma(d[i], 1.0000001, 0.5);
            } else if (OP == 7) {  // F2F.F32.F64 + F2F.F64.F32
                f[i] = (float)d[i]; d[i] = d[i] + (double)f[i];
            } else if (OP == 8) {  // IMAD (32-bit) only
                a[i] = a[i] * 0x9E3779B1u + b[i];
            } else if (OP == 9) {  // FSEL + ISETP
                f[i] = (a[i] > b[i]) ? f[i] * 1.5f : f[i]; a[i] = a[i] * 0x9E3779B1u + 1;
            } else if (OP == 10) {  // mad.wide.u32 carry-compare: d = k*1 + {~t, acc}
                uint64_t c = ((uint64_t)b[i] << 32) | (uint32_t)(~seed - i);
                uint64_t r; asm volatile("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(a[i]), "l"(c));
                b[i] = (uint32_t)(r >> 32); a[i] = a[i] * 0x9E3779B1u + 1;
            } else if (OP == 11) {  // I2F.F64.U32
                d[i] = d[i] + (double)a[i]; a[i] += 77;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __float_as_uint(f[i]) + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double ops_per_iter8, int sm, double mhz) {
    uint32_t* out; cudaMalloc(&out, sm * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sm * 8, 256>>>(out, 12345u); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<sm * 8, 256>>>(out, 12345u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_iters = (double)sm * 8 * 8 * ITER * 8;     // warps * iters * 8 slots
    double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-44s %8.3f ms  %7.3f slot-iters/clk/SM  (x%.0f nominal instr = %.2f instr/clk/SM)\n", name, ms,
           warp_iters / clk / sm, ops_per_iter8, warp_iters / clk / sm * ops_per_iter8);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount; double mhz = p.clockRate / 1e3;
    printf("%s SMs=%d clock=%.0f MHz (nominal; rates assume this clock)\n", p.name, sm, mhz);
    run<0>("IMAD.WIDE.U32 + LOP3", 2, sm, mhz);
    run<1>("LOP3 x2", 2, sm, mhz);
    run<2>("philox round (2 IMAD.WIDE + 2 LOP3)", 4, sm, mhz);
    run<3>("ISETP + pred add x2", 4, sm, mhz);
    run<4>("I2FP.F32.U32 + FFMA + IADD", 3, sm, mhz);
    run<5>("IMAD.HI.U32 (+IADD)", 2, sm, mhz);
    run<6>("DFMA", 1, sm, mhz);
    run<7>("F2F.F32.F64 + F2F.F64.F32 + DADD", 3, sm, mhz);
    run<8>("IMAD", 1, sm, mhz);
    run<9>("ISETP + FMUL + FSEL + IMAD", 4, sm, mhz);
    run<10>("mad.wide carry-compare + IMAD", 2, sm, mhz);
    run<11>("I2F.F64.U32 + DADD + IADD", 3, sm, mhz);
    return 0;
}
