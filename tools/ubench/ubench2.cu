// ubench2.cu -- issue-rate probes for the integer instruction forms used by the track kernel (sm_100a).
// Scratch measurement tool.  Each probe runs 8 independent dependency chains per thread, 16 warps per
// SMSP, and reports warp-instructions per clock per SMSP (1.0 = the issue limit).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t one) {
    uint32_t a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i * 77 + threadIdx.x; c[i] = seed * 5 + i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); }
            if (OP == 1) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
            if (OP == 2) { asm volatile("{.reg .u32 t; add.cc.u32 t, %1, %2; addc.u32 %0, %0, 0;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); b[i] ^= a[i]; }
            if (OP == 3) { asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); b[i] ^= a[i]; }
            if (OP == 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); }
            if (OP == 5) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[i]), "r"(b[i])); a[i] = (uint32_t)p ^ (uint32_t)(p >> 32); }
            if (OP == 6) { asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(a[i]) : "r"(b[i])); }
            if (OP == 7) { asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %2; selp.u32 %0, %1, %0, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); b[i] += a[i]; }
            if (OP == 8) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); }
            if (OP == 9) {  // philox-like round + one compare-count per round: 2 IMAD.WIDE, 2 LOP3, 1 IADD3.CC, 0.5 IADD3.X
                uint64_t p0, p1;
                asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p0) : "r"(a[i]));
                asm volatile("mul.wide.u32 %0, %1, 0xCD9E8D57;" : "=l"(p1) : "r"(b[i]));
                a[i] = (uint32_t)(p1 >> 32) ^ (uint32_t)p0 ^ seed; b[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p1 ^ one;
                asm volatile("{.reg .u32 t; add.cc.u32 t, %1, %2; addc.u32 %0, %0, 0;}" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
            }
            if (OP == 10) {  // philox-like round only
                uint64_t p0, p1;
                asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p0) : "r"(a[i]));
                asm volatile("mul.wide.u32 %0, %1, 0xCD9E8D57;" : "=l"(p1) : "r"(b[i]));
                a[i] = (uint32_t)(p1 >> 32) ^ (uint32_t)p0 ^ seed; b[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p1 ^ one;
            }
            if (OP == 11) {  // philox-like round + 2 extra LOP3 + 1 IADD3 (plain)
                uint64_t p0, p1;
                asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p0) : "r"(a[i]));
                asm volatile("mul.wide.u32 %0, %1, 0xCD9E8D57;" : "=l"(p1) : "r"(b[i]));
                a[i] = (uint32_t)(p1 >> 32) ^ (uint32_t)p0 ^ seed; b[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p1 ^ one;
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));
            }
            if (OP == 12) { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b[i])), "f"(__uint_as_float(c[i]))); a[i] = __float_as_uint(f); }
            if (OP == 13) {  // 1 IMAD.WIDE + 2 FFMA + 2 LOP3
                uint64_t p0; asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p0) : "r"(a[i]));
                a[i] = (uint32_t)(p0 >> 32) ^ (uint32_t)p0 ^ seed;
                float f = __uint_as_float(b[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(0.9999f), "f"(0.25f));
                b[i] = __float_as_uint(f);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sm, double mhz) {
    uint32_t* out; cudaMalloc(&out, sm * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sm * 8, 256>>>(out, 12345u, 1u); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<sm * 8, 256>>>(out, 12345u, 1u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double slot_iters_per_smsp = 16.0 * ITER * 8;   // 16 warps per SMSP
    double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-58s %8.3f ms  %7.2f clk per slot-iteration per SMSP\n", name, ms, clk / slot_iters_per_smsp);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sm = p.multiProcessorCount; double mhz = p.clockRate / 1e3;
    printf("%s SMs=%d clock=%.0f MHz\n", p.name, sm, mhz);
    run<0>("LOP3", sm, mhz);
    run<1>("IADD", sm, mhz);
    run<2>("add.cc+addc (+LOP3)", sm, mhz);
    run<3>("setp + @p add (+LOP3)", sm, mhz);
    run<4>("IMAD", sm, mhz);
    run<5>("IMAD.WIDE + LOP3", sm, mhz);
    run<6>("SHF funnel", sm, mhz);
    run<7>("setp + selp (+IADD)", sm, mhz);
    run<8>("IMAD.HI", sm, mhz);
    run<9>("philox round + add.cc/addc", sm, mhz);
    run<10>("philox round", sm, mhz);
    run<11>("philox round + LOP3 + IADD", sm, mhz);
    run<12>("FFMA", sm, mhz);
    run<13>("IMAD.WIDE + 2 FFMA + 2 LOP3", sm, mhz);
    return 0;
}
