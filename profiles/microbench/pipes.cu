// Pipe-throughput microbenchmark for sm_100a: FFMA vs packed FFMA2 (fma.rn.f32x2), IMAD.WIDE, LOP3, MUFU, and mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, uint32_t seed) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float m = 1.0001f + seed * 1e-9f, c = 1e-6f;
    unsigned long long p0, p1, p2, p3;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    unsigned long long mm, cc;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    uint32_t x0 = threadIdx.x + seed, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3;
    float h[16];
#pragma unroll
    for (int i = 0; i < 16; i++) h[i] = a0 + i;
#pragma unroll 4
    for (int i = 0; i < ITERS; i++) {
        if (MODE == 0) {          // 8 FFMA
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        } else if (MODE == 1) {   // 4 FFMA2 (= 8 FMAs)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
        } else if (MODE == 2) {   // 4 IMAD.WIDE (philox-like)
            unsigned long long q0 = (unsigned long long)x0 * 0xD2511F53u, q1 = (unsigned long long)x1 * 0xCD9E8D57u;
            unsigned long long q2 = (unsigned long long)x2 * 0xD2511F53u, q3 = (unsigned long long)x3 * 0xCD9E8D57u;
            x0 = (uint32_t)(q0 >> 32) ^ (uint32_t)q1; x1 = (uint32_t)(q1 >> 32) ^ (uint32_t)q2;
            x2 = (uint32_t)(q2 >> 32) ^ (uint32_t)q3; x3 = (uint32_t)(q3 >> 32) ^ (uint32_t)q0;
        } else if (MODE == 3) {   // 4 FFMA2 + 4 IMAD.WIDE + 4 LOP3 interleaved
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
            unsigned long long q0 = (unsigned long long)x0 * 0xD2511F53u, q1 = (unsigned long long)x1 * 0xCD9E8D57u;
            unsigned long long q2 = (unsigned long long)x2 * 0xD2511F53u, q3 = (unsigned long long)x3 * 0xCD9E8D57u;
            x0 = (uint32_t)(q0 >> 32) ^ (uint32_t)q1; x1 = (uint32_t)(q1 >> 32) ^ (uint32_t)q2;
            x2 = (uint32_t)(q2 >> 32) ^ (uint32_t)q3; x3 = (uint32_t)(q3 >> 32) ^ (uint32_t)q0;
        } else if (MODE == 4) {   // 8 FFMA + 4 IMAD.WIDE + 4 LOP3
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            unsigned long long q0 = (unsigned long long)x0 * 0xD2511F53u, q1 = (unsigned long long)x1 * 0xCD9E8D57u;
            unsigned long long q2 = (unsigned long long)x2 * 0xD2511F53u, q3 = (unsigned long long)x3 * 0xCD9E8D57u;
            x0 = (uint32_t)(q0 >> 32) ^ (uint32_t)q1; x1 = (uint32_t)(q1 >> 32) ^ (uint32_t)q2;
            x2 = (uint32_t)(q2 >> 32) ^ (uint32_t)q3; x3 = (uint32_t)(q3 >> 32) ^ (uint32_t)q0;
        } else if (MODE == 5) {   // 4 MUFU.EX2
            a0 = exp2f(a0) ; a1 = exp2f(a1); a2 = exp2f(a2); a3 = exp2f(a3);
            asm volatile("" : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3));
        } else if (MODE == 6) {   // 8 FADD (alu or fma pipe?)
            a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
            asm volatile("" : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));
        } else if (MODE == 8) {   // 4 x (mul.lo + xor): 32-bit IMAD
            x0 = (x0 * 0xD2511F53u) ^ x1; x1 = (x1 * 0xCD9E8D57u) ^ x2; x2 = (x2 * 0xD2511F53u) ^ x3; x3 = (x3 * 0xCD9E8D57u) ^ x0;
        } else if (MODE == 9) {   // 4 x (mul.hi + xor): IMAD.HI
            x0 = __umulhi(x0, 0xD2511F53u) ^ x1; x1 = __umulhi(x1, 0xCD9E8D57u) ^ x2; x2 = __umulhi(x2, 0xD2511F53u) ^ x3; x3 = __umulhi(x3, 0xCD9E8D57u) ^ x0;
        } else if (MODE == 10) {  // 4 x (mul.wide.u16 + xor)
            unsigned t0, t1, t2, t3;
            asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(t0) : "h"((unsigned short)x0), "h"((unsigned short)0x1F53));
            asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(t1) : "h"((unsigned short)x1), "h"((unsigned short)0x8D57));
            asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(t2) : "h"((unsigned short)x2), "h"((unsigned short)0x1F53));
            asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(t3) : "h"((unsigned short)x3), "h"((unsigned short)0x8D57));
            x0 = t0 ^ x1; x1 = t1 ^ x2; x2 = t2 ^ x3; x3 = t3 ^ x0;
        } else if (MODE == 11) {  // 8 LOP3/IADD only (ALU pipe reference)
            x0 = (x0 ^ x1) + 0x9E3779B9u; x1 = (x1 ^ x2) + 0xBB67AE85u; x2 = (x2 ^ x3) + 0x9E3779B9u; x3 = (x3 ^ x0) + 0xBB67AE85u;
        } else if (MODE == 12 || MODE == 13) {   // 4 independent legacy HMMA.1688.F32.TF32 (mma.sync m16n8k8), MODE 13: + 8 FFMA
            uint32_t b0 = x0 & 0xffffe000u, b1 = x1 & 0xffffe000u;
#define HMMA(d0, d1, d2, d3) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b0), "r"(b1))
            HMMA(h[0], h[1], h[2], h[3]); HMMA(h[4], h[5], h[6], h[7]); HMMA(h[8], h[9], h[10], h[11]); HMMA(h[12], h[13], h[14], h[15]);
            if (MODE == 13) {
                a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
                a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            }
        } else if (MODE == 14 || MODE == 15) {  // 4 DEPENDENT HMMAs (one accumulator): latency
            uint32_t b0 = x0 & 0xffffe000u, b1 = x1 & 0xffffe000u;
            HMMA(h[0], h[1], h[2], h[3]); HMMA(h[0], h[1], h[2], h[3]); HMMA(h[0], h[1], h[2], h[3]); HMMA(h[0], h[1], h[2], h[3]);
        } else if (MODE == 7) {   // 8 FFMA + 8 LOP3 (dual issue across pipes?)
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            x0 = (x0 ^ x1) + 0x9E3779B9u; x1 = (x1 ^ x2) + 0xBB67AE85u; x2 = (x2 ^ x3) + 0x9E3779B9u; x3 = (x3 ^ x0) + 0xBB67AE85u;
        }
    }
    float lo, hi; float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
#pragma unroll
    for (int i = 0; i < 16; i++) s += h[i];
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0 ^ p1 ^ p2 ^ p3));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + lo + hi + (float)(x0 ^ x1 ^ x2 ^ x3);
}
template <int MODE> void run(const char* name, double ops_per_iter, float* out, int threads = 256) {
    const int blocks = threads == 256 ? 148 * 8 : 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, threads>>>(out, r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_iters = 5.0 * blocks * (threads / 32) * ITERS;
    double t = ms * 1e-3;
    printf("%-34s %8.3f ms  %7.2f warp-iters/ns  => %6.2f listed-ops per SM per clk @1.9GHz (x32 lanes)\n", name, ms, warp_iters / t * 1e-9,
           warp_iters * ops_per_iter / t / 148 / 1.9e9);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("8 FFMA", 8, out);
    run<1>("4 FFMA2 (8 fma)", 4, out);
    run<2>("4 IMAD.WIDE + 4 LOP", 8, out);
    run<3>("4 FFMA2 + 4 IMAD.WIDE + 4 LOP", 12, out);
    run<4>("8 FFMA + 4 IMAD.WIDE + 4 LOP", 16, out);
    run<5>("4 MUFU.EX2", 4, out);
    run<6>("8 FADD", 8, out);
    run<7>("8 FFMA + 4 LOP3 + 4 IADD", 16, out);
    run<8>("4 IMAD.lo + 4 LOP", 8, out);
    run<9>("4 IMAD.HI + 4 LOP", 8, out);
    run<10>("4 mul.wide.u16 + 4 LOP", 8, out);
    run<11>("4 LOP + 4 IADD", 8, out);
    run<12>("4 HMMA.1688.TF32 (independent)", 4, out);
    run<13>("4 HMMA.1688.TF32 + 8 FFMA", 12, out);
    run<14>("4 HMMA.1688.TF32 (one chain)", 4, out);
    run<15>("4 HMMA (one chain), 1 warp/SMSP", 4, out, 4 * 32);
    return 0;
}
