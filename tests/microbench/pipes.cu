// Micro-benchmark of the sm_100a integer pipes that bound the NTT butterfly.
// One 1024-thread CTA per SM; every thread runs 8 independent dependency chains of one instruction kind;
// throughput is reported as thread-instructions per clock per SM from in-kernel %clock64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int KIND> __global__ void __launch_bounds__(1024) k(uint64_t *out, long long *cycles) {
    uint32_t a[8], b = threadIdx.x * 2654435761u + 12345u, c = blockIdx.x + 77u;
    uint64_t w[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = b + i * 1315423911u; w[i] = ((uint64_t)a[i] << 32) | (a[i] ^ c); d[i] = 1.0 + i * 1e-9; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
            if (KIND == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (KIND == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 5) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 31 | c));
            if (KIND == 6) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 31 | c));
            if (KIND == 7) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(1e-12));
            if (KIND == 8) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 31 | c));
            if (KIND == 9) asm volatile("{.reg .pred p; setp.ge.u64 p, %0, %1; @p sub.u64 %0, %0, %1;}" : "+l"(w[i]) : "l"((uint64_t)b << 31 | c));
            // rounding a double to an integer-valued double (FRND): which pipe, and does it overlap with DFMA?
            if (KIND == 10) asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(d[i]));
            // the FP64 butterfly's quotient step two ways, 8 FP64-class instructions per chain step:
            //   11: seven DFMA + one FRND (q = rint(h * pinv));  12: eight DFMA (q = fma(h, pinv, MAGIC) - MAGIC)
            if (KIND == 11) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tcvt.rni.f64.f64 %0, %0;\n\t"
                             "fma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;"
                             : "+d"(d[i]) : "d"(1.0000001), "d"(1e-12));
            }
            if (KIND == 12) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\t"
                             "fma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;\n\tfma.rn.f64 %0, %0, %1, %2;"
                             : "+d"(d[i]) : "d"(1.0000001), "d"(1e-12));
            }
            // the conversions of SmallAcc::reduce
            if (KIND == 13) asm volatile("{.reg .f64 t; cvt.rn.f64.u64 t, %0; cvt.rzi.u64.f64 %0, t;}" : "+l"(w[i]));
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + w[i] + (uint64_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND> void run(const char *name, int sms, uint64_t *out, long long *cyc) {
    k<KIND><<<sms, 1024>>>(out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<sms, 1024>>>(out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double ops = 1024.0 * ITERS * 8;
    printf("{\"kind\": \"%s\", \"thread_inst_per_clk_per_sm\": %.2f, \"cycles\": %.0f, \"ms\": %.4f, \"implied_mhz\": %.0f}\n",
           name, ops / avg, avg, ms, avg / (ms * 1e3));
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)sms * 1024 * 8); cudaMalloc(&cyc, sms * sizeof(long long));
    printf("{\"sms\": %d}\n", sms);
    run<0>("imad.lo.u32", sms, out, cyc);
    run<1>("imad.wide.u32", sms, out, cyc);
    run<2>("imad.hi.u32", sms, out, cyc);
    run<3>("iadd.u32", sms, out, cyc);
    run<4>("lop3.b32", sms, out, cyc);
    run<5>("mul.hi.u64", sms, out, cyc);
    run<6>("mul.lo.u64", sms, out, cyc);
    run<7>("dfma", sms, out, cyc);
    run<8>("add.u64", sms, out, cyc);
    run<9>("cond_sub.u64", sms, out, cyc);
    run<10>("cvt.rni.f64.f64", sms, out, cyc);
    run<11>("7 dfma + 1 cvt.rni per step (count: steps)", sms, out, cyc);
    run<12>("8 dfma per step (count: steps)", sms, out, cyc);
    run<13>("cvt.f64.u64 + cvt.u64.f64 (count: pairs)", sms, out, cyc);
    return 0;
}
