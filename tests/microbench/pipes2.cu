// Micro-benchmark (round 2) of what bounds the FP64 butterfly on sm_100a: issue rate of DFMA / DADD / DMUL by operand form,
// the exact FP64 modular butterfly in registers at several occupancies, and FP64 + integer butterflies sharing an SM.
// One CTA per SM; throughput from in-kernel clock64 (per SM clock, independent of DVFS).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define D __device__ __forceinline__

// ---- raw pipe rates -------------------------------------------------------------------------------------------------------
template <int KIND> __global__ void __launch_bounds__(1024) pipe(double *out, long long *cycles, double c1, double c2, int iters) {
    double d[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = 1.0 + (threadIdx.x + i) * 1e-9; e[i] = 0.5 + i * 1e-7; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));           // 3 distinct register operands
            if (KIND == 1) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(d[i]));                                // one register
            if (KIND == 2) asm volatile("fma.rn.f64 %0, %0, %1, 0d4338000000000000;" : "+d"(d[i]) : "d"(c1));      // immediate addend
            if (KIND == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c1));
            if (KIND == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c1));
            if (KIND == 5) asm volatile("add.rn.f64 %0, %0, 0d4338000000000000;" : "+d"(d[i]));
            if (KIND == 6) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) & 7])); // 3 distinct, per-chain
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- FP64 exact modular butterfly (the library's ArithF64) ----------------------------------------------------------------
struct ModF { double p, pinv; };
D double mulmod(double y, double w, double wp, const ModF &m) {
    const double kMagic = 6755399441055744.0;
    double h = __dmul_rn(y, w);
    double l = __fma_rn(y, w, -h);
    double q = __dadd_rn(__fma_rn(y, wp, kMagic), -kMagic);
    double r = __fma_rn(-q, m.p, h);
    return __dadd_rn(r, l);
}
D void bf64(double &X, double &Y, double w, double wp, const ModF &m) {
    double t = mulmod(Y, w, wp, m), x = X;
    X = __dadd_rn(x, t); Y = __dadd_rn(x, -t);
}
// ---- integer butterfly (the library's ArithU64Fast) -------------------------------------------------------------------------
D u32 lo32(u64 x) { return (u32)x; } D u32 hi32(u64 x) { return (u32)(x >> 32); }
D u64 mk(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
D u64 madw(u32 a, u32 b, u64 c) { return (u64)a * b + c; }
struct ModI { u64 p, p4; u32 np0, np1, H; };
D u64 mul_acc(u64 add, u64 Y, u64 w, u64 wq, const ModI &m) {
    u32 y0 = lo32(Y), y1 = hi32(Y), w0 = lo32(w), w1 = hi32(w), v0 = lo32(wq), v1 = hi32(wq);
    u64 A = (u64)y1 * v0, B = (u64)y0 * v1;
    u64 q = madw(y1, v1, (u64)hi32(A) + (u64)hi32(B));
    u32 q0 = lo32(q), q1 = hi32(q);
    u64 r = madw(y0, w0, add);
    r = madw(q0, m.np0, r);
    u32 rh = hi32(r);
    rh = y1 * w0 + rh; rh = y0 * w1 + rh; rh = q1 * m.np0 + rh; rh = q0 * m.np1 + rh;
    return mk(lo32(r), rh);
}
D void bfi(u64 &X, u64 &Y, u64 w, u64 wq, const ModI &m) {
    u64 x = X; if (hi32(x) >= m.H) x -= m.p4;
    u64 xn = mul_acc(x, Y, w, wq, m);
    X = xn; Y = (x + m.p4) + (x - xn);
}

// MODE 0: all warps FP64; 1: all warps integer; 2: even warps FP64, odd warps integer. R = log2 of values per thread.
template <int MODE, int R, int MAXT> __global__ void __launch_bounds__(MAXT) bfly(double *dd, const double2 *twd, const ulonglong2 *twi, ModF mf, ModI mi,
                                                                        int reps, long long *cycles, int nthreads) {
    constexpr int K = 1 << R;
    const bool fp = MODE == 0 || (MODE == 2 && ((threadIdx.x >> 5) & 1) == 0);
    double x[K]; u64 y[K];
    for (int i = 0; i < K; i++) { x[i] = dd[(i * nthreads + threadIdx.x) & 4095]; y[i] = (u64)(x[i] * 1e6) + threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    if (fp) {
#pragma unroll 1
        for (int rep = 0; rep < reps; rep++) {
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int half = K >> (u + 1);
#pragma unroll
                for (int j = 0; j < (1 << u); j++) {
                    double2 w = twd[((1 << u) + j + rep) & 63];
#pragma unroll
                    for (int r = 0; r < half; r++) { int k = (j << (R - u)) | r; bf64(x[k], x[k + half], w.x, w.y, mf); }
                }
            }
        }
    } else {
#pragma unroll 1
        for (int rep = 0; rep < reps; rep++) {
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int half = K >> (u + 1);
#pragma unroll
                for (int j = 0; j < (1 << u); j++) {
                    ulonglong2 w = twi[((1 << u) + j + rep) & 63];
#pragma unroll
                    for (int r = 0; r < half; r++) { int k = (j << (R - u)) | r; bfi(y[k], y[k + half], w.x, w.y, mi); }
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < K; i++) s += x[i] + (double)y[i];
    dd[4096 + blockIdx.x * 1024 + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int sms;
static double avg_cycles(long long *cyc) {
    long long h[256]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double a = 0; for (int i = 0; i < sms; i++) a += h[i]; return a / sms;
}
template <int KIND> void runp(const char *name, double *out, long long *cyc) {
    for (int threads : {1024, 512, 256}) {
        pipe<KIND><<<sms, threads>>>(out, cyc, 1.0000001, 1e-12, 2048);
        cudaDeviceSynchronize();
        pipe<KIND><<<sms, threads>>>(out, cyc, 1.0000001, 1e-12, 2048);
        cudaDeviceSynchronize();
        double a = avg_cycles(cyc);
        printf("{\"kind\": \"%s\", \"threads_per_sm\": %d, \"thread_inst_per_clk_per_sm\": %.2f}\n", name, threads, (double)threads * 2048 * 8 / a);
    }
}
template <int MODE, int R, int MAXT = 1024> void runb(const char *name, double *dd, double2 *twd, ulonglong2 *twi, ModF mf, ModI mi, long long *cyc) {
    for (int threads : {1024, 768, 512, 256}) {
        if (threads > MAXT) continue;
        const int reps = 64;
        bfly<MODE, R, MAXT><<<sms, threads>>>(dd, twd, twi, mf, mi, reps, cyc, threads);
        cudaDeviceSynchronize();
        bfly<MODE, R, MAXT><<<sms, threads>>>(dd, twd, twi, mf, mi, reps, cyc, threads);
        cudaError_t e = cudaDeviceSynchronize();
        double a = avg_cycles(cyc);
        double bf = (double)threads * ((1 << R) / 2) * R * reps;
        printf("{\"variant\": \"%s\", \"radix_log\": %d, \"threads_per_sm\": %d, \"butterflies_per_clk_per_sm\": %.3f, \"err\": %d}\n", name, R, threads, bf / a, (int)e);
    }
}
int main() {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *dd; double2 *twd; ulonglong2 *twi; long long *cyc;
    cudaMalloc(&dd, (4096 + 256 * 1024) * 8); cudaMalloc(&twd, 64 * 16); cudaMalloc(&twi, 64 * 16); cudaMalloc(&cyc, 256 * 8);
    const u64 p40 = 1099510054913ull, p60 = 0xffffffffffc0001ull;
    double hd[4096]; for (int i = 0; i < 4096; i++) hd[i] = (double)((i * 2654435761u) % 1000003);
    cudaMemcpy(dd, hd, sizeof(hd), cudaMemcpyHostToDevice);
    double2 htd[64]; ulonglong2 hti[64];
    for (int i = 0; i < 64; i++) {
        u64 w = (0x123456789abcdefull * (i + 3)) % p40; htd[i].x = (double)w; htd[i].y = (double)w / (double)p40;
        u64 v = (0x123456789abcdefull * (i + 3)) % p60; hti[i].x = v; hti[i].y = (u64)(((unsigned __int128)v << 64) / p60);
    }
    cudaMemcpy(twd, htd, sizeof(htd), cudaMemcpyHostToDevice); cudaMemcpy(twi, hti, sizeof(hti), cudaMemcpyHostToDevice);
    ModF mf{(double)p40, 1.0 / (double)p40};
    ModI mi; mi.p = p60; mi.p4 = 4 * p60; u64 np = 0 - p60; mi.np0 = (u32)np; mi.np1 = (u32)(np >> 32); mi.H = (u32)((4 * p60) >> 32) + 1;
    printf("{\"sms\": %d}\n", sms);
    runp<0>("dfma r,r,r (shared operands)", dd, cyc);
    runp<1>("dfma one register", dd, cyc);
    runp<2>("dfma r,r,imm", dd, cyc);
    runp<3>("dadd r,r", dd, cyc);
    runp<4>("dmul r,r", dd, cyc);
    runp<5>("dadd r,imm", dd, cyc);
    runp<6>("dfma 3 distinct per chain", dd, cyc);
    runb<0, 3>("fp64", dd, twd, twi, mf, mi, cyc);
    runb<0, 4, 512>("fp64", dd, twd, twi, mf, mi, cyc);
    runb<1, 3>("int60", dd, twd, twi, mf, mi, cyc);
    runb<1, 4, 512>("int60", dd, twd, twi, mf, mi, cyc);
    runb<2, 3>("mixed fp64/int60 by warp (butterflies of both kinds counted)", dd, twd, twi, mf, mi, cyc);
    return 0;
}
