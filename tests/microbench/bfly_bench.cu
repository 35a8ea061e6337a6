// Butterfly-variant throughput experiment (development aid for DESIGN.md §butterfly): 32 coefficients per thread,
// 5 radix-2 layers (80 butterflies) repeated REPS times in registers; reports butterflies per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define D __device__ __forceinline__
D u64 mk(u32 lo, u32 hi){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
D u32 lo32(u64 x){ return (u32)x; } D u32 hi32(u64 x){ return (u32)(x>>32); }
D u64 madw(u32 a, u32 b, u64 c){ u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
D u64 mulw(u32 a, u32 b){ u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
D u32 madlo(u32 a, u32 b, u32 c){ u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
struct P { u64 p, p2, p4; u32 np0, np1; u32 H; };

D u64 shoup_acc(u64 add, u64 Y, u64 w, u64 wq, const P& m) {
    u32 y0 = lo32(Y), y1 = hi32(Y), w0 = lo32(w), w1 = hi32(w), v0 = lo32(wq), v1 = hi32(wq);
    u64 A = mulw(y1, v0), B = mulw(y0, v1);
    u64 q = madw(y1, v1, (u64)hi32(A) + (u64)hi32(B));
    u32 q0 = lo32(q), q1 = hi32(q);
    u64 r = madw(y0, w0, add);
    r = madw(q0, m.np0, r);
    u32 rh = hi32(r);
    rh = madlo(y1, w0, rh); rh = madlo(y0, w1, rh); rh = madlo(q1, m.np0, rh); rh = madlo(q0, m.np1, rh);
    return mk(lo32(r), rh);
}
template <int V> D void bfly(u64& X, u64& Y, u64 w, u64 wq, const P& m) {
    if (V == 0) { // compiler 64-bit ops, Harvey
        u64 x = X - (X >= m.p2 ? m.p2 : 0);
        u64 t = Y * w - __umul64hi(Y, wq) * m.p;
        X = x + t; Y = x - t + m.p2;
    } else if (V == 1) {
        u64 x = X; if (hi32(x) >= m.H) x -= m.p4;
        u64 t = shoup_acc(0, Y, w, wq, m);
        X = x + t; Y = x + m.p4 - t;
    } else if (V == 2) {
        u64 x = X; if (hi32(x) >= m.H) x -= m.p4;
        u64 xn = shoup_acc(x, Y, w, wq, m);
        X = xn; Y = (x + m.p4) + (x - xn);
    } else if (V == 3) { // exact hi via compiler, fused lo
        u64 x = X; if (hi32(x) >= m.H) x -= m.p4;
        u64 q = __umul64hi(Y, wq);
        u64 t = Y * w - q * m.p;
        X = x + t; Y = x + m.p4 - t;
    }
}
template <int V> __global__ void __launch_bounds__(512) k(u64* data, const ulonglong2* tw, P m, int reps, long long* cycles) {
    u64 x[32];
    for (int i = 0; i < 32; i++) x[i] = data[(blockIdx.x * 32 + i) * 512 + threadIdx.x];
    __syncthreads();
    long long t0 = clock64();
    for (int rep = 0; rep < reps; rep++) {
        #pragma unroll
        for (int u = 0; u < 5; u++) {
            const int half = 16 >> u;
            #pragma unroll
            for (int j = 0; j < (1 << u); j++) {
                ulonglong2 w = tw[((1 << u) + j + rep) & 63];
                #pragma unroll
                for (int r = 0; r < half; r++) {
                    int kk = (j << (5 - u)) | r;
                    bfly<V>(x[kk], x[kk + half], w.x, w.y, m);
                }
            }
        }
    }
    long long t1 = clock64();
    for (int i = 0; i < 32; i++) data[(blockIdx.x * 32 + i) * 512 + threadIdx.x] = x[i];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int KIND> __global__ void __launch_bounds__(1024) pipe(u64 *out, long long *cycles) {
    u64 w[8]; u32 b = threadIdx.x * 2654435761u + 12345u;
    for (int i = 0; i < 8; i++) w[i] = ((u64)(b + i * 77u) << 32) | (b ^ (i * 1315423911u));
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < 2048; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) w[i] = mulw(lo32(w[i]), hi32(w[i]));                  // IMAD.WIDE, RZ addend
            if (KIND == 1) { u64 t = mulw(lo32(w[i]), b); w[i] = t + w[i]; }      // product + 64-bit add
            if (KIND == 2) { u32 h; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(lo32(w[i])), "r"(hi32(w[i]))); w[i] = mk(h, lo32(w[i]) ^ b); }
        }
    }
    long long t1 = clock64();
    u64 s = 0; for (int i = 0; i < 8; i++) s += w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int V> void run(const char* name, int sms, u64* data, ulonglong2* tw, P m, long long* cyc) {
    const int reps = 50;
    k<V><<<sms, 512>>>(data, tw, m, reps, cyc);
    cudaDeviceSynchronize();
    k<V><<<sms, 512>>>(data, tw, m, reps, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double bf = 512.0 * 80 * reps;
    printf("{\"variant\": \"%s\", \"butterflies_per_clk_per_sm\": %.3f, \"clk_per_warp_butterfly_per_smsp\": %.2f}\n", name, bf / avg, avg / (bf / 32 / 4));
}
template <int K> void runp(const char* name, int sms, u64* out, long long* cyc) {
    pipe<K><<<sms, 1024>>>(out, cyc); cudaDeviceSynchronize();
    pipe<K><<<sms, 1024>>>(out, cyc); cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    printf("{\"kind\": \"%s\", \"thread_inst_per_clk_per_sm\": %.2f}\n", name, 1024.0 * 2048 * 8 / avg);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64* data; ulonglong2* tw; long long* cyc;
    cudaMalloc(&data, (size_t)sms * 32 * 512 * 8); cudaMalloc(&tw, 64 * 16); cudaMalloc(&cyc, 256 * 8);
    u64 p = 0xffffffffffc0001ull;
    ulonglong2 htw[64];
    for (int i = 0; i < 64; i++) { u64 w = (0x123456789abcdefull * (i + 3)) % p; htw[i].x = w; htw[i].y = (u64)(((unsigned __int128)w << 64) / p); }
    cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice);
    cudaMemset(data, 1, (size_t)sms * 32 * 512 * 8);
    P m; m.p = p; m.p2 = 2 * p; m.p4 = 4 * p; u64 np = 0 - p; m.np0 = (u32)np; m.np1 = (u32)(np >> 32); m.H = (u32)((4 * p) >> 32) + 1;
    run<0>("compiler_u64_harvey", sms, data, tw, m, cyc);
    run<1>("approxq_split", sms, data, tw, m, cyc);
    run<2>("approxq_chainX", sms, data, tw, m, cyc);
    run<3>("compiler_mulhi_hicmp", sms, data, tw, m, cyc);
    runp<0>("imad.wide.u32 (rz addend, dependent)", sms, data, cyc);
    runp<1>("imad.wide + add64", sms, data, cyc);
    runp<2>("mul.hi.u32", sms, data, cyc);
    return 0;
}
