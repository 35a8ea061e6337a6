// Generic CTA launcher.
//
// Every kernel in this library is a functor F with
//     static constexpr int kThreads;                      // CTA size (compile time, feeds __launch_bounds__)
//     __device__ void operator()(u32 bid, u64 *smem) const // the work of CTA `bid`
// Inside, code that all threads of the CTA execute between two barriers is wrapped in TB_PHASE(t, ...),
// where `t` is the thread index. On the GPU a phase is the body followed by __syncthreads(); nothing is
// carried in registers across phases.
//
// TROYB_HOST_EMUL (tests/emul only, never part of the product library) re-targets the same functors to a
// sequential CPU loop over CTAs and threads so that index arithmetic can be debugged without a GPU.
#pragma once
#include "context.h"
#include "modarith.cuh"

namespace troyb {

inline unsigned long long &launch_counter() {
    static unsigned long long n = 0;
    return n;
}

struct ProfRecord {
    const char *name;
    u32 grid;
    cudaEvent_t e0, e1;
};
struct Profiler {
    bool enabled = false;
    std::vector<ProfRecord> records;
};
inline Profiler &profiler() {
    static Profiler p;
    return p;
}

#if defined(TROYB_HOST_EMUL)

#define TB_PHASE(NT, t, ...)                       \
    for (u32 t = 0; t < (u32)(NT); t++) {          \
        __VA_ARGS__                                \
    }

template <class F> inline cudaError_t launch_cta(const F &f, u32 grid, size_t smem_bytes, cudaStream_t) {
    std::vector<u64> sm(smem_bytes / 8 + 1);
    for (u32 b = 0; b < grid; b++) f(b, sm.data());
    return cudaSuccess;
}

#else

#define TB_PHASE(NT, t, ...)                       \
    {                                              \
        const u32 t = threadIdx.x;                 \
        __VA_ARGS__                                \
    }                                              \
    __syncthreads();

template <class F> __global__ void __launch_bounds__(F::kThreads) k_cta(const F f) {
    extern __shared__ __align__(16) u64 tb_smem[];
    f(blockIdx.x, tb_smem);
}

template <class F> inline cudaError_t launch_cta(const F &f, u32 grid, size_t smem_bytes, cudaStream_t stream) {
    if (grid == 0) return cudaSuccess;
    if (smem_bytes > 48 * 1024) {
        static bool configured = false; // per functor type
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(k_cta<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
            configured = true;
        }
    }
    Profiler &prof = profiler();
    if (prof.enabled) {
        // bench.py's per-kernel timing pass: CUDA events on the launching stream around this one launch
        ProfRecord r;
        r.name = f.name();
        r.grid = grid;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, stream);
        k_cta<F><<<grid, F::kThreads, smem_bytes, stream>>>(f);
        cudaEventRecord(r.e1, stream);
        prof.records.push_back(r);
    } else {
        k_cta<F><<<grid, F::kThreads, smem_bytes, stream>>>(f);
    }
    launch_counter()++;
    return cudaGetLastError();
}

#endif

} // namespace troyb
