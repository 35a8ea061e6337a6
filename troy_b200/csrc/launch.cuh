// Generic CTA launcher.
//
// Every kernel in this library is a functor F with
//     static constexpr int kThreads;                      // CTA size (compile time, feeds __launch_bounds__)
//     __device__ void operator()(u32 bid, u64 *smem) const // the work of CTA `bid`
// Inside, code that all threads of the CTA execute between two barriers is wrapped in TB_PHASE(t, ...),
// where `t` is the thread index. On the GPU a phase is the body followed by __syncthreads(); nothing is
// carried in registers across phases.
//
// Cluster kernels (launch_cluster) are functors with F::kStages stages separated by cluster-wide barriers:
//     template <int S> __device__ void stage(u32 bid, u32 grid, u32 rank, u32 nranks, u64 *smem, const Peers &peers) const
// `peers` reads / writes 64-bit words of the shared memory of any CTA of the cluster (distributed shared memory).
// A last barrier after the final stage keeps every CTA's shared memory alive until its peers have finished reading it.
//
// TROYB_HOST_EMUL (tests/emul only, never part of the product library) re-targets the same functors to a
// sequential CPU loop over CTAs and threads so that index arithmetic can be debugged without a GPU.
#pragma once
#include "context.h"
#include "modarith.cuh"
#include <atomic>
#include <mutex>

namespace troyb {

// Process-wide launch bookkeeping. Host threads may launch concurrently (one evaluator per thread, ctypes releases the
// GIL): the counter is atomic and the profiler's record list is guarded by a mutex.
inline std::atomic<unsigned long long> &launch_counter() {
    static std::atomic<unsigned long long> n{0};
    return n;
}

struct ProfRecord {
    const char *name;
    u32 grid;
    cudaEvent_t e0, e1;
};
struct Profiler {
    std::atomic<bool> enabled{false};
    std::mutex mu;
    std::vector<ProfRecord> records;
    void push(const ProfRecord &r) {
        std::lock_guard<std::mutex> g(mu);
        records.push_back(r);
    }
};
inline Profiler &profiler() {
    static Profiler p;
    return p;
}

// cudaFuncSetAttribute applies to the CURRENT device only: remember per kernel which devices have been configured
// (bit d of the mask = device d; one process may drive several GPUs, one context per device).
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    template <class Fn> cudaError_t ensure(Fn configure) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = 1ull << (dev & 63);
        if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
        e = configure(); // idempotent: two threads racing here both set the same attribute
        if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

#if defined(TROYB_HOST_EMUL)

#define TB_PHASE(NT, t, ...)                       \
    for (u32 t = 0; t < (u32)(NT); t++) {          \
        __VA_ARGS__                                \
    }

template <class F> inline cudaError_t launch_cta(const F &f, u32 grid, size_t smem_bytes, cudaStream_t) {
    std::vector<u64> sm(smem_bytes / 8 + 1);
    for (u32 b = 0; b < grid; b++) f(b, sm.data());
    launch_counter()++;
    return cudaSuccess;
}

struct Peers {
    u64 *const *tiles;
    u32 map(u32 rank) const { return rank; } // handle of a peer's tile
    u64 ld(u32 handle, u32 word) const { return tiles[handle][word]; }
    void st(u32 handle, u32 word, u64 v) const { tiles[handle][word] = v; }
};

inline void l2_prefetch(const void *, u32) {}
inline void cluster_arrive_started() {}
inline void cluster_wait_started() {}

template <class F, int S> inline void emul_stages(const F &f, u32 b0, u32 grid, u32 nranks, std::vector<u64 *> &ptr, const Peers &peers) {
    if constexpr (S < F::kStages) {
        for (u32 r = 0; r < nranks; r++) f.template stage<S>(b0 + r, grid, r, nranks, ptr[r], peers);
        emul_stages<F, S + 1>(f, b0, grid, nranks, ptr, peers);
    }
}

template <class F> inline cudaError_t launch_cluster(const F &f, u32 grid, u32 nranks, size_t smem_bytes, cudaStream_t) {
    std::vector<std::vector<u64>> sm(nranks, std::vector<u64>(smem_bytes / 8 + 1));
    std::vector<u64 *> ptr(nranks);
    for (u32 r = 0; r < nranks; r++) ptr[r] = sm[r].data();
    Peers peers{ptr.data()};
    for (u32 b0 = 0; b0 < grid; b0 += nranks) emul_stages<F, 0>(f, b0, grid, nranks, ptr, peers);
    launch_counter()++;
    return cudaSuccess;
}

#else

#define TB_PHASE(NT, t, ...)                       \
    {                                              \
        const u32 t = threadIdx.x;                 \
        __VA_ARGS__                                \
    }                                              \
    __syncthreads();

// Resident CTAs per SM the register allocation should allow: F::kMinCtas when the functor declares it (0 = leave the
// choice to ptxas, which is not the same as asking for 1).
template <class F, class = void> struct MinCtas { static constexpr int value = 0; };
template <class F> struct MinCtas<F, decltype((void)F::kMinCtas)> { static constexpr int value = F::kMinCtas; };

template <class F> __global__ void __launch_bounds__(F::kThreads, MinCtas<F>::value) k_cta(const F f) {
    extern __shared__ __align__(1024) u64 tb_smem[];
    f(blockIdx.x, tb_smem);
}

template <class F> inline cudaError_t launch_cta(const F &f, u32 grid, size_t smem_bytes, cudaStream_t stream) {
    if (grid == 0) return cudaSuccess;
    if (smem_bytes > 48 * 1024) {
        static DeviceOnce once; // per functor type
        cudaError_t e = once.ensure([] { return cudaFuncSetAttribute(k_cta<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
        if (e != cudaSuccess) return e;
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
        prof.push(r);
    } else {
        k_cta<F><<<grid, F::kThreads, smem_bytes, stream>>>(f);
    }
    launch_counter()++;
    return cudaGetLastError();
}

// ---- thread-block clusters + distributed shared memory ---------------------------------------------------------------
// Distributed-shared-memory accesses carry no "memory" clobber: within a stage a CTA's own tile is touched either only
// through Peers or only through ordinary shared-memory accesses, and stages are separated by cluster_barrier(), which
// does clobber memory. The compiler is therefore free to batch global loads around them.
struct Peers {
    u32 base; // shared::cta address of this CTA's dynamic shared memory (same offset in every CTA of the cluster)
    // shared::cluster address of the start of CTA `rank`'s tile
    TB_D u32 map(u32 rank) const {
        u32 r;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(base), "r"(rank));
        return r;
    }
    TB_D u64 ld(u32 handle, u32 word) const {
        u64 v;
        asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(handle + word * 8u));
        return v;
    }
    TB_D void st(u32 handle, u32 word, u64 v) const {
        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(handle + word * 8u), "l"(v));
    }
};

// Asks the L2 to fetch `bytes` (multiple of 16) at p (16-byte aligned); returns at once.
TB_D void l2_prefetch(const void *p, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes));
}

// Split cluster barrier without memory ordering: "every CTA of the cluster has started". A CTA's shared memory exists only
// once it runs, so a peer may store into it (distributed shared memory) only after this handshake; the arrive is issued
// first thing and the wait just before the first remote store, with the CTA's own global loads in between.
TB_D void cluster_arrive_started() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
TB_D void cluster_wait_started() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }

TB_D void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <class F, int S> TB_D void run_stages(const F &f, u32 rank, u32 nranks, u64 *smem, const Peers &peers) {
    if constexpr (S < F::kStages) {
        f.template stage<S>(blockIdx.x, gridDim.x, rank, nranks, smem, peers);
        if (nranks > 1 && (S + 1 < F::kStages || F::kTrailingBarrier) && f.barrier_after(S)) {
            if (S + 1 < F::kStages) {
                cluster_barrier();
            } else {
                // trailing barrier: "nobody reads my shared memory any more" - execution ordering only, no memory fence
                cluster_arrive_started();
                cluster_wait_started();
            }
        }
        run_stages<F, S + 1>(f, rank, nranks, smem, peers);
    }
}

// The functor travels as a __grid_constant__ parameter: kernels may take the address of its members (tensor maps).
template <class F> __global__ void __launch_bounds__(F::kThreads, F::kMinCtas) k_cluster(const __grid_constant__ F f) {
    extern __shared__ __align__(1024) u64 tb_smem[];
    u32 rank, nranks;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
    const Peers peers{(u32)__cvta_generic_to_shared(tb_smem)};
    run_stages<F, 0>(f, rank, nranks, tb_smem, peers);
}

template <class F> inline cudaError_t launch_cluster(const F &f, u32 grid, u32 nranks, size_t smem_bytes, cudaStream_t stream) {
    if (grid == 0) return cudaSuccess;
    if (grid % nranks) return cudaErrorInvalidValue;
    static DeviceOnce once; // per functor type
    {
        cudaError_t e0 = once.ensure([] { return cudaFuncSetAttribute(k_cluster<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
        if (e0 != cudaSuccess) return e0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(F::kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nranks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    Profiler &prof = profiler();
    cudaError_t e;
    if (prof.enabled) {
        ProfRecord r;
        r.name = f.name();
        r.grid = grid;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, stream);
        e = cudaLaunchKernelEx(&cfg, k_cluster<F>, f);
        cudaEventRecord(r.e1, stream);
        prof.push(r);
    } else {
        e = cudaLaunchKernelEx(&cfg, k_cluster<F>, f);
    }
    launch_counter()++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

#endif

} // namespace troyb
