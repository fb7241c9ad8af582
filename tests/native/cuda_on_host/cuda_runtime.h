// tests/native/cuda_on_host/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the library's .cu sources (libcxxdes_b200/csrc, unchanged
// except for the three textual rewrites of build.py: kernel launches, `extern __shared__`, inline PTX loads/stores)
// and run their kernels on the CPU, one block after another, every CUDA thread of a block as a fibre:
//   * __syncthreads / __syncwarp / votes / shuffles are real rendez-vous points between the fibres of a block / warp
//     (a lane that reaches one waits until every lane named in the mask has arrived -- lanes run in any order in between);
//   * shared memory is a per-block arena, "32-bit shared addresses" are offsets from an origin just below it;
//   * device memory is host memory (poisoned, not zeroed, as on the device), streams run in host order.
// What this buys: the `-m gpu` parity tests' LOGIC -- the kernels' state machines, heaps, hand-overs, saved state --
// can be run against the oracle where no GPU exists (tests/test_kernels_on_host.py).  What it does not show: timing,
// memory ordering between warps, register/shared-memory limits, the PTX slot of the headline kernel (its C++ twin runs
// instead, as in tests/native/mm1_lane_host.cpp).  The resulting library is built under tests/native/_build only, is
// never installed next to libcxxdes_b200.so and libcxxdes_b200 never loads it.
#pragma once
#define CUDA_ON_HOST 1
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>
#include <utility>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static      // blocks run one after another; a kernel's static shared variables are reused
#define __constant__

// ---- vector types ------------------------------------------------------------------------------------------------
struct __attribute__((aligned(8))) uint2 { unsigned int x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned int x, y, z, w; };
struct uint3 { unsigned int x, y, z; };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned long long x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x((unsigned int)x_), y(y_), z(z_) {}
};

// ---- runtime API (runtime.cpp) -------------------------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cuda_host_stream;
struct cuda_host_event;
typedef cuda_host_stream *cudaStream_t;
typedef cuda_host_event *cudaEvent_t;
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin, sharedMemPerMultiprocessor, totalGlobalMem;
    int regsPerMultiprocessor, maxThreadsPerMultiProcessor;
};

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int priority);
cudaError_t cudaDeviceGetStreamPriorityRange(int *low, int *high);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- the running thread ----------------------------------------------------------------------------------------------
namespace cuda_host {

struct block_ctx;
struct thread_ctx {
    uint3 thread_idx;
    block_ctx *block;
};
struct block_ctx {
    uint3 idx, dim, grid;
    unsigned char *dynamic_shared;
};
extern thread_ctx *current;          // one OS thread runs every fibre: a plain global
extern uintptr_t shared_origin;      // 32-bit "shared addresses" are offsets from here (4 KiB below the running block's arena)

inline unsigned char *dynamic_shared() { return current->block->dynamic_shared; }
inline uintptr_t shared_base() { return shared_origin; }
template <class T> inline T *shared_pointer(uint32_t addr) { return (T *)(shared_base() + addr); }

// rendez-vous of the lanes named in `mask` (runtime.cpp); `op` names the call for the divergence check.  Returns the 32
// deposited words, valid until this lane's next rendez-vous.
enum { OP_SYNCWARP = 1, OP_VOTE = 2, OP_SHFL = 3 };
const uint64_t *warp_gather(unsigned mask, uint64_t value, int op);
void block_barrier();
void yield();
int lane_id();

struct launch_config {
    dim3 grid, block;
    size_t shared_bytes;
};
void run_grid(const launch_config &cfg, void (*thread_entry)(void *), void *closure, const char *what);

template <class... P>
struct launcher {
    launch_config cfg;
    void (*kernel)(P...);
    const char *name;
    struct closure_t {
        void (*kernel)(P...);
        std::tuple<P...> args;
    };
    static void entry(void *c) {
        closure_t *cl = (closure_t *)c;
        std::apply(cl->kernel, cl->args);   // by-value parameters: every thread gets its own copies
    }
    template <size_t... I, class... A>
    static void fill(std::tuple<P...> &t, std::index_sequence<I...>, A &&...a) {
        ((std::get<I>(t) = std::forward<A>(a)), ...);
    }
    // Trailing parameters the call leaves out are value-initialised: a function pointer carries no default arguments,
    // and the one kernel that has any (mm1_ahead_kernel: done_flag = nullptr, list_capacity = 0) defaults them to zero.
    template <class... A>
    void operator()(A &&...a) const {
        static_assert(sizeof...(A) <= sizeof...(P), "more arguments than kernel parameters");
        closure_t cl{kernel, std::tuple<P...>{}};
        fill(cl.args, std::index_sequence_for<A...>{}, std::forward<A>(a)...);
        run_grid(cfg, &entry, &cl, name);
    }
};
template <class... P>
inline launcher<P...> launch(dim3 grid, dim3 block, size_t shared_bytes, cudaStream_t, void (*kernel)(P...), const char *name) {
    return launcher<P...>{launch_config{grid, block, shared_bytes}, kernel, name};
}

// inline PTX rewritten by build.py
template <class T> inline T lds(uint32_t addr) { return *shared_pointer<T>(addr); }
template <class T> inline void sts(uint32_t addr, T v) { *shared_pointer<T>(addr) = v; }
template <class T> inline T ldg(const void *p) { T v; memcpy(&v, p, sizeof(T)); return v; }
template <class T> inline void stg(const void *p, T v) { memcpy((void *)p, &v, sizeof(T)); }
uint64_t globaltimer();

}  // namespace cuda_host

#define threadIdx (::cuda_host::current->thread_idx)
#define blockIdx (::cuda_host::current->block->idx)
#define blockDim (::cuda_host::current->block->dim)
#define gridDim (::cuda_host::current->block->grid)
static const int warpSize = 32;

// ---- warp and block primitives ---------------------------------------------------------------------------------------
static inline void __syncthreads() { ::cuda_host::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { ::cuda_host::warp_gather(mask, 0, ::cuda_host::OP_SYNCWARP); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t *w = ::cuda_host::warp_gather(mask, pred != 0, ::cuda_host::OP_VOTE);
    unsigned b = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && w[l]) b |= 1u << l;
    return b;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { return 0xFFFFFFFFu; }
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *w = ::cuda_host::warp_gather(mask, raw, ::cuda_host::OP_SHFL);
    const int lane = ::cuda_host::lane_id();
    const int from = (lane & ~(width - 1)) | (src & (width - 1));
    T r;
    memcpy(&r, &w[from], sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *w = ::cuda_host::warp_gather(mask, raw, ::cuda_host::OP_SHFL);
    const int lane = ::cuda_host::lane_id();
    int from = lane + (int)delta;
    if (from > (lane | (width - 1))) from = lane;   // past the end of the segment: own value
    T r;
    memcpy(&r, &w[from], sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    return __shfl_sync(mask, v, (::cuda_host::lane_id() ^ lane_mask), width);
}

// one OS thread: atomics are plain read-modify-writes
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { ::cuda_host::yield(); }

// ---- arithmetic intrinsics ---------------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline long long __double_as_longlong(double d) { long long b; memcpy(&b, &d, 8); return b; }
static inline double __longlong_as_double(long long b) { double d; memcpy(&d, &b, 8); return d; }
static inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; memcpy(&d, &b, 8); return d;
}
static inline int __double2loint(double d) { unsigned long long b; memcpy(&b, &d, 8); return (int)(unsigned)b; }
static inline int __double2hiint(double d) { unsigned long long b; memcpy(&b, &d, 8); return (int)(unsigned)(b >> 32); }
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)(uint32_t)((uintptr_t)p - ::cuda_host::shared_base()); }
