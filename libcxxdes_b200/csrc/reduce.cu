// reduce.cu -- per-GPU reduction of the per-replication result columns to the accumulator struct
// that the final all-reduce sums across GPUs.  Deterministic: fixed grid, each block
// reduces a fixed slice with a fixed tree, and one block folds the block partials in order.
#include "kernels.h"

namespace cxb {

namespace {

constexpr int RED_THREADS = 256;

struct partial {
    unsigned long long n_events, n_errors, hash_sum;
    long long end_sum;
    long long i64_sum[CXXDES_B200_N_I64];
    double f64_sum[CXXDES_B200_N_F64];
    double f64_sumsq[CXXDES_B200_N_F64];
};

__device__ __forceinline__ void zero(partial &p) {
    p.n_events = p.n_errors = p.hash_sum = 0;
    p.end_sum = 0;
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) p.i64_sum[k] = 0;
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) p.f64_sum[k] = p.f64_sumsq[k] = 0.0;
}

__device__ __forceinline__ void add(partial &a, const partial &b) {
    a.n_events += b.n_events;
    a.n_errors += b.n_errors;
    a.hash_sum += b.hash_sum;
    a.end_sum += b.end_sum;
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) a.i64_sum[k] += b.i64_sum[k];
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) {
        a.f64_sum[k] += b.f64_sum[k];
        a.f64_sumsq[k] += b.f64_sumsq[k];
    }
}

__device__ __forceinline__ partial shuffle_down(const partial &p, int delta) {
    partial q;
    q.n_events = __shfl_down_sync(0xFFFFFFFFu, p.n_events, delta);
    q.n_errors = __shfl_down_sync(0xFFFFFFFFu, p.n_errors, delta);
    q.hash_sum = __shfl_down_sync(0xFFFFFFFFu, p.hash_sum, delta);
    q.end_sum = __shfl_down_sync(0xFFFFFFFFu, p.end_sum, delta);
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) q.i64_sum[k] = __shfl_down_sync(0xFFFFFFFFu, p.i64_sum[k], delta);
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) {
        q.f64_sum[k] = __shfl_down_sync(0xFFFFFFFFu, p.f64_sum[k], delta);
        q.f64_sumsq[k] = __shfl_down_sync(0xFFFFFFFFu, p.f64_sumsq[k], delta);
    }
    return q;
}

__device__ __forceinline__ partial block_reduce(partial p) {
    __shared__ partial warp_part[RED_THREADS / 32];
    for (int d = 16; d > 0; d >>= 1) add(p, shuffle_down(p, d));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x < 32) {
        partial q;
        zero(q);
        if (threadIdx.x < RED_THREADS / 32) q = warp_part[threadIdx.x];
        for (int d = 16; d > 0; d >>= 1) add(q, shuffle_down(q, d));
        p = q;
    }
    return p;  // valid in thread 0
}

__global__ void __launch_bounds__(RED_THREADS) reduce_columns_kernel(device_columns c, uint64_t n, partial *parts) {
    // contiguous slice per block, coalesced within the slice
    uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    uint64_t begin = per_block * blockIdx.x;
    uint64_t end = begin + per_block < n ? begin + per_block : n;
    partial p;
    zero(p);
    for (uint64_t r = begin + threadIdx.x; r < end; r += RED_THREADS) {
        p.n_events += c.n_events[r];
        // (a replication that run_until / run_for stopped at the horizon is not an error: cxxdes_b200_fetch ignores
        // the UNFINISHED bit the same way)
        p.n_errors += (c.status[r] & ~(uint32_t)CXXDES_B200_REP_UNFINISHED) != 0;
        p.hash_sum += c.trace_hash[r];
        p.end_sum += c.end_time[r];
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) p.i64_sum[k] += c.i64[k][r];
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) {
            double v = c.f64[k][r];
            p.f64_sum[k] += v;
            p.f64_sumsq[k] += v * v;
        }
    }
    p = block_reduce(p);
    if (threadIdx.x == 0) parts[blockIdx.x] = p;
}

__global__ void __launch_bounds__(RED_THREADS)
reduce_final_kernel(const partial *parts, int n_parts, uint64_t n, cxxdes_b200_accumulators *acc) {
    partial p;
    zero(p);
    // each thread folds a strided subset in index order; the tree that follows is fixed
    for (int b = threadIdx.x; b < n_parts; b += RED_THREADS) add(p, parts[b]);
    p = block_reduce(p);
    if (threadIdx.x == 0) {
        acc->n_replications = n;
        acc->n_events = p.n_events;
        acc->n_errors = p.n_errors;
        acc->trace_hash_sum = p.hash_sum;
        acc->end_time_sum = p.end_sum;
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) acc->i64_sum[k] = p.i64_sum[k];
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) {
            acc->f64_sum[k] = p.f64_sum[k];
            acc->f64_sumsq[k] = p.f64_sumsq[k];
        }
    }
}

}  // namespace

size_t reduce_scratch_bytes(int sm_count) { return sizeof(partial) * (size_t)sm_count * 2; }

cudaError_t launch_reduce(const device_columns &cols, uint64_t n, cxxdes_b200_accumulators *device_acc,
                          void *scratch, int sm_count, cudaStream_t stream) {
    int blocks = sm_count * 2;
    partial *parts = (partial *)scratch;
    reduce_columns_kernel<<<blocks, RED_THREADS, 0, stream>>>(cols, n, parts);
    reduce_final_kernel<<<1, RED_THREADS, 0, stream>>>(parts, blocks, n, device_acc);
    return cudaGetLastError();
}

// run() after run_until(T): the reference left every replication at now_ >= T (environment.ipp:194), and the
// events that run() then dispatches are all later than T, so the final clock is max(natural end, T).  The
// kernels that replay a replication from tick 0 report the natural end; this pass applies the floor.
__global__ void clock_floor_kernel(int64_t *end_time, uint64_t n, int64_t floor) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        if (end_time[r] < floor) end_time[r] = floor;
}

cudaError_t launch_clock_floor(int64_t *end_time, uint64_t n, int64_t floor, int sm_count, cudaStream_t stream) {
    clock_floor_kernel<<<sm_count * 4, 256, 0, stream>>>(end_time, n, floor);
    return cudaGetLastError();
}

// run_for(dt) when nothing is scheduled any more: run_until(now() + dt) only moves each replication's own clock
// (environment.ipp:190-196,205-209); saturates instead of wrapping.
__global__ void clock_add_kernel(int64_t *end_time, uint64_t n, int64_t dt) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t t = end_time[r];
        end_time[r] = t > INT64_MAX - dt ? INT64_MAX : t + dt;
    }
}

cudaError_t launch_clock_add(int64_t *end_time, uint64_t n, int64_t dt, int sm_count, cudaStream_t stream) {
    clock_add_kernel<<<sm_count * 4, 256, 0, stream>>>(end_time, n, dt);
    return cudaGetLastError();
}

}  // namespace cxb
