// device/models_common.cuh -- launch-time structures shared by the model kernels.
#pragma once
#include "engine.cuh"

namespace cxb {

// Result columns in device memory (same layout as cxxdes_b200_columns, device pointers).
struct device_columns {
    int64_t *end_time;
    uint64_t *n_events;
    uint64_t *trace_hash;
    uint32_t *status;
    double *f64[CXXDES_B200_N_F64];
    int64_t *i64[CXXDES_B200_N_I64];
};

// What every model kernel needs to know about the ensemble shard it runs.
struct shard {
    uint64_t seed;
    uint64_t first_replication;  // global id of local replication 0
    uint64_t n_replications;     // local count
    clock_config clk;
    int64_t run_until;           // < 0: run to exhaustion (environment::run)
    unsigned long long *work_counter;  // replication-level work stealing
    const log_entry *log_table;  // global copy of the log table
    unsigned long long *retry_count;   // number of entries in retry_list
    uint32_t *retry_list;        // local replication ids that must be re-run by the fallback kernel
    // M/M/1 runs in up to three passes: list A (queue outgrew the shared-memory window -> second pass
    // of the fast kernel with the two-level ring), then retry_list above (-> 64-bit fallback kernel)
    unsigned long long *retry_a_count, *retry_a_cursor;
    uint32_t *retry_a_list;
};

// A retry-list entry that has not been written (yet): lists read while they are still being appended to
// (mm1_ahead.cu, POLL) are filled with this value before the launch.  Local replication ids are < 2^32 - 1.
constexpr uint32_t LIST_UNSET = 0xFFFFFFFFu;

struct rep_out {
    int64_t end_time;
    uint64_t n_events;
    uint64_t trace_hash;
    uint32_t status;
    double f64[CXXDES_B200_N_F64];
    int64_t i64[CXXDES_B200_N_I64];
};

__device__ __forceinline__ void store_result(const device_columns &c, uint64_t r, const rep_out &o) {
    c.end_time[r] = o.end_time;
    c.n_events[r] = o.n_events;
    c.trace_hash[r] = o.trace_hash;
    c.status[r] = o.status;
#pragma unroll
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) c.f64[k][r] = o.f64[k];
#pragma unroll
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) c.i64[k][r] = o.i64[k];
}

// Next replication for this lane (persistent kernel, lane-level stealing).
__device__ __forceinline__ bool steal(const shard &s, uint64_t &local_rep) {
    unsigned long long r = atomicAdd(s.work_counter, 1ULL);
    local_rep = r;
    return r < s.n_replications;
}

}  // namespace cxb
