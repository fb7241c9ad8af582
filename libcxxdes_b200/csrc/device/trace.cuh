// device/trace.cuh -- event trace of ONE replication (cxxdes_b200_trace): the (time, priority, kind, process) of every
// token environment::step pops (environment.ipp:117-146), the sequence the trace digest folds, written out in order.
//
// The reference has nothing but debug prints here (token.ipp:50-60); the oracle harness records the same four fields
// from env.next_event() before each env.step() (oracle/ref/harness.hpp `dump`), so a digest mismatch can be localised
// by diffing the two lists.  Tracing is a separate instantiation of the full-size kernels (ENV = traced_environment);
// the kernels a run() launches carry none of it.
#pragma once
#include "engine.cuh"

namespace cxb {

struct trace_sink {
    int64_t *time;
    int64_t *priority;
    uint32_t *kind;      // token_kind (shared/trace_hash.hpp)
    uint32_t *proc;      // process ordinal, PROC_NONE for a null-target token
    uint64_t max_events;
};

struct traced_environment : environment {
    trace_sink sink;

    __device__ __forceinline__ void record(uint64_t at, int64_t time, int64_t priority, uint32_t kind, uint32_t proc) const {
        if (at >= sink.max_events) return;
        sink.time[at] = time;
        sink.priority[at] = priority;
        sink.kind[at] = kind;
        sink.proc[at] = proc & 0xFFFFu;
    }

    // environment::step_pop, plus one row of the trace
    __device__ __forceinline__ token step_pop() {
        const token t = environment::step_pop();
        const uint32_t target = tag_target(t.tag);
        const uint32_t kind = tag_handler_plus1(t.tag) ? TOKEN_HANDLER : (target != TAG_NO_TARGET ? TOKEN_RESUME : TOKEN_NOOP);
        record(n_events - 1, key_time(t.key), key_prio(t.key), kind, target);
        return t;
    }
};

__device__ __forceinline__ void attach_trace(environment &, const trace_sink &) {}
__device__ __forceinline__ void attach_trace(traced_environment &env, const trace_sink &sink) { env.sink = sink; }

}  // namespace cxb
