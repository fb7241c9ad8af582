// device/process.cuh -- process runtime semantics (start / completion tokens, priority inheritance).
//
// A reference process is a coroutine<> plus its coroutine_data (coroutine.ipp:1-357,
// coroutine_data.ipp:1-228).  The C++ coroutine frame cannot run on the device, so each
// process body is lowered to a resume-point state machine (`pc`); what remains of the
// runtime is the token traffic, which this header reproduces.
#pragma once
#include "engine.cuh"

namespace cxb {

struct process {
    int64_t priority;  // coroutine_data::priority_ (inherit until bound)
    int64_t latency;   // coroutine_data::latency_
    int pc;
    bool bound;
    bool complete;
    // completion_tokens_ (coroutine_data.ipp:179-197, reserve(2)): registered, not scheduled.
    int n_completion;
    uint32_t ctag[2];
    int64_t cprio[2];
    int64_t clat[2];

    __device__ __forceinline__ void init(int64_t prio = PRIO_INHERIT, int64_t lat = 0) {
        priority = prio;
        latency = lat;
        pc = 0;
        bound = false;
        complete = false;
        n_completion = 0;
    }
};

// coroutine_data::bind_, environment.ipp:281-307; inheritance await_transform.ipp:50-52.
__device__ __forceinline__ void bind(environment &env, process &p, uint32_t index, int64_t inherited_priority) {
    if (p.bound) return;
    p.bound = true;
    if (p.priority == PRIO_INHERIT) p.priority = inherited_priority;
    env.schedule(env.now + p.latency, p.priority, make_tag(index));
}

// coroutine::await_suspend, coroutine.ipp:194-207.
__device__ __forceinline__ void await_completion(environment &env, process &child, uint32_t awaiter_tag,
                                                 int64_t return_latency = 0,
                                                 int64_t return_priority = PRIO_INHERIT) {
    if (child.n_completion >= 2) {
        env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
        return;
    }
    int k = child.n_completion++;
    child.ctag[k] = awaiter_tag;
    child.cprio[k] = return_priority == PRIO_INHERIT ? child.priority : return_priority;
    child.clat[k] = return_latency;
}

// coroutine_data_::do_return, environment.ipp:343-348 with schedule_completion_ :321-329.
__device__ __forceinline__ void do_return(environment &env, process &p) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (k < p.n_completion) env.schedule(p.clat[k] + env.now, p.cprio[k], p.ctag[k]);
    p.n_completion = 0;
    p.complete = true;
}

// environment::bind(coroutine), coroutine.ipp:352-357: start token + null-target completion.
__device__ __forceinline__ void bind_root(environment &env, process &p, uint32_t index) {
    bind(env, p, index, 0);
    if (!p.complete) await_completion(env, p, make_tag(TAG_NO_TARGET));
}

// timeout_base::await_suspend (relative), timeout.ipp:21-32.
__device__ __forceinline__ void timeout(environment &env, const process &p, uint32_t index, int64_t ticks,
                                        uint32_t handler_plus1 = 0) {
    env.schedule(env.now + ticks, p.priority, make_tag(index, handler_plus1));
}

}  // namespace cxb
