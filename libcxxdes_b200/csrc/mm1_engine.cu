// mm1_engine.cu -- M/M/1 (examples/producer_consumer.cpp:9-59) on the generic device engine.
//
// This is the straightforward lowering: three processes (co_main, producer, consumer) as
// resume-point state machines over the engine primitives, one thread per replication, event
// heap in shared memory, sync::queue<double> as a ring in global memory (so it cannot
// overflow below `queue_capacity`).  It exists as the reference lowering the fused kernel
// (mm1_fused.cu) is cross-checked against, and it handles the degenerate n_packets < 2 cases.
#include "device/models_common.cuh"
#include "device/process.cuh"
#include "device/trace.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int HEAP_CAP = 4;  // measured peak is 2; 4 leaves slack, overflow is flagged

struct mm1_state {
    process main, prod, cons;
    any_all all;       // the all_of of co_main
    wait_list waiters; // queue<double>::event_
    uint64_t i, n;     // producer loop index, consumer count
    uint32_t head, count;  // ring
    int64_t x_tick;
    double total_latency, avg_latency;
    uint32_t max_queue;
};

}  // namespace

// ENV = traced_environment: the same kernel writing the event trace of its replications (cxxdes_b200_trace).
template <typename ENV>
__global__ void __launch_bounds__(128)
mm1_engine_kernel(cxxdes_b200_mm1_params prm, shard sh, device_columns out, int64_t *ring_base,
                  uint32_t ring_cap, trace_sink sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    lane_array<uint64_t> hkey = carve<uint64_t>(cursor, HEAP_CAP);
    lane_array<uint32_t> htag = carve<uint32_t>(cursor, HEAP_CAP);
    lane_array<uint32_t> wwho = carve<uint32_t>(cursor, 2);
    lane_array<int32_t> wlat = carve<int32_t>(cursor, 2);

    // ring[slot][thread]: this thread's slots are `total_threads` apart (coalesced across lanes)
    const uint64_t total_threads = (uint64_t)gridDim.x * blockDim.x;
    int64_t *ring = ring_base + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);

    uint64_t rep;
    while (steal(sh, rep)) {
        ENV env;
        env.init(hkey, htag, HEAP_CAP);
        attach_trace(env, sink);
        mm1_state st;
        st.main.init();
        st.prod.init();
        st.cons.init();
        st.waiters.init(wwho, wlat, 2);
        st.i = st.n = 0;
        st.head = st.count = 0;
        st.x_tick = 0;
        st.total_latency = st.avg_latency = 0.0;
        st.max_queue = 0;
        exp_stream lambda, mu;
        lambda.init(sh.seed, sh.first_replication + rep, 1, prm.lambda);
        mu.init(sh.seed, sh.first_replication + rep, 2, prm.mu);

        bind_root(env, st.main, 0);  // simulation::bind, simulation.hpp:76-81

        // environment::run / run_until, environment.ipp:179-196
        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            token t = env.step_pop();
            if (tag_handler_plus1(t.tag)) {
                st.all.invoke(env, t);
                continue;
            }
            uint32_t target = tag_target(t.tag);
            if (target == 0) {
                // co_main: co_await (producer() && consumer());  producer_consumer.cpp:56-58
                if (st.main.pc == 0) {
                    bind(env, st.prod, 1, st.main.priority);  // all_of::await_bind, any_of.ipp:41-48
                    bind(env, st.cons, 2, st.main.priority);
                    st.all.arm(true, 2, 2);                    // await_ready/suspend, :50-84
                    await_completion(env, st.prod, make_tag(0, 1));
                    await_completion(env, st.cons, make_tag(0, 1));
                    st.main.pc = 1;
                } else {
                    do_return(env, st.main);
                }
            } else if (target == 1) {
                // producer, producer_consumer.cpp:31-36
                if (st.i < prm.n_packets) {
                    // q.put(now_seconds()): emplace; wake  (queue.hpp:48-53).  The payload is a
                    // pure function of the tick, so the tick is stored and converted at pop time.
                    if (st.count >= ring_cap) {
                        env.status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
                    } else {
                        uint32_t slot = st.head + st.count;
                        if (slot >= ring_cap) slot -= ring_cap;
                        ring[(uint64_t)slot * total_threads] = env.now;
                        ++st.count;
                    }
                    st.waiters.wake(env);
                    if (st.count > st.max_queue) st.max_queue = st.count;
                    timeout(env, st.prod, 1, real_to_sim(lambda.draw(table), sh.clk));  // :34
                    ++st.i;
                } else {
                    do_return(env, st.prod);
                }
            } else if (target == 2) {
                // consumer, producer_consumer.cpp:38-54
                if (st.cons.pc == 2)
                    st.total_latency += (sim_to_seconds(env.now, sh.clk) - sim_to_seconds(st.x_tick, sh.clk));
                // q.pop(): while (!can_pop) wait; front; pop; wake  (queue.hpp:58-65)
                if (st.count == 0) {
                    st.waiters.wait(env, 2, st.cons.priority);
                    st.cons.pc = 1;
                } else {
                    st.x_tick = ring[(uint64_t)st.head * total_threads];
                    if (++st.head >= ring_cap) st.head = 0;
                    --st.count;
                    st.waiters.wake(env);
                    ++st.n;
                    if (st.n == prm.n_packets) {
                        st.avg_latency = st.total_latency / (double)prm.n_packets;  // :46
                        do_return(env, st.cons);
                    } else {
                        timeout(env, st.cons, 2, real_to_sim(mu.draw(table), sh.clk));  // :50
                        st.cons.pc = 2;
                    }
                }
            }
            // TAG_NO_TARGET: null-target completion, nothing to do (still one event)
            if (env.status) break;
        }
        if (sh.run_until >= 0) {
            if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
            env.now = env.now > sh.run_until ? env.now : sh.run_until;  // environment.ipp:194
        }
        rep_out o;
        o.end_time = env.now;
        o.n_events = env.n_events;
        o.trace_hash = env.hash;
        o.status = env.status;
        o.f64[0] = st.avg_latency;
        o.f64[1] = st.total_latency;
        o.i64[0] = (int64_t)st.n;
        o.i64[1] = (int64_t)st.max_queue;
        store_result(out, rep, o);
    }
}

size_t mm1_engine_smem_bytes(int threads) {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 16 +
           (size_t)threads * (HEAP_CAP * (8 + 4) + 2 * (4 + 4)) + 64;
}

cudaError_t launch_mm1_engine(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                              int64_t *ring, uint32_t ring_cap, int blocks, int threads, cudaStream_t stream) {
    mm1_engine_kernel<environment><<<blocks, threads, mm1_engine_smem_bytes(threads), stream>>>(prm, sh, out, ring, ring_cap,
                                                                                             trace_sink{});
    return cudaGetLastError();
}

// One replication (sh.n_replications == 1), one thread, with its event trace.
cudaError_t launch_mm1_trace(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out, int64_t *ring,
                             uint32_t ring_cap, const trace_sink &sink, cudaStream_t stream) {
    mm1_engine_kernel<traced_environment><<<1, 32, mm1_engine_smem_bytes(32), stream>>>(prm, sh, out, ring, ring_cap, sink);
    return cudaGetLastError();
}

}  // namespace cxb
