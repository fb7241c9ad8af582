// models_aloha.cu -- pure ALOHA (examples/aloha.cpp:25-85) on the device engine.
//
// One thread per replication.  The event heap (up to N + 2 tokens, real key ties: 0.3-0.5 % of
// pops) lives in shared memory and performs libstdc++'s exact __push_heap / __adjust_heap moves
// (device/engine.cuh), because tie order among equal (time, priority) keys is observable in the
// trace digest.  Station state is bit-packed in registers: the std::set<bool*> of active
// transmissions is a 64-bit mask, the per-iteration `collided` flags another, the resume point
// of each station two more; the loop counters sit in shared memory next to the heap.
//
// Lowering, processes: co_main = 0, station s = 1 + s, all priorities 0:
//   co_main   all_of.range(stations): bind every station in index order (start tokens), one
//             shared all_of handler on their completion tokens (any_of.ipp:41-84,197-209)
//   station   for i < pps: { collided = false; insert; check_collisions; timeout(frame_time);
//                            erase; if (!collided) ++succ; timeout(exponential()) }
#include "device/models_common.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int ALOHA_THREADS = 128;

struct aloha_rep_params {
    float lambda;
    uint64_t pps;
};

// main(), aloha.cpp:96,101: float arithmetic, size_t operands converted to float.
__device__ __forceinline__ aloha_rep_params aloha_params_of(const cxxdes_b200_aloha_params &p, uint64_t grep) {
    aloha_rep_params r;
    if (p.sweep_steps > 1) {
        const uint64_t num_steps = p.sweep_steps;
        const uint64_t step = grep % num_steps;
        const float begin = p.lambda_begin, end = p.lambda_end;
        const float num = (end - begin) * (float)step;
        r.lambda = begin + num / (float)(num_steps - 1);
        r.pps = (uint64_t)(r.lambda * p.packets_scale);
    } else {
        r.lambda = p.lambda_begin;
        r.pps = p.packets_per_station;
    }
    return r;
}

}  // namespace

__global__ void __launch_bounds__(ALOHA_THREADS)
aloha_kernel(cxxdes_b200_aloha_params prm, shard sh, device_columns out) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    const int N = (int)prm.num_stations;
    const int heap_cap = N + 2;
    lane_array<uint64_t> hkey = carve<uint64_t>(cursor, heap_cap);
    lane_array<uint32_t> htag = carve<uint32_t>(cursor, heap_cap);
    lane_array<uint32_t> loop_i = carve<uint32_t>(cursor, N);

    const int64_t frame_ticks = real_to_sim_f32(prm.frame_time, sh.clk);  // env.timeout(float), aloha.cpp:60
    const uint32_t HANDLER_TAG = make_tag(0, 1);

    uint64_t rep;
    while (steal(sh, rep)) {
        const uint64_t grep = sh.first_replication + rep;
        const aloha_rep_params rp = aloha_params_of(prm, grep);
        // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
        const divisor rate = make_divisor((double)(rp.lambda / (float)prm.num_stations));
        stream_addr addr;
        addr.seed = sh.seed;
        addr.replication = grep;

        environment env;
        env.init(hkey, htag, heap_cap);
        uint64_t active = 0, collided = 0, pc_tx = 0, pc_idle = 0;  // pc_tx: awaiting the frame timeout; pc_idle: awaiting the inter-arrival timeout
        uint64_t successful = 0;
        uint32_t main_pc = 0;
        any_all all;
        float s_result = 0.0f, g_result = 0.0f;
        for (int s = 0; s < N; ++s) loop_i[s] = 0;

        // environment::bind(co_main): start token + null-target completion (coroutine.ipp:352-357)
        env.schedule(0, 0, make_tag(0));

        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            token t = env.step_pop();
            if (tag_handler_plus1(t.tag)) {
                all.invoke(env, t);
            } else {
                const uint32_t target = tag_target(t.tag);
                if (target == 0) {
                    if (main_pc == 0) {
                        // co_main, aloha.cpp:39-44: bind stations in index order, one all_of handler
                        for (int s = 0; s < N; ++s) env.schedule(env.now, 0, make_tag(1 + s));
                        all.arm(true, (uint32_t)N, (uint32_t)N);
                        main_pc = 1;
                    } else {
                        // :46-47  S = succ / now_seconds() (size_t / double -> float), G = lambda * frame_time
                        s_result = (float)((double)successful / sim_to_seconds(env.now, sh.clk));
                        g_result = rp.lambda * prm.frame_time;
                        env.schedule(env.now, 0, make_tag(TAG_NO_TARGET));  // co_main returns
                    }
                } else if (target != TAG_NO_TARGET) {
                    // station, aloha.cpp:51-71
                    const int s = (int)target - 1;
                    const uint64_t bit = 1ULL << s;
                    if (pc_tx & bit) {
                        // frame finished: erase, count, draw the waiting time (:62-69)
                        pc_tx &= ~bit;
                        active &= ~bit;
                        if (!(collided & bit)) ++successful;
                        addr.stream = target;
                        const double wait = exponential(addr, (uint64_t)loop_i[s], rate, table);
                        env.schedule(env.now + real_to_sim(wait, sh.clk), 0, make_tag(target));
                        pc_idle |= bit;
                    } else {
                        uint32_t i = loop_i[s];
                        if (pc_idle & bit) {  // the for loop's ++i after the inter-arrival timeout
                            pc_idle &= ~bit;
                            loop_i[s] = ++i;
                        }
                        if (i < rp.pps) {
                            // collided = false; insert; check_collisions (:55-59,73-78)
                            collided &= ~bit;
                            active |= bit;
                            if (__popcll(active) > 1) collided |= active;
                            env.schedule(env.now + frame_ticks, 0, make_tag(target));
                            pc_tx |= bit;
                        } else {
                            env.schedule(env.now, 0, HANDLER_TAG);  // co_return -> completion token with the all_of handler
                        }
                    }
                }
            }
            if (env.status) break;
        }
        if (sh.run_until >= 0) {
            if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
            env.now = env.now > sh.run_until ? env.now : sh.run_until;
        }
        rep_out o;
        o.end_time = env.now;
        o.n_events = env.n_events;
        o.trace_hash = env.hash;
        o.status = env.status;
        o.f64[0] = (double)s_result;
        o.f64[1] = (double)g_result;
        o.i64[0] = (int64_t)successful;
        o.i64[1] = (int64_t)rp.pps;
        store_result(out, rep, o);
    }
}

size_t aloha_smem_bytes(uint32_t num_stations) {
    const size_t n = num_stations;
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 +
           (size_t)ALOHA_THREADS * ((n + 2) * (8 + 4) + n * 4);
}

cudaError_t launch_aloha(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                         int sm_count, size_t smem_optin, cudaStream_t stream) {
    const size_t smem = aloha_smem_bytes(prm.num_stations);
    if (smem > smem_optin) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(aloha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    aloha_kernel<<<sm_count * per_sm, ALOHA_THREADS, smem, stream>>>(prm, sh, out);
    return cudaGetLastError();
}

}  // namespace cxb
