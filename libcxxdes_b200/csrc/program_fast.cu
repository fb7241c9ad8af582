// program_fast.cu -- first pass of the process-program interpreter: per-replication state in SHARED memory.
//
// program_kernel.cu keeps a replication's process table, wait lists, queues and handler pool in fixed-size
// per-thread arrays (local memory): general, but every interpreted operation waits on L2 (ncu, profiles/
// r01_program_kernel_v2: 24 warps stalled on the long scoreboard per issued instruction, 16 of 32 lanes active).
// This kernel runs the same operation semantics (same citations; the code below restates program_kernel.cu's
// `interp` on a different data layout) with two changes:
//
//  * Layout.  The state is sized FROM THE PROGRAM at create time (program_fast_plan): processes, wait lists,
//    handler pool, semaphores and statistics are 32/64-bit words in shared memory, `[word][thread]` so that a
//    warp access is conflict free wherever each lane's process is.  Only queue payloads (one 8-byte item per
//    put/pop) stay in global memory, `[slot][thread]`.  M/M/1 as a program: 272 B per replication instead of
//    ~9 KB, 768 resident replications per SM.
//  * Control flow.  One loop iteration = one event for every lane: the warp pops together, the resumed
//    bodies run (this is where lanes diverge: one path per (process, pc) in the warp), and the timer the body
//    ended on -- delay / until / env.timeout(exponential) -- is only NOTED by the body: the variate is drawn
//    and the token pushed after the lanes have re-converged, so the expensive part of an event (Philox, log,
//    divide, heap push) is executed once per warp, not once per path.  The token order is unchanged: a timer
//    is the last thing a body does before it suspends.  A lane whose replication ends takes the next one at
//    once (lane-level work stealing).  [A first version with ONE operation per loop iteration and lanes free
//    to drift apart ran 7 of 32 lanes per instruction and was slower than the local-memory kernel.]
//
// Anything this layout cannot hold (heap / wait-list / queue capacity, priorities outside 16 bits, a
// semaphore above 2^32-2, loops nested deeper than planned, ...) raises a status bit, and the replication is
// handed to program_kernel.cu through the retry list: that kernel, with its full-size structures, produces the
// result (or the authoritative error status).  Never a silently different trace.
#include <cstdlib>

#include "device/models_common.cuh"
#include "device/program_tags.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int PF_THREADS = 128;
constexpr int PF_HANDLERS = 8;     // at most: live any_of/all_of handlers per replication (pool, by serial)
constexpr int PF_QUEUE_CAP = 128;  // items per queue (global memory) unless cxxdes_b200_config.queue_capacity asks for more
constexpr uint32_t PF_NONE = 0xFFFFFFFFu;
constexpr int32_t INHERIT32 = INT32_MAX;
constexpr int64_t INHERIT64 = INT64_MAX;
constexpr uint32_t PF_PHASE_SAVED = 1, PF_PHASE_DONE = 2, PF_PHASE_REPLAY = 3;   // word 0 of a saved state; 0 = not started
constexpr int PF_MAX_OPS_PER_EVENT = 100000;  // a body that never suspends (same guard as program_kernel.cu)

// process word 0: pc | loop_depth << 16 | n_comp << 18 | bound << 20 | complete << 21
constexpr uint32_t P_PC = 0xFFFFu, P_DEPTH_SHIFT = 16, P_NCOMP_SHIFT = 18, P_BOUND = 1u << 20, P_COMPLETE = 1u << 21;
// handler word: serial+1 (24 bits) | remaining << 24 (6 bits) | is_all << 30 | armed << 31.  `total` is not kept:
// a handler only exists when no child was ready, so any_of fires on the first token and all_of on the last.
constexpr uint32_t H_SERIAL = 0xFFFFFFu, H_REM_SHIFT = 24, H_ALL = 1u << 30, H_ARMED = 1u << 31;

struct machine {
    environment env;
    const device_program &prg;
    const shard &sh;
    const program_fast_layout &L;
    const log_entry *table;
    uint32_t *W;    // this thread's column of 32-bit state words
    uint64_t *D;    // ... of 64-bit state words: f64[0..1], i64[0..1], then one register per process
    int64_t *items; // queue payloads in global memory, this thread's column
    uint32_t item_stride;
    stream_addr addr;
    uint32_t rate_base;   // parameter sweep: first rate of this replication's row

    __device__ machine(const device_program &p, const shard &s, const program_fast_layout &l, const log_entry *t)
        : prg(p), sh(s), L(l), table(t) {}

    __device__ __forceinline__ uint32_t &w(int i) { return W[i * PF_THREADS]; }
    __device__ __forceinline__ uint64_t &d(int i) { return D[i * PF_THREADS]; }
    __device__ __forceinline__ int pbase(uint32_t p) const { return L.w_procs + (int)p * L.pw; }
    __device__ __forceinline__ int64_t prio_of(uint32_t p) { return (int64_t)(int32_t)w(pbase(p) + 1); }

    __device__ void fresh(uint32_t p) {
        const int64_t pr = prg.procs[p].priority;
        w(pbase(p)) = prg.procs[p].entry;
        w(pbase(p) + 1) = (uint32_t)(pr == INHERIT64 ? INHERIT32 : (int32_t)pr);
    }

    // coroutine_data::bind_, environment.ipp:281-307
    __device__ void bind(uint32_t p, int64_t inherited) {
        const int b = pbase(p);
        const uint32_t w0 = w(b);
        if (w0 & P_BOUND) return;
        w(b) = w0 | P_BOUND;
        int64_t pr = (int64_t)(int32_t)w(b + 1);
        if (pr == INHERIT32) {
            pr = inherited;
            w(b + 1) = (uint32_t)(int32_t)pr;
        }
        env.schedule(env.now + prg.procs[p].latency, pr, ptag(p));
    }

    // coroutine::await_suspend, coroutine.ipp:194-207 (the completion token's priority and latency are
    // functions of the child alone, so only the tag is stored)
    __device__ void register_completion(uint32_t child, uint32_t tag) {
        const int b = pbase(child);
        const uint32_t w0 = w(b);
        const uint32_t k = (w0 >> P_NCOMP_SHIFT) & 3u;
        if (k >= 2) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        w(b + 2 + (int)k) = tag;
        w(b) = w0 + (1u << P_NCOMP_SHIFT);
    }

    // coroutine_data_::do_return, environment.ipp:321-348
    __device__ void do_return(uint32_t p) {
        const int b = pbase(p);
        const uint32_t w0 = w(b);
        const uint32_t n = (w0 >> P_NCOMP_SHIFT) & 3u;
        if (n) {
            const int64_t rp = prg.procs[p].return_priority;
            const int64_t pr = rp == INHERIT64 ? (int64_t)(int32_t)w(b + 1) : rp;
            const int64_t at = prg.procs[p].return_latency + env.now;
            for (uint32_t k = 0; k < n; ++k) env.schedule(at, pr, w(b + 2 + (int)k));
        }
        w(b) = (w0 & ~(3u << P_NCOMP_SHIFT)) | P_COMPLETE;
    }

    // waiters of a queue / semaphore / mutex: `while (!cond) co_await event_.wait()` -- latency 0, the waiting
    // process's priority (queue.hpp:48-65, semaphore.hpp:58-78, mutex.hpp:91-99).  list = count word + entries.
    __device__ void wait_plain(int list, uint32_t p) {
        const uint32_t n = w(list);
        if ((int)n >= L.wait_cap) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        w(list + 1 + (int)n) = p;
        w(list) = n + 1;
    }
    // wake_awaitable::await_ready, event.hpp:125-134
    __device__ void wake_plain(int list) {
        const uint32_t n = w(list);
        if (n == 0) return;
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t p = w(list + 1 + (int)k);
            env.schedule(env.now, prio_of(p), ptag(p));
        }
        w(list) = 0;
    }
    // sync::event proper: entries {process | priority, latency}  (event.hpp:87-139)
    __device__ void wait_event(int list, uint32_t p, int64_t latency, int64_t prio) {
        const uint32_t n = w(list);
        if ((int)n >= L.wait_cap || prio < -32768 || prio > 32767 || latency < 0 || latency > INT32_MAX) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        w(list + 1 + 2 * (int)n) = p | ((uint32_t)(prio + 32768) << 16);
        w(list + 2 + 2 * (int)n) = (uint32_t)(int32_t)latency;
        w(list) = n + 1;
    }
    __device__ void wake_event(int list) {
        const uint32_t n = w(list);
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t e = w(list + 1 + 2 * (int)k);
            const int64_t lat = (int64_t)(int32_t)w(list + 2 + 2 * (int)k);
            env.schedule(env.now + lat, (int64_t)(e >> 16) - 32768, ptag(e & 0xFFFFu));
        }
        w(list) = 0;
    }

    __device__ __forceinline__ int64_t op_priority(const cxxdes_b200_op &op, uint32_t self) {
        return op.b == CXXDES_B200_INHERIT ? prio_of(self) : (int64_t)op.b;
    }

    // any_of / all_of: any_of.ipp:41-92.  Returns true when the awaiting process suspends.
    __device__ bool composition(uint32_t self, const cxxdes_b200_op &head, uint32_t first_child) {
        const bool is_all = head.code == CXXDES_B200_OP_ALL_OF;
        const uint32_t total = head.a;
        // await_bind: every child in argument order (coroutine children get their start token here); what they
        // inherit is the composition's priority: .priority(p), else the awaiting process's (any_of.ipp:41-48)
        const int64_t self_prio = head.b == CXXDES_B200_INHERIT ? prio_of(self) : (int64_t)head.b;
        for (uint32_t c = 0; c < total; ++c) {
            const cxxdes_b200_op &ch = prg.ops[first_child + c];
            if (ch.code == CXXDES_B200_OP_CHILD_PROC) bind(ch.a, self_prio);
        }
        // await_ready
        uint32_t ready_mask = 0, remaining = total;
        for (uint32_t c = 0; c < total; ++c) {
            const cxxdes_b200_op &ch = prg.ops[first_child + c];
            bool ready = false;
            if (ch.code == CXXDES_B200_OP_CHILD_PROC) ready = (w(pbase(ch.a)) & P_COMPLETE) != 0;
            else if (ch.code == CXXDES_B200_OP_CHILD_UNTIL) ready = env.now >= ch.imm;
            if (ready) {
                ready_mask |= 1u << c;
                --remaining;
            }
        }
        if (is_all ? remaining == 0 : remaining < total) return false;
        // await_suspend: a new handler holds the output token; non-ready children in order
        const uint32_t serial = w(L.w_serial);
        w(L.w_serial) = serial + 1;
        if (serial + 1 > PTAG_MAX_SERIAL) {
            env.status |= CXXDES_B200_REP_BAD_PROGRAM;
            return true;
        }
        // storage: any handler that has fired (its late tokens are recognised by serial, device/program_tags.cuh)
        int hw = -1;
        for (int k = 0; k < L.n_handlers; ++k)
            if (!(w(L.w_handlers + k) & H_ARMED)) {
                hw = L.w_handlers + k;
                break;
            }
        if (hw < 0) {   // more live compositions than the pool holds
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return true;
        }
        if (head.imm < 0 || head.imm > INT32_MAX) {   // .return_latency() beyond the 32-bit word of this layout
            env.status |= CXXDES_B200_REP_TIME_RANGE;
            return true;
        }
        w(hw) = (serial + 1) | (remaining << H_REM_SHIFT) | (is_all ? H_ALL : 0u) | H_ARMED;
        w(hw + L.n_handlers) = (uint32_t)head.imm;   // .return_latency(): added to the satisfying child token's time
        const uint32_t tag = ptag(self, serial + 1);
        for (uint32_t c = 0; c < total; ++c) {
            if (ready_mask & (1u << c)) continue;
            const cxxdes_b200_op &ch = prg.ops[first_child + c];
            if (ch.code == CXXDES_B200_OP_CHILD_PROC) register_completion(ch.a, tag);
            else if (ch.code == CXXDES_B200_OP_CHILD_DELAY)
                env.schedule(env.now + ch.imm, ch.b == CXXDES_B200_INHERIT ? self_prio : (int64_t)ch.b, tag);
            else
                env.schedule(ch.imm, ch.b == CXXDES_B200_INHERIT ? self_prio : (int64_t)ch.b, tag);
        }
        return true;
    }

    // custom_handler::invoke, any_of.ipp:9-26
    __device__ void invoke_handler(const token &t) {
        const uint32_t hp1 = ptag_serial_plus1(t.tag);
        int hw = -1;
        uint32_t h = 0;
        for (int k = 0; k < L.n_handlers; ++k) {
            const uint32_t x = w(L.w_handlers + k);
            if ((x & H_SERIAL) == hp1) {
                hw = L.w_handlers + k;
                h = x;
            }
        }
        if (hw < 0 || !(h & H_ARMED)) return;  // the handler already fired (and its storage may have been reused)
        const uint32_t remaining = ((h >> H_REM_SHIFT) & 63u) - 1u;
        h = (h & ~(63u << H_REM_SHIFT)) | ((remaining & 63u) << H_REM_SHIFT);
        const bool cond = (h & H_ALL) ? remaining == 0 : true;
        if (cond) {
            env.schedule((int64_t)w(hw + L.n_handlers) + key_time(t.key), key_prio(t.key), ptag(ptag_target(t.tag)));
            h &= ~H_ARMED;
        }
        w(hw) = h;
    }

    // A timer a body ended on, noted for the converged part of the iteration.
    struct timer_note {
        uint32_t kind;   // 0 none, 1 delay in ticks, 2 exponential(rates[a]), 3 real seconds (bits)
        uint32_t a;
        int64_t value;   // ticks / bits of the double
        int64_t prio;
    };

    // One operation of process p (one step of coroutine_data::resume, coroutine_data.ipp:20-29).  w0 is the
    // process's word 0, kept in a register by the caller while the body runs.
    // Returns true while p keeps running, false when it suspended or returned.
    __device__ bool exec(uint32_t p, uint32_t &w0, timer_note &timer, int &jumps_left) {
        const int b = pbase(p);
        const uint32_t pc = w0 & P_PC;   // in range: program_check() proves that no path runs off the end of the program
        const cxxdes_b200_op op = prg.ops[pc];
        bool running = true;
        uint32_t next_pc = pc + 1;
        switch (op.code) {
            // the operations that touch process words of OTHER processes (or, in odd programs, this one)
            // go through shared memory: write w0 back first, read it again afterwards
            case CXXDES_B200_OP_RETURN:
                w(b) = w0;
                do_return(p);
                w0 = w(b);
                return false;
            case CXXDES_B200_OP_DELAY:
                timer.kind = 1;
                timer.value = op.imm;
                timer.prio = op_priority(op, p);
                running = false;
                break;
            case CXXDES_B200_OP_UNTIL:
                if (env.now < op.imm) {
                    timer.kind = 1;
                    timer.value = op.imm - env.now;
                    timer.prio = op_priority(op, p);
                    running = false;
                }
                break;
            case CXXDES_B200_OP_DELAY_EXP:
                timer.kind = 2;
                timer.a = op.a;
                timer.prio = op_priority(op, p);
                running = false;
                break;
            case CXXDES_B200_OP_DELAY_REAL:
                timer.kind = 3;
                timer.value = op.imm;
                timer.prio = op_priority(op, p);
                running = false;
                break;
            case CXXDES_B200_OP_FRESH:
                w(b) = w0;
                fresh(op.a);
                w0 = w(b);
                break;
            case CXXDES_B200_OP_AWAIT:
                w(b) = w0;
                bind(op.a, prio_of(p));
                if (!(w(pbase(op.a)) & P_COMPLETE)) {
                    register_completion(op.a, ptag(p));
                    running = false;
                }
                w0 = w(b);
                break;
            case CXXDES_B200_OP_ASYNC:
                w(b) = w0;
                bind(op.a, prio_of(p));
                if (!(w(pbase(op.a)) & P_COMPLETE)) register_completion(op.a, ptag(PTAG_NONE));
                w0 = w(b);
                break;
            case CXXDES_B200_OP_ALL_OF:
            case CXXDES_B200_OP_ANY_OF:
                next_pc = pc + 1 + op.a;
                w(b) = w0;
                running = !composition(p, op, pc + 1);
                w0 = w(b);
                break;
            case CXXDES_B200_OP_Q_PUT: {
                const int qw = L.w_queues + op.a;
                const uint32_t q = w(qw);
                const uint32_t head = q & 0xFFFFu, count = q >> 16;
                const uint64_t max_size = prg.queue_max[op.a];
                if (max_size != 0 && count >= max_size) {  // while (!can_put()) co_await event_.wait();
                    wait_plain(L.w_qwait + op.a * L.wait_words, p);
                    return false;   // pc unchanged: the put is retried after the wake-up
                }
                if (count >= (uint32_t)L.queue_cap) {
                    env.status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
                    return false;
                }
                const int64_t v = op.b ? (int64_t)d(L.d_regs + (int)p) : env.now;
                items[(uint32_t)(op.a * L.queue_cap + (int)((head + count) & (uint32_t)(L.queue_cap - 1))) * item_stride] = v;
                w(qw) = head | ((count + 1) << 16);
                wake_plain(L.w_qwait + op.a * L.wait_words);
                break;
            }
            case CXXDES_B200_OP_Q_POP: {
                const int qw = L.w_queues + op.a;
                const uint32_t q = w(qw);
                uint32_t head = q & 0xFFFFu, count = q >> 16;
                if (count == 0) {
                    wait_plain(L.w_qwait + op.a * L.wait_words, p);
                    return false;
                }
                d(L.d_regs + (int)p) = (uint64_t)items[(uint32_t)(op.a * L.queue_cap + (int)head) * item_stride];
                head = (head + 1) & (uint32_t)(L.queue_cap - 1);
                if (--count == 0) head = 0;   // an empty ring restarts at slot 0: short queues stay in one line
                w(qw) = head | (count << 16);
                wake_plain(L.w_qwait + op.a * L.wait_words);
                break;
            }
            case CXXDES_B200_OP_SEM_DOWN: {
                const uint32_t v = w(L.w_sems + op.a);
                if (v == 0) {
                    wait_plain(L.w_swait + op.a * L.wait_words, p);
                    return false;
                }
                w(L.w_sems + op.a) = v - 1;
                wake_plain(L.w_swait + op.a * L.wait_words);
                break;
            }
            case CXXDES_B200_OP_SEM_UP: {
                const uint32_t v = w(L.w_sems + op.a);
                if ((uint64_t)v >= prg.sem_max[op.a]) {
                    wait_plain(L.w_swait + op.a * L.wait_words, p);
                    return false;
                }
                if (v >= 0xFFFFFFFEu) {   // beyond the 32-bit counter of this layout
                    env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                    return false;
                }
                w(L.w_sems + op.a) = v + 1;
                wake_plain(L.w_swait + op.a * L.wait_words);
                break;
            }
            case CXXDES_B200_OP_MTX_ACQUIRE: {
                const uint32_t m = w(L.w_mtx);
                if (m & (1u << op.a)) {
                    wait_plain(L.w_mwait + op.a * L.wait_words, p);
                    return false;
                }
                w(L.w_mtx) = m | (1u << op.a);
                break;
            }
            case CXXDES_B200_OP_MTX_RELEASE:
                w(L.w_mtx) = w(L.w_mtx) & ~(1u << op.a);
                wake_plain(L.w_mwait + op.a * L.wait_words);
                break;
            case CXXDES_B200_OP_EV_WAIT:
                wait_event(L.w_ewait + op.a * L.ewait_words, p, op.imm, op_priority(op, p));
                running = false;
                break;
            case CXXDES_B200_OP_EV_WAKE:
                wake_event(L.w_ewait + op.a * L.ewait_words);
                break;
            case CXXDES_B200_OP_LOOP: {
                const uint32_t depth = (w0 >> P_DEPTH_SHIFT) & 3u;
                if (op.imm <= 0) {
                    next_pc = (uint32_t)op.b;
                } else if ((int)depth >= L.loop_words || op.imm > 0xFFFFFFFFLL) {
                    env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                    return false;
                } else {
                    w(b + L.o_loop + (int)depth) = (uint32_t)op.imm;
                    w0 += 1u << P_DEPTH_SHIFT;
                }
                break;
            }
            case CXXDES_B200_OP_END_LOOP: {
                const uint32_t depth = (w0 >> P_DEPTH_SHIFT) & 3u;
                if (depth == 0) {   // (program_check pairs the loops; never touch the word below the counters)
                    env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                    return false;
                }
                const int lw = b + L.o_loop + (int)depth - 1;
                const uint32_t left = w(lw) - 1;
                w(lw) = left;
                if (left > 0) {
                    next_pc = (uint32_t)op.b;
                    // a body that loops without ever suspending (same guard as program_kernel.cu; only a backward
                    // jump can keep a body running for more than n_ops operations)
                    if (--jumps_left == 0) env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                } else {
                    w0 -= 1u << P_DEPTH_SHIFT;
                }
                break;
            }
            case CXXDES_B200_OP_SET_REG_NOW:
                d(L.d_regs + (int)p) = (uint64_t)env.now;
                break;
            case CXXDES_B200_OP_LAZY_TIMEOUT:   // timeout.ipp:106-163: the deadline is fixed, nothing is scheduled
                d(L.d_regs + (int)p) = (uint64_t)(env.now + op.imm);
                break;
            case CXXDES_B200_OP_UNTIL_REG: {
                const int64_t at = (int64_t)d(L.d_regs + (int)p);
                if (env.now < at) {
                    timer.kind = 1;
                    timer.value = at - env.now;
                    timer.prio = op_priority(op, p);
                    running = false;
                }
                break;
            }
            case CXXDES_B200_OP_STAT_SECONDS: {
                const double x = bits_as_double(d(op.a));
                d(op.a) = double_bits(x + (sim_to_seconds(env.now, sh.clk) - sim_to_seconds((int64_t)d(L.d_regs + (int)p), sh.clk)));
                break;
            }
            case CXXDES_B200_OP_STAT_TICKS: {
                const double x = bits_as_double(d(op.a));
                d(op.a) = double_bits(x + (double)(env.now - (int64_t)d(L.d_regs + (int)p)));
                break;
            }
            case CXXDES_B200_OP_COUNT:
                d(CXXDES_B200_N_F64 + op.a) += (uint64_t)op.imm;
                break;
            case CXXDES_B200_OP_IF: {   // plain C++ between the awaits: a read of the model's own state, no token
                const uint32_t idx = op.a & 0xFFu, src = (op.a >> 8) & 0xFu, cmp = (op.a >> 12) & 0x7u;
                int64_t v;
                if (src == CXXDES_B200_SRC_COUNTER) v = (int64_t)d(CXXDES_B200_N_F64 + (int)idx);
                else if (src == CXXDES_B200_SRC_QUEUE_SIZE) v = (int64_t)(w(L.w_queues + (int)idx) >> 16);
                else if (src == CXXDES_B200_SRC_SEM_VALUE) v = (int64_t)w(L.w_sems + (int)idx);
                else if (src == CXXDES_B200_SRC_NOW) v = env.now;
                else if (src == CXXDES_B200_SRC_REGISTER) v = (int64_t)d(L.d_regs + (int)p);
                else v = (w(L.w_mtx) >> idx) & 1u;
                const bool taken = cmp == CXXDES_B200_CMP_EQ ? v == op.imm : cmp == CXXDES_B200_CMP_NE ? v != op.imm
                                 : cmp == CXXDES_B200_CMP_LT ? v < op.imm : cmp == CXXDES_B200_CMP_LE ? v <= op.imm
                                 : cmp == CXXDES_B200_CMP_GT ? v > op.imm : v >= op.imm;
                if (!taken) next_pc = (uint32_t)op.b;
                break;
            }
            case CXXDES_B200_OP_JUMP:
                next_pc = (uint32_t)op.b;
                break;
            case CXXDES_B200_OP_MARK:
                d(CXXDES_B200_N_F64 + 0) += 1;
                d(CXXDES_B200_N_F64 + 1) = d(CXXDES_B200_N_F64 + 1) * 1000003ULL + (uint64_t)(env.now * op.imm + (int64_t)op.b);
                break;
            default:
                env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                return false;
        }
        w0 = (w0 & ~P_PC) | (next_pc & P_PC);
        return running;
    }

    __device__ static __forceinline__ uint64_t double_bits(double x) { return (uint64_t)__double_as_longlong(x); }

    __device__ void start(uint64_t rep) {
        env.now = 0;
        env.n_events = 0;
        env.hash = TRACE_HASH_INIT;
        env.status = 0;
        env.n = 0;
        addr.seed = sh.seed;
        addr.replication = sh.first_replication + rep;
        rate_base = prg.sweep_steps ? (uint32_t)((sh.first_replication + rep) % prg.sweep_steps) * prg.n_rates : 0u;
        for (int k = 0; k < L.n_words; ++k) w(k) = 0;
        for (int k = 0; k < L.n_dwords; ++k) d(k) = 0;
        for (uint32_t p = 0; p < prg.n_procs; ++p) fresh(p);
        for (uint32_t s = 0; s < prg.n_sems; ++s) w(L.w_sems + (int)s) = (uint32_t)prg.sem_init[s];
        // environment::bind(coroutine) for every root, coroutine.ipp:352-357
        for (uint32_t r = 0; r < prg.n_roots; ++r) {
            const uint32_t p = prg.roots[r];
            bind(p, 0);
            if (!(w(pbase(p)) & P_COMPLETE)) register_completion(p, ptag(PTAG_NONE));
        }
    }

    // ---- run_until / run_for in pieces (environment.ipp:190-214): a replication that stops at the horizon
    // with tokens still scheduled is written to global memory and the next launch continues from it.
    // Saved state, `[word][replication]`: 8 header words {phase, heap size, now, n_events, digest}, the 32-bit
    // state words, the 64-bit ones, the live heap entries; queue payloads `[queue][slot][replication]`.
    __device__ __forceinline__ int heap_words_at() const { return 8 + L.n_words + 2 * L.n_dwords; }

    __device__ void save(const program_fast_state &st, uint64_t rep) {
        uint32_t *s = st.words + rep;
        const size_t n = st.n_replications;
        s[1 * n] = (uint32_t)env.n;
        s[2 * n] = (uint32_t)env.now;
        s[3 * n] = (uint32_t)((uint64_t)env.now >> 32);
        s[4 * n] = (uint32_t)env.n_events;
        s[5 * n] = (uint32_t)(env.n_events >> 32);
        s[6 * n] = (uint32_t)env.hash;
        s[7 * n] = (uint32_t)(env.hash >> 32);
        for (int k = 0; k < L.n_words; ++k) s[(size_t)(8 + k) * n] = w(k);
        for (int k = 0; k < L.n_dwords; ++k) {
            const uint64_t v = d(k);
            s[(size_t)(8 + L.n_words + 2 * k) * n] = (uint32_t)v;
            s[(size_t)(8 + L.n_words + 2 * k + 1) * n] = (uint32_t)(v >> 32);
        }
        const int hw = heap_words_at();
        for (int k = 0; k < env.n; ++k) {
            const uint64_t key = env.hkey[k];
            s[(size_t)(hw + 3 * k) * n] = (uint32_t)key;
            s[(size_t)(hw + 3 * k + 1) * n] = (uint32_t)(key >> 32);
            s[(size_t)(hw + 3 * k + 2) * n] = env.htag[k];
        }
        for (uint32_t q = 0; q < prg.n_queues; ++q) {
            const uint32_t qw = w(L.w_queues + (int)q);
            const uint32_t head = qw & 0xFFFFu, count = qw >> 16;
            for (uint32_t k = 0; k < count; ++k) {
                const size_t slot = (size_t)q * L.queue_cap + ((head + k) & (uint32_t)(L.queue_cap - 1));
                st.queue_items[slot * n + rep] = items[slot * item_stride];
            }
        }
        s[0] = PF_PHASE_SAVED;
    }

    __device__ void restore(const program_fast_state &st, uint64_t rep) {
        const uint32_t *s = st.words + rep;
        const size_t n = st.n_replications;
        env.n = (int)s[1 * n];
        env.now = (int64_t)((uint64_t)s[2 * n] | ((uint64_t)s[3 * n] << 32));
        env.n_events = (uint64_t)s[4 * n] | ((uint64_t)s[5 * n] << 32);
        env.hash = (uint64_t)s[6 * n] | ((uint64_t)s[7 * n] << 32);
        env.status = 0;
        addr.seed = sh.seed;
        addr.replication = sh.first_replication + rep;
        rate_base = prg.sweep_steps ? (uint32_t)((sh.first_replication + rep) % prg.sweep_steps) * prg.n_rates : 0u;
        for (int k = 0; k < L.n_words; ++k) w(k) = s[(size_t)(8 + k) * n];
        for (int k = 0; k < L.n_dwords; ++k)
            d(k) = (uint64_t)s[(size_t)(8 + L.n_words + 2 * k) * n] | ((uint64_t)s[(size_t)(8 + L.n_words + 2 * k + 1) * n] << 32);
        const int hw = heap_words_at();
        for (int k = 0; k < env.n; ++k) {
            env.hkey[k] = (uint64_t)s[(size_t)(hw + 3 * k) * n] | ((uint64_t)s[(size_t)(hw + 3 * k + 1) * n] << 32);
            env.htag[k] = s[(size_t)(hw + 3 * k + 2) * n];
        }
        for (uint32_t q = 0; q < prg.n_queues; ++q) {
            const uint32_t qw = w(L.w_queues + (int)q);
            const uint32_t head = qw & 0xFFFFu, count = qw >> 16;
            for (uint32_t k = 0; k < count; ++k) {
                const size_t slot = (size_t)q * L.queue_cap + ((head + k) & (uint32_t)(L.queue_cap - 1));
                items[slot * item_stride] = st.queue_items[slot * n + rep];
            }
        }
    }
};

}  // namespace

// STATEFUL: the run_until / run_for variant that keeps replications in global memory between launches (a separate
// instantiation, so that the plain run() path carries none of it).
template <bool STATEFUL>
__global__ void __launch_bounds__(PF_THREADS)
program_fast_kernel(device_program prg_in, shard sh, device_columns out, program_fast_layout L, int64_t *queue_items,
                    program_fast_state st) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    device_program prg = prg_in;
    {
        // one contiguous blob on the device (c_api.cu upload_program): ops first, 16-byte aligned pieces
        const uint4 *src = (const uint4 *)prg_in.ops;
        uint4 *dst = (uint4 *)cursor;
        for (uint32_t k = threadIdx.x; k < L.staged_bytes / 16; k += blockDim.x) dst[k] = src[k];
        const unsigned char *base = (const unsigned char *)prg_in.ops;
        auto moved = [&](const void *ptr) { return (const void *)(cursor + ((const unsigned char *)ptr - base)); };
        prg.ops = (const cxxdes_b200_op *)moved(prg_in.ops);
        prg.procs = (const cxxdes_b200_process *)moved(prg_in.procs);
        prg.roots = (const uint32_t *)moved(prg_in.roots);
        prg.queue_max = (const uint64_t *)moved(prg_in.queue_max);
        prg.sem_init = (const uint64_t *)moved(prg_in.sem_init);
        prg.sem_max = (const uint64_t *)moved(prg_in.sem_max);
        prg.rates = (const divisor *)moved(prg_in.rates);
        cursor += L.staged_bytes;
    }
    stage_log_table(table, sh.log_table);   // ends with __syncthreads(): the staged program is visible too
    lane_array<uint64_t> hkey = carve<uint64_t>(cursor, L.heap_cap);
    lane_array<uint64_t> dwords = carve<uint64_t>(cursor, L.n_dwords);
    lane_array<uint32_t> htag = carve<uint32_t>(cursor, L.heap_cap);
    lane_array<uint32_t> words = carve<uint32_t>(cursor, L.n_words);

    machine m(prg, sh, L, table);
    hkey.stride = PF_THREADS;   // compile-time stride: heap slot addresses become shifts folded into the LDS/STS
    htag.stride = PF_THREADS;
    m.env.init(hkey, htag, L.heap_cap);
    m.W = words.base;
    m.D = dwords.base;
    m.item_stride = gridDim.x * PF_THREADS;
    m.items = queue_items + (size_t)blockIdx.x * PF_THREADS + threadIdx.x;
    environment &env = m.env;

    bool active = false;
    uint64_t rep = 0;
    for (;;) {
        if (!active) {
            if (!steal(sh, rep)) break;
            if (STATEFUL) {
                const uint32_t phase = st.resume ? st.words[rep] : 0u;
                if (phase == PF_PHASE_DONE) {   // finished in an earlier piece: now_ = max(now_, t), environment.ipp:194
                    if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) out.end_time[rep] = sh.run_until;
                    continue;
                }
                if (phase == PF_PHASE_REPLAY) {   // left this layout earlier: the full-size kernel replays it from tick 0
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                    continue;
                }
                if (phase == PF_PHASE_SAVED) m.restore(st, rep);
                else m.start(rep);
            } else {
                m.start(rep);
            }
            active = true;
        }
        const bool more = !env.status && env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until);
        if (!more) {
            if (env.status) {   // this layout could not hold the replication: the full-size kernel redoes it
                const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                sh.retry_list[at] = (uint32_t)rep;
                if (STATEFUL) st.words[rep] = PF_PHASE_REPLAY;
            } else {
                if (sh.run_until >= 0) {
                    if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
                    env.now = env.now > sh.run_until ? env.now : sh.run_until;
                }
                if (STATEFUL) {
                    if (env.status & CXXDES_B200_REP_UNFINISHED) m.save(st, rep);
                    else st.words[rep] = PF_PHASE_DONE;
                }
                rep_out o;
                o.end_time = env.now;
                o.n_events = env.n_events;
                o.trace_hash = env.hash;
                o.status = env.status;
                for (int k = 0; k < CXXDES_B200_N_F64; ++k) o.f64[k] = bits_as_double(m.d(k));
                for (int k = 0; k < CXXDES_B200_N_I64; ++k) o.i64[k] = (int64_t)m.d(CXXDES_B200_N_F64 + k);
                store_result(out, rep, o);
            }
            active = false;
            continue;
        }
        // environment::step, environment.ipp:117-146 (the digest names processes by ordinal)
        const token t = env.pop_top();
        const int64_t time = key_time(t.key);
        env.now = time > env.now ? time : env.now;
        ++env.n_events;
        const uint32_t target = ptag_target(t.tag);
        const uint32_t hp1 = ptag_serial_plus1(t.tag);
        const uint32_t kind = hp1 ? TOKEN_HANDLER : (target != PTAG_NONE ? TOKEN_RESUME : TOKEN_NOOP);
        const uint32_t ordinal = target != PTAG_NONE ? prg.procs[target].ordinal : PROC_NONE;
        env.hash = trace_hash_step(env.hash, time, key_prio(t.key), kind, ordinal);
        machine::timer_note timer;
        timer.kind = 0;
        if (hp1) {
            m.invoke_handler(t);
        } else if (target != PTAG_NONE) {
            // coroutine_data::resume: run the body until it suspends or returns
            const int b = m.pbase(target);
            uint32_t w0 = m.w(b);
            int jumps_left = PF_MAX_OPS_PER_EVENT;
            while (m.exec(target, w0, timer, jumps_left)) {
                if (env.status) break;
            }
            m.w(b) = w0;
        }
        // the timer the body ended on, for all lanes of the warp at once
        if (timer.kind) {
            int64_t ticks = timer.value;
            if (timer.kind >= 2) {
                double v;
                if (timer.kind == 2) {
                    const int b = m.pbase(target);
                    const uint32_t draw = m.w(b + L.o_draws);
                    m.w(b + L.o_draws) = draw + 1;
                    m.addr.stream = prg.procs[target].ordinal;
                    v = exponential(m.addr, (uint64_t)draw, prg.rates[m.rate_base + timer.a], table);
                } else {
                    v = bits_as_double((uint64_t)timer.value);
                }
                ticks = real_to_sim(v, sh.clk);
            }
            env.schedule(env.now + ticks, timer.prio, ptag(target));
        }
    }
}

// Size the shared-memory state from the (host copy of the) program.  valid = false: run program_kernel.cu alone.
program_fast_layout program_fast_plan(const cxxdes_b200_program &p, uint32_t blob_bytes, size_t smem_optin,
                                      uint32_t queue_capacity) {
    program_fast_layout L{};
    L.valid = false;
    if (getenv("CXXDES_B200_PRG_NO_FAST")) return L;
    if (blob_bytes > 16 * 1024) return L;
    if (p.n_procs > 72) return L;     // (more processes than this layout was built and measured for: the last pass of program_kernel.cu)
    bool uses_reg = false, uses_draws = false;
    int depth = 0, max_depth = 0, max_timers = 0, n_compositions = 0;
    for (uint32_t i = 0; i < p.n_ops; ++i) {
        const cxxdes_b200_op &op = p.ops[i];
        switch (op.code) {
            case CXXDES_B200_OP_Q_POP: case CXXDES_B200_OP_SET_REG_NOW: case CXXDES_B200_OP_STAT_SECONDS:
            case CXXDES_B200_OP_STAT_TICKS: case CXXDES_B200_OP_LAZY_TIMEOUT: case CXXDES_B200_OP_UNTIL_REG:
                uses_reg = true;
                break;
            case CXXDES_B200_OP_Q_PUT:
                if (op.b) uses_reg = true;
                break;
            case CXXDES_B200_OP_IF:
                if (((op.a >> 8) & 0xFu) == CXXDES_B200_SRC_REGISTER) uses_reg = true;
                break;
            case CXXDES_B200_OP_CHILD_ALL_OF: case CXXDES_B200_OP_CHILD_ANY_OF: case CXXDES_B200_OP_CHILD_EV_WAIT:
                return L;   // nested compositions, event waits inside compositions: the full-size kernel alone
            case CXXDES_B200_OP_DELAY_EXP:
                uses_draws = true;
                break;
            case CXXDES_B200_OP_LOOP:
                if (++depth > max_depth) max_depth = depth;
                break;
            case CXXDES_B200_OP_END_LOOP:
                if (depth > 0) --depth;
                break;
            case CXXDES_B200_OP_RETURN:
                depth = 0;
                break;
            case CXXDES_B200_OP_ALL_OF: case CXXDES_B200_OP_ANY_OF: {
                int timers = 0;
                if (op.a > 32) return L;   // (the ready flags of the children are one 32-bit word here)
                ++n_compositions;
                for (uint32_t c = 1; c <= op.a && i + c < p.n_ops; ++c)
                    if (p.ops[i + c].code != CXXDES_B200_OP_CHILD_PROC) ++timers;
                if (timers > max_timers) max_timers = timers;
                break;
            }
            default:
                break;
        }
    }
    if (max_depth > 2) max_depth = 2;   // deeper nesting is rejected at run time by either kernel
    for (uint32_t i = 0; i < p.n_procs; ++i) {
        const int64_t pr = p.procs[i].priority;
        if (pr != INT64_MAX && (pr < INT32_MIN || pr >= INT32_MAX)) return L;
    }
    for (uint32_t s = 0; s < p.n_sems; ++s)
        if (p.sem_init[s] > 0xFFFFFFFEull) return L;
    if (p.n_mutexes > 32) return L;

    // one resume / start token per process, the roots' completion tokens, timers of compositions (live + stale)
    L.heap_cap = (int)p.n_procs + 2 * max_timers + 2;
    if (L.heap_cap < 4) L.heap_cap = 4;
    if (L.heap_cap > 64) L.heap_cap = 64;
    // a process waits in one composition at a time, and a handler that has fired gives its storage back at once
    L.n_handlers = n_compositions < (int)p.n_procs ? n_compositions : (int)p.n_procs;
    if (L.n_handlers < 1) L.n_handlers = 1;
    if (L.n_handlers > PF_HANDLERS) L.n_handlers = PF_HANDLERS;
    L.wait_cap = (int)p.n_procs;
    L.wait_words = 1 + L.wait_cap;
    L.ewait_words = 1 + 2 * L.wait_cap;
    L.loop_words = max_depth;
    int pw = 4;
    L.o_draws = pw; pw += uses_draws ? 1 : 0;
    L.o_loop = pw; pw += max_depth;
    L.pw = pw;
    int w = 0;
    L.w_serial = w++;
    L.w_mtx = w++;
    L.w_handlers = w; w += 2 * L.n_handlers;   // handler word + return latency
    L.w_queues = w; w += (int)p.n_queues;
    L.w_sems = w; w += (int)p.n_sems;
    L.w_procs = w; w += pw * (int)p.n_procs;
    L.w_qwait = w; w += L.wait_words * (int)p.n_queues;
    L.w_swait = w; w += L.wait_words * (int)p.n_sems;
    L.w_mwait = w; w += L.wait_words * (int)p.n_mutexes;
    L.w_ewait = w; w += L.ewait_words * (int)p.n_events;
    L.n_words = w;
    L.d_regs = CXXDES_B200_N_F64 + CXXDES_B200_N_I64;
    L.n_dwords = L.d_regs + (uses_reg ? (int)p.n_procs : 0);
    if (!uses_reg) L.d_regs = 0;   // never read
    L.staged_bytes = (blob_bytes + 15u) & ~15u;
    const size_t per_thread = (size_t)L.heap_cap * 12 + (size_t)L.n_dwords * 8 + (size_t)L.n_words * 4;
    L.smem_bytes = sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 + L.staged_bytes + per_thread * PF_THREADS;
    if (L.smem_bytes > smem_optin) return L;
    int per_sm = (int)((227 * 1024) / (L.smem_bytes + 1024));
    if (per_sm > 12) per_sm = 12;   // 1536 threads of <= 42 registers
    if (const char *env = getenv("CXXDES_B200_PRG_BLOCKS")) { int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
    if (per_sm < 2) return L;       // too few resident warps to be worth it: the full-size kernel alone
    L.blocks_per_sm = per_sm;
    // queue payloads live in global memory, so a heavily loaded model can ask for deeper queues than the
    // full-size kernel's 128 (power of two, at most 32768: head and count share one 32-bit word)
    L.queue_cap = PF_QUEUE_CAP;
    while (L.queue_cap < (int)queue_capacity && L.queue_cap < 32768) L.queue_cap *= 2;
    L.queue_item_bytes_per_thread = (size_t)p.n_queues * L.queue_cap * sizeof(int64_t);
    L.valid = true;
    return L;
}

size_t program_fast_queue_bytes(const program_fast_layout &L, int sm_count) {
    if (!L.valid) return 0;
    return L.queue_item_bytes_per_thread * (size_t)sm_count * L.blocks_per_sm * PF_THREADS;
}

size_t program_fast_state_words(const program_fast_layout &L) {
    return 8 + (size_t)L.n_words + 2 * (size_t)L.n_dwords + 3 * (size_t)L.heap_cap;
}
size_t program_fast_state_queue_bytes(const program_fast_layout &L, uint64_t n_replications) {
    return L.queue_item_bytes_per_thread * n_replications;
}

cudaError_t launch_program_fast(const device_program &prg, const program_fast_layout &L, const shard &sh,
                                const device_columns &out, int64_t *queue_items, const program_fast_state &st,
                                int sm_count, cudaStream_t stream) {
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<sm_count * L.blocks_per_sm, PF_THREADS, L.smem_bytes, stream>>>(prg, sh, out, L, queue_items, st);
        return cudaGetLastError();
    };
    return st.words ? go(program_fast_kernel<true>) : go(program_fast_kernel<false>);
}

}  // namespace cxb
