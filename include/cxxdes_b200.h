/* cxxdes_b200.h -- C ABI of the B200 ensemble discrete-event engine.
 *
 * libcxxdes (the reference) is a header-only C++20 library with no plugin or
 * FFI layer; its only boundary is the C++ API
 *     CXXDES_SIMULATION(model) { coroutine<> co_main(); };  model.run();
 * This header is the C-ABI a binding for that API's *batched* run path calls:
 * plain pointers and sizes, no C++/torch types, no exceptions.  The C++ mirror
 * of the reference vocabulary (ensemble<Model>::run / run_for / now / ...)
 * lives in include/cxxdes_b200/ensemble.hpp and is a thin wrapper over these
 * functions.  Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Threading: an ensemble handle is not thread-safe (like environment,
 * include/cxxdes/core/impl/environment.ipp:9-10); different handles may be
 * driven from different threads / processes (one per GPU).
 */
#ifndef CXXDES_B200_H_INCLUDED
#define CXXDES_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (the reference throws std::runtime_error; see cxxdes_b200_last_error) ---- */
enum {
    CXXDES_B200_OK = 0,
    CXXDES_B200_ERR_INVALID = 1,   /* bad argument / unsupported configuration            */
    CXXDES_B200_ERR_CUDA = 2,      /* CUDA runtime failure (message in last_error)         */
    CXXDES_B200_ERR_STATE = 3,     /* call not valid in this state (e.g. reconfigure after use,
                                      environment.ipp:43-65 throws in the reference)       */
    CXXDES_B200_ERR_CAPACITY = 4   /* a replication exceeded a device-side capacity; its
                                      status word says which (never silent corruption)     */
};

/* ---- time units: cxxdes::time_utils::time_unit_type, include/cxxdes/misc/time.hpp:18-25 ---- */
enum {
    CXXDES_B200_SECONDS = 0,
    CXXDES_B200_MILLISECONDS = 1,
    CXXDES_B200_MICROSECONDS = 2,
    CXXDES_B200_NANOSECONDS = 3,
    CXXDES_B200_PICOSECONDS = 4
};

/* env.time_unit(x) / env.time_precision(x): environment.ipp:43-70.
 * A time quantity is {t, u} exactly like cxxdes::core::time (misc/time.hpp:31-37). */
typedef struct cxxdes_b200_clock {
    int64_t unit_t;
    int32_t unit_u;
    int32_t precision_u;
    int64_t precision_t;
} cxxdes_b200_clock;

/* ---- models (process bodies lowered to device state machines) ---- */
enum {
    CXXDES_B200_MODEL_MM1 = 1,     /* examples/producer_consumer.cpp:9-59                  */
    CXXDES_B200_MODEL_ALOHA = 2,   /* examples/aloha.cpp:25-85                             */
    CXXDES_B200_MODEL_MMC = 3,     /* resource + any_of balking (examples/resource.cpp pattern) */
    CXXDES_B200_MODEL_ARCHSIM = 4, /* examples/basic_arch_sim.cpp:13-110                   */
    CXXDES_B200_MODEL_PROGRAM = 5  /* any model given as a process program (see below)     */
};

/* producer_consumer_example(lambda, mu, n_packets): examples/producer_consumer.cpp:10-20 */
typedef struct cxxdes_b200_mm1_params {
    double lambda;
    double mu;
    uint64_t n_packets;
} cxxdes_b200_mm1_params;

/* aloha_config: examples/aloha.cpp:13-18; the per-replication offered-load sweep
 * reproduces main() (examples/aloha.cpp:87-103): replication r uses step
 * i = r % sweep_steps, lambda_i = begin + (end-begin)*i/(steps-1) in float,
 * packets_per_station = size_t(lambda_i * packets_scale).  sweep_steps <= 1 keeps
 * lambda_begin and packets_per_station for every replication. */
typedef struct cxxdes_b200_aloha_params {
    uint32_t num_stations;
    uint32_t sweep_steps;
    float lambda_begin;
    float lambda_end;
    float frame_time;
    float packets_scale;
    uint64_t packets_per_station;
} cxxdes_b200_aloha_params;

/* M/M/c with balking customers built from sync::resource, async and any_of
 * (SURVEY Appendix E model; primitives include/cxxdes/sync/resource.hpp:37-100,
 * core/impl/any_of.ipp, core/impl/async.ipp). */
typedef struct cxxdes_b200_mmc_params {
    double lambda;
    double mu;
    double patience;
    uint32_t servers;
    uint32_t reserved;
    uint64_t n_customers;
} cxxdes_b200_mmc_params;

/* memory-hierarchy model, examples/basic_arch_sim.cpp:83-109: mem2{latency, capacity}
 * backed by mem1{latency}; n_loads addresses drawn as philox % addr_space,
 * n_requesters processes share the hierarchy. */
typedef struct cxxdes_b200_archsim_params {
    int64_t l2_latency;
    int64_t l1_latency;
    uint32_t l2_capacity;
    uint32_t addr_space;
    uint32_t n_requesters;
    uint32_t reserved;
    uint64_t n_loads;
} cxxdes_b200_archsim_params;

/* ---- process programs: generic lowering of coroutine bodies -------------------------------
 * A model that has no hand-written kernel is given as a *process program*: a static table of
 * processes (coroutine<> objects: ordinal for the trace digest, .priority(), .latency(),
 * .return_latency(), .return_priority() -- coroutine.ipp:106-176) and, per process, a straight
 * list of operations, each the lowering of one reference awaitable.  One device kernel interprets
 * the program for every replication.  Operation semantics (what is pushed, in which order) follow
 * SURVEY 3.3; the citations name the reference code each one reproduces.                        */
enum {
    CXXDES_B200_OP_RETURN = 0,      /* co_return: completion tokens released (environment.ipp:321-348)        */
    CXXDES_B200_OP_DELAY = 1,       /* co_await delay(imm) / timeout(imm) / yield(): timeout.ipp:21-32,180-182;
                                       b = explicit priority, or CXXDES_B200_INHERIT                          */
    CXXDES_B200_OP_UNTIL = 2,       /* co_await until(imm): ready if now >= imm (timeout.ipp:14-29)           */
    CXXDES_B200_OP_DELAY_EXP = 3,   /* co_await env.timeout(exponential(rates[a])): stream = process ordinal  */
    CXXDES_B200_OP_DELAY_REAL = 4,  /* co_await env.timeout(double): imm holds the bits of the double         */
    CXXDES_B200_OP_FRESH = 5,       /* process a becomes a new, unstarted coroutine object again               */
    CXXDES_B200_OP_AWAIT = 6,       /* co_await process a (coroutine.ipp:179-207, environment.ipp:281-307)     */
    CXXDES_B200_OP_ASYNC = 7,       /* co_await async(process a): async.ipp:21-28                              */
    CXXDES_B200_OP_ALL_OF = 8,      /* co_await all_of(children...): a = child count, children follow;
                                       b = .priority(p) or INHERIT (what the children inherit), imm = .return_latency() */
    CXXDES_B200_OP_ANY_OF = 9,      /* co_await any_of(children...)            (any_of.ipp:3-92,233-247)       */
    CXXDES_B200_OP_CHILD_PROC = 10, /*   child: process a                                                      */
    CXXDES_B200_OP_CHILD_DELAY = 11,/*   child: delay(imm), b = priority or INHERIT                            */
    CXXDES_B200_OP_CHILD_UNTIL = 12,/*   child: until(imm)                                                     */
    CXXDES_B200_OP_Q_PUT = 13,      /* co_await queue[a].put(now or reg): sync/queue.hpp:48-53; b=1 puts reg   */
    CXXDES_B200_OP_Q_POP = 14,      /* reg = co_await queue[a].pop(): sync/queue.hpp:58-65                     */
    CXXDES_B200_OP_SEM_DOWN = 15,   /* co_await semaphore[a].down(): sync/semaphore.hpp:70-78                  */
    CXXDES_B200_OP_SEM_UP = 16,     /* co_await semaphore[a].up(): sync/semaphore.hpp:58-66                    */
    CXXDES_B200_OP_MTX_ACQUIRE = 17,/* co_await mutex[a].acquire(): sync/mutex.hpp:91-99                       */
    CXXDES_B200_OP_MTX_RELEASE = 18,/* co_await handle.release(): sync/mutex.hpp:70-79                         */
    CXXDES_B200_OP_EV_WAIT = 19,    /* co_await event[a].wait(imm latency, b priority): sync/event.hpp:136-139 */
    CXXDES_B200_OP_EV_WAKE = 20,    /* co_await event[a].wake(): sync/event.hpp:125-134                        */
    CXXDES_B200_OP_LOOP = 21,       /* for (k = 0; k < imm; ++k) {: b = index of the op after the matching END */
    CXXDES_B200_OP_END_LOOP = 22,   /* }: b = index of the first op of the body                                */
    CXXDES_B200_OP_SET_REG_NOW = 23,/* reg = now()                                                              */
    CXXDES_B200_OP_STAT_SECONDS = 24,/* f64[a] += now_seconds() - seconds(reg)   (producer_consumer.cpp:52)    */
    CXXDES_B200_OP_STAT_TICKS = 25, /* f64[a] += (double)(now() - reg)                                          */
    CXXDES_B200_OP_COUNT = 26,      /* i64[a] += imm                                                             */
    CXXDES_B200_OP_MARK = 27,       /* observation point (the prints of the reference examples): ++i64[0];
                                       i64[1] = i64[1] * 1000003 + (now() * imm + b)   (wrapping)              */
    CXXDES_B200_OP_LAZY_TIMEOUT = 28,/* reg = now() + imm: `auto t = co_await lazy_timeout(imm)` fixes an absolute
                                       deadline without suspending (timeout.ipp:104-172)                        */
    CXXDES_B200_OP_UNTIL_REG = 29,  /* co_await t / co_await instant(reg): until(reg), b = priority or INHERIT    */
    CXXDES_B200_OP_IF = 30,         /* if (!(value cmp imm)) goto b: structured if / else on the model's own state.
                                       a = index | source << 8 | comparison << 12 (CXXDES_B200_SRC_*, _CMP_*);
                                       b = first op after the then-block (forward)                               */
    CXXDES_B200_OP_JUMP = 31,       /* goto b (forward): the end of a then-block jumps over the else-block       */
    CXXDES_B200_OP_CHILD_ALL_OF = 32,/*   child: a nested all_of -- `(a && b) || c`, `a && b && c` = all_of(all_of(a, b), c)
                                       in the reference (compositions.ipp); a = its own child count, its children
                                       follow in place (pre-order), b / imm as for ALL_OF.  When the inner condition
                                       holds, its output token goes to the OUTER handler (any_of.ipp:66-84)        */
    CXXDES_B200_OP_CHILD_ANY_OF = 33,/*   child: a nested any_of                                                   */
    CXXDES_B200_OP_CHILD_EV_WAIT = 34/*   child: event[a].wait(imm latency, b priority) -- `evt.wait() || timeout(t)`, waiting
                                       for a signal with a time limit (sync/event.hpp:27-74,136-139): the waiter token
                                       carries the composition's handler                                           */
};
enum {  /* what an IF looks at: plain reads of public state, no events */
    CXXDES_B200_SRC_COUNTER = 0,    /* i64[index]                                        */
    CXXDES_B200_SRC_QUEUE_SIZE = 1, /* queue[index].size()        (sync/queue.hpp:68-70) */
    CXXDES_B200_SRC_SEM_VALUE = 2,  /* semaphore[index].value()   (semaphore.hpp:52-54)  */
    CXXDES_B200_SRC_NOW = 3,        /* now()                                             */
    CXXDES_B200_SRC_REGISTER = 4,   /* the process register                              */
    CXXDES_B200_SRC_MUTEX = 5       /* mutex[index].is_acquired() (mutex.hpp:101-104)    */
};
enum { CXXDES_B200_CMP_EQ = 0, CXXDES_B200_CMP_NE, CXXDES_B200_CMP_LT, CXXDES_B200_CMP_LE, CXXDES_B200_CMP_GT, CXXDES_B200_CMP_GE };
#define CXXDES_B200_INHERIT INT32_MAX /* priority_consts::inherit (defs.ipp:35) in the 32-bit b field */

typedef struct cxxdes_b200_op {
    uint16_t code;
    uint16_t a;
    int32_t b;
    int64_t imm;
} cxxdes_b200_op;

typedef struct cxxdes_b200_process {
    uint32_t entry;          /* index of the first op of the body                                   */
    uint32_t ordinal;        /* name in the trace digest and Philox stream id (< 2^16)             */
    int64_t priority;        /* coroutine::priority(x); INT64_MAX = inherit from the awaiter        */
    int64_t latency;         /* coroutine::latency(x): start token at now + latency                 */
    int64_t return_latency;  /* coroutine::return_latency(x)                                        */
    int64_t return_priority; /* coroutine::return_priority(x); INT64_MAX = the process priority     */
} cxxdes_b200_process;

/* Host pointers; cxxdes_b200_create copies everything to the device. */
typedef struct cxxdes_b200_program {
    const cxxdes_b200_op *ops;
    const cxxdes_b200_process *procs;
    const uint32_t *roots;      /* processes bound with env.bind(), in this order (coroutine.ipp:352-357) */
    const uint64_t *queue_max;  /* per queue: max_size, 0 = unbounded (sync/queue.hpp:38)                 */
    const uint64_t *sem_init;   /* per semaphore: initial value                                           */
    const uint64_t *sem_max;    /* per semaphore: max (resource{c} = {c, c}, sync/resource.hpp:78)        */
    const double *rates;        /* rates of the exponential draws: max(1, sweep_steps) rows of n_rates     */
    uint32_t n_ops, n_procs, n_roots, n_queues, n_sems, n_mutexes, n_events, n_rates;
    uint32_t sweep_steps;       /* parameter sweep inside ONE ensemble (the loop of the examples' main(),
                                   producer_consumer.cpp:67-73, aloha.cpp:93-103): replication r draws with row
                                   (first_replication + r) % sweep_steps of `rates`; 0 = one row              */
    uint32_t reserved;
} cxxdes_b200_program;

/* ---- per-replication results, structure of arrays (one entry per replication) ----
 * Any pointer may be NULL (column not wanted).  f64/i64 columns are model outputs:
 *   MM1    f64[0]=avg_latency f64[1]=total_latency            (producer_consumer.cpp:24,29,45-52)
 *          i64[0]=packets consumed i64[1]=max queue length
 *   ALOHA  f64[0]=S (float widened) f64[1]=G (float widened)  (aloha.cpp:46-47)
 *          i64[0]=successful_transmissions i64[1]=packets_per_station
 *   MMC    f64[0]=total_wait  i64[0]=served i64[1]=balked
 *   ARCH   f64[0]=sum of load latencies (ticks)  i64[0]=hits i64[1]=misses
 */
#define CXXDES_B200_N_F64 2
#define CXXDES_B200_N_I64 2
typedef struct cxxdes_b200_columns {
    int64_t *end_time;    /* env.now() after the run: environment.ipp:24-26           */
    uint64_t *n_events;   /* number of environment::step() calls that returned true   */
    uint64_t *trace_hash; /* digest of every popped token (shared/trace_hash.hpp)     */
    uint32_t *status;     /* 0 = ok, else CXXDES_B200_REP_* bits                      */
    double *f64[CXXDES_B200_N_F64];
    int64_t *i64[CXXDES_B200_N_I64];
} cxxdes_b200_columns;

enum {
    CXXDES_B200_REP_QUEUE_OVERFLOW = 1u,  /* sync::queue ring capacity exceeded   */
    CXXDES_B200_REP_HEAP_OVERFLOW = 2u,   /* event heap capacity exceeded         */
    CXXDES_B200_REP_WAITER_OVERFLOW = 4u, /* sync::event waiter list exceeded     */
    CXXDES_B200_REP_TIME_RANGE = 8u,      /* tick value outside the packed range  */
    CXXDES_B200_REP_UNFINISHED = 16u,     /* run_until stopped before the end     */
    CXXDES_B200_REP_BAD_PROGRAM = 32u     /* process program did something the interpreter does not support */
};

/* ---- ensemble accumulators (what the final all-reduce sums) ---- */
typedef struct cxxdes_b200_accumulators {
    uint64_t n_replications;
    uint64_t n_events;
    uint64_t n_errors;
    uint64_t trace_hash_sum; /* wrapping sum of per-replication digests ("checksum of checksums") */
    int64_t end_time_sum;
    int64_t i64_sum[CXXDES_B200_N_I64];
    double f64_sum[CXXDES_B200_N_F64];
    double f64_sumsq[CXXDES_B200_N_F64];
} cxxdes_b200_accumulators;

typedef struct cxxdes_b200_ensemble cxxdes_b200_ensemble; /* opaque */

typedef struct cxxdes_b200_config {
    int32_t model;               /* CXXDES_B200_MODEL_*                                     */
    int32_t device;              /* CUDA device ordinal                                     */
    uint64_t seed;
    uint64_t first_replication;  /* global id of replication 0 of this shard (multi-GPU)    */
    uint64_t n_replications;
    cxxdes_b200_clock clock;
    const void *params;          /* cxxdes_b200_<model>_params                              */
    size_t params_size;
    void *stream;                /* cudaStream_t to launch on, NULL = a stream owned by the handle */
    uint32_t flags;              /* CXXDES_B200_FLAG_*                                      */
    uint32_t queue_capacity;     /* 0 = default; ring entries per replication (M/M/1), items per
                                    sync::queue of a process program (default 128)           */
} cxxdes_b200_config;

enum {
    CXXDES_B200_FLAG_GENERIC_ENGINE = 1u, /* use the generic token-heap engine kernel even when a
                                             fused model kernel exists (parity cross-check)    */
    CXXDES_B200_FLAG_MM1_LOCKSTEP = 2u,   /* M/M/1: first pass by the draw-on-demand kernel instead
                                             of the generate-ahead kernel (A/B measurements)   */
    CXXDES_B200_FLAG_COOP_HEAP = 4u,      /* ALOHA: first pass by the kernel whose event heaps are operated by groups of 8
                                             lanes (sub-warp cooperative push / pop with ballots) instead of one thread
                                             per replication (A/B measurements; identical results) */
    CXXDES_B200_FLAG_WIDE_TOKENS = 8u     /* ALOHA, M/M/c: first pass on 8-byte tokens (time, tag) even where the kernel with
                                             4-byte tokens (time relative to a moving epoch + a short code) applies;
                                             memory hierarchy: 64-bit stamps and 32-bit addresses in shared memory even
                                             where 32 / 16 bits provably suffice (A/B measurements; identical results) */
};

/* Number of CUDA devices visible (0 and an error string when there is none). */
int cxxdes_b200_device_count(int *count);

/* Construct an ensemble of `n_replications` independent simulations of one model:
 * replaces constructing one CXXDES_SIMULATION object per replication
 * (core/simulation.hpp:31-85) plus env.time_unit()/time_precision()
 * (environment.ipp:43-70).  Copies `params` to the device. */
int cxxdes_b200_create(const cxxdes_b200_config *cfg, cxxdes_b200_ensemble **out);

/* Validate a process program without a GPU: CXXDES_B200_OK, or CXXDES_B200_ERR_INVALID with the reason in
 * cxxdes_b200_last_error() (operand ranges, composition structure, forward-only IF / JUMP targets, the last
 * operation a RETURN, ...).  cxxdes_b200_create applies the same check. */
int cxxdes_b200_check_program(const cxxdes_b200_program *program);

/* simulation::run() for every replication: bind co_main once, then
 * `while (step());` (core/simulation.hpp:51-54,76-81; environment.ipp:117-146,179-182).
 * Asynchronous on the handle's stream.  On an ensemble that has already run to its end the reference's run()
 * finds nothing to dispatch; this call computes the same results again (a benchmark loop calls it once per step)
 * and leaves every replication's clock where the earlier run_until / run_for calls had put it. */
int cxxdes_b200_run(cxxdes_b200_ensemble *e);

/* environment::run_until(t) / run_for(dt) in ticks for every replication
 * (environment.ipp:190-214).  Every model keeps its replications' state on the device
 * between calls and continues from it (so does a later cxxdes_b200_run); only the
 * replications a kernel hands to its retry kernel are replayed from tick 0 to the new
 * horizon, with identical results. */
int cxxdes_b200_run_until(cxxdes_b200_ensemble *e, int64_t t);
int cxxdes_b200_run_for(cxxdes_b200_ensemble *e, int64_t dt);

/* Block until everything launched on the handle's stream has finished. */
int cxxdes_b200_sync(cxxdes_b200_ensemble *e);

/* Copy per-replication results to HOST arrays (sim.now(), model statistics:
 * core/simulation.hpp:33-45).  Implies a sync.  Returns CXXDES_B200_ERR_CAPACITY
 * if any replication reported a non-zero status. */
int cxxdes_b200_fetch(cxxdes_b200_ensemble *e, const cxxdes_b200_columns *host_columns);

/* DEVICE pointers of the result columns (for device-side consumers / NCCL). */
int cxxdes_b200_device_columns(cxxdes_b200_ensemble *e, cxxdes_b200_columns *device_columns);

/* Reduce the per-replication columns of this shard on the device and return
 * the accumulators (host copy) and, optionally, the device address of the
 * accumulator struct so the caller's collective (ncclAllReduce over NVLink)
 * can sum shards without a host round trip. */
int cxxdes_b200_reduce(cxxdes_b200_ensemble *e, cxxdes_b200_accumulators *host_out,
                       void **device_accumulators);

/* cxxdes_b200_reduce, then sum the accumulators of all shards: one ncclAllReduce pair (7 x 64-bit integer
 * words, 4 x double) over `nccl_comm` -- an ncclComm_t of the caller with one rank per shard -- on the handle's
 * stream, in place on the device struct; `host_out` receives the ensemble totals (every rank the same).  This is
 * the only exchange of the path: replications share nothing.  NCCL is resolved when this is first
 * called (the libnccl.so.2 already loaded in the process, else the system one): the library itself does not
 * link against it.  Every rank of the communicator must make the call. */
int cxxdes_b200_reduce_allgpus(cxxdes_b200_ensemble *e, void *nccl_comm, cxxdes_b200_accumulators *host_out);

/* Device time of the most recent run/run_until launch in milliseconds (CUDA
 * events recorded on the launching stream), and how many kernels this library
 * launched on the handle so far. */
int cxxdes_b200_last_run_ms(cxxdes_b200_ensemble *e, float *ms);
int cxxdes_b200_launch_count(cxxdes_b200_ensemble *e, uint64_t *launches);

/* Diagnostics of the most recent run / run_until / run_for launch: how many replications left the first-pass
 * kernel's compact representation and were redone (results are identical either way; this only tells where the
 * time went).  counts[0] = redone by the model's second pass (M/M/1: queue outgrew the shared-memory window ->
 * two-level ring; M/M/c and process programs: heap, wait list or queue outgrew the fixed capacities of the
 * kernels -> the last pass, all of them in global memory; for M/M/c
 * sized for the model's worst case, because the reference's containers are unbounded: std::priority_queue,
 * environment.ipp:263, std::vector of waiters, sync/event.hpp:87-139.  CXXDES_B200_DEEP_BYTES, read at create, bounds
 * the memory of that pass -- default 256 MiB, 0 = no such pass, the replication keeps its status flag),
 * counts[1] = of those, taken by the helper grid that runs next to the first pass,
 * counts[2] = redone by the full-size 64-bit kernel, counts[3] = the longest time, in nanoseconds, the helper grid spent on
 * one replication (0: none).  Implies a sync. */
int cxxdes_b200_pass_counts(cxxdes_b200_ensemble *e, uint64_t counts[4]);

/* Event trace of ONE replication (index within this shard): the replication is run again from tick 0 -- to the end,
 * or to `until` >= 0 as environment::run_until would -- on the model's full-size kernel with a sink attached to
 * environment::step (environment.ipp:117-146), and the first `max_events` popped tokens are returned in order as
 * (time, priority, kind = CXXDES_B200_TOKEN_*, process ordinal or 0xFFFF for a null-target token): the four public
 * token fields (token.ipp:23-35) the trace digest folds, i.e. what the reference would print under
 * CXXDES_DEBUG_TOKEN (token.ipp:50-60).  *n_events = events of the whole run, *trace_hash = its digest: equal to the
 * replication's trace_hash column means the list IS the sequence the ensemble's kernel dispatched; a digest mismatch
 * against a CPU run is localised by diffing the two lists.  Host arrays of max_events entries; result columns and
 * saved state of the handle are not touched.  Implies a sync. */
enum { CXXDES_B200_TOKEN_RESUME = 0, CXXDES_B200_TOKEN_HANDLER = 1, CXXDES_B200_TOKEN_NOOP = 2 };
int cxxdes_b200_trace(cxxdes_b200_ensemble *e, uint64_t replication, int64_t until, uint64_t max_events, int64_t *time,
                      int64_t *priority, uint32_t *kind, uint32_t *proc, uint64_t *n_written, uint64_t *n_events,
                      uint64_t *trace_hash);

/* Name of the first-pass kernel instantiation a plain run() of this handle launches, e.g.
 * "mm1_ahead_kernel<256,3,32,4,0,0,0>+helper<32,16,32,4>" (template arguments as in csrc/): the key under which
 * profiles/ stores measured per-kernel figures.  NUL-terminated, truncated to buf_size. */
int cxxdes_b200_kernel_name(cxxdes_b200_ensemble *e, char *buf, size_t buf_size);

/* The same answer for the ALOHA, M/M/c and memory-hierarchy models WITHOUT a handle or a device: their first-pass
 * kernel follows from cfg->params, cfg->clock and cfg->flags alone (token width, layout: which ranges the parameters
 * provably or almost surely stay in).  M/M/1 and process programs are planned from the device's resources as well:
 * CXXDES_B200_ERR_STATE.  Lets a host, or a test on a machine without a GPU, see which path a configuration takes. */
int cxxdes_b200_plan_kernel(const cxxdes_b200_config *cfg, char *buf, size_t buf_size);

/* environment::reset() (environment.ipp:154-176) + rebind: replications start over. */
int cxxdes_b200_reset(cxxdes_b200_ensemble *e);

int cxxdes_b200_destroy(cxxdes_b200_ensemble *e);

/* Message of the last error on the calling thread (the reference reports these
 * conditions by throwing std::runtime_error; the C++ wrapper rethrows). */
const char *cxxdes_b200_last_error(void);

/* Library/ABI version and the SM architecture the kernels were built for. */
int cxxdes_b200_version(int *abi_version, int *sm_arch);

#ifdef __cplusplus
}
#endif
#endif /* CXXDES_B200_H_INCLUDED */
