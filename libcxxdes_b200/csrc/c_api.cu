// c_api.cu -- extern "C" boundary of the ensemble engine (include/cxxdes_b200.h).
//
// Owns the device buffers behind the opaque handle, picks the kernel for the model, launches on
// the caller's stream (or one it owns) and brackets every run with CUDA events recorded on that
// same stream.  No CPU implementation of the dispatch loop exists on this side of the boundary:
// if no CUDA device is present every entry point fails with CXXDES_B200_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "cxxdes_b200.h"
#include "kernels.h"

using namespace cxb;

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(CXXDES_B200_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// time::count, include/cxxdes/misc/time.hpp:75-86 (integer arithmetic, evaluated once on the host).
int64_t time_count(int64_t t, int u, int64_t prec_t, int prec_u) {
    static const int64_t conv[5] = {1, 1000, 1000000, 1000000000, 1000000000000LL};
    if (u < prec_u) return conv[prec_u - u] * t / prec_t;
    return (t / prec_t) / conv[u - prec_u];
}

union model_params {
    cxxdes_b200_mm1_params mm1;
    cxxdes_b200_aloha_params aloha;
    cxxdes_b200_mmc_params mmc;
    cxxdes_b200_archsim_params arch;
    cxxdes_b200_program program;  // host pointers are only valid during cxxdes_b200_create
};

size_t params_size_of(int model) {
    switch (model) {
        case CXXDES_B200_MODEL_MM1: return sizeof(cxxdes_b200_mm1_params);
        case CXXDES_B200_MODEL_ALOHA: return sizeof(cxxdes_b200_aloha_params);
        case CXXDES_B200_MODEL_MMC: return sizeof(cxxdes_b200_mmc_params);
        case CXXDES_B200_MODEL_ARCHSIM: return sizeof(cxxdes_b200_archsim_params);
        case CXXDES_B200_MODEL_PROGRAM: return sizeof(cxxdes_b200_program);
        default: return 0;
    }
}

}  // namespace

struct cxxdes_b200_ensemble {
    cxxdes_b200_config cfg;
    model_params params;
    clock_config clk;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool timed = false;
    uint64_t launches = 0;
    int64_t horizon = -1;  // last run_until target; -1 = never run, INT64_MAX = ran to exhaustion
    // The clock of an exhausted ensemble: the horizon a run_until / run_for had reached before the run() that exhausted it
    // (now_ stays there if the model ended earlier, environment.ipp:194), and what run_until (first: 0) / run_for (first:
    // 1) calls made of it afterwards, in order.  A run() that recomputes the ensemble applies them again.
    int64_t end_floor = -1;
    std::vector<std::pair<int, int64_t>> after_end;
    bool used = false;

    // device memory
    unsigned char *col_block = nullptr;  // all result columns, one allocation
    device_columns cols{};
    unsigned long long *counters = nullptr;  // [0] work, [1] retry count, [2] fallback cursor, [3] list A count, [4] list A cursor,
                                             // [5] M/M/1 first pass done (flag), [6] list A entries taken by the overlapped helper,
                                             // [7] longest time (ns) the helper spent on one replication
    uint32_t *retry_list = nullptr;
    log_entry *log_table = nullptr;
    int64_t *ring = nullptr;
    size_t ring_bytes = 0;
    void *ahead_gring = nullptr;  // second ring level of the generate-ahead M/M/1 kernel
    size_t ahead_gring_bytes = 0;
    uint32_t *mm1_state = nullptr;  // per-replication state between run_until / run_for pieces (M/M/1)
    bool mm1_state_valid = false;   // written by the previous launch of this handle
    cxxdes_b200_accumulators *acc = nullptr;
    void *reduce_scratch = nullptr;

    unsigned char *program_blob = nullptr;  // device copy of a process program
    uint32_t program_blob_bytes = 0;
    device_program dprog{};
    program_fast_layout prog_fast{};        // shared-memory first pass (program_fast.cu), sized from the program
    int64_t *prog_queue_items = nullptr;    // its queue payloads, [queue][slot][thread]
    // replications between run_until / run_for pieces, `[word][replication]` as the handle's model kernel defines it
    // (program_fast.cu, models_aloha.cu, models_mmc.cu, models_archsim.cu); allocated by the first piece
    uint32_t *saved_words = nullptr;
    int64_t *saved_queue_items = nullptr;   // process programs: queue payloads
    bool saved_valid = false;               // written by the previous launch of this handle
    // last pass of the models whose containers are unbounded in the reference: heap and wait lists in global memory,
    // sized for the model's worst case (M/M/c: models_mmc.cu mmc_deep_plan)
    mmc_deep mmc_deep_pass{};
    program_deep prog_deep_pass{};

    // M/M/1: the second pass runs next to the first one on a stream of the handle's own (mm1_fused.cu)
    cudaStream_t helper_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    bool use_fused = false;
    mm1_fused_plan mm1_plan{};
    int engine_blocks = 0, engine_threads = 128;
    uint32_t engine_ring = 0;
};

namespace {

int set_device(const cxxdes_b200_ensemble *e) {
    CUDA_TRY(cudaSetDevice(e->cfg.device));
    return CXXDES_B200_OK;
}

// Global memory the last, unbounded pass of a model may use for its heaps and wait lists (CXXDES_B200_DEEP_BYTES, read at
// create; 0 switches the pass off and a replication that outgrows the on-chip capacities keeps its status flag).
size_t deep_pass_budget() {
    size_t budget = (size_t)256 << 20;
    if (const char *text = getenv("CXXDES_B200_DEEP_BYTES")) budget = (size_t)strtoull(text, nullptr, 10);
    const size_t most = (size_t)8 << 30;   // entries are indexed [entry * threads] in 32 bits: 8 GiB / 12 bytes stays below 2^31
    return budget < most ? budget : most;
}

int allocate(cxxdes_b200_ensemble *e) {
    const uint64_t n = e->cfg.n_replications;
    // columns: 8-byte ones first, the 4-byte status last; every column 256-byte aligned
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t col8 = pad(n * 8), col4 = pad(n * 4);
    const size_t total = col8 * (3 + CXXDES_B200_N_F64 + CXXDES_B200_N_I64) + col4;
    CUDA_TRY(cudaMalloc(&e->col_block, total));
    CUDA_TRY(cudaMemsetAsync(e->col_block, 0, total, e->stream));
    unsigned char *p = e->col_block;
    e->cols.end_time = (int64_t *)p; p += col8;
    e->cols.n_events = (uint64_t *)p; p += col8;
    e->cols.trace_hash = (uint64_t *)p; p += col8;
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) { e->cols.f64[k] = (double *)p; p += col8; }
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) { e->cols.i64[k] = (int64_t *)p; p += col8; }
    e->cols.status = (uint32_t *)p;
    CUDA_TRY(cudaMalloc(&e->counters, 8 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&e->retry_list, 2 * sizeof(uint32_t) * (n ? n : 1)));  // two lists (models_common.cuh)
    CUDA_TRY(cudaMalloc(&e->log_table, sizeof(LOG_TABLE_HOST)));
    CUDA_TRY(cudaMemcpyAsync(e->log_table, LOG_TABLE_HOST, sizeof(LOG_TABLE_HOST), cudaMemcpyHostToDevice,
                             e->stream));
    CUDA_TRY(cudaMalloc(&e->acc, sizeof(cxxdes_b200_accumulators)));
    CUDA_TRY(cudaMalloc(&e->reduce_scratch, reduce_scratch_bytes(e->sm_count)));
    if (e->ring_bytes) CUDA_TRY(cudaMalloc(&e->ring, e->ring_bytes));
    if (e->ahead_gring_bytes) CUDA_TRY(cudaMalloc(&e->ahead_gring, e->ahead_gring_bytes));
    // (the last passes are an extra: when their memory cannot be had the ensemble works without them and a replication
    // that outgrows the fixed capacities keeps its status flag)
    if (e->cfg.model == CXXDES_B200_MODEL_MMC) {
        if (const size_t bytes = mmc_deep_plan(e->params.mmc, e->sm_count, deep_pass_budget(), e->mmc_deep_pass))
            if (cudaMalloc(&e->mmc_deep_pass.base, bytes) != cudaSuccess) {
                (void)cudaGetLastError();
                e->mmc_deep_pass = mmc_deep{};
            }
    }
    if (e->cfg.model == CXXDES_B200_MODEL_PROGRAM) {
        if (const size_t bytes = program_deep_plan(e->dprog, e->sm_count, deep_pass_budget(), e->prog_deep_pass))
            if (cudaMalloc(&e->prog_deep_pass.base, bytes) != cudaSuccess) {
                (void)cudaGetLastError();
                e->prog_deep_pass = program_deep{};
            }
        if (program_needs_last_pass(e->dprog) && !e->prog_deep_pass.base)
            return fail(CXXDES_B200_ERR_INVALID, "program: more than 4 sync objects of a kind need the last pass's memory "
                                                 "(CXXDES_B200_DEEP_BYTES too small, or the allocation failed)");
    }
    return CXXDES_B200_OK;
}

// Deep-copy a process program to the device (one allocation) and build the device-side view.
int upload_program(cxxdes_b200_ensemble *e) {
    const cxxdes_b200_program &p = e->params.program;
    const char *why = program_check(p);
    if (why) return fail(CXXDES_B200_ERR_INVALID, "%s", why);
    if ((p.n_queues && !p.queue_max) || (p.n_sems && (!p.sem_init || !p.sem_max)) || (p.n_rates && !p.rates))
        return fail(CXXDES_B200_ERR_INVALID, "program: a counted array is NULL");
    auto pad = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t s_ops = pad(sizeof(cxxdes_b200_op) * p.n_ops), s_procs = pad(sizeof(cxxdes_b200_process) * p.n_procs),
                 s_roots = pad(4 * p.n_roots), s_q = pad(8 * p.n_queues), s_s = pad(8 * p.n_sems),
                 s_r = pad(sizeof(divisor) * p.n_rates * (p.sweep_steps ? p.sweep_steps : 1));
    const size_t total = s_ops + s_procs + s_roots + s_q + 2 * s_s + s_r + 16;
    unsigned char *host = new (std::nothrow) unsigned char[total]();
    if (!host) return fail(CXXDES_B200_ERR_INVALID, "out of host memory");
    size_t off = 0;
    auto put = [&](const void *src, size_t bytes, size_t padded) {
        size_t at = off;
        if (bytes) memcpy(host + off, src, bytes);
        off += padded;
        return at;
    };
    const size_t o_ops = put(p.ops, sizeof(cxxdes_b200_op) * p.n_ops, s_ops);
    const size_t o_procs = put(p.procs, sizeof(cxxdes_b200_process) * p.n_procs, s_procs);
    const size_t o_roots = put(p.roots, 4 * p.n_roots, s_roots);
    const size_t o_q = put(p.queue_max, 8 * p.n_queues, s_q);
    const size_t o_si = put(p.sem_init, 8 * p.n_sems, s_s);
    const size_t o_sm = put(p.sem_max, 8 * p.n_sems, s_s);
    const size_t o_r = off;
    for (uint32_t k = 0; k < p.n_rates * (p.sweep_steps ? p.sweep_steps : 1); ++k) {
        if (!(p.rates[k] > 0.0)) {
            delete[] host;
            return fail(CXXDES_B200_ERR_INVALID, "program: rates must be > 0");
        }
        divisor d = make_divisor(p.rates[k]);  // IEEE 1.0 / rate on the host: same bits as on the device
        memcpy(host + off + k * sizeof(divisor), &d, sizeof(divisor));
    }
    cudaError_t ce = cudaMalloc(&e->program_blob, total);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->program_blob, host, total, cudaMemcpyHostToDevice);
    delete[] host;
    if (ce != cudaSuccess) return fail(CXXDES_B200_ERR_CUDA, "program upload: %s", cudaGetErrorString(ce));
    e->program_blob_bytes = (uint32_t)total;
    e->prog_fast = program_fast_plan(p, (uint32_t)total, e->smem_optin, e->cfg.queue_capacity);
    if (const size_t qb = program_fast_queue_bytes(e->prog_fast, e->sm_count)) {
        ce = cudaMalloc(&e->prog_queue_items, qb);
        if (ce != cudaSuccess) return fail(CXXDES_B200_ERR_CUDA, "program queues: %s", cudaGetErrorString(ce));
    }
    device_program &d = e->dprog;
    d.ops = (const cxxdes_b200_op *)(e->program_blob + o_ops);
    d.procs = (const cxxdes_b200_process *)(e->program_blob + o_procs);
    d.roots = (const uint32_t *)(e->program_blob + o_roots);
    d.queue_max = (const uint64_t *)(e->program_blob + o_q);
    d.sem_init = (const uint64_t *)(e->program_blob + o_si);
    d.sem_max = (const uint64_t *)(e->program_blob + o_sm);
    d.rates = (const divisor *)(e->program_blob + o_r);
    d.n_ops = p.n_ops; d.n_procs = p.n_procs; d.n_roots = p.n_roots; d.n_queues = p.n_queues;
    d.n_sems = p.n_sems; d.n_mutexes = p.n_mutexes; d.n_events = p.n_events; d.n_rates = p.n_rates;
    d.sweep_steps = p.sweep_steps;
    memset(&e->params.program, 0, sizeof(e->params.program));  // the host pointers are not ours to keep
    return CXXDES_B200_OK;
}

int plan(cxxdes_b200_ensemble *e) {
    const cxxdes_b200_config &c = e->cfg;
    switch (c.model) {
        case CXXDES_B200_MODEL_MM1: {
            const cxxdes_b200_mm1_params &p = e->params.mm1;
            if (!(p.lambda > 0.0) || !(p.mu > 0.0)) return fail(CXXDES_B200_ERR_INVALID, "mm1: rates must be > 0");
            e->use_fused = !(c.flags & CXXDES_B200_FLAG_GENERIC_ENGINE) && p.n_packets >= 2 &&
                           p.n_packets < (1ULL << 30);
            if (e->use_fused) {
                e->mm1_plan = plan_mm1_fused(p, c.queue_capacity, e->sm_count, e->smem_optin, e->clk.tick_mult, e->clk.ticks_per_unit,
                                             (c.flags & CXXDES_B200_FLAG_MM1_LOCKSTEP) != 0, c.n_replications);
                e->ring_bytes = (size_t)e->mm1_plan.fb_blocks * e->mm1_plan.fb_threads * e->mm1_plan.fb_ring *
                                sizeof(int64_t);
                if (!(c.flags & CXXDES_B200_FLAG_MM1_LOCKSTEP)) e->ahead_gring_bytes = mm1_ahead_gring_bytes(e->mm1_plan);
            } else {
                e->engine_threads = 128;
                e->engine_blocks = e->sm_count * 4;
                uint64_t cap = c.queue_capacity ? c.queue_capacity : 1024;
                if (cap > p.n_packets) cap = p.n_packets;
                if (cap < 2) cap = 2;
                e->engine_ring = (uint32_t)cap;
                e->ring_bytes = (size_t)e->engine_blocks * e->engine_threads * cap * sizeof(int64_t);
            }
            return CXXDES_B200_OK;
        }
        case CXXDES_B200_MODEL_ALOHA: {
            const cxxdes_b200_aloha_params &p = e->params.aloha;
            if (p.num_stations < 1 || p.num_stations > 64)
                return fail(CXXDES_B200_ERR_INVALID, "aloha: num_stations must be in [1, 64]");
            if (!(p.lambda_begin > 0.0f) || !(p.frame_time >= 0.0f))
                return fail(CXXDES_B200_ERR_INVALID, "aloha: lambda must be > 0 and frame_time >= 0");
            if (p.packets_per_station >= (1ULL << 32))
                return fail(CXXDES_B200_ERR_INVALID, "aloha: packets_per_station must be < 2^32");
            if (aloha_smem_bytes(p.num_stations) > e->smem_optin)
                return fail(CXXDES_B200_ERR_INVALID, "aloha: shared memory budget exceeded");
            return CXXDES_B200_OK;
        }
        case CXXDES_B200_MODEL_MMC: {
            const cxxdes_b200_mmc_params &p = e->params.mmc;
            if (!(p.lambda > 0.0) || !(p.mu > 0.0) || !(p.patience >= 0.0) || p.servers < 1)
                return fail(CXXDES_B200_ERR_INVALID, "mmc: rates must be > 0, patience >= 0, servers >= 1");
            if (p.n_customers > 32000)
                return fail(CXXDES_B200_ERR_INVALID, "mmc: n_customers must be <= 32000 (16-bit process ordinals)");
            e->ring_bytes = (size_t)mmc_blocks(e->sm_count) * mmc_threads() * (p.n_customers ? p.n_customers : 1) *
                            sizeof(uint64_t);
            return CXXDES_B200_OK;
        }
        case CXXDES_B200_MODEL_ARCHSIM: {
            if (!archsim_supported(e->params.arch))
                return fail(CXXDES_B200_ERR_INVALID,
                            "archsim: need 1..16 requesters, capacity 1..64, addr_space >= 1, latencies >= 0");
            if (archsim_smem_bytes(e->params.arch) > e->smem_optin)
                return fail(CXXDES_B200_ERR_INVALID, "archsim: shared memory budget exceeded");
            return CXXDES_B200_OK;
        }
        case CXXDES_B200_MODEL_PROGRAM:
            return upload_program(e);
        default:
            return fail(CXXDES_B200_ERR_INVALID, "model %d is not available in this build", c.model);
    }
}

// State kept between pieces: wanted from the first run_until / run_for on (and by a run() that follows one).
// Returns the buffer (nullptr: the plain run() kernels) and whether it holds the previous launch's replications.
int saved_state_for(cxxdes_b200_ensemble *e, int64_t until, size_t words_per_replication, uint32_t **words, int *resume) {
    *words = nullptr;
    *resume = 0;
    if (until < 0 && !e->saved_valid) return CXXDES_B200_OK;
    if (!e->saved_words) CUDA_TRY(cudaMalloc(&e->saved_words, words_per_replication * 4 * e->cfg.n_replications));
    *words = e->saved_words;
    *resume = e->saved_valid ? 1 : 0;
    return CXXDES_B200_OK;
}

int launch(cxxdes_b200_ensemble *e, int64_t until) {
    int rc = set_device(e);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(e->counters, 0, 8 * sizeof(unsigned long long), e->stream));
    shard sh{};
    sh.seed = e->cfg.seed;
    sh.first_replication = e->cfg.first_replication;
    sh.n_replications = e->cfg.n_replications;
    sh.clk = e->clk;
    sh.run_until = until;
    sh.work_counter = e->counters + 0;
    sh.retry_count = e->counters + 1;
    sh.retry_list = e->retry_list;
    sh.retry_a_count = e->counters + 3;
    sh.retry_a_cursor = e->counters + 4;
    sh.retry_a_list = e->retry_list + e->cfg.n_replications;
    sh.log_table = e->log_table;
    CUDA_TRY(cudaEventRecord(e->ev_begin, e->stream));
    int launches = 0;
    // whether THIS launch was given the state buffers (and so leaves them describing its replications): a run() on
    // the plain kernels after reset(), or a piece that ran on a kernel without saved state, must not make a later
    // call resume from records an earlier launch wrote
    bool wrote_mm1_state = false, wrote_saved = false;
    switch (e->cfg.model) {
        case CXXDES_B200_MODEL_MM1:
            if (e->use_fused) {
                // run_until / run_for continue where the previous piece stopped (environment.ipp:190-214)
                // when the generate-ahead kernel runs the model; otherwise a piece replays from tick 0
                const bool resumable = !(e->cfg.flags & CXXDES_B200_FLAG_MM1_LOCKSTEP) && !e->mm1_plan.ah_direct_global;
                if (resumable && !e->mm1_state && (until >= 0 || e->mm1_state_valid)) {
                    CUDA_TRY(cudaMalloc(&e->mm1_state, mm1_ahead_state_bytes(e->cfg.n_replications)));
                }
                const int resume = resumable && e->mm1_state && e->mm1_state_valid;
                wrote_mm1_state = resumable && e->mm1_state != nullptr;
                const mm1_overlap overlap{e->helper_stream, e->ev_fork, e->ev_join, (unsigned int *)(e->counters + 5)};
                CUDA_TRY(launch_mm1_fused(e->params.mm1, sh, e->cols, e->mm1_plan, e->ring, e->counters + 2,
                                          e->ahead_gring, resumable ? e->mm1_state : nullptr, resume, e->stream, &launches,
                                          (e->cfg.flags & CXXDES_B200_FLAG_MM1_LOCKSTEP) != 0, overlap));
            } else {
                CUDA_TRY(launch_mm1_engine(e->params.mm1, sh, e->cols, e->ring, e->engine_ring, e->engine_blocks,
                                           e->engine_threads, e->stream));
                launches = 1;
            }
            break;
        case CXXDES_B200_MODEL_ALOHA:
        {
            // run_until / run_for continue where the previous piece stopped when the 32-bit kernel runs the model
            aloha_saved st{};
            st.n_replications = e->cfg.n_replications;
            const bool coop = (e->cfg.flags & CXXDES_B200_FLAG_COOP_HEAP) != 0;   // keeps no state: pieces replay from tick 0
            if (!coop && aloha_keeps_state(e->params.aloha, sh)) {
                rc = saved_state_for(e, until, aloha_saved_words(e->params.aloha), &st.words, &st.resume);
                if (rc) return rc;
            }
            wrote_saved = st.words != nullptr;
            CUDA_TRY(launch_aloha(e->params.aloha, sh, e->cols, e->sm_count, e->smem_optin, e->counters + 2, e->stream,
                                  &launches, st, coop, (e->cfg.flags & CXXDES_B200_FLAG_WIDE_TOKENS) != 0));
            break;
        }
        case CXXDES_B200_MODEL_MMC: {
            mmc_saved st{};
            st.n_replications = e->cfg.n_replications;
            if (mmc_keeps_state(e->params.mmc, sh)) {
                rc = saved_state_for(e, until, mmc_saved_words(e->params.mmc), &st.words, &st.resume);
                if (rc) return rc;
            }
            wrote_saved = st.words != nullptr;
            CUDA_TRY(launch_mmc(e->params.mmc, sh, e->cols, (uint64_t *)e->ring, e->sm_count, e->counters + 2, e->stream,
                                &launches, st, (e->cfg.flags & CXXDES_B200_FLAG_WIDE_TOKENS) != 0, e->mmc_deep_pass));
            break;
        }
        case CXXDES_B200_MODEL_ARCHSIM: {
            archsim_saved st{};
            st.n_replications = e->cfg.n_replications;
            rc = saved_state_for(e, until, archsim_saved_words(e->params.arch), &st.words, &st.resume);
            if (rc) return rc;
            wrote_saved = st.words != nullptr;
            CUDA_TRY(launch_archsim(e->params.arch, sh, e->cols, e->sm_count, e->stream, st,
                                    (e->cfg.flags & CXXDES_B200_FLAG_WIDE_TOKENS) != 0));
            launches = 1;
            break;
        }
        case CXXDES_B200_MODEL_PROGRAM: {
            // run_until / run_for continue where the previous piece stopped (environment.ipp:190-214): the
            // shared-memory interpreter saves what is still running at the horizon and resumes it
            program_fast_state st{};
            st.n_replications = e->cfg.n_replications;
            if (e->prog_fast.valid) {
                rc = saved_state_for(e, until, program_fast_state_words(e->prog_fast), &st.words, &st.resume);
                if (rc) return rc;
                if (st.words && !e->saved_queue_items)
                    if (const size_t qb = program_fast_state_queue_bytes(e->prog_fast, e->cfg.n_replications))
                        CUDA_TRY(cudaMalloc(&e->saved_queue_items, qb));
                st.queue_items = e->saved_queue_items;
            }
            wrote_saved = st.words != nullptr;
            CUDA_TRY(launch_program(e->dprog, e->program_blob_bytes, sh, e->cols, e->sm_count, e->counters + 2, e->stream,
                                    &launches, e->prog_fast, e->prog_queue_items, st, e->prog_deep_pass));
            break;
        }
        default:
            return fail(CXXDES_B200_ERR_INVALID, "model %d is not available in this build", e->cfg.model);
    }
    if (until < 0) {
        // run() after run_until(horizon): now_ stays at the horizon where the model ended before it.  On an ensemble that
        // is already exhausted the reference's run() finds nothing to dispatch; here the same results are computed again
        // (what a benchmark loop times) and the clock is brought back to where the earlier calls had left it.
        if (e->horizon != INT64_MAX) {
            e->end_floor = e->horizon;
            e->after_end.clear();
        }
        if (e->end_floor >= 0) {
            CUDA_TRY(launch_clock_floor(e->cols.end_time, e->cfg.n_replications, e->end_floor, e->sm_count, e->stream));
            ++launches;
        }
        for (const std::pair<int, int64_t> &op : e->after_end) {
            if (op.first == 0) CUDA_TRY(launch_clock_floor(e->cols.end_time, e->cfg.n_replications, op.second, e->sm_count, e->stream));
            else CUDA_TRY(launch_clock_add(e->cols.end_time, e->cfg.n_replications, op.second, e->sm_count, e->stream));
            ++launches;
        }
    }
    CUDA_TRY(cudaEventRecord(e->ev_end, e->stream));
    e->launches += launches;
    e->timed = true;
    e->used = true;
    e->mm1_state_valid = wrote_mm1_state;
    e->saved_valid = wrote_saved;
    e->horizon = until < 0 ? INT64_MAX : until;
    return CXXDES_B200_OK;
}

}  // namespace

extern "C" {

int cxxdes_b200_device_count(int *count) {
    if (!count) return fail(CXXDES_B200_ERR_INVALID, "count is NULL");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return CXXDES_B200_OK;
}

int cxxdes_b200_version(int *abi_version, int *sm_arch) {
    if (abi_version) *abi_version = 1;
    if (sm_arch) *sm_arch = 100;
    return CXXDES_B200_OK;
}

const char *cxxdes_b200_last_error(void) { return g_error; }

int cxxdes_b200_create(const cxxdes_b200_config *cfg, cxxdes_b200_ensemble **out) {
    if (!cfg || !out) return fail(CXXDES_B200_ERR_INVALID, "cfg/out is NULL");
    *out = nullptr;
    size_t want = params_size_of(cfg->model);
    if (want == 0) return fail(CXXDES_B200_ERR_INVALID, "unknown model %d", cfg->model);
    if (!cfg->params || cfg->params_size != want)
        return fail(CXXDES_B200_ERR_INVALID, "params_size %zu does not match model %d (%zu)", cfg->params_size,
                    cfg->model, want);
    if ((cfg->flags & CXXDES_B200_FLAG_COOP_HEAP) && cfg->model != CXXDES_B200_MODEL_ALOHA)
        return fail(CXXDES_B200_ERR_INVALID, "CXXDES_B200_FLAG_COOP_HEAP: only the ALOHA model has a cooperative-heap kernel");
    if (cfg->n_replications == 0 || cfg->n_replications > 0xFFFFFFFFull)
        return fail(CXXDES_B200_ERR_INVALID, "n_replications must be in [1, 2^32)");
    const cxxdes_b200_clock &ck = cfg->clock;
    if (ck.unit_u < 0 || ck.unit_u > 4 || ck.precision_u < 0 || ck.precision_u > 4 || ck.precision_t <= 0 ||
        ck.unit_t <= 0)
        return fail(CXXDES_B200_ERR_INVALID, "invalid clock");
    int n_dev = 0;
    CUDA_TRY(cudaGetDeviceCount(&n_dev));
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(CXXDES_B200_ERR_CUDA, "device %d not present (%d visible)", cfg->device, n_dev);

    cxxdes_b200_ensemble *e = new (std::nothrow) cxxdes_b200_ensemble();
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "out of host memory");
    e->cfg = *cfg;
    memcpy(&e->params, cfg->params, want);
    e->cfg.params = nullptr;
    e->clk.ticks_per_unit = time_count(ck.unit_t, ck.unit_u, ck.precision_t, ck.precision_u);
    e->clk.tick_mult = ck.precision_t;
    static const double conv[5] = {1, 1e-3, 1e-6, 1e-9, 1e-12};
    e->clk.seconds_per_tick_unit = conv[ck.precision_u];

    int rc = set_device(e);
    if (rc) { delete e; return rc; }
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, cfg->device);
    if (ce != cudaSuccess) { delete e; return fail(CXXDES_B200_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce)); }
    if (prop.major != 10) {
        delete e;
        return fail(CXXDES_B200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                    prop.major, prop.minor);
    }
    e->sm_count = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    if (cfg->stream) {
        e->stream = (cudaStream_t)cfg->stream;
    } else {
        ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
        if (ce != cudaSuccess) { delete e; return fail(CXXDES_B200_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)); }
        e->own_stream = true;
    }
    rc = plan(e);
    if (rc == CXXDES_B200_OK) rc = allocate(e);
    if (rc == CXXDES_B200_OK) {
        ce = cudaEventCreate(&e->ev_begin);
        if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev_end);
        if (ce == cudaSuccess && e->use_fused) ce = configure_mm1_fused(e->mm1_plan);
        if (ce == cudaSuccess && e->use_fused) ce = configure_mm1_ahead(e->mm1_plan);
        // (CXXDES_B200_MM1_NO_OVERLAP=1: second pass after the first one, as before -- for A/B measurements)
        const char *no_overlap = getenv("CXXDES_B200_MM1_NO_OVERLAP");
        if (ce == cudaSuccess && e->use_fused && !(cfg->flags & CXXDES_B200_FLAG_MM1_LOCKSTEP) && !(no_overlap && no_overlap[0] == '1')) {
            // highest priority: its few small blocks are to be placed next to the first pass as soon as they are
            // launched (and streams of different priority do not share a hardware work queue, behind whose head -- the
            // first pass -- the helper would otherwise wait: seen with several handles in one process)
            int prio_low = 0, prio_high = 0;
            ce = cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
            if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&e->helper_stream, cudaStreamNonBlocking, prio_high);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
        }
        if (ce != cudaSuccess) rc = fail(CXXDES_B200_ERR_CUDA, "setup: %s", cudaGetErrorString(ce));
    }
    if (rc != CXXDES_B200_OK) {
        cxxdes_b200_destroy(e);
        return rc;
    }
    *out = e;
    return CXXDES_B200_OK;
}

int cxxdes_b200_check_program(const cxxdes_b200_program *program) {
    if (!program) return fail(CXXDES_B200_ERR_INVALID, "program is NULL");
    if (const char *why = program_check(*program)) return fail(CXXDES_B200_ERR_INVALID, "%s", why);
    if ((program->n_queues && !program->queue_max) || (program->n_sems && (!program->sem_init || !program->sem_max)) ||
        (program->n_rates && !program->rates))
        return fail(CXXDES_B200_ERR_INVALID, "program: a counted array is NULL");
    for (uint32_t k = 0; k < program->n_rates * (program->sweep_steps ? program->sweep_steps : 1); ++k)
        if (!(program->rates[k] > 0.0)) return fail(CXXDES_B200_ERR_INVALID, "program: rates must be > 0");
    return CXXDES_B200_OK;
}

int cxxdes_b200_run(cxxdes_b200_ensemble *e) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    // Already at its end and every replication's state kept from the pieces that led there: nothing to dispatch, results
    // and clocks stay (the reference's run() on an exhausted environment).  Without kept state the ensemble is computed
    // again from tick 0 (launch() restores the clocks).
    if (e->horizon == INT64_MAX && (e->saved_valid || e->mm1_state_valid)) return CXXDES_B200_OK;
    return launch(e, -1);
}

// Every first-pass kernel keeps each replication's state between calls and continues from it, as
// environment::run_until does (M/M/1: 128 bytes per replication; programs, ALOHA, M/M/c, arch-sim: their
// heap and model state).  The retry kernels replay a
// replication from tick 0 up to the new target, which yields exactly the state the reference
// reaches by continuing (every replication is a deterministic function of its seed).
int cxxdes_b200_run_until(cxxdes_b200_ensemble *e, int64_t t) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    if (t < 0) return fail(CXXDES_B200_ERR_INVALID, "run_until: t must be >= 0");
    if (e->horizon == INT64_MAX) {
        // already exhausted: nothing left to dispatch, but the clock still follows, now_ = max(now_, t)
        // (environment.ipp:194)
        int rc = set_device(e);
        if (rc) return rc;
        CUDA_TRY(launch_clock_floor(e->cols.end_time, e->cfg.n_replications, t, e->sm_count, e->stream));
        ++e->launches;
        if (!e->after_end.empty() && e->after_end.back().first == 0) {      // max(max(now, a), b) = max(now, max(a, b))
            if (t > e->after_end.back().second) e->after_end.back().second = t;
        } else {
            e->after_end.emplace_back(0, t);
        }
        return CXXDES_B200_OK;
    }
    if (e->horizon > t) t = e->horizon;                  // now_ = max(now_, t), environment.ipp:194
    return launch(e, t);
}

int cxxdes_b200_run_for(cxxdes_b200_ensemble *e, int64_t dt) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    if (dt < 0) return fail(CXXDES_B200_ERR_INVALID, "run_for: dt must be >= 0");
    int64_t base = e->horizon < 0 ? 0 : e->horizon;
    if (base == INT64_MAX) {
        // exhausted: run_until(now() + dt) with each replication's own now() (environment.ipp:205-209)
        int rc = set_device(e);
        if (rc) return rc;
        if (!e->after_end.empty() && e->after_end.back().first == 1 && dt > INT64_MAX - e->after_end.back().second)
            return fail(CXXDES_B200_ERR_INVALID, "run_for: now() + dt overflows the 64-bit clock");
        CUDA_TRY(launch_clock_add(e->cols.end_time, e->cfg.n_replications, dt, e->sm_count, e->stream));
        ++e->launches;
        if (!e->after_end.empty() && e->after_end.back().first == 1) e->after_end.back().second += dt;   // (a loop of run_for calls stays one entry)
        else e->after_end.emplace_back(1, dt);
        return CXXDES_B200_OK;
    }
    if (dt > INT64_MAX - 1 - base) return fail(CXXDES_B200_ERR_INVALID, "run_for: now() + dt overflows the 64-bit clock");
    return cxxdes_b200_run_until(e, base + dt);
}

int cxxdes_b200_sync(cxxdes_b200_ensemble *e) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    int rc = set_device(e);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return CXXDES_B200_OK;
}

int cxxdes_b200_fetch(cxxdes_b200_ensemble *e, const cxxdes_b200_columns *h) {
    if (!e || !h) return fail(CXXDES_B200_ERR_INVALID, "handle/columns is NULL");
    if (!e->used) return fail(CXXDES_B200_ERR_STATE, "fetch before run");
    int rc = set_device(e);
    if (rc) return rc;
    const uint64_t n = e->cfg.n_replications;
    auto copy = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (!dst) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream);
    };
    CUDA_TRY(copy(h->end_time, e->cols.end_time, n * 8));
    CUDA_TRY(copy(h->n_events, e->cols.n_events, n * 8));
    CUDA_TRY(copy(h->trace_hash, e->cols.trace_hash, n * 8));
    CUDA_TRY(copy(h->status, e->cols.status, n * 4));
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) CUDA_TRY(copy(h->f64[k], e->cols.f64[k], n * 8));
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) CUDA_TRY(copy(h->i64[k], e->cols.i64[k], n * 8));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (h->status) {
        for (uint64_t r = 0; r < n; ++r)
            if (h->status[r] & ~(uint32_t)CXXDES_B200_REP_UNFINISHED)
                return fail(CXXDES_B200_ERR_CAPACITY, "replication %llu reported status 0x%x",
                            (unsigned long long)(e->cfg.first_replication + r), h->status[r]);
    }
    return CXXDES_B200_OK;
}

int cxxdes_b200_device_columns(cxxdes_b200_ensemble *e, cxxdes_b200_columns *d) {
    if (!e || !d) return fail(CXXDES_B200_ERR_INVALID, "handle/columns is NULL");
    d->end_time = e->cols.end_time;
    d->n_events = e->cols.n_events;
    d->trace_hash = e->cols.trace_hash;
    d->status = e->cols.status;
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) d->f64[k] = e->cols.f64[k];
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) d->i64[k] = e->cols.i64[k];
    return CXXDES_B200_OK;
}

int cxxdes_b200_reduce(cxxdes_b200_ensemble *e, cxxdes_b200_accumulators *host_out, void **device_acc) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    if (!e->used) return fail(CXXDES_B200_ERR_STATE, "reduce before run");
    int rc = set_device(e);
    if (rc) return rc;
    CUDA_TRY(launch_reduce(e->cols, e->cfg.n_replications, e->acc, e->reduce_scratch, e->sm_count, e->stream));
    e->launches += 2;
    if (device_acc) *device_acc = e->acc;
    if (host_out) {
        CUDA_TRY(cudaMemcpyAsync(host_out, e->acc, sizeof(*host_out), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return CXXDES_B200_OK;
}

// The final exchange of a multi-GPU run.  NCCL entry points are looked up at the first call, so that a process
// which never calls this (or which already carries its own NCCL, e.g. the one bundled with PyTorch) is untouched.
namespace {
struct nccl_api {
    int (*all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    const char *(*error_string)(int) = nullptr;
    bool ok = false;
};
const nccl_api &nccl() {
    static nccl_api api = [] {
        nccl_api a;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.all_reduce = (decltype(a.all_reduce))dlsym(h, "ncclAllReduce");
        a.group_start = (decltype(a.group_start))dlsym(h, "ncclGroupStart");
        a.group_end = (decltype(a.group_end))dlsym(h, "ncclGroupEnd");
        a.error_string = (decltype(a.error_string))dlsym(h, "ncclGetErrorString");
        a.ok = a.all_reduce && a.group_start && a.group_end && a.error_string;
        return a;
    }();
    return api;
}
constexpr int NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
static_assert(offsetof(cxxdes_b200_accumulators, f64_sum) == 7 * 8, "accumulator layout: 7 integer words first");
static_assert(sizeof(cxxdes_b200_accumulators) == 11 * 8, "accumulator layout: then 4 doubles");
}  // namespace

int cxxdes_b200_reduce_allgpus(cxxdes_b200_ensemble *e, void *nccl_comm, cxxdes_b200_accumulators *host_out) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    if (!nccl_comm) return fail(CXXDES_B200_ERR_INVALID, "reduce_allgpus: nccl_comm is NULL");
    if (!e->used) return fail(CXXDES_B200_ERR_STATE, "reduce before run");
    const nccl_api &api = nccl();
    if (!api.ok) return fail(CXXDES_B200_ERR_CUDA, "reduce_allgpus: libnccl.so.2 not found: %s", dlerror());
    int rc = set_device(e);
    if (rc) return rc;
    CUDA_TRY(launch_reduce(e->cols, e->cfg.n_replications, e->acc, e->reduce_scratch, e->sm_count, e->stream));
    e->launches += 2;
    // integer words sum exactly (the digest sum wraps, as on one GPU); the FP64 sums are order dependent at 1e-16
    unsigned char *acc = (unsigned char *)e->acc;
    int nrc = api.group_start();
    if (nrc == 0) nrc = api.all_reduce(acc, acc, 7, NCCL_UINT64, NCCL_SUM, nccl_comm, e->stream);
    if (nrc == 0) nrc = api.all_reduce(acc + 7 * 8, acc + 7 * 8, 4, NCCL_FLOAT64, NCCL_SUM, nccl_comm, e->stream);
    const int erc = api.group_end();
    if (nrc == 0) nrc = erc;
    if (nrc != 0) return fail(CXXDES_B200_ERR_CUDA, "reduce_allgpus: NCCL: %s", api.error_string(nrc));
    if (host_out) {
        CUDA_TRY(cudaMemcpyAsync(host_out, e->acc, sizeof(*host_out), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return CXXDES_B200_OK;
}

int cxxdes_b200_last_run_ms(cxxdes_b200_ensemble *e, float *ms) {
    if (!e || !ms) return fail(CXXDES_B200_ERR_INVALID, "handle/ms is NULL");
    if (!e->timed) return fail(CXXDES_B200_ERR_STATE, "no run recorded");
    int rc = set_device(e);
    if (rc) return rc;
    CUDA_TRY(cudaEventSynchronize(e->ev_end));
    CUDA_TRY(cudaEventElapsedTime(ms, e->ev_begin, e->ev_end));
    return CXXDES_B200_OK;
}

// Event trace of one replication: a one-replication shard re-run from tick 0 on the model's full-size kernel
// (64-bit ticks, generic heap) with a sink attached to environment::step.  Temporary device buffers, freed before
// returning; the handle's result columns and saved state are not touched.
int cxxdes_b200_trace(cxxdes_b200_ensemble *e, uint64_t replication, int64_t until, uint64_t max_events, int64_t *time,
                      int64_t *priority, uint32_t *kind, uint32_t *proc, uint64_t *n_written, uint64_t *n_events,
                      uint64_t *trace_hash) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    if (replication >= e->cfg.n_replications) return fail(CXXDES_B200_ERR_INVALID, "trace: replication out of range");
    if (max_events && (!time || !priority || !kind || !proc)) return fail(CXXDES_B200_ERR_INVALID, "trace: a column is NULL");
    int rc = set_device(e);
    if (rc) return rc;
    struct buffers {
        unsigned char *cols = nullptr, *rows = nullptr;
        int64_t *ring = nullptr;
        unsigned long long *counters = nullptr;
        ~buffers() { cudaFree(cols); cudaFree(rows); cudaFree(ring); cudaFree(counters); }
    } b;
    const size_t rows_bytes = (max_events ? max_events : 1) * 24;
    CUDA_TRY(cudaMalloc(&b.cols, 8 * 256));
    CUDA_TRY(cudaMalloc(&b.rows, rows_bytes));
    CUDA_TRY(cudaMalloc(&b.counters, 8 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(b.cols, 0, 8 * 256, e->stream));
    CUDA_TRY(cudaMemsetAsync(b.counters, 0, 8 * sizeof(unsigned long long), e->stream));
    device_columns out{};
    out.end_time = (int64_t *)(b.cols + 0 * 256);
    out.n_events = (uint64_t *)(b.cols + 1 * 256);
    out.trace_hash = (uint64_t *)(b.cols + 2 * 256);
    out.status = (uint32_t *)(b.cols + 3 * 256);
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) out.f64[k] = (double *)(b.cols + (4 + k) * 256);
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) out.i64[k] = (int64_t *)(b.cols + (6 + k) * 256);
    trace_sink sink{};
    sink.time = (int64_t *)b.rows;
    sink.priority = sink.time + max_events;
    sink.kind = (uint32_t *)(sink.priority + max_events);
    sink.proc = sink.kind + max_events;
    sink.max_events = max_events;
    shard sh{};
    sh.seed = e->cfg.seed;
    sh.first_replication = e->cfg.first_replication + replication;
    sh.n_replications = 1;
    sh.clk = e->clk;
    sh.run_until = until < 0 ? -1 : until;
    sh.work_counter = b.counters + 0;
    sh.retry_count = b.counters + 1;
    sh.retry_list = (uint32_t *)(b.counters + 6);
    sh.retry_a_count = b.counters + 3;
    sh.retry_a_cursor = b.counters + 4;
    sh.retry_a_list = (uint32_t *)(b.counters + 7);
    sh.log_table = e->log_table;
    switch (e->cfg.model) {
        case CXXDES_B200_MODEL_MM1: {
            uint64_t cap = e->params.mm1.n_packets < 4096 ? e->params.mm1.n_packets : 4096;   // sync::queue ring of the engine kernel
            if (cap < 2) cap = 2;
            CUDA_TRY(cudaMalloc(&b.ring, cap * 32 * sizeof(int64_t)));
            CUDA_TRY(launch_mm1_trace(e->params.mm1, sh, out, b.ring, (uint32_t)cap, sink, e->stream));
            break;
        }
        case CXXDES_B200_MODEL_ALOHA:
            CUDA_TRY(launch_aloha_trace(e->params.aloha, sh, out, sink, e->stream));
            break;
        case CXXDES_B200_MODEL_MMC:
            CUDA_TRY(launch_mmc_trace(e->params.mmc, sh, out, (uint64_t *)e->ring, sink, e->stream, e->mmc_deep_pass));
            break;
        case CXXDES_B200_MODEL_ARCHSIM:
            CUDA_TRY(launch_archsim_trace(e->params.arch, sh, out, sink, e->stream));
            break;
        case CXXDES_B200_MODEL_PROGRAM:
            CUDA_TRY(launch_program_trace(e->dprog, e->program_blob_bytes, sh, out, sink, e->stream, e->prog_deep_pass));
            break;
        default:
            return fail(CXXDES_B200_ERR_INVALID, "unknown model");
    }
    ++e->launches;
    uint64_t total = 0, digest = 0;
    uint32_t status = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, out.n_events, 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(&digest, out.trace_hash, 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(&status, out.status, 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    const uint64_t n = total < max_events ? total : max_events;
    if (n) {
        CUDA_TRY(cudaMemcpy(time, sink.time, n * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(priority, sink.priority, n * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(kind, sink.kind, n * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(proc, sink.proc, n * 4, cudaMemcpyDeviceToHost));
    }
    if (n_written) *n_written = n;
    if (n_events) *n_events = total;
    if (trace_hash) *trace_hash = digest;
    if (status & ~(uint32_t)CXXDES_B200_REP_UNFINISHED)
        return fail(CXXDES_B200_ERR_CAPACITY, "trace: replication %llu reported status 0x%x",
                    (unsigned long long)sh.first_replication, status);
    return CXXDES_B200_OK;
}

// ALOHA, M/M/c, memory hierarchy: the first-pass kernel follows from the parameters, the clock and the flags alone.
static void host_planned_kernel_name(int model, const model_params &params, const shard &sh, uint32_t flags, char *buf,
                                     size_t buf_size) {
    const bool wide = (flags & CXXDES_B200_FLAG_WIDE_TOKENS) != 0;
    switch (model) {
        case CXXDES_B200_MODEL_ALOHA:
            snprintf(buf, buf_size, !aloha_keeps_state(params.aloha, sh) ? "aloha_kernel"
                                    : (flags & CXXDES_B200_FLAG_COOP_HEAP) ? "aloha_coop_kernel"
                                    : (!wide && aloha_compact_applies(params.aloha, sh)) ? "aloha_compact_kernel<false>"
                                                                                         : "aloha_fast_kernel<false>");
            break;
        case CXXDES_B200_MODEL_MMC:
            snprintf(buf, buf_size, !mmc_keeps_state(params.mmc, sh) ? "mmc_kernel"
                                    : (!wide && mmc_narrow_applies(params.mmc, sh)) ? "mmc_fast_kernel<false,narrow>"
                                                                                    : "mmc_fast_kernel<false,wide>");
            break;
        default:
            snprintf(buf, buf_size, "archsim_kernel<%s%s>", params.arch.n_requesters == 1 ? "1 requester" : "k requesters",
                     archsim_uses_narrow_layout(params.arch, wide) ? ",narrow" : "");
            break;
    }
}

// The same decision without a handle or a device (include/cxxdes_b200.h).
int cxxdes_b200_plan_kernel(const cxxdes_b200_config *cfg, char *buf, size_t buf_size) {
    if (!cfg || !buf || !buf_size) return fail(CXXDES_B200_ERR_INVALID, "cfg/buf is NULL");
    const size_t want = params_size_of(cfg->model);
    if (want == 0) return fail(CXXDES_B200_ERR_INVALID, "unknown model %d", cfg->model);
    if (!cfg->params || cfg->params_size != want)
        return fail(CXXDES_B200_ERR_INVALID, "params_size %zu does not match model %d (%zu)", cfg->params_size, cfg->model, want);
    if (cfg->model != CXXDES_B200_MODEL_ALOHA && cfg->model != CXXDES_B200_MODEL_MMC && cfg->model != CXXDES_B200_MODEL_ARCHSIM)
        return fail(CXXDES_B200_ERR_STATE, "the kernel of this model is planned from the device's resources: create a handle and ask cxxdes_b200_kernel_name");
    const cxxdes_b200_clock &ck = cfg->clock;
    if (ck.unit_u < 0 || ck.unit_u > 4 || ck.precision_u < 0 || ck.precision_u > 4 || ck.precision_t <= 0 || ck.unit_t <= 0)
        return fail(CXXDES_B200_ERR_INVALID, "invalid clock");
    model_params params;
    memcpy(&params, cfg->params, want);
    shard sh{};
    sh.clk.ticks_per_unit = time_count(ck.unit_t, ck.unit_u, ck.precision_t, ck.precision_u);
    sh.clk.tick_mult = ck.precision_t;
    static const double conv[5] = {1, 1e-3, 1e-6, 1e-9, 1e-12};
    sh.clk.seconds_per_tick_unit = conv[ck.precision_u];
    sh.run_until = -1;
    host_planned_kernel_name(cfg->model, params, sh, cfg->flags, buf, buf_size);
    return CXXDES_B200_OK;
}

// Which first-pass kernel instantiation the handle's runs use (profiles are keyed by this string).
int cxxdes_b200_kernel_name(cxxdes_b200_ensemble *e, char *buf, size_t buf_size) {
    if (!e || !buf || !buf_size) return fail(CXXDES_B200_ERR_INVALID, "handle/buf is NULL");
    shard sh{};
    sh.clk = e->clk;
    sh.run_until = -1;
    switch (e->cfg.model) {
        case CXXDES_B200_MODEL_MM1:
            if (!e->use_fused) snprintf(buf, buf_size, "mm1_engine_kernel");
            else if (e->cfg.flags & CXXDES_B200_FLAG_MM1_LOCKSTEP)
                snprintf(buf, buf_size, e->mm1_plan.direct_global ? "mm1_fused_kernel<false,128,1>" : "mm1_fused_kernel<true,256,4>");
            else if (e->mm1_plan.ah_direct_global) snprintf(buf, buf_size, "mm1_fused_kernel<false,128,1>");
            else {
                // <THREADS, MIN_BLOCKS, RING_SLOTS, EVENT_SLOTS, MODE, SPILL, STATE> of a plain run()
                const mm1_fused_plan &p = e->mm1_plan;
                const int mode = (e->clk.tick_mult != 1 ? 2 : 0) | (p.ah_wide ? 4 : 0);
                snprintf(buf, buf_size, "mm1_ahead_kernel<%d,%d,%u,4,%d,%d,0>%s", p.ah_threads, p.ah_blocks / e->sm_count,
                         p.ah_ring_slots, mode, p.ah_spill ? 1 : 0, (!p.ah_spill && e->helper_stream) ? "+helper<32,16,32,4>" : "");
            }
            break;
        case CXXDES_B200_MODEL_ALOHA:
        case CXXDES_B200_MODEL_MMC:
        case CXXDES_B200_MODEL_ARCHSIM:
            host_planned_kernel_name(e->cfg.model, e->params, sh, e->cfg.flags, buf, buf_size);
            break;
        case CXXDES_B200_MODEL_PROGRAM:
            snprintf(buf, buf_size, e->prog_fast.valid ? "program_fast_kernel<false>"
                                    : program_needs_last_pass(e->dprog) ? "program_kernel<last pass>" : "program_kernel");
            break;
        default:
            return fail(CXXDES_B200_ERR_INVALID, "unknown model");
    }
    return CXXDES_B200_OK;
}

int cxxdes_b200_pass_counts(cxxdes_b200_ensemble *e, uint64_t counts[4]) {
    if (!e || !counts) return fail(CXXDES_B200_ERR_INVALID, "handle/counts is NULL");
    if (!e->used) return fail(CXXDES_B200_ERR_STATE, "pass_counts before run");
    int rc = set_device(e);
    if (rc) return rc;
    unsigned long long c[8];
    CUDA_TRY(cudaMemcpyAsync(c, e->counters, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    counts[0] = c[3];
    counts[1] = c[6] & 0xFFFFFFFFull;
    counts[2] = c[1];
    counts[3] = c[7];
    return CXXDES_B200_OK;
}

int cxxdes_b200_launch_count(cxxdes_b200_ensemble *e, uint64_t *launches) {
    if (!e || !launches) return fail(CXXDES_B200_ERR_INVALID, "handle/launches is NULL");
    *launches = e->launches;
    return CXXDES_B200_OK;
}

int cxxdes_b200_reset(cxxdes_b200_ensemble *e) {
    if (!e) return fail(CXXDES_B200_ERR_INVALID, "handle is NULL");
    e->horizon = -1;
    e->end_floor = -1;
    e->after_end.clear();
    e->used = false;
    e->timed = false;
    e->mm1_state_valid = false;
    e->saved_valid = false;
    return CXXDES_B200_OK;
}

int cxxdes_b200_destroy(cxxdes_b200_ensemble *e) {
    if (!e) return CXXDES_B200_OK;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->col_block);
    cudaFree(e->counters);
    cudaFree(e->retry_list);
    cudaFree(e->log_table);
    cudaFree(e->ring);
    cudaFree(e->ahead_gring);
    cudaFree(e->mm1_state);
    cudaFree(e->acc);
    cudaFree(e->reduce_scratch);
    cudaFree(e->program_blob);
    cudaFree(e->prog_queue_items);
    cudaFree(e->saved_words);
    cudaFree(e->saved_queue_items);
    cudaFree(e->mmc_deep_pass.base);
    cudaFree(e->prog_deep_pass.base);
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->helper_stream) { cudaStreamSynchronize(e->helper_stream); cudaStreamDestroy(e->helper_stream); }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return CXXDES_B200_OK;
}

}  // extern "C"
