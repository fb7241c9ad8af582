// kernels.h -- host-callable launchers of the model kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>

#include "device/models_common.cuh"
#include "device/trace.cuh"

namespace cxb {

// ---- M/M/1 on the generic engine (mm1_engine.cu) ----
size_t mm1_engine_smem_bytes(int threads);
cudaError_t launch_mm1_engine(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                              int64_t *ring, uint32_t ring_cap, int blocks, int threads, cudaStream_t stream);

// ---- fused M/M/1 (mm1_fused.cu) ----
struct mm1_fused_plan {
    int threads;          // threads per block
    int blocks;           // persistent grid
    uint32_t smem_ring;   // ring entries per replication in shared memory (fast kernel)
    size_t smem_bytes;
    int fb_threads, fb_blocks;  // fallback kernel (global-memory ring, 64-bit ticks)
    uint32_t fb_ring;           // ring entries per fallback thread
    bool direct_global;         // skip the shared-memory kernel (high utilisation / huge queues)
    // generate-ahead kernel (mm1_ahead.cu), the default first pass
    int ah_geometry;            // which compiled (threads, blocks/SM, ring slots) variant
    int ah_threads, ah_blocks;
    uint32_t ah_ring_slots;
    size_t ah_smem_bytes;
    uint32_t ah_ring_cap;       // packets (queued + generated ahead) a replication's ring may hold
    uint32_t ah_lead_cap;       // packets that may be generated ahead of the producer
    bool ah_spill;              // the FIRST pass already uses the two-level ring (heavily loaded queues)
    uint32_t ah_spill_cap;      // packets the two-level ring may hold
    uint32_t ah_gring_slots;    // its slots per replication (power of two)
    bool ah_wide;               // the clock passes 2^30 ticks: lanes keep times relative to an epoch (MODE_WIDE)
    bool ah_direct_global;
};
mm1_fused_plan plan_mm1_fused(const cxxdes_b200_mm1_params &prm, uint32_t queue_capacity, int sm_count,
                              size_t smem_per_block_optin, int64_t tick_mult, int64_t ticks_per_unit, bool lockstep,
                              uint64_t n_replications);
// What the second pass needs to run NEXT TO the first one instead of after it (helper_stream == nullptr: after it).
struct mm1_overlap {
    cudaStream_t helper_stream;     // a stream of the handle's own
    cudaEvent_t fork, join;         // main stream -> helper stream before the first pass, and back after the helper
    unsigned int *done_flag;        // device word: 0 while the first pass runs, non-zero after it; done_flag[2]
                                    // counts the replications the helper took
};
cudaError_t launch_mm1_fused(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                             const mm1_fused_plan &plan, int64_t *fb_ring, unsigned long long *fb_counter,
                             void *ahead_gring, uint32_t *ahead_state, int resume, cudaStream_t stream, int *launches,
                             bool lockstep, const mm1_overlap &overlap);
cudaError_t configure_mm1_fused(const mm1_fused_plan &plan);

// ---- fused M/M/1, variates generated ahead of the event loop (mm1_ahead.cu) ----
void plan_mm1_ahead(mm1_fused_plan &plan, const cxxdes_b200_mm1_params &prm, uint32_t queue_capacity, int sm_count,
                    int64_t tick_mult, int64_t ticks_per_unit, uint64_t n_replications);
cudaError_t configure_mm1_ahead(const mm1_fused_plan &plan);
size_t mm1_ahead_gring_bytes(const mm1_fused_plan &plan);
size_t mm1_ahead_state_bytes(uint64_t n_replications);   // run_until in pieces: 128 B per replication
cudaError_t launch_mm1_ahead(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                             const mm1_fused_plan &plan, bool spill, void *gring, uint32_t *state, int resume,
                             const uint32_t *list, const unsigned long long *list_count, unsigned long long *list_cursor,
                             uint32_t *fail_list, unsigned long long *fail_count, cudaStream_t stream, int leave_blocks = 0);

cudaError_t launch_mm1_ahead_helper(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                                    const mm1_fused_plan &plan, void *gring, const uint32_t *list, uint32_t list_capacity,
                                    unsigned long long *list_cursor, unsigned int *done_flag, uint32_t *fail_list,
                                    unsigned long long *fail_count, cudaStream_t stream);

// ---- ALOHA (models_aloha.cu) ----
size_t aloha_smem_bytes(uint32_t num_stations);
struct aloha_saved {          // run_until / run_for in pieces (nullptr words = the plain run() kernel)
    uint32_t *words;          // [word][replication], word 0 = phase
    uint64_t n_replications;
    int resume;               // 0: every replication starts from tick 0 (first piece, or after reset)
};
size_t aloha_saved_words(const cxxdes_b200_aloha_params &prm);
bool aloha_keeps_state(const cxxdes_b200_aloha_params &prm, const shard &sh);   // the 32-bit kernel runs the model
bool aloha_compact_applies(const cxxdes_b200_aloha_params &prm, const shard &sh);   // ... on 4-byte tokens
cudaError_t launch_aloha(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                         int sm_count, size_t smem_optin, unsigned long long *list_cursor, cudaStream_t stream,
                         int *launches, const aloha_saved &st, bool coop, bool wide_tokens);
// first pass with sub-warp cooperative heaps (models_aloha_coop.cu); the caller runs the retry list afterwards
cudaError_t launch_aloha_coop(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                              int sm_count, cudaStream_t stream);

// ---- M/M/c with balking (models_mmc.cu) ----
int mmc_blocks(int sm_count);
int mmc_threads();
struct mmc_saved {            // run_until / run_for in pieces (nullptr words = the plain run() kernel)
    uint32_t *words;          // [word][replication], word 0 = phase
    uint64_t n_replications;
    int resume;               // 0: every replication starts from tick 0 (first piece, or after reset)
};
bool mmc_narrow_applies(const cxxdes_b200_mmc_params &prm, const shard &sh);   // first pass on 4-byte tokens
bool mmc_keeps_state(const cxxdes_b200_mmc_params &prm, const shard &sh);   // the 32-bit kernel runs the model
size_t mmc_saved_words(const cxxdes_b200_mmc_params &prm);
// The last pass: heap and wait list of a replication in global memory, sized for the model's worst case (what the
// reference's unbounded containers could hold).  base == nullptr: no such pass.
struct mmc_deep {
    unsigned char *base;
    int threads;                // of the deep grid; its arrays are laid out [entry][thread]
    int heap_cap, waiter_cap;
};
size_t mmc_deep_plan(const cxxdes_b200_mmc_params &prm, int sm_count, size_t budget_bytes, mmc_deep &d);   // bytes to allocate
cudaError_t launch_mmc(const cxxdes_b200_mmc_params &prm, const shard &sh, const device_columns &out,
                       uint64_t *customer_state, int sm_count, unsigned long long *list_cursor, cudaStream_t stream,
                       int *launches, const mmc_saved &st, bool wide_tokens, const mmc_deep &deep);

// ---- memory hierarchy (models_archsim.cu) ----
bool archsim_supported(const cxxdes_b200_archsim_params &p);
size_t archsim_smem_bytes(const cxxdes_b200_archsim_params &p);
struct archsim_saved {        // run_until / run_for in pieces (nullptr words = the plain run() kernel)
    uint32_t *words;          // [word][replication], word 0 = phase
    uint64_t n_replications;
    int resume;               // 0: every replication starts from tick 0 (first piece, or after reset)
};
size_t archsim_saved_words(const cxxdes_b200_archsim_params &p);
bool archsim_uses_narrow_layout(const cxxdes_b200_archsim_params &prm, bool wide_layout);   // 32-bit stamps, 16-bit addresses
cudaError_t launch_archsim(const cxxdes_b200_archsim_params &prm, const shard &sh, const device_columns &out,
                           int sm_count, cudaStream_t stream, const archsim_saved &st, bool wide_layout);

// ---- process programs (program_kernel.cu) ----
struct device_program {  // cxxdes_b200_program with device pointers; rates as {d, RN(1/d)}
    const cxxdes_b200_op *ops;
    const cxxdes_b200_process *procs;
    const uint32_t *roots;
    const uint64_t *queue_max;
    const uint64_t *sem_init;
    const uint64_t *sem_max;
    const divisor *rates;
    uint32_t n_ops, n_procs, n_roots, n_queues, n_sems, n_mutexes, n_events, n_rates;
    uint32_t sweep_steps;   // rows of `rates` (0 = 1): replication r draws with row (first_replication + r) % sweep_steps
};
const char *program_check(const cxxdes_b200_program &p);  // nullptr when the interpreter supports it
// First pass with the per-replication state in shared memory (program_fast.cu): word offsets sized from the program.
struct program_fast_layout {
    bool valid;                 // false: program_kernel.cu alone runs the program
    int heap_cap;               // tokens a replication may have scheduled at once
    int n_handlers;             // any_of / all_of handler pool
    int n_words, n_dwords;      // 32-bit / 64-bit state words per replication
    int w_serial, w_mtx, w_handlers, w_queues, w_sems, w_procs, w_qwait, w_swait, w_mwait, w_ewait;
    int pw, o_draws, o_loop, loop_words;      // words per process, offsets of the optional ones
    int wait_cap, wait_words, ewait_words;    // wait lists: count word + entries
    int d_regs;                 // first 64-bit word of the per-process registers
    int queue_cap;              // items per queue (power of two), payloads in global memory
    uint32_t staged_bytes;      // program text staged into shared memory
    size_t smem_bytes;
    int blocks_per_sm;
    size_t queue_item_bytes_per_thread;
};
// run_until / run_for in pieces: replications saved between launches (nullptr words = the plain run() kernel)
struct program_fast_state {
    uint32_t *words;          // [word][replication], word 0 = phase
    int64_t *queue_items;     // [queue][slot][replication]
    uint64_t n_replications;
    int resume;               // 0: every replication starts from tick 0 (first piece, or after reset)
};
size_t program_fast_state_words(const program_fast_layout &L);
size_t program_fast_state_queue_bytes(const program_fast_layout &L, uint64_t n_replications);
program_fast_layout program_fast_plan(const cxxdes_b200_program &host_program, uint32_t blob_bytes, size_t smem_optin,
                                      uint32_t queue_capacity);
size_t program_fast_queue_bytes(const program_fast_layout &L, int sm_count);
cudaError_t launch_program_fast(const device_program &prg, const program_fast_layout &L, const shard &sh,
                                const device_columns &out, int64_t *queue_items, const program_fast_state &st,
                                int sm_count, cudaStream_t stream);
// blob_bytes: size of the contiguous device copy of the program that starts at prg.ops
// The last pass of the interpreter: heap, wait lists and queue rings of a replication in global memory (a slab per
// thread).  base == nullptr: no such pass.
struct program_deep {
    unsigned char *base;
    int threads;                       // of the deep grid
    int heap_cap, waiter_cap, queue_cap;
    size_t bytes_per_thread;
};
size_t program_deep_plan(const device_program &prg, int sm_count, size_t budget_bytes, program_deep &d);   // bytes to allocate
bool program_needs_last_pass(const device_program &prg);   // more than 4 sync objects of a kind: program_kernel's own arrays cannot hold it
cudaError_t launch_program(const device_program &prg, uint32_t blob_bytes, const shard &sh, const device_columns &out,
                           int sm_count, unsigned long long *list_cursor, cudaStream_t stream, int *launches,
                           const program_fast_layout &fast, int64_t *queue_items, const program_fast_state &fast_state, const program_deep &deep);

// ---- reduction of result columns to accumulators (reduce.cu) ----
size_t reduce_scratch_bytes(int sm_count);
cudaError_t launch_reduce(const device_columns &cols, uint64_t n, cxxdes_b200_accumulators *device_acc,
                          void *scratch, int sm_count, cudaStream_t stream);


// ---- event trace of one replication (cxxdes_b200_trace): sh.n_replications == 1, full-size kernels ----
cudaError_t launch_mm1_trace(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out, int64_t *ring,
                             uint32_t ring_cap, const trace_sink &sink, cudaStream_t stream);
cudaError_t launch_aloha_trace(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                               const trace_sink &sink, cudaStream_t stream);
cudaError_t launch_mmc_trace(const cxxdes_b200_mmc_params &prm, const shard &sh, const device_columns &out,
                             uint64_t *customer_state, const trace_sink &sink, cudaStream_t stream, const mmc_deep &deep);
cudaError_t launch_archsim_trace(const cxxdes_b200_archsim_params &prm, const shard &sh, const device_columns &out,
                                 const trace_sink &sink, cudaStream_t stream);
cudaError_t launch_program_trace(const device_program &prg, uint32_t blob_bytes, const shard &sh, const device_columns &out,
                                 const trace_sink &sink, cudaStream_t stream, const program_deep &deep);

cudaError_t launch_clock_floor(int64_t *end_time, uint64_t n, int64_t floor, int sm_count, cudaStream_t stream);
cudaError_t launch_clock_add(int64_t *end_time, uint64_t n, int64_t dt, int sm_count, cudaStream_t stream);

}  // namespace cxb
