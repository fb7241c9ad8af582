// cxxdes_b200/program.hpp -- host C++ builder for process programs (cxxdes_b200_program).
//
// A libcxxdes model is a set of coroutine bodies made of `co_await <awaitable>` statements.  A model that
// has no hand-written kernel is handed to the B200 engine as a *process program*: the same bodies, written
// once more as calls on a `body` object -- one call per co_await, with the reference's own names for the
// awaitables -- and interpreted on the device for every replication (csrc/program_fast.cu, program_kernel.cu).
//
//     reference (examples/producer_consumer.cpp:31-58)        this header
//     -----------------------------------------------        -----------------------------------------------
//     sync::queue<double> q;                                  queue q = prog.add_queue();
//     coroutine<> producer() {                                prog.body_of(producer, [&](body &co) {
//       for (i = 0; i < n_packets; ++i) {                       co.loop(n_packets, [&] {
//         co_await q.put(now_seconds());                          co.await(q.put());
//         co_await env.timeout(lambda()); } }                     co.await(env_timeout(exponential{lam})); }); });
//     coroutine<> co_main() {                                 prog.body_of(co_main, [&](body &co) {
//       co_await (producer() && consumer()); }                  co.await(producer && consumer); });
//
// `program` only records; nothing here runs a simulation on the host (there is no CPU path).  Give the
// program to ensemble<models::process_program> (ensemble.hpp) to run it.  Errors are std::runtime_error /
// std::invalid_argument as in the reference (environment.ipp:44-45).  Header-only, C++17, plain g++.
#pragma once
#include <cstring>
#include <initializer_list>
#include <memory>
#include <limits>
#include <utility>

#include "cxxdes_b200/ensemble.hpp"

namespace cxxdes {
namespace b200 {

// priority_consts::inherit (defs.ipp:31-40): "take the priority of the awaiting process"
constexpr priority_type inherit = std::numeric_limits<priority_type>::max();

class program;
class body;

namespace detail {
inline std::int32_t prio32(priority_type p) {
    if (p == inherit) return CXXDES_B200_INHERIT;
    if (p < std::numeric_limits<std::int32_t>::min() || p >= CXXDES_B200_INHERIT)
        throw std::invalid_argument("cxxdes_b200: explicit awaitable priorities must fit 32 bits");
    return (std::int32_t)p;
}
}  // namespace detail

// ---- awaitable descriptors (what the reference's factory functions return) ----------------------------
struct delay_t {  // delay(t) / timeout(t) / yield(): timeout.ipp:106-143,180-182
    time_integral ticks;
    priority_type prio = inherit;
    delay_t priority(priority_type p) const { return {ticks, p}; }
};
struct until_t {  // until(t): timeout.ipp:14-29
    time_integral ticks;
    priority_type prio = inherit;
    until_t priority(priority_type p) const { return {ticks, p}; }
};
inline delay_t delay(time_integral ticks) { return {ticks}; }
inline delay_t timeout(time_integral ticks) { return {ticks}; }
inline delay_t yield() { return {0}; }
inline until_t until(time_integral ticks) { return {ticks}; }
inline until_t instant(time_integral ticks) { return {ticks}; }  // one functor in the reference (timeout.ipp:101)
struct lazy_timeout_t {  // lazy_timeout(t) / lazy_delay(t): timeout.ipp:104-172
    time_integral ticks;
};
inline lazy_timeout_t lazy_timeout(time_integral ticks) { return {ticks}; }
inline lazy_timeout_t lazy_delay(time_integral ticks) { return {ticks}; }
struct lazy_instant_t {  // what `co_await lazy_timeout(t)` returned: the instant kept in the process register
    priority_type prio = inherit;
    lazy_instant_t priority(priority_type p) const { return {p}; }
};
constexpr lazy_instant_t lazy_instant{};

struct exponential {  // examples/random_variable.hpp:17-40, drawn from the Philox stream of the awaiting process
    double rate;
    std::vector<double> per_point;   // a swept rate: one value per sweep point (program::sweep)
    exponential(double r) : rate(r) {}
    exponential(std::vector<double> points) : rate(0), per_point(std::move(points)) {}
};
struct env_timeout_t {  // env.timeout(double seconds) / env.timeout(random variable): timeout.ipp:184-187
    bool random;
    double value;
    std::vector<double> per_point;
};
inline env_timeout_t env_timeout(double seconds) { return {false, seconds, {}}; }
inline env_timeout_t env_timeout(exponential rv) { return {true, rv.rate, std::move(rv.per_point)}; }

struct process {  // a coroutine<> object of the model
    std::uint32_t id;
};
struct async_t {  // async(p): async.ipp:9-49
    process p;
};
inline async_t async(process p) { return {p}; }

struct composition;
struct simple_op;
struct child {  // an operand of all_of / any_of / && / ||: a process, a timer, or another composition
    enum kind_t { PROC, DELAY, UNTIL, NESTED, EV_WAIT } kind;
    std::uint32_t proc;
    time_integral ticks;
    priority_type prio;
    std::shared_ptr<const composition> nested;
    child(process p) : kind(PROC), proc(p.id), ticks(0), prio(inherit) {}
    child(delay_t d) : kind(DELAY), proc(0), ticks(d.ticks), prio(d.prio) {}
    child(until_t u) : kind(UNTIL), proc(0), ticks(u.ticks), prio(u.prio) {}
    child(const composition &c);
    child(const simple_op &op);          // evt.wait(latency, priority): `evt.wait() || timeout(t)`
};
struct composition {  // any_of.ipp:233-247; compositions.ipp (operator&&, operator||)
    bool all;
    std::vector<child> children;
    priority_type prio = inherit;      // .priority(p): what the children inherit (any_of.ipp:30-33,41-48)
    time_integral latency = 0;         // .return_latency(l): added to the token that resumes the awaiting process
    composition priority(priority_type p) const {
        composition c = *this;
        c.prio = p;
        return c;
    }
    composition return_latency(time_integral l) const {
        composition c = *this;
        c.latency = l;
        return c;
    }
};
template <typename... C>
composition all_of(C... c) { return composition{true, {child(c)...}}; }
template <typename... C>
composition any_of(C... c) { return composition{false, {child(c)...}}; }
inline composition all_of_range(const std::vector<process> &ps) {  // all_of.range(first, last)
    composition c{true, {}};
    for (process p : ps) c.children.emplace_back(p);
    return c;
}
inline composition any_of_range(const std::vector<process> &ps) {
    composition c{false, {}};
    for (process p : ps) c.children.emplace_back(p);
    return c;
}
inline child::child(const composition &c) : kind(NESTED), proc(0), ticks(0), prio(inherit), nested(std::make_shared<composition>(c)) {}
// a && b, a || b (compositions.ipp).  A chain nests exactly as in the reference: `a && b && c` is
// all_of(all_of(a, b), c) -- the inner composition has a handler of its own, which is an event more than the flat
// all_of(a, b, c).  Write all_of(a, b, c) when the flat form is meant.
inline composition operator&&(child a, child b) { return composition{true, {a, b}}; }
inline composition operator||(child a, child b) { return composition{false, {a, b}}; }
inline composition operator&&(const composition &a, child b) { return composition{true, {child(a), b}}; }
inline composition operator||(const composition &a, child b) { return composition{false, {child(a), b}}; }
inline composition operator&&(const composition &a, const composition &b) { return composition{true, {child(a), child(b)}}; }
inline composition operator||(const composition &a, const composition &b) { return composition{false, {child(a), child(b)}}; }
inline composition operator&&(child a, const composition &b) { return composition{true, {a, child(b)}}; }
inline composition operator||(child a, const composition &b) { return composition{false, {a, child(b)}}; }

// ---- conditions of if_then / if_then_else: plain reads of the model's own state (no events) --------------
struct condition {
    std::uint16_t a;     // index | source << 8 | comparison << 12 (cxxdes_b200.h)
    std::int64_t value;
};
struct operand {         // something an `if` can look at; compare it with an integer to get a condition
    std::uint16_t src, index;
    condition cmp(int c, std::int64_t v) const { return {(std::uint16_t)(index | (src << 8) | (c << 12)), v}; }
    condition operator==(std::int64_t v) const { return cmp(CXXDES_B200_CMP_EQ, v); }
    condition operator!=(std::int64_t v) const { return cmp(CXXDES_B200_CMP_NE, v); }
    condition operator<(std::int64_t v) const { return cmp(CXXDES_B200_CMP_LT, v); }
    condition operator<=(std::int64_t v) const { return cmp(CXXDES_B200_CMP_LE, v); }
    condition operator>(std::int64_t v) const { return cmp(CXXDES_B200_CMP_GT, v); }
    condition operator>=(std::int64_t v) const { return cmp(CXXDES_B200_CMP_GE, v); }
};
inline operand counter(int k) { return {CXXDES_B200_SRC_COUNTER, (std::uint16_t)k}; }   // i64[k], what co.count(k) adds to
inline operand now() { return {CXXDES_B200_SRC_NOW, 0}; }                                // env.now()
inline operand process_register() { return {CXXDES_B200_SRC_REGISTER, 0}; }

// sync objects hand out awaitable descriptors the way sync::queue / semaphore / mutex / event do
struct simple_op {
    std::uint16_t code, a;
    std::int32_t b;
    std::int64_t imm;
};
inline child::child(const simple_op &op) : kind(EV_WAIT), proc(op.a), ticks(op.imm), prio(op.b == CXXDES_B200_INHERIT ? inherit : (priority_type)op.b) {
    if (op.code != CXXDES_B200_OP_EV_WAIT)
        throw std::invalid_argument("cxxdes_b200: of the sync operations only evt.wait() can be a child of all_of / any_of");
}
struct queue {  // sync::queue<T>: sync/queue.hpp:38-90
    std::uint16_t id;
    simple_op put() const { return {CXXDES_B200_OP_Q_PUT, id, 0, 0}; }           // q.put(now)
    simple_op put_register() const { return {CXXDES_B200_OP_Q_PUT, id, 1, 0}; }  // q.put(<process register>)
    simple_op pop() const { return {CXXDES_B200_OP_Q_POP, id, 0, 0}; }           // register = q.pop()
    operand size() const { return {CXXDES_B200_SRC_QUEUE_SIZE, id}; }            // q.size() in a condition
};
struct semaphore {  // sync::semaphore: sync/semaphore.hpp:58-78
    std::uint16_t id;
    simple_op up() const { return {CXXDES_B200_OP_SEM_UP, id, 0, 0}; }
    simple_op down() const { return {CXXDES_B200_OP_SEM_DOWN, id, 0, 0}; }
    operand value() const { return {CXXDES_B200_SRC_SEM_VALUE, id}; }            // sem.value() in a condition
};
struct resource {  // sync::resource{count}: acquire = down, handle.release = up (sync/resource.hpp:37-100)
    std::uint16_t id;
    simple_op acquire() const { return {CXXDES_B200_OP_SEM_DOWN, id, 0, 0}; }
    simple_op release() const { return {CXXDES_B200_OP_SEM_UP, id, 0, 0}; }
    operand available() const { return {CXXDES_B200_SRC_SEM_VALUE, id}; }        // free units in a condition
};
struct mutex {  // sync::mutex: sync/mutex.hpp:31-109
    std::uint16_t id;
    simple_op acquire() const { return {CXXDES_B200_OP_MTX_ACQUIRE, id, 0, 0}; }
    simple_op release() const { return {CXXDES_B200_OP_MTX_RELEASE, id, 0, 0}; }
    operand is_acquired() const { return {CXXDES_B200_SRC_MUTEX, id}; }          // m.is_acquired() in a condition (0 / 1)
};
struct event {  // sync::event: sync/event.hpp:87-139
    std::uint16_t id;
    simple_op wait(time_integral latency = 0, priority_type prio = inherit) const {
        return {CXXDES_B200_OP_EV_WAIT, id, detail::prio32(prio), latency};
    }
    simple_op wake() const { return {CXXDES_B200_OP_EV_WAKE, id, 0, 0}; }
};

// ---- the program ---------------------------------------------------------------------------------------
class program {
public:
    struct options {  // coroutine<>::priority / latency / return_latency / return_priority (coroutine.ipp:106-176)
        priority_type priority = inherit;
        time_integral latency = 0;
        time_integral return_latency = 0;
        priority_type return_priority = inherit;
    };

    // A coroutine<> object.  `ordinal` names it in the trace digest and selects its Philox stream.
    process add_process(std::uint32_t ordinal, options o) {
        procs_.push_back({NO_ENTRY, ordinal, o.priority, o.latency, o.return_latency, o.return_priority});
        return {(std::uint32_t)procs_.size() - 1};
    }
    process add_process(std::uint32_t ordinal) { return add_process(ordinal, options{}); }
    process add_process(std::uint32_t ordinal, priority_type priority) {
        options o;
        o.priority = priority;
        return add_process(ordinal, o);
    }
    // A parameter sweep inside ONE ensemble -- the loop over load points of the examples' main()
    // (producer_consumer.cpp:67-73, aloha.cpp:93-103): replication r uses point r % steps of every rate that was
    // given one value per point, `exponential{std::vector<double>}`.
    void sweep(std::uint32_t steps) { sweep_steps_ = steps; }

    // env.bind(p): simulation<> binds co_main; several roots are allowed (examples/clocks.cpp)
    void bind(process p) { roots_.push_back(p.id); }

    queue add_queue(std::uint64_t max_size = 0) {
        queue_max_.push_back(max_size);
        return {(std::uint16_t)(queue_max_.size() - 1)};
    }
    semaphore add_semaphore(std::uint64_t value = 0, std::uint64_t max = std::numeric_limits<std::uint64_t>::max()) {
        sem_init_.push_back(value);
        sem_max_.push_back(max);
        return {(std::uint16_t)(sem_init_.size() - 1)};
    }
    resource add_resource(std::uint64_t count) { return {add_semaphore(count, count).id}; }
    // a rate for draws that are emitted by index (cxxdes_b200/record: `simple_op{CXXDES_B200_OP_DELAY_EXP, index, ...}`)
    std::uint16_t add_rate(double value) { return rate(value); }
    mutex add_mutex() { return {(std::uint16_t)n_mutexes_++}; }
    event add_event() { return {(std::uint16_t)n_events_++}; }

    // sequential(a, b, ...) / the comma operator: an unnamed coroutine that awaits its arguments in order
    // (sequential.ipp:1-20).  Await the returned process or use it as a child of all_of / any_of; it is named
    // `parent_ordinal | 0x8000`, the way the oracle harness reports library-created coroutines.
    process sequential(std::uint32_t parent_ordinal, std::initializer_list<child> steps);

    // Define the body of process p: `fn(body &)` issues the co_await statements in order; co_return is implied.
    template <typename F>
    void body_of(process p, F &&fn);

    // The C-ABI view (pointers into this object: keep it alive and unmodified while the view is in use).
    cxxdes_b200_program c_struct() const {
        for (std::size_t i = 0; i < procs_.size(); ++i)
            if (procs_[i].entry == NO_ENTRY)
                throw std::runtime_error("cxxdes_b200: process " + std::to_string(i) + " has no body");
        if (roots_.empty()) throw std::runtime_error("cxxdes_b200: no process bound to the environment (program::bind)");
        cxxdes_b200_program p{};
        p.ops = ops_.data();
        p.procs = procs_.data();
        p.roots = roots_.data();
        p.queue_max = queue_max_.empty() ? &zero_ : queue_max_.data();
        p.sem_init = sem_init_.empty() ? &zero_ : sem_init_.data();
        p.sem_max = sem_max_.empty() ? &zero_ : sem_max_.data();
        const std::size_t rows = sweep_steps_ ? sweep_steps_ : 1;
        rate_table_.clear();
        for (std::size_t step = 0; step < rows; ++step)
            for (const std::vector<double> &r : rates_) {
                if (r.size() != 1 && r.size() != rows)
                    throw std::runtime_error("cxxdes_b200: a swept rate needs one value per sweep point");
                rate_table_.push_back(r.size() == 1 ? r[0] : r[step]);
            }
        p.rates = rate_table_.empty() ? &zero_rate_ : rate_table_.data();
        p.sweep_steps = sweep_steps_;
        p.n_ops = (std::uint32_t)ops_.size();
        p.n_procs = (std::uint32_t)procs_.size();
        p.n_roots = (std::uint32_t)roots_.size();
        p.n_queues = (std::uint32_t)queue_max_.size();
        p.n_sems = (std::uint32_t)sem_init_.size();
        p.n_mutexes = n_mutexes_;
        p.n_events = n_events_;
        p.n_rates = (std::uint32_t)rates_.size();
        return p;
    }

    // Validate against the interpreters' rules without a GPU; throws std::runtime_error with the reason.
    void check() const {
        const cxxdes_b200_program p = c_struct();
        detail::check(cxxdes_b200_check_program(&p));
    }

    const std::vector<cxxdes_b200_op> &ops() const { return ops_; }
    const std::vector<cxxdes_b200_process> &processes() const { return procs_; }

private:
    friend class body;
    static constexpr std::uint32_t NO_ENTRY = 0xFFFFFFFFu;
    std::uint32_t emit(std::uint16_t code, std::uint16_t a = 0, std::int32_t b = 0, std::int64_t imm = 0) {
        ops_.push_back({code, a, b, imm});
        return (std::uint32_t)ops_.size() - 1;
    }
    std::uint16_t rate(double value) {
        rates_.push_back({value});
        return (std::uint16_t)(rates_.size() - 1);
    }
    std::uint16_t rate(const std::vector<double> &per_point) {
        rates_.push_back(per_point);
        return (std::uint16_t)(rates_.size() - 1);
    }
    std::vector<cxxdes_b200_op> ops_;
    std::vector<cxxdes_b200_process> procs_;
    std::vector<std::uint32_t> roots_;
    std::vector<std::uint64_t> queue_max_, sem_init_, sem_max_;
    std::vector<std::vector<double>> rates_;      // one value, or one per sweep point
    mutable std::vector<double> rate_table_;      // [sweep point][rate], built by c_struct()
    std::uint32_t sweep_steps_ = 0;
    std::uint32_t n_mutexes_ = 0, n_events_ = 0;
    std::uint64_t zero_ = 0;
    double zero_rate_ = 0.0;
};

// One coroutine body being written.  Every await() is one `co_await` of the reference.
class body {
public:
    body &await(delay_t d) {  // co_await delay(t) / timeout(t) / yield()
        prog_.emit(CXXDES_B200_OP_DELAY, 0, detail::prio32(d.prio), d.ticks);
        return *this;
    }
    body &await(until_t u) {  // co_await until(t)
        prog_.emit(CXXDES_B200_OP_UNTIL, 0, detail::prio32(u.prio), u.ticks);
        return *this;
    }
    body &await(lazy_timeout_t t) {  // auto t = co_await lazy_timeout(x): fixes the deadline now() + x, no event
        prog_.emit(CXXDES_B200_OP_LAZY_TIMEOUT, 0, 0, t.ticks);
        return *this;
    }
    body &await(lazy_instant_t t) {  // co_await t
        prog_.emit(CXXDES_B200_OP_UNTIL_REG, 0, detail::prio32(t.prio), 0);
        return *this;
    }
    body &await(env_timeout_t t) {  // co_await env.timeout(x)
        if (t.random) {
            prog_.emit(CXXDES_B200_OP_DELAY_EXP, t.per_point.empty() ? prog_.rate(t.value) : prog_.rate(t.per_point),
                       CXXDES_B200_INHERIT, 0);
        } else {
            std::int64_t bits;
            std::memcpy(&bits, &t.value, sizeof bits);
            prog_.emit(CXXDES_B200_OP_DELAY_REAL, 0, CXXDES_B200_INHERIT, bits);
        }
        return *this;
    }
    body &await(process p) {  // co_await p
        prog_.emit(CXXDES_B200_OP_AWAIT, (std::uint16_t)p.id);
        return *this;
    }
    body &await(async_t a) {  // co_await async(p)
        prog_.emit(CXXDES_B200_OP_ASYNC, (std::uint16_t)a.p.id);
        return *this;
    }
    body &await(const composition &c) {  // co_await all_of(...) / any_of(...) / (a && b) / (a || b), nested or not
        emit_composition(c, c.all ? CXXDES_B200_OP_ALL_OF : CXXDES_B200_OP_ANY_OF);
        return *this;
    }
    body &await(simple_op op) {  // co_await q.put(..) / q.pop() / sem.up() / mtx.acquire() / evt.wait() ...
        prog_.emit(op.code, op.a, op.b, op.imm);
        return *this;
    }

    // for (k = 0; k < count; ++k) { fn(); }   (at most two levels deep per process)
    template <typename F>
    body &loop(std::int64_t count, F &&fn) {
        const std::uint32_t head = prog_.emit(CXXDES_B200_OP_LOOP, 0, 0, count);
        fn();
        const std::uint32_t end = prog_.emit(CXXDES_B200_OP_END_LOOP, 0, (std::int32_t)(head + 1), 0);
        prog_.ops_[head].b = (std::int32_t)(end + 1);
        return *this;
    }

    // if (cond) { then_(); } [else { else_(); }] -- the plain C++ control flow between the awaits, on the model's
    // own state: co.if_then(q.size() < 5, [&] { co.await(q.put()); });
    template <typename Then>
    body &if_then(condition c, Then &&then_) {
        const std::uint32_t at = prog_.emit(CXXDES_B200_OP_IF, c.a, 0, c.value);
        then_();
        prog_.ops_[at].b = (std::int32_t)prog_.ops_.size();
        return *this;
    }
    template <typename Then, typename Else>
    body &if_then_else(condition c, Then &&then_, Else &&else_) {
        const std::uint32_t at = prog_.emit(CXXDES_B200_OP_IF, c.a, 0, c.value);
        then_();
        const std::uint32_t jump = prog_.emit(CXXDES_B200_OP_JUMP);
        prog_.ops_[at].b = (std::int32_t)prog_.ops_.size();
        else_();
        prog_.ops_[jump].b = (std::int32_t)prog_.ops_.size();
        return *this;
    }

    // co_return in the middle of a body: `while (true) { ...; if (n == n_packets) co_return; ... }`
    body &co_return_() {
        prog_.emit(CXXDES_B200_OP_RETURN);
        return *this;
    }

    // the same coroutine function called again: p becomes a new, unstarted coroutine object
    body &fresh(process p) {
        prog_.emit(CXXDES_B200_OP_FRESH, (std::uint16_t)p.id);
        return *this;
    }
    // model statistics (the plain C++ statements between the co_awaits of the examples)
    body &set_register_now() {  // x = now()
        prog_.emit(CXXDES_B200_OP_SET_REG_NOW);
        return *this;
    }
    body &stat_seconds(int k) {  // f64[k] += now_seconds() - seconds(register): producer_consumer.cpp:52
        prog_.emit(CXXDES_B200_OP_STAT_SECONDS, (std::uint16_t)k);
        return *this;
    }
    body &stat_ticks(int k) {  // f64[k] += now() - register
        prog_.emit(CXXDES_B200_OP_STAT_TICKS, (std::uint16_t)k);
        return *this;
    }
    body &count(int k, std::int64_t amount = 1) {  // i64[k] += amount
        prog_.emit(CXXDES_B200_OP_COUNT, (std::uint16_t)k, 0, amount);
        return *this;
    }
    body &mark(std::int64_t scale, std::int32_t tag) {  // an observation (the prints of the examples)
        prog_.emit(CXXDES_B200_OP_MARK, 0, tag, scale);
        return *this;
    }

private:
    friend class program;
    explicit body(program &p) : prog_(p) {}
    void emit_composition(const composition &c, std::uint16_t code) {   // children follow in place, pre-order
        if (c.children.empty() || c.children.size() > 64)
            throw std::invalid_argument("cxxdes_b200: a composition takes 1..64 children");
        prog_.emit(code, (std::uint16_t)c.children.size(), detail::prio32(c.prio), c.latency);
        for (const child &ch : c.children) {
            if (ch.kind == child::PROC) prog_.emit(CXXDES_B200_OP_CHILD_PROC, (std::uint16_t)ch.proc);
            else if (ch.kind == child::NESTED)
                emit_composition(*ch.nested, ch.nested->all ? CXXDES_B200_OP_CHILD_ALL_OF : CXXDES_B200_OP_CHILD_ANY_OF);
            else if (ch.kind == child::EV_WAIT)
                prog_.emit(CXXDES_B200_OP_CHILD_EV_WAIT, (std::uint16_t)ch.proc, detail::prio32(ch.prio), ch.ticks);
            else
                prog_.emit(ch.kind == child::DELAY ? CXXDES_B200_OP_CHILD_DELAY : CXXDES_B200_OP_CHILD_UNTIL, 0,
                           detail::prio32(ch.prio), ch.ticks);
        }
    }
    program &prog_;
};

template <typename F>
void program::body_of(process p, F &&fn) {
    if (p.id >= procs_.size()) throw std::invalid_argument("cxxdes_b200: unknown process");
    if (procs_[p.id].entry != NO_ENTRY) throw std::runtime_error("cxxdes_b200: process body defined twice");
    procs_[p.id].entry = (std::uint32_t)ops_.size();
    body b(*this);
    fn(b);
    emit(CXXDES_B200_OP_RETURN);
}

inline process program::sequential(std::uint32_t parent_ordinal, std::initializer_list<child> steps) {
    process p = add_process(parent_ordinal | 0x8000u);
    body_of(p, [&](body &co) {
        for (const child &st : steps) {
            if (st.kind == child::PROC) co.await(process{st.proc});
            else if (st.kind == child::DELAY) co.await(delay_t{st.ticks, st.prio});
            else co.await(until_t{st.ticks, st.prio});
        }
    });
    return p;
}

namespace models {
// Any model given as a process program; runs on the interpreter kernel.
struct process_program {
    static constexpr int id = CXXDES_B200_MODEL_PROGRAM;
    using params_type = cxxdes_b200_program;
    program prog;
    params_type params() const { return prog.c_struct(); }
};
}  // namespace models

}  // namespace b200
}  // namespace cxxdes
