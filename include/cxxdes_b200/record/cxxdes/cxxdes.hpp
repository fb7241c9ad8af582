// cxxdes_b200/record/cxxdes/cxxdes.hpp -- a RECORDING implementation of the libcxxdes API.
//
// Put `-I <repo>/include/cxxdes_b200/record` in front of the reference's include directory and a model written for
// libcxxdes -- `#include <cxxdes/cxxdes.hpp>`, `CXXDES_SIMULATION(x) { coroutine<> co_main(); ... }`, real C++20
// coroutine bodies with `co_await timeout(t)`, `co_await q.pop()`, `co_await all_of(p(1), p(2).priority(3))` ... --
// compiles against THIS header unchanged.  Nothing is simulated here: the coroutine bodies are run once on the host,
// every `co_await` is noted instead of suspending, and the result is the model as a process program
// (cxxdes_b200/program.hpp -> cxxdes_b200_program) for the B200 ensemble engine:
//
//     my_model sim{args...};
//     cxxdes::b200::program prog = sim.record();                       // the model, lowered
//     auto ens = sim.ensemble(1u << 20, /*seed*/ 7);  ens->run();      // a million replications on the GPU
//
// How a body becomes a program.  The names below mirror the reference's (core/simulation.hpp:31-85, coroutine.ipp,
// timeout.ipp, any_of.ipp, async.ipp, sync/*.hpp); each awaitable's `await_ready` is true, so a body runs straight
// through, `for` / `while` loops unrolled by the host; repeated runs of awaits are then folded back into counted loops
// (a body that never ends, `while (true) { co_await timeout(p); }`, becomes a loop of 2^30 turns).  Control flow may
// depend on the model's parameters and on plain counters -- whatever has the same value in every replication -- but
// NOT on simulated state: `q.pop()` returns a default-constructed value, `now()` returns 0.  Statistics computed in
// plain C++ between the awaits (`total += now_seconds() - x`) are therefore not recorded; write those models with the
// builder (program.hpp: co.stat_seconds(k)).  Random variates: cxxdes_b200/random_variable.hpp (`exponential_rv`)
// returns a tagged value that `env.timeout(x)` turns into a draw from the awaiting process's Philox stream.
// `subroutine<T>` bodies run inside the process that awaits them (subroutine.ipp: no tokens of their own), so their awaits
// are noted in that process; `_Co_with(x) { ... };` and `sequential(a, b, ...)` / the comma operator are, as in the
// reference (co_with.ipp:27-35, sequential.ipp:1-20), coroutines of the library -- helper processes of the program.
// So are `>>` and `flag_done`.  `co_await this_coroutine()` / `this_environment()` give views for a body's own use.
// Not provided: arithmetic on time literals (`1_s + 3 * b`; single literals such as `10_x`, `5_ms` are), exceptions through
// tokens, interrupts.  An `environment` may also be driven by hand, as examples/event.cpp and clocks.cpp do: env.bind(p)
// makes p a root process, env.run() / run_until(t) / run_for(t) run ONE replication of what has been bound on the GPU
// (simulation<>::run / run_for do the same).  Coroutine handles are copyable (the same process awaited from two others, or
// again after it returned); a coroutine object that is never awaited or bound is a process nothing starts.
// tests/test_record_examples.py: twenty of the reference's own examples, source files unchanged.
//
// Process ordinals (names in the trace digest, Philox stream ids) are given in creation order: co_main is 0, every
// coroutine object created after it 1, 2, ...  C++20, header-only, g++ >= 11.
#pragma once
#include <coroutine>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "cxxdes_b200/program.hpp"

namespace cxxdes {

namespace record {

namespace b = ::cxxdes::b200;

// one recorded co_await
struct step {
    enum kind_t { DELAY, UNTIL, REAL, EXP, AWAIT, ASYNC, COMPOSITION, SIMPLE } kind;
    std::int64_t ticks = 0;
    b::priority_type prio = b::inherit;
    double real = 0.0;
    std::uint32_t index = 0;                          // process / rate
    std::shared_ptr<const b::composition> comp;
    b::simple_op op{0, 0, 0, 0};
};
inline bool same(const b::composition &x, const b::composition &y) {
    if (x.all != y.all || x.prio != y.prio || x.latency != y.latency || x.children.size() != y.children.size()) return false;
    for (std::size_t i = 0; i < x.children.size(); ++i) {
        const b::child &c = x.children[i], &d = y.children[i];
        if (c.kind != d.kind || c.proc != d.proc || c.ticks != d.ticks || c.prio != d.prio) return false;
        if (c.kind == b::child::NESTED && !same(*c.nested, *d.nested)) return false;
    }
    return true;
}
inline bool same(const step &x, const step &y) {
    if (x.kind != y.kind || x.ticks != y.ticks || x.prio != y.prio || x.index != y.index) return false;
    if (x.kind == step::REAL && x.real != y.real) return false;
    if (x.kind == step::SIMPLE && (x.op.code != y.op.code || x.op.a != y.op.a || x.op.b != y.op.b || x.op.imm != y.op.imm)) return false;
    if (x.kind == step::COMPOSITION && !same(*x.comp, *y.comp)) return false;
    return true;
}

// a body after folding: steps and counted loops
struct node {
    std::int64_t count = 0;          // > 0: a loop over `body`
    std::vector<node> body;
    step s{};
};

constexpr std::size_t MAX_STEPS = 1u << 18;        // a body still running after this many awaits is taken to be endless
constexpr std::int64_t ENDLESS = std::int64_t(1) << 30;

struct proc_info {
    std::uint32_t ordinal = 0;
    b::program::options opt;
    std::coroutine_handle<> handle;
    bool started = false;
    bool endless = false;
    std::vector<step> steps;
};

struct context {
    std::vector<proc_info> procs;
    std::vector<std::uint32_t> stack;                 // processes whose bodies are being run (innermost last)
    std::vector<std::uint32_t> roots;
    std::vector<std::uint64_t> queue_max, sem_init, sem_max;
    std::uint32_t n_mutexes = 0, n_events = 0;
    std::vector<double> rates;
    b::time unit = b::one_second, precision = b::one_second;
    std::int64_t last_now = 0;                        // the clock after the last env.run*() that ran on a device

    struct endless_body {};

    std::uint32_t current() const {
        if (stack.empty()) throw std::runtime_error("cxxdes (recording): co_await outside a process body");
        return stack.back();
    }
    void note(step s) {
        proc_info &p = procs[current()];
        if (p.steps.size() >= MAX_STEPS) {
            p.endless = true;
            throw endless_body{};
        }
        p.steps.push_back(std::move(s));
    }

    // run the body of process `id` now (once): every co_await it makes lands in procs[id].steps
    void run_body(std::uint32_t id) {
        if (procs[id].started) return;
        procs[id].started = true;
        stack.push_back(id);
        try {
            procs[id].handle.resume();
        } catch (const endless_body &) {
            // (the frame is destroyed with its coroutine object)
        }
        stack.pop_back();
    }
};

inline context *&current_context() {
    static thread_local context *c = nullptr;
    return c;
}
inline context &ctx() {
    if (!current_context()) throw std::runtime_error("cxxdes (recording): no environment");
    return *current_context();
}

// env.run() / run_until(t) / run_for(t) of an environment driven by hand (examples/event.cpp, clocks.cpp): what was asked
// for, in order; the recorded model outlives a local `environment` through the shared context
struct run_request {
    std::shared_ptr<context> model;
    std::int64_t until;            // < 0: run()
};
inline std::vector<run_request> &run_requests() {
    static std::vector<run_request> *r = new std::vector<run_request>;   // (read by exit handlers: never destroyed)
    return *r;
}

// ---- folding repeated runs of steps back into loops (at most two deep, like the interpreters) ----
inline bool same_run(const std::vector<step> &v, std::size_t a, std::size_t c, std::size_t len) {
    for (std::size_t k = 0; k < len; ++k)
        if (!same(v[a + k], v[c + k])) return false;
    return true;
}
inline std::vector<node> fold(const std::vector<step> &v, std::size_t begin, std::size_t end, int depth, bool endless_tail) {
    std::vector<node> out;
    std::size_t i = begin;
    while (i < end) {
        std::size_t best_p = 0, best_k = 0;
        if (depth < 2) {
            const std::size_t max_p = (end - i) / 2 < 64 ? (end - i) / 2 : 64;
            for (std::size_t p = 1; p <= max_p; ++p) {
                std::size_t k = 1;
                while (i + (k + 1) * p <= end && same_run(v, i, i + k * p, p)) ++k;
                if (k >= 2 && k * p > best_k * best_p) { best_p = p; best_k = k; }
            }
        }
        if (best_k >= 2) {
            node loop;
            loop.count = (std::int64_t)best_k;
            loop.body = fold(v, i, i + best_p, depth + 1, false);
            i += best_k * best_p;
            // an endless body was cut off somewhere inside a turn of its last loop
            if (endless_tail && depth == 0 && end - i < best_p && same_run(v, i - best_p, i, end - i)) {
                loop.count = ENDLESS;
                i = end;
            }
            out.push_back(std::move(loop));
        } else {
            node leaf;
            leaf.s = v[i++];
            out.push_back(std::move(leaf));
        }
    }
    return out;
}

inline void emit(b::body &co, const std::vector<node> &nodes) {
    for (const node &n : nodes) {
        if (n.count > 0) {
            co.loop(n.count, [&] { emit(co, n.body); });
            continue;
        }
        const step &s = n.s;
        switch (s.kind) {
            case step::DELAY: co.await(b::delay_t{s.ticks, s.prio}); break;
            case step::UNTIL: co.await(b::until_t{s.ticks, s.prio}); break;
            case step::REAL: co.await(b::env_timeout(s.real)); break;
            case step::EXP: co.await(b::simple_op{CXXDES_B200_OP_DELAY_EXP, (std::uint16_t)s.index, CXXDES_B200_INHERIT, 0}); break;
            case step::AWAIT: co.await(b::process{s.index}); break;
            case step::ASYNC: co.await(b::async(b::process{s.index})); break;
            case step::COMPOSITION: co.await(*s.comp); break;
            case step::SIMPLE: co.await(s.op); break;
        }
    }
}

}  // namespace record

namespace core {

using time_integral = std::intmax_t;
using priority_type = std::intmax_t;
using real_type = double;
using time = ::cxxdes::b200::time;
namespace time_ops = ::cxxdes::b200::time_ops;

struct priority_consts {
    static constexpr priority_type highest = INT64_MIN, inherit = ::cxxdes::b200::inherit, lowest = INT64_MAX - 1, zero = 0;
};

// ---- awaitables: every one notes itself in the body being run and is ready at once ---------------------------
template <typename R = void>
struct noted {
    R value{};
    bool await_ready() const noexcept { return true; }
    void await_suspend(std::coroutine_handle<>) const noexcept {}
    R await_resume() { return std::move(value); }
};
template <>
struct noted<void> {
    bool await_ready() const noexcept { return true; }
    void await_suspend(std::coroutine_handle<>) const noexcept {}
    void await_resume() const noexcept {}
};

struct timeout_awaitable {   // what timeout(t, priority) / delay / until / instant / yield() return (timeout.ipp:56-182)
    time_integral ticks;
    priority_type prio = priority_consts::inherit;
    bool absolute = false;
};
inline timeout_awaitable delay(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio, false}; }
inline timeout_awaitable timeout(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio, false}; }
inline timeout_awaitable yield() { return {0}; }
inline timeout_awaitable until(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio, true}; }
inline timeout_awaitable instant(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio, true}; }
// the same with a time literal (`5_ms`, `10_x`): environment::real_to_sim of a time-expression node (environment.ipp:73-76)
// with the unit and precision of the environment being recorded -- single literals, not the expression DSL around them
namespace detail {
inline time_integral ticks_of(time x) { return ::cxxdes::b200::count(x, record::ctx().precision); }
inline time_integral ticks_of(time_ops::unitless_time x) { return x.t * ::cxxdes::b200::count(record::ctx().unit, record::ctx().precision); }
template <typename X>
using time_literal = decltype(ticks_of(std::declval<X>()));
}  // namespace detail
template <typename X, typename = detail::time_literal<X>>
timeout_awaitable delay(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio, false}; }
template <typename X, typename = detail::time_literal<X>>
timeout_awaitable timeout(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio, false}; }
template <typename X, typename = detail::time_literal<X>>
timeout_awaitable until(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio, true}; }
template <typename X, typename = detail::time_literal<X>>
timeout_awaitable instant(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio, true}; }

// lazy_timeout(t) / lazy_delay(t) (timeout.ipp:104-177): `auto t = co_await lazy_timeout(x)` fixes the deadline now() + x
// without an event, `co_await t` waits for it.  One pending deadline per process (the process register of the program).
struct lazy_instant_awaitable {
    priority_type prio;
    noted<> cxxdes_record() const;
};
struct lazy_timeout_awaitable {
    time_integral ticks;
    priority_type prio;
    noted<lazy_instant_awaitable> cxxdes_record() const;
};
inline lazy_timeout_awaitable lazy_timeout(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio}; }
inline lazy_timeout_awaitable lazy_delay(time_integral t, priority_type prio = priority_consts::inherit) { return {t, prio}; }
template <typename X, typename = detail::time_literal<X>>
lazy_timeout_awaitable lazy_timeout(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio}; }
template <typename X, typename = detail::time_literal<X>>
lazy_timeout_awaitable lazy_delay(X x, priority_type prio = priority_consts::inherit) { return {detail::ticks_of(x), prio}; }
inline noted<lazy_instant_awaitable> lazy_timeout_awaitable::cxxdes_record() const {
    record::step s;
    s.kind = record::step::SIMPLE;
    s.op = {CXXDES_B200_OP_LAZY_TIMEOUT, 0, 0, ticks};
    record::ctx().note(std::move(s));
    return {lazy_instant_awaitable{prio}};
}
inline noted<> lazy_instant_awaitable::cxxdes_record() const {
    record::step s;
    s.kind = record::step::SIMPLE;
    s.op = {CXXDES_B200_OP_UNTIL_REG, 0, ::cxxdes::b200::detail::prio32(prio), 0};
    record::ctx().note(std::move(s));
    return {};
}

struct env_timeout_awaitable {   // env.timeout(x)   (timeout.ipp:184-187)
    bool random;
    double value;
    std::uint32_t rate;
};

template <typename T>
class coroutine;
template <typename T>
class subroutine;
class environment;

// co_await this_coroutine() / this_environment() (await_transform.ipp:21-25,66-72): no event; what they return is a
// view for a body's own use (prints): the process's priority option as given, a clock that reads 0 while recording
struct this_coroutine {};
struct this_environment {};
struct coroutine_view {
    priority_type priority_ = 0;
    time_integral latency_ = 0;
    priority_type priority() const noexcept { return priority_; }
    time_integral latency() const noexcept { return latency_; }
    const coroutine_view *operator->() const noexcept { return this; }
};
struct environment_view {   // what `co_await this_environment()` hands back: env->now(), env->timeout(x)
    time_integral now() const noexcept { return 0; }
    real_type now_seconds() const noexcept { return 0.0; }
    env_timeout_awaitable timeout(double seconds) const { return {false, seconds, 0}; }
    const environment_view *operator->() const noexcept { return this; }
};

namespace detail {
// what a body may co_await (await_transform.ipp): shared by the promises of coroutine<T> and subroutine<T>
struct await_ops {
    std::suspend_always initial_suspend() const noexcept { return {}; }
    std::suspend_always final_suspend() const noexcept { return {}; }
    void unhandled_exception() { throw; }

    // co_await <something>: note it in the process whose body is being run, do not suspend
    noted<> await_transform(timeout_awaitable t) {
        record::step s;
        s.kind = t.absolute ? record::step::UNTIL : record::step::DELAY;
        s.ticks = t.ticks;
        s.prio = t.prio;
        record::ctx().note(std::move(s));
        return {};
    }
    noted<> await_transform(env_timeout_awaitable t) {
        record::step s;
        s.kind = t.random ? record::step::EXP : record::step::REAL;
        s.real = t.value;
        s.index = t.rate;
        record::ctx().note(std::move(s));
        return {};
    }
    template <typename U>
    noted<U> await_transform(coroutine<U> &&child);
    template <typename U>
    noted<U> await_transform(coroutine<U> &child);
    template <typename U>
    noted<U> await_transform(subroutine<U> &&sub);
    noted<coroutine_view> await_transform(this_coroutine) const {
        const record::proc_info &p = record::ctx().procs[record::ctx().current()];
        coroutine_view v;
        v.priority_ = p.opt.priority == ::cxxdes::b200::inherit ? 0 : p.opt.priority;
        v.latency_ = p.opt.latency;
        return {v};
    }
    noted<environment_view> await_transform(this_environment) const { return {}; }
    template <typename A>
    auto await_transform(A &&a) -> decltype(std::forward<A>(a).cxxdes_record()) {   // sync operations, compositions, async
        return std::forward<A>(a).cxxdes_record();
    }
    // co_yield x is co_await x (await_transform.ipp:74-77): what `_Co_with` expands to
    template <typename A>
    auto yield_value(A &&a) -> decltype(this->await_transform(std::forward<A>(a))) {
        return await_transform(std::forward<A>(a));
    }
};
// a coroutine<T> is a process of its own
struct promise_base : await_ops {
    std::uint32_t id;
    promise_base() {
        record::context &c = record::ctx();
        id = (std::uint32_t)c.procs.size();
        c.procs.emplace_back();
        c.procs.back().ordinal = id;
    }
};
}  // namespace detail

// ---- coroutine<T>: a process of the model (coroutine.ipp) --------------------------------------------------------
template <typename T = void>
class coroutine {
public:
    struct promise_type : detail::promise_base {
        T result{};
        coroutine get_return_object() { return coroutine{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        template <typename V>
        void return_value(V &&v) { result = std::forward<V>(v); }
    };
    coroutine() noexcept : h_(nullptr) {}            // (an empty handle, to be assigned a process later)
    // a handle on the process, copyable as in the reference (coroutine.ipp:30-62: a counted reference to coroutine_data);
    // the frame lives as long as one handle does
    coroutine(coroutine &&o) noexcept : h_(std::exchange(o.h_, nullptr)), frame_(std::move(o.frame_)) {}
    coroutine(const coroutine &) = default;
    coroutine &operator=(const coroutine &) = default;
    coroutine &operator=(coroutine &&o) noexcept {
        if (this != &o) {
            h_ = std::exchange(o.h_, nullptr);
            frame_ = std::move(o.frame_);
        }
        return *this;
    }
    std::uint32_t id() const { return h_.promise().id; }
    // coroutine.ipp:106-176
    coroutine &&priority(priority_type p) && { info().opt.priority = p; return std::move(*this); }
    coroutine &&latency(time_integral t) && { info().opt.latency = t; return std::move(*this); }
    coroutine &&return_latency(time_integral t) && { info().opt.return_latency = t; return std::move(*this); }
    coroutine &&return_priority(priority_type p) && { info().opt.return_priority = p; return std::move(*this); }
    coroutine &priority(priority_type p) & { info().opt.priority = p; return *this; }
    coroutine &latency(time_integral t) & { info().opt.latency = t; return *this; }
    coroutine &return_latency(time_integral t) & { info().opt.return_latency = t; return *this; }
    coroutine &return_priority(priority_type p) & { info().opt.return_priority = p; return *this; }
    T run_now() {   // run the body (once) and hand back what it returned
        record::ctx().procs[id()].handle = h_;
        record::ctx().run_body(id());
        return h_.promise().result;
    }
private:
    explicit coroutine(std::coroutine_handle<promise_type> h)
        : h_(h), frame_(h.address(), [](void *a) { std::coroutine_handle<>::from_address(a).destroy(); }) {
        record::ctx().procs[h.promise().id].handle = h;
    }
    record::proc_info &info() { return record::ctx().procs[id()]; }
    std::coroutine_handle<promise_type> h_;
    std::shared_ptr<void> frame_;
};
template <>
class coroutine<void> {
public:
    struct promise_type : detail::promise_base {
        coroutine get_return_object() { return coroutine{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        void return_void() const noexcept {}
    };
    coroutine() noexcept : h_(nullptr) {}            // (an empty handle, to be assigned a process later)
    // a handle on the process, copyable as in the reference (coroutine.ipp:30-62: a counted reference to coroutine_data);
    // the frame lives as long as one handle does
    coroutine(coroutine &&o) noexcept : h_(std::exchange(o.h_, nullptr)), frame_(std::move(o.frame_)) {}
    coroutine(const coroutine &) = default;
    coroutine &operator=(const coroutine &) = default;
    coroutine &operator=(coroutine &&o) noexcept {
        if (this != &o) {
            h_ = std::exchange(o.h_, nullptr);
            frame_ = std::move(o.frame_);
        }
        return *this;
    }
    std::uint32_t id() const { return h_.promise().id; }
    coroutine &&priority(priority_type p) && { info().opt.priority = p; return std::move(*this); }
    coroutine &&latency(time_integral t) && { info().opt.latency = t; return std::move(*this); }
    coroutine &&return_latency(time_integral t) && { info().opt.return_latency = t; return std::move(*this); }
    coroutine &&return_priority(priority_type p) && { info().opt.return_priority = p; return std::move(*this); }
    coroutine &priority(priority_type p) & { info().opt.priority = p; return *this; }
    coroutine &latency(time_integral t) & { info().opt.latency = t; return *this; }
    coroutine &return_latency(time_integral t) & { info().opt.return_latency = t; return *this; }
    coroutine &return_priority(priority_type p) & { info().opt.return_priority = p; return *this; }
    void run_now() { record::ctx().run_body(id()); }
private:
    explicit coroutine(std::coroutine_handle<promise_type> h)
        : h_(h), frame_(h.address(), [](void *a) { std::coroutine_handle<>::from_address(a).destroy(); }) {
        record::ctx().procs[h.promise().id].handle = h;
    }
    record::proc_info &info() { return record::ctx().procs[id()]; }
    std::coroutine_handle<promise_type> h_;
    std::shared_ptr<void> frame_;
};

// ---- subroutine<T> (subroutine.ipp): a coroutine frame on the call stack of the process that awaits it -- no start
// or completion token, no process of its own.  Recording: its body runs where it is awaited and its awaits are noted in
// the process being run.
template <typename T = void>
class subroutine {
public:
    struct promise_type : detail::await_ops {
        T result{};
        subroutine get_return_object() { return subroutine{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        template <typename V>
        void return_value(V &&v) { result = std::forward<V>(v); }
    };
    subroutine(subroutine &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    subroutine(const subroutine &) = delete;
    ~subroutine() { if (h_) h_.destroy(); }
    T run_now() {
        if (!h_.done()) h_.resume();
        return h_.promise().result;
    }
    // std::move(s).as_coroutine() (subroutine.ipp:54-55): the subroutine as a process of its own
    coroutine<T> as_coroutine() && {
        return [](subroutine s) -> coroutine<T> { co_return co_await std::move(s); }(std::move(*this));
    }
private:
    explicit subroutine(std::coroutine_handle<promise_type> h) : h_(h) {}
    std::coroutine_handle<promise_type> h_;
};
template <>
class subroutine<void> {
public:
    struct promise_type : detail::await_ops {
        subroutine get_return_object() { return subroutine{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        void return_void() const noexcept {}
    };
    subroutine(subroutine &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    subroutine(const subroutine &) = delete;
    ~subroutine() { if (h_) h_.destroy(); }
    void run_now() {
        if (!h_.done()) h_.resume();
    }
    coroutine<void> as_coroutine() && {
        return [](subroutine s) -> coroutine<void> { co_await std::move(s); }(std::move(*this));
    }
private:
    explicit subroutine(std::coroutine_handle<promise_type> h) : h_(h) {}
    std::coroutine_handle<promise_type> h_;
};

namespace detail {
template <typename U>
noted<U> await_ops::await_transform(subroutine<U> &&sub) {   // co_await sub(): inline (subroutine.ipp:31-50)
    record::ctx().current();            // (must be inside a process body)
    if constexpr (std::is_void_v<U>) {
        sub.run_now();
        return {};
    } else {
        return noted<U>{sub.run_now()};
    }
}
template <typename U>
noted<U> await_ops::await_transform(coroutine<U> &child) {   // co_await p   (coroutine.ipp:179-207)
    record::step s;
    s.kind = record::step::AWAIT;
    s.index = child.id();
    record::ctx().note(std::move(s));
    if constexpr (std::is_void_v<U>) {
        child.run_now();
        return {};
    } else {
        return noted<U>{child.run_now()};
    }
}
template <typename U>
noted<U> await_ops::await_transform(coroutine<U> &&child) {
    // (the awaited temporary dies with the full expression; its body has been run by then and its steps are kept)
    return await_transform(static_cast<coroutine<U> &>(child));
}
}  // namespace detail

// ---- async(p)   (async.ipp:9-49) ----------------------------------------------------------------------------------
template <typename U>
struct async_awaitable {
    std::uint32_t id;
    noted<> cxxdes_record() const {
        record::step s;
        s.kind = record::step::ASYNC;
        s.index = id;
        record::ctx().note(std::move(s));
        return {};
    }
};
template <typename U>
async_awaitable<U> async(coroutine<U> &&c) {
    c.run_now();
    return {c.id()};
}
template <typename U>
async_awaitable<U> async(coroutine<U> &c) {
    c.run_now();
    return {c.id()};
}

// ---- all_of / any_of / && / ||   (any_of.ipp, compositions.ipp) ----------------------------------------------------
struct composition_awaitable {
    ::cxxdes::b200::composition comp;
    composition_awaitable priority(priority_type p) const { return {comp.priority(p)}; }
    composition_awaitable return_latency(time_integral t) const { return {comp.return_latency(t)}; }
    noted<> cxxdes_record() const {
        record::step s;
        s.kind = record::step::COMPOSITION;
        s.comp = std::make_shared<::cxxdes::b200::composition>(comp);
        record::ctx().note(std::move(s));
        return {};
    }
};
namespace detail {
template <typename U>
::cxxdes::b200::child as_child(coroutine<U> &&c) { c.run_now(); return ::cxxdes::b200::process{c.id()}; }
template <typename U>
::cxxdes::b200::child as_child(coroutine<U> &c) { c.run_now(); return ::cxxdes::b200::process{c.id()}; }
inline ::cxxdes::b200::child as_child(timeout_awaitable t) {
    if (t.absolute) return ::cxxdes::b200::until_t{t.ticks, t.prio};
    return ::cxxdes::b200::delay_t{t.ticks, t.prio};
}
inline ::cxxdes::b200::child as_child(const composition_awaitable &c) { return ::cxxdes::b200::child(c.comp); }
template <typename X>
using child_of = decltype(as_child(std::declval<X>()));
}  // namespace detail
namespace detail {
template <bool ALL>
struct composition_functor {   // any_of.ipp:163-231: all_of(a, b, ...), all_of.range(first, last), ...
    template <typename... A>
    [[nodiscard]] composition_awaitable operator()(A &&...a) const {
        return {::cxxdes::b200::composition{ALL, {as_child(std::forward<A>(a))...}}};
    }
    template <typename... A>
    [[nodiscard]] composition_awaitable by_value(A &&...a) const { return (*this)(std::forward<A>(a)...); }
    template <typename... A>
    [[nodiscard]] composition_awaitable by_reference(A &&...a) const { return (*this)(std::forward<A>(a)...); }
    template <typename... A>
    [[nodiscard]] composition_awaitable copy_rvalues(A &&...a) const { return (*this)(std::forward<A>(a)...); }
    template <typename Iterator>
    [[nodiscard]] composition_awaitable range(Iterator first, Iterator last) const {
        ::cxxdes::b200::composition c{ALL, {}};
        for (; first != last; ++first) c.children.push_back(as_child(*first));
        return {c};
    }
    template <typename Iterator>
    [[nodiscard]] composition_awaitable range_copy(Iterator first, Iterator last) const { return range(first, last); }
    template <typename Iterator>
    [[nodiscard]] composition_awaitable range_no_copy(Iterator first, Iterator last) const { return range(first, last); }
};
}  // namespace detail
inline constexpr detail::composition_functor<true> all_of;
inline constexpr detail::composition_functor<false> any_of;
template <typename X, typename Y, typename = detail::child_of<X>, typename = detail::child_of<Y>>
composition_awaitable operator&&(X &&x, Y &&y) { return all_of(std::forward<X>(x), std::forward<Y>(y)); }
template <typename X, typename Y, typename = detail::child_of<X>, typename = detail::child_of<Y>>
composition_awaitable operator||(X &&x, Y &&y) { return any_of(std::forward<X>(x), std::forward<Y>(y)); }

// ---- sequential(a, b, ...) and the comma operator (sequential.ipp:1-84, compositions.ipp:24-27): a coroutine of the
// library that awaits its arguments in order -- in the program, a helper process ------------------------------------
struct sequential_helper {
    template <typename... Ts>
    static coroutine<void> seq_proc_tuple(Ts... ts) {
        ((co_await std::move(ts)), ...);
        co_return;
    }
    template <typename Iterator>
    static coroutine<void> seq_proc_range(Iterator begin, Iterator end) {
        for (Iterator it = begin; it != end; ++it) co_await (*it);
        co_return;
    }
    struct functor {
        template <typename... Ts>
        [[nodiscard]] auto operator()(Ts &&...ts) const { return by_value(std::forward<Ts>(ts)...); }
        template <typename... Ts>
        [[nodiscard]] auto by_value(Ts &&...ts) const {
            return seq_proc_tuple<std::remove_cvref_t<Ts>...>(std::forward<Ts>(ts)...);
        }
        template <typename... Ts>
        [[nodiscard]] auto copy_rvalues(Ts &&...ts) const { return seq_proc_tuple<Ts...>(std::forward<Ts>(ts)...); }
        template <typename Iterator>
        [[nodiscard]] auto range_no_copy(Iterator first, Iterator last) const { return seq_proc_range(first, last); }
    };
};
inline constexpr sequential_helper::functor sequential;
template <typename X, typename Y, typename = detail::child_of<X>, typename = detail::child_of<Y>>
auto operator,(X &&x, Y &&y) { return sequential(std::forward<X>(x), std::forward<Y>(y)); }

// ---- p >> out, flag_done(flag)   (compositions.ipp:29-37): library coroutines again; while recording `out` receives
// what the awaited body returned on the host (the same in every replication, or it is not recordable) ---------------
template <typename U, typename Output>
coroutine<void> operator>>(coroutine<U> &&a, Output &output) {
    output = co_await a;
}
template <typename U, typename Output>
coroutine<void> operator>>(coroutine<U> &a, Output &output) {
    output = co_await a;
}
inline coroutine<void> flag_done(bool &flag) {
    flag = true;
    co_return;
}

// ---- a random variate on its way to env.timeout (cxxdes_b200/random_variable.hpp) ----------------------------------
struct variate_tag {
    std::uint32_t rate;   // index into the program's rates
};

// ---- environment (environment.ipp:12-279): here, the thing that owns the recording ---------------------------------
class environment {
public:
    environment() : c_(std::make_shared<record::context>()) { record::current_context() = c_.get(); }
    environment(const environment &) = delete;
    // 0 while the bodies are being recorded; after run() / run_until() / run_for() the clock of the replication that ran
    time_integral now() const noexcept { return now_; }
    real_type now_seconds() const noexcept { return now_seconds_; }
    void time_unit(time x) { c_->unit = x; }
    void time_precision(time x) { c_->precision = x; }
    time time_unit() const { return c_->unit; }
    time time_precision() const { return c_->precision; }
    env_timeout_awaitable timeout(double seconds) const { return {false, seconds, 0}; }
    env_timeout_awaitable timeout(variate_tag v) const { return {true, 0.0, v.rate}; }
    template <typename U>
    void bind(coroutine<U> &&c) {   // environment::bind (coroutine.ipp:352-357): a root process
        // (a root bound after the environment has run would start at the clock of that moment: a recorded program's roots
        // all start at tick 0)
        if (horizon_ >= 0 || ran_) throw std::runtime_error("cxxdes (recording): bind() after run() / run_until() / run_for() is not supported");
        record::current_context() = c_.get();
        c_->roots.push_back(c.id());
        c.run_now();
    }
    std::uint32_t add_rate(double rate) {
        c_->rates.push_back(rate);
        return (std::uint32_t)c_->rates.size() - 1;
    }
    record::context &recording() { return *c_; }

    // environment::run / run_until / run_for (environment.ipp:179-214) of an environment driven by hand: ONE replication
    // of what has been bound so far, on the GPU (the clock an ensemble handle keeps between calls is kept here as the
    // horizon reached).  -DCXXDES_B200_RECORD_ONLY: only note the request (record::run_requests(); tests without a device).
    void run() { request(-1); }
    void run_until(time_integral t) { request(t > horizon_ ? t : horizon_); }
    void run_for(time_integral t) { request((horizon_ < 0 ? 0 : horizon_) + t); }

    // the recorded model as a process program
    ::cxxdes::b200::program program() const { return program_of(*c_); }
    static ::cxxdes::b200::program program_of(const record::context &c) {
        namespace b = ::cxxdes::b200;
        b::program prog;
        for (std::uint64_t m : c.queue_max) prog.add_queue(m);
        for (std::size_t k = 0; k < c.sem_init.size(); ++k) prog.add_semaphore(c.sem_init[k], c.sem_max[k]);
        for (std::uint32_t k = 0; k < c.n_mutexes; ++k) prog.add_mutex();
        for (std::uint32_t k = 0; k < c.n_events; ++k) prog.add_event();
        std::vector<b::process> ps;
        for (const record::proc_info &p : c.procs) ps.push_back(prog.add_process(p.ordinal, p.opt));
        for (std::uint32_t r : c.roots) prog.bind(ps[r]);
        for (double r : c.rates) prog.add_rate(r);      // (the recorded draws refer to them by index)
        for (std::size_t k = 0; k < c.procs.size(); ++k) {
            const record::proc_info &p = c.procs[k];
            if (!p.started) {
                // a coroutine object that was created but never awaited or bound never runs (examples/interrupt_test.cpp
                // makes three and awaits one): a process without a body that nothing starts
                prog.body_of(ps[k], [](b::body &) {});
                continue;
            }
            const std::vector<record::node> nodes = record::fold(p.steps, 0, p.steps.size(), 0, p.endless);
            prog.body_of(ps[k], [&](b::body &co) { record::emit(co, nodes); });
        }
        return prog;
    }

private:
    void request(std::int64_t until) {
        ran_ = true;
        record::run_requests().push_back({c_, until});
#if !defined(CXXDES_B200_RECORD_ONLY)
        using ens_t = ::cxxdes::b200::ensemble<::cxxdes::b200::models::process_program>;
        if (!ensemble_) {
            ensemble_ = std::make_shared<ens_t>(::cxxdes::b200::models::process_program{program()}, 1, 0x5EED);
            ensemble_->env.time_unit(c_->unit);
            ensemble_->env.time_precision(c_->precision);
        }
        if (until < 0) ensemble_->run();
        else ensemble_->run_until(until);
        now_ = ensemble_->now()[0];
        now_seconds_ = ensemble_->now_seconds()[0];
        c_->last_now = now_;
#endif
        if (until >= 0) horizon_ = until;
    }
    std::shared_ptr<record::context> c_;
    std::int64_t horizon_ = -1;
    bool ran_ = false;
    time_integral now_ = 0;
    real_type now_seconds_ = 0.0;
#if !defined(CXXDES_B200_RECORD_ONLY)
    std::shared_ptr<::cxxdes::b200::ensemble<::cxxdes::b200::models::process_program>> ensemble_;
#endif
};

// ---- simulation<Derived> (core/simulation.hpp:31-85) ----------------------------------------------------------------
template <typename Derived>
struct simulation {
    auto now() const noexcept { return env.now(); }
    auto now_seconds() const noexcept { return env.now_seconds(); }
    environment env;

    // the model as a process program (runs co_main's body and everything it awaits, once, on the host)
    ::cxxdes::b200::program record() {
        bind();
        return env.program();
    }
    // an ensemble of R replications of the recorded model on the GPU
    std::unique_ptr<::cxxdes::b200::ensemble<::cxxdes::b200::models::process_program>> ensemble(std::uint64_t n_replications,
                                                                                                 std::uint64_t seed = 0x5EED) {
        using ens_t = ::cxxdes::b200::ensemble<::cxxdes::b200::models::process_program>;
        auto ens = std::make_unique<ens_t>(::cxxdes::b200::models::process_program{record()}, n_replications, seed);
        ens->env.time_unit(env.time_unit());
        ens->env.time_precision(env.time_precision());
        return ens;
    }

    // simulation<>::run / run_for (core/simulation.hpp:51-62): ONE replication of the recorded model on the GPU
    void run() { bind(); env.run(); }
    void run_for(::cxxdes::b200::time_integral t) { bind(); env.run_for(t); }

protected:
    void bind() {
        if (!bound_) {
            env.bind(static_cast<Derived &>(*this).co_main());
            bound_ = true;
        }
    }

private:
    bool bound_ = false;
};

#define CXXDES_SIMULATION(name) struct name : cxxdes::core::simulation<name>

namespace detail {
struct under_helper {};
}

}  // namespace core

// ---- sync primitives (sync/*.hpp): objects register themselves with the environment under construction --------------
namespace sync {

template <typename T>
class queue {   // sync/queue.hpp:38-90
public:
    explicit queue(std::size_t max_size = 0) : id_((std::uint16_t)record::ctx().queue_max.size()) {
        record::ctx().queue_max.push_back(max_size);
    }
    struct put_awaitable {
        std::uint16_t id;
        core::noted<> cxxdes_record() const { return note({CXXDES_B200_OP_Q_PUT, id, 0, 0}); }
    };
    struct pop_awaitable {
        std::uint16_t id;
        core::noted<T> cxxdes_record() const {
            note({CXXDES_B200_OP_Q_POP, id, 0, 0});
            return {};
        }
    };
    template <typename V>
    put_awaitable put(V &&) const { return {id_}; }
    pop_awaitable pop() const { return {id_}; }
private:
    static core::noted<> note(::cxxdes::b200::simple_op op) {
        record::step s;
        s.kind = record::step::SIMPLE;
        s.op = op;
        record::ctx().note(std::move(s));
        return {};
    }
    std::uint16_t id_;
};

namespace detail {
inline core::noted<> note_simple(::cxxdes::b200::simple_op op) {
    record::step s;
    s.kind = record::step::SIMPLE;
    s.op = op;
    record::ctx().note(std::move(s));
    return {};
}
struct simple_awaitable {
    ::cxxdes::b200::simple_op op;
    core::noted<> cxxdes_record() const { return note_simple(op); }
};
// `evt.wait() || timeout(t)`: an event wait as a child of a composition (found by the compositions through ADL; the
// builder refuses any other sync operation there)
inline ::cxxdes::b200::child as_child(const simple_awaitable &a) { return ::cxxdes::b200::child(a.op); }
}  // namespace detail

template <typename T = std::size_t>
class semaphore {   // sync/semaphore.hpp:58-78
public:
    explicit semaphore(std::size_t value = 0, std::size_t max = ~std::size_t(0)) : id_((std::uint16_t)record::ctx().sem_init.size()) {
        record::ctx().sem_init.push_back(value);
        record::ctx().sem_max.push_back(max);
    }
    detail::simple_awaitable up() const { return {{CXXDES_B200_OP_SEM_UP, id_, 0, 0}}; }
    detail::simple_awaitable down() const { return {{CXXDES_B200_OP_SEM_DOWN, id_, 0, 0}}; }
protected:
    std::uint16_t id_;
};

class resource {   // sync/resource.hpp:37-100
public:
    explicit resource(std::size_t count) : sem_(count, count) {}
    struct handle {
        detail::simple_awaitable release_op;
        detail::simple_awaitable release() const { return release_op; }
    };
    struct acquire_awaitable {
        detail::simple_awaitable down, up;
        core::noted<handle> cxxdes_record() const {
            down.cxxdes_record();
            return {handle{up}};
        }
    };
    acquire_awaitable acquire() const { return {sem_.down(), sem_.up()}; }
private:
    semaphore<> sem_;
};

class mutex {   // sync/mutex.hpp:31-109
public:
    mutex() : id_((std::uint16_t)record::ctx().n_mutexes++) {}
    struct handle {
        std::uint16_t id;
        detail::simple_awaitable release() const { return {{CXXDES_B200_OP_MTX_RELEASE, id, 0, 0}}; }
    };
    struct acquire_awaitable {
        std::uint16_t id;
        core::noted<handle> cxxdes_record() const {
            detail::note_simple({CXXDES_B200_OP_MTX_ACQUIRE, id, 0, 0});
            return {handle{id}};
        }
    };
    acquire_awaitable acquire() const { return {id_}; }
private:
    std::uint16_t id_;
};

class event {   // sync/event.hpp:87-139
public:
    event() : id_((std::uint16_t)record::ctx().n_events++) {}
    detail::simple_awaitable wait(core::time_integral latency = 0, core::priority_type prio = core::priority_consts::inherit) const {
        return {{CXXDES_B200_OP_EV_WAIT, id_, ::cxxdes::b200::detail::prio32(prio), latency}};
    }
    detail::simple_awaitable wake() const { return {{CXXDES_B200_OP_EV_WAKE, id_, 0, 0}}; }
private:
    std::uint16_t id_;
};

}  // namespace sync
}  // namespace cxxdes

// ---- _Co_with(x) { body };   (co_with.ipp:19-35): a coroutine that acquires x, runs the body -- a subroutine -- and
// releases; awaited with co_yield.  Same text as the reference's: in the program, a helper process
// {acquire, the body's awaits, release} that the enclosing process awaits. ------------------------------------------------
namespace cxxdes::core::detail {
template <typename A>
concept recorded_acquirable = requires(A a) { a.acquire().cxxdes_record().await_resume().release().cxxdes_record(); };
}
template <cxxdes::core::detail::recorded_acquirable A, typename F>
cxxdes::core::coroutine<void> operator+(A &a, F &&f) {
    auto handle = co_await a.acquire();
    co_await std::forward<F>(f)();
    co_await handle.release();
}
#define _Co_with(x) co_yield (x) + [&]() mutable -> cxxdes::core::subroutine<void>

// _Coroutine(...) { body }: an anonymous coroutine<void> from a lambda body (under_coroutine.ipp:11-17)
template <typename F>
auto operator+(cxxdes::core::detail::under_helper, F f) {
    return f();
}
#define _Coroutine(...) cxxdes::core::detail::under_helper{} + [&](__VA_ARGS__) -> coroutine<void>
