// oracle/port/engine.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Coroutine-free CPU restatement of the reference's event-dispatch path, so the
// checker also exists on machines where /root/reference does not (the GPU box).
// PINNED: validated event-by-event against the real reference engine
// (oracle/_ref) in tests/test_oracle_vs_reference.py and against the golden
// vectors under tests/golden/ that were produced by the reference itself.
//
// Every function cites the reference code it restates (paths relative to the
// reference root).  Plain C++17 + the shared arithmetic header; no CUDA.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "cxxdes_b200.h"
#include "cxxdes_b200/shared/philox.hpp"
#include "cxxdes_b200/shared/trace_hash.hpp"
#include "cxxdes_b200/shared/variates.hpp"

namespace port_oracle {

constexpr int64_t PRIO_INHERIT = INT64_MAX;  // priority_consts::inherit, defs.ipp:35
constexpr int32_t NO_TARGET = -1;
constexpr int32_t NO_HANDLER = -1;

// token: include/cxxdes/core/impl/token.ipp:6-62 ({time, priority, coro_data, handler}).
struct token {
    int64_t time = 0;
    int64_t priority = 0;
    int32_t target = NO_TARGET;    // process index (coro_data), -1 = null
    int32_t handler = NO_HANDLER;  // any_of/all_of handler index, -1 = null
};

// environment::token_comp, environment.ipp:255-260 (exception tokens do not exist here):
// true when a is strictly LATER than b on (time, priority).
inline bool token_later(const token &a, const token &b) {
    return a.time > b.time || (a.time == b.time && a.priority > b.priority);
}

// std::priority_queue<token*, std::vector<token*>, token_comp> (environment.ipp:263) as
// implemented by libstdc++ 13: stl_queue.h:738-778, stl_heap.h:135-148 (__push_heap),
// :224-249 (__adjust_heap), :254-270 (__pop_heap).  Tie order among equal keys is a
// property of exactly this algorithm.
struct ref_heap {
    std::vector<token> v;

    bool empty() const { return v.empty(); }
    size_t size() const { return v.size(); }
    const token &top() const { return v.front(); }

    static void sift_up(std::vector<token> &v, size_t hole, size_t top_index, const token &x) {
        while (hole > top_index) {
            size_t parent = (hole - 1) / 2;
            if (!token_later(v[parent], x)) break;
            v[hole] = v[parent];
            hole = parent;
        }
        v[hole] = x;
    }

    void push(const token &x) {
        v.push_back(x);
        sift_up(v, v.size() - 1, 0, x);
    }

    void pop() {
        size_t n = v.size();
        if (n > 1) {
            token val = v[n - 1];
            v[n - 1] = v[0];
            size_t len = n - 1;
            size_t hole = 0;
            size_t child = 0;
            while (child < (len - 1) / 2) {  // len >= 1 here
                child = 2 * (child + 1);
                if (token_later(v[child], v[child - 1])) --child;
                v[hole] = v[child];
                hole = child;
            }
            if ((len & 1) == 0 && child == (len - 2) / 2) {
                child = 2 * (child + 1);
                v[hole] = v[child - 1];
                hole = child - 1;
            }
            sift_up(v, hole, 0, val);
        }
        v.pop_back();
    }
};

// One process = one coroutine<> with its coroutine_data (coroutine_data.ipp:1-228).
struct process {
    uint32_t ordinal = 0;            // name used in the trace digest
    int64_t priority = PRIO_INHERIT; // coroutine_data::priority_
    int64_t latency = 0;             // coroutine_data::latency_
    bool bound = false;
    bool complete = false;
    int pc = 0;                      // resume point of the lowered body
    std::vector<token> completion;   // completion_tokens_, coroutine_data.ipp:179-197
};

// sync::event: include/cxxdes/sync/event.hpp:87-139.
struct event {
    std::vector<token> waiters;
};

// any_all_helper::custom_handler: any_of.ipp:3-27.
struct any_all_handler {
    bool is_all = false;
    size_t total = 0;
    size_t remaining = 0;
    bool has_completion = false;
    token completion_tkn;
};

struct trace_event {
    int64_t time;
    int64_t priority;
    uint32_t kind;
    uint32_t proc;
};

struct engine {
    int64_t now = 0;
    ref_heap heap;
    std::vector<process> procs;
    std::vector<any_all_handler> handlers;
    uint64_t n_events = 0;
    uint64_t hash = cxxdes_b200::TRACE_HASH_INIT;
    size_t peak_heap = 0;
    std::vector<trace_event> *dump = nullptr;
    size_t dump_limit = 0;

    int32_t add_process(uint32_t ordinal, int64_t priority = PRIO_INHERIT, int64_t latency = 0) {
        process p;
        p.ordinal = ordinal;
        p.priority = priority;
        p.latency = latency;
        procs.push_back(p);
        return (int32_t)procs.size() - 1;
    }

    // environment::schedule_token, environment.ipp:94-98.
    void schedule(const token &t) {
        heap.push(t);
        if (heap.size() > peak_heap) peak_heap = heap.size();
    }

    // coroutine_data::bind_, environment.ipp:281-307: start token {now + latency, priority};
    // priority inherits from the awaiting process when left at `inherit`
    // (await_transform.ipp:50-52).  Binding twice is a no-op.
    void bind(int32_t p, int64_t inherited_priority) {
        process &pr = procs[p];
        if (pr.bound) return;
        pr.bound = true;
        if (pr.priority == PRIO_INHERIT) pr.priority = inherited_priority;
        schedule(token{now + pr.latency, pr.priority, p, NO_HANDLER});
    }

    // coroutine::await_suspend, coroutine.ipp:194-207: a completion token is created and
    // registered on the child, not scheduled.  Returns its index in child.completion.
    size_t await_completion(int32_t child, int32_t awaiter, int32_t handler = NO_HANDLER,
                            int64_t return_latency = 0, int64_t return_priority = PRIO_INHERIT) {
        process &c = procs[child];
        token t{return_latency, return_priority, awaiter, handler};
        if (t.priority == PRIO_INHERIT) t.priority = c.priority;
        c.completion.push_back(t);
        return c.completion.size() - 1;
    }

    // coroutine_data_::do_return, environment.ipp:343-348 + schedule_completion_ :321-329.
    void do_return(int32_t p) {
        process &pr = procs[p];
        for (token t : pr.completion) {
            t.time += now;
            schedule(t);
        }
        pr.completion.clear();
        pr.complete = true;
    }

    // environment::bind(coroutine), coroutine.ipp:352-357: start + null-target completion.
    void bind_root(int32_t p) {
        bind(p, 0);
        if (!procs[p].complete) await_completion(p, NO_TARGET);
    }

    // timeout_base::await_suspend, timeout.ipp:21-32 (relative delay).
    void timeout(int32_t p, int64_t ticks, int64_t priority = PRIO_INHERIT,
                 int32_t handler = NO_HANDLER) {
        if (priority == PRIO_INHERIT) priority = procs[p].priority;
        schedule(token{now + ticks, priority, p, handler});
    }

    // wait_awaitable::await_suspend, event.hpp:136-139.
    void wait(event &e, int32_t p, int64_t latency = 0, int64_t priority = PRIO_INHERIT) {
        if (priority == PRIO_INHERIT) priority = procs[p].priority;
        e.waiters.push_back(token{latency, priority, p, NO_HANDLER});
    }

    // wake_awaitable::await_ready, event.hpp:125-134.
    void wake(event &e) {
        for (token t : e.waiters) {
            t.time += now;
            schedule(t);
        }
        e.waiters.clear();
    }

    int32_t add_handler(bool is_all, size_t total, size_t remaining, const token &out) {
        any_all_handler h;
        h.is_all = is_all;
        h.total = total;
        h.remaining = remaining;
        h.has_completion = true;
        h.completion_tkn = out;
        handlers.push_back(h);
        return (int32_t)handlers.size() - 1;
    }

    // custom_handler::invoke, any_of.ipp:9-26; conditions any_of.ipp:233-247.
    void invoke_handler(const token &tkn) {
        any_all_handler &h = handlers[tkn.handler];
        --h.remaining;
        bool cond = h.is_all ? (h.remaining == 0) : (h.remaining < h.total);
        if (h.has_completion && cond) {
            token out = h.completion_tkn;
            out.time += tkn.time;
            out.priority = tkn.priority;
            out.target = tkn.target;
            schedule(out);
            h.has_completion = false;
        }
    }

    void observe(const token &t) {
        uint32_t kind = t.handler != NO_HANDLER
                            ? cxxdes_b200::TOKEN_HANDLER
                            : (t.target != NO_TARGET ? cxxdes_b200::TOKEN_RESUME
                                                     : cxxdes_b200::TOKEN_NOOP);
        uint32_t proc = t.target != NO_TARGET ? procs[t.target].ordinal : cxxdes_b200::PROC_NONE;
        hash = cxxdes_b200::trace_hash_step(hash, t.time, t.priority, kind, proc);
        ++n_events;
        if (dump && dump->size() < dump_limit)
            dump->push_back(trace_event{t.time, t.priority, kind, proc});
    }

    // environment::step, environment.ipp:117-146.
    template <typename Model>
    bool step(Model &m) {
        if (heap.empty()) return false;
        token t = heap.top();
        observe(t);
        heap.pop();
        now = t.time > now ? t.time : now;
        if (t.handler != NO_HANDLER)
            invoke_handler(t);
        else if (t.target != NO_TARGET)
            m.resume(t.target);
        return true;
    }

    // environment::run, environment.ipp:179-182; run_until, :190-196.
    template <typename Model>
    void run(Model &m, int64_t until) {
        if (until < 0) {
            while (step(m)) {
            }
        } else {
            while (!heap.empty() && heap.top().time <= until) step(m);
            now = now > until ? now : until;
        }
    }
};

}  // namespace port_oracle
