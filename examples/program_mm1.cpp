// examples/program_mm1.cpp -- a model written against the host C++ builder (cxxdes_b200/program.hpp) instead of
// picking one of the built-in kernels: the M/M/1 queue of the reference's examples/producer_consumer.cpp:31-58,
// statement by statement, run as a process program on the interpreter kernel and -- as a cross-check -- by the
// hand-written M/M/1 kernel.  Both must produce the same event trace for every replication.
//
//   g++ -std=c++17 -O2 -Iinclude examples/program_mm1.cpp -Llibcxxdes_b200/_lib -lcxxdes_b200
//       -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o program_mm1
#include <cstdio>
#include <cstdlib>

#include "cxxdes_b200/program.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

static program producer_consumer_example(double lambda, double mu, std::int64_t n_packets) {
    program prog;
    queue q = prog.add_queue();                              // sync::queue<double> q;
    process co_main = prog.add_process(0);
    process producer = prog.add_process(1), consumer = prog.add_process(2);
    prog.bind(co_main);                                      // simulation<>::run() binds co_main
    prog.body_of(producer, [&](body &co) {                   // coroutine<> producer() {
        co.loop(n_packets, [&] {                             //   for (i = 0; i < n_packets; ++i) {
            co.await(q.put());                               //     co_await q.put(now_seconds());
            co.await(env_timeout(exponential{lambda}));      //     co_await env.timeout(lambda()); } }
        });
    });
    prog.body_of(consumer, [&](body &co) {                   // coroutine<> consumer() {
        co.loop(1 << 30, [&] {                               //   while (true) {
            co.await(q.pop()).count(0);                      //     auto x = co_await q.pop(); ++n;
            co.if_then(counter(0) == n_packets,              //     if (n == n_packets) {
                       [&] { co.co_return_(); });            //       avg_latency = total_latency / n; co_return; }
            co.await(env_timeout(exponential{mu}));          //     co_await env.timeout(mu());
            co.stat_seconds(1);                              //     total_latency += now_seconds() - x; }
        });
    });
    prog.body_of(co_main, [&](body &co) {                    // coroutine<> co_main() {
        co.await(producer && consumer);                      //   co_await (producer() && consumer()); }
    });
    return prog;
}

int main(int argc, char **argv) {
    const std::uint64_t replications = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4096;
    const std::int64_t n_packets = argc > 2 ? std::strtoll(argv[2], nullptr, 10) : 1000;
    const double lambda = 5.0, mu = 10.0;
    try {
        ensemble<models::process_program> prog{{producer_consumer_example(lambda, mu, n_packets)}, replications};
        prog.env.time_unit(1_s);
        prog.env.time_precision(1_ms);
        prog.run();
        ensemble<models::producer_consumer> fused{{lambda, mu, (std::uint64_t)n_packets}, replications};
        fused.env.time_unit(1_s);
        fused.env.time_precision(1_ms);
        fused.run();
        std::uint64_t different = 0;
        for (std::uint64_t r = 0; r < replications; ++r)
            different += prog.trace_hash()[r] != fused.trace_hash()[r] || prog.now()[r] != fused.now()[r] ||
                         prog.n_events()[r] != fused.n_events()[r] || prog.stat(1)[r] != fused.stat(1)[r];
        auto acc = prog.reduce();
        std::printf("replications %llu, events %llu, mean total latency per packet %.6f s (theory %.6f), "
                    "interpreter %.3e events/s, hand-written kernel %.3e events/s, differing replications %llu\n",
                    (unsigned long long)replications, (unsigned long long)acc.n_events,
                    acc.f64_sum[1] / (double)acc.n_replications / (double)n_packets, 1.0 / (mu - lambda),
                    acc.n_events / (prog.last_run_ms() * 1e-3), acc.n_events / (fused.last_run_ms() * 1e-3),
                    (unsigned long long)different);
        return different ? 1 : 0;
    } catch (std::runtime_error const &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
