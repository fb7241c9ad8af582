// examples/mm1k_loss.cpp -- a model with a decision between its awaits, written with the host C++ builder: the
// M/M/1/K loss system.  The producer looks at the buffer before it puts a packet,
//     if (q.size() < k) co_await q.put(now_seconds()); else ++dropped;
// and the fraction of dropped packets is compared with the textbook blocking probability
//     P = (1 - rho) rho^N / (1 - rho^(N + 1)),   N = k + 1  (k waiting, one in service).
// A sweep over the offered load runs inside ONE ensemble: replication r uses load point r % n_points.
//
//   g++ -std=c++17 -O2 -Iinclude examples/mm1k_loss.cpp -Llibcxxdes_b200/_lib -lcxxdes_b200
//       -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o mm1k_loss
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "cxxdes_b200/program.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

int main(int argc, char **argv) {
    const std::uint64_t replications = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 20000;
    const std::int64_t n_packets = argc > 2 ? std::strtoll(argv[2], nullptr, 10) : 4000;
    const std::int64_t k = 4;
    const double mu = 10.0;
    const std::vector<double> lambdas = {2.0, 4.0, 6.0, 8.0, 9.5};
    try {
        program prog;
        prog.sweep((std::uint32_t)lambdas.size());
        queue q = prog.add_queue();
        process co_main = prog.add_process(0), producer = prog.add_process(1), consumer = prog.add_process(2);
        prog.bind(co_main);
        prog.body_of(producer, [&](body &co) {
            co.loop(n_packets, [&] {
                co.if_then_else(q.size() < k, [&] { co.await(q.put()); },     // room: co_await q.put(now_seconds());
                                [&] { co.count(1); });                        // full: ++dropped;
                co.await(env_timeout(exponential{lambdas}));
            });
        });
        prog.body_of(consumer, [&](body &co) {
            co.loop(1 << 30, [&] {                                            // serves until the run ends
                co.await(q.pop());
                co.await(env_timeout(exponential{mu})).count(0);
            });
        });
        prog.body_of(co_main, [&](body &co) {
            co.await(async(consumer));
            co.await(producer);
        });

        ensemble<models::process_program> ens{{prog}, replications};
        ens.env.time_unit(1_s);
        ens.env.time_precision(1_ms);
        ens.run();
        const auto &dropped = ens.count(1);
        std::printf("lambda, rho, dropped fraction, theory, replications\n");
        int bad = 0;
        for (std::size_t point = 0; point < lambdas.size(); ++point) {
            double sum = 0;
            std::uint64_t n = 0;
            for (std::uint64_t r = point; r < replications; r += lambdas.size(), ++n) sum += (double)dropped[r];
            const double rho = lambdas[point] / mu, cap = (double)(k + 1);
            const double theory = (1 - rho) * std::pow(rho, cap) / (1 - std::pow(rho, cap + 1));
            const double got = sum / ((double)n * (double)n_packets);
            std::printf("%.1f, %.2f, %.5f, %.5f, %llu\n", lambdas[point], rho, got, theory, (unsigned long long)n);
            bad += std::fabs(got - theory) > 0.004 + 0.03 * theory;
        }
        return bad ? 1 : 0;
    } catch (std::runtime_error const &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
