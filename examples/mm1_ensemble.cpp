// examples/mm1_ensemble.cpp -- the M/M/1 sweep of the reference's examples/producer_consumer.cpp:61-75,
// each load point run as an ensemble of independent replications on a B200.
//
//   g++ -std=c++17 -O2 -Iinclude examples/mm1_ensemble.cpp -Llibcxxdes_b200/_lib -lcxxdes_b200
//       -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o mm1_ensemble
#include <cstdio>
#include <cstdlib>

#include "cxxdes_b200/ensemble.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

int main(int argc, char **argv) {
    const std::uint64_t replications = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4096;
    const std::uint64_t n_packets = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 10000;
    const double lambda_start = 0.5, lambda_end = 10.0, mu = lambda_end;
    const int n_steps = 5;
    try {
        std::printf("lambda, mu, rho, avg_latency (mean over %llu replications), ci95, events/s\n",
                    (unsigned long long)replications);
        for (int i = 0; i < n_steps; ++i) {
            double lambda = lambda_start + (lambda_end - 0.5 - lambda_start) * i / (n_steps - 1);
            ensemble<models::producer_consumer> ens{{lambda, mu, n_packets}, replications};
            ens.env.time_unit(1_s);
            ens.env.time_precision(1_ms);
            ens.run();
            auto acc = ens.reduce();
            const estimate latency = mean_ci(acc, 0);
            std::printf("%.3f, %.3f, %.3f, %.6f, %.6f, %.3e\n", lambda, mu, lambda / mu, latency.mean, latency.ci95,
                        acc.n_events / (ens.last_run_ms() * 1e-3));
        }
    } catch (std::runtime_error const &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
