// examples/aloha_sweep.cpp -- the offered-load sweep of the reference's examples/aloha.cpp:87-109 ("G [Erlang],
// S [Erlang]", 50 load points) as ONE ensemble: replication r runs load point r mod 50, so every row of the table
// is a mean over replications/50 independent runs instead of a single run, with its 95 % interval.
//
//   g++ -std=c++17 -O2 -Iinclude examples/aloha_sweep.cpp -Llibcxxdes_b200/_lib -lcxxdes_b200
//       -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o aloha_sweep
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "cxxdes_b200/ensemble.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

int main(int argc, char **argv) {
    const std::uint64_t replications = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 5000;
    const std::uint32_t num_stations = argc > 2 ? (std::uint32_t)std::atoi(argv[2]) : 64;
    const float packets_scale = argc > 3 ? (float)std::atof(argv[3]) : 1000.0f;   // packets_per_station = size_t(lambda * 1000)
    const std::uint32_t num_steps = 50;
    try {
        models::aloha model;
        model.num_stations = num_stations;
        model.sweep_steps = num_steps;       // lambda = begin + (end - begin) * i / (num_steps - 1), i = r mod num_steps
        model.lambda_begin = 0.1f;
        model.lambda_end = 3.0f;
        model.frame_time = 1.0f;
        model.packets_scale = packets_scale;
        model.packets_per_station = 0;       // 0: derived from lambda as in the example's main()
        ensemble<models::aloha> ens{model, replications};
        ens.env.time_unit(1_s);
        ens.env.time_precision(1_ms);
        ens.run();
        const auto &S = ens.stat(0), &G = ens.stat(1);
        std::printf("G [Erlang], S [Erlang], ci95, replications\n");
        for (std::uint32_t i = 0; i < num_steps; ++i) {
            double sum = 0, sumsq = 0;
            std::uint64_t n = 0;
            for (std::uint64_t r = i; r < replications; r += num_steps, ++n) {
                sum += S[r];
                sumsq += S[r] * S[r];
            }
            if (n == 0) continue;
            const double mean = sum / (double)n;
            const double var = n > 1 ? std::fmax(0.0, (sumsq - (double)n * mean * mean) / (double)(n - 1)) : 0.0;
            std::printf("%g, %g, %.3g, %llu\n", G[i], mean, 1.959964 * std::sqrt(var / (double)n), (unsigned long long)n);
        }
    } catch (std::runtime_error const &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
