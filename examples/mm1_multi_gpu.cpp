// examples/mm1_multi_gpu.cpp -- one M/M/1 ensemble sharded over every GPU of the node from a C++ host:
// one thread per GPU, shard g runs the global replications [g*R, (g+1)*R), and the only exchange is the NCCL
// all-reduce of the accumulator struct at the end (ensemble::reduce_allgpus).  Results do not depend on the
// number of GPUs: the Philox key holds the global replication id.
//
//   g++ -std=c++17 -O2 -pthread -Iinclude -I/usr/local/cuda/include examples/mm1_multi_gpu.cpp
//       -Llibcxxdes_b200/_lib -lcxxdes_b200 -lnccl -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o mm1_multi_gpu
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "cxxdes_b200/ensemble.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

int main(int argc, char **argv) {
    const std::uint64_t per_gpu = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 65536;
    const std::uint64_t n_packets = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1000;
    int n_gpus = 0;
    if (cxxdes_b200_device_count(&n_gpus) != CXXDES_B200_OK || n_gpus < 1) {
        std::fprintf(stderr, "error: %s\n", cxxdes_b200_last_error());
        return 2;
    }
    if (argc > 3 && std::atoi(argv[3]) < n_gpus) n_gpus = std::atoi(argv[3]);
    std::vector<ncclComm_t> comms(n_gpus);
    if (ncclCommInitAll(comms.data(), n_gpus, nullptr) != ncclSuccess) {
        std::fprintf(stderr, "error: ncclCommInitAll failed\n");
        return 2;
    }
    std::vector<cxxdes_b200_accumulators> total(n_gpus), local(n_gpus);
    std::vector<float> ms(n_gpus);
    std::vector<std::string> errors(n_gpus);
    std::vector<std::thread> threads;
    for (int g = 0; g < n_gpus; ++g)
        threads.emplace_back([&, g] {
            try {
                ensemble<models::producer_consumer> ens{{5.0, 10.0, n_packets}, per_gpu, /*seed*/ 0x5EED,
                                                        /*first_replication*/ (std::uint64_t)g * per_gpu, /*device*/ g};
                ens.env.time_unit(1_s);
                ens.env.time_precision(1_ms);
                ens.run();
                local[g] = ens.reduce();
                total[g] = ens.reduce_allgpus(comms[g]);
                ms[g] = ens.last_run_ms();
            } catch (std::runtime_error const &e) {
                errors[g] = e.what();
            }
        });
    for (auto &t : threads) t.join();
    for (int g = 0; g < n_gpus; ++g) ncclCommDestroy(comms[g]);
    for (int g = 0; g < n_gpus; ++g)
        if (!errors[g].empty()) {
            std::fprintf(stderr, "error on GPU %d: %s\n", g, errors[g].c_str());
            return 2;
        }
    // every rank holds the same totals, and they are the sums of the shard accumulators
    std::uint64_t events = 0, hash = 0;
    float slowest = 0;
    for (int g = 0; g < n_gpus; ++g) {
        events += local[g].n_events;
        hash += local[g].trace_hash_sum;
        if (ms[g] > slowest) slowest = ms[g];
    }
    int bad = 0;
    for (int g = 0; g < n_gpus; ++g)
        bad += total[g].n_events != events || total[g].trace_hash_sum != hash ||
               total[g].n_replications != per_gpu * (std::uint64_t)n_gpus || total[g].f64_sum[0] != total[0].f64_sum[0];
    const estimate latency = mean_ci(total[0], 0);
    std::printf("gpus %d, replications %llu, events %llu, digest sum %016llx, avg_latency %.6f +- %.6f, "
                "%.3e events/s (slowest shard %.3f ms), inconsistent ranks %d\n",
                n_gpus, (unsigned long long)total[0].n_replications, (unsigned long long)events, (unsigned long long)hash,
                latency.mean, latency.ci95, events / (slowest * 1e-3), (double)slowest, bad);
    return bad ? 1 : 0;
}
