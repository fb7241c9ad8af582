// tests/native/plan_kernel_host.cpp -- the C++ host mirror's planned_kernel(): which first-pass kernel a model takes,
// asked before anything is created (no handle, no device).  Prints one line per case; tests/test_host_cpp.py compares.
#include <cstdio>
#include <stdexcept>

#include "cxxdes_b200/ensemble.hpp"

using namespace cxxdes::b200;
using namespace cxxdes::b200::time_ops;

int main() {
    {
        ensemble<models::aloha> ens{{64, 50, 0.1f, 3.0f, 1.0f, 1000.0f, 0}, 1u << 18};
        ens.env.time_unit(1_s);
        ens.env.time_precision(1_ms);
        std::printf("aloha %s\n", ens.planned_kernel().c_str());
    }
    {
        ensemble<models::mmc_balking> ens{{9.0, 1.0, 0.5, 8, 1023}, 1u << 20};
        ens.env.time_unit(1_s);
        ens.env.time_precision(1_ms);
        std::printf("mmc %s\n", ens.planned_kernel().c_str());
    }
    {
        ensemble<models::memory_hierarchy> ens{{1, 10, 16, 32, 4, 256}, 1u << 22};
        std::printf("archsim %s\n", ens.planned_kernel().c_str());
    }
    try {
        ensemble<models::producer_consumer> ens{{5.0, 10.0, 10000}, 1u << 20};
        std::printf("mm1 %s\n", ens.planned_kernel().c_str());
        return 1;
    } catch (const std::runtime_error &e) {
        std::printf("mm1 throws: %s\n", e.what());
    }
    return 0;
}
