// tests/native/example_recorder.cpp -- TEST INFRASTRUCTURE.  ONE OF THE REFERENCE'S OWN EXAMPLES, its source file
// included where it lies and unchanged (-DEXAMPLE_FILE="</root/reference/examples/x.cpp>"; only its main() is renamed),
// compiled against the RECORDING <cxxdes/cxxdes.hpp> (-I include/cxxdes_b200/record): the coroutine bodies are recorded
// into a process program, written as JSON to argv[1].  The example's prints go to stdout as the bodies run on the host
// (now() reads 0 there).  Two shapes of example:
//   -DEXAMPLE_SIM=its CXXDES_SIMULATION class [-DEXAMPLE_ARGS=constructor arguments] [-DEXAMPLE_UNTIL=t]: the harness
//       makes the object and records it (what .run() / .run_for(t) of the example's main() would run);
//   -DEXAMPLE_MAIN: the example drives an `environment` by hand (env.bind(...); env.run() / env.run_for(t)): its main()
//       runs as it is, and the model and horizon of its last env.run*() are taken from the recorder's log of such
//       requests (-DCXXDES_B200_RECORD_ONLY: noted, not run on a GPU) when the process exits ($EXAMPLE_OUT).
#if !defined(EXAMPLE_MAIN)
#define main cxxdes_example_main
#endif
#include EXAMPLE_FILE
#undef main

#include <cstdlib>

#include "program_dump.hpp"

static int write_program(const char *path, const cxxdes::b200::program &prog, cxxdes::b200::time unit, cxxdes::b200::time precision,
                         long long until, long long now = 0) {
    std::FILE *out = std::fopen(path, "w");
    if (!out) return 2;
    std::fprintf(out, "{\n\"until\": %lld,\n\"now\": %lld,\n", until, now);
    dump_program(out, "program", prog, unit, precision, true);
    std::fprintf(out, "}\n");
    std::fclose(out);
    return 0;
}

#if defined(EXAMPLE_MAIN)
// The example's own main() is THE main (several of them end without a return statement, which only main may do): the
// program is written when the process exits, to the file $EXAMPLE_OUT names.
static void write_at_exit() {
    const char *path = std::getenv("EXAMPLE_OUT");
    if (!path || cxxdes::record::run_requests().empty()) return;
    const cxxdes::record::run_request &req = cxxdes::record::run_requests().back();
    // (without -DCXXDES_B200_RECORD_ONLY the example's env.run*() ran one replication on the device: "now" is its clock)
    write_program(path, cxxdes::core::environment::program_of(*req.model), req.model->unit, req.model->precision, req.until,
                  req.model->last_now);
}
static const int registered_at_start = std::atexit(write_at_exit);
#else
int main(int argc, char **argv) {
    if (argc < 2) return 2;
#if defined(EXAMPLE_ARGS)
    EXAMPLE_SIM sim{EXAMPLE_ARGS};
#else
    EXAMPLE_SIM sim;
#endif
    const cxxdes::b200::program prog = sim.record();
#if defined(EXAMPLE_UNTIL)
    const long long until = EXAMPLE_UNTIL;
#else
    const long long until = -1;
#endif
    return write_program(argv[1], prog, sim.env.time_unit(), sim.env.time_precision(), until);
}
#endif
