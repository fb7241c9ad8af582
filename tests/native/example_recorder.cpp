// tests/native/example_recorder.cpp -- TEST INFRASTRUCTURE.  ONE OF THE REFERENCE'S OWN EXAMPLES, its source file
// included where it lies and unchanged (-DEXAMPLE_FILE="</root/reference/examples/x.cpp>", -DEXAMPLE_SIM=its simulation
// class; only its main() is renamed), compiled against the RECORDING <cxxdes/cxxdes.hpp>
// (-I include/cxxdes_b200/record): the coroutine bodies are recorded into a process program, written as JSON to argv[1].
// The example's prints go to stdout as the bodies run on the host (now() reads 0 there).
#define main cxxdes_example_main
#include EXAMPLE_FILE
#undef main

#include "program_dump.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    EXAMPLE_SIM sim;
    const cxxdes::b200::program prog = sim.record();
    std::FILE *out = std::fopen(argv[1], "w");
    if (!out) return 2;
    std::fprintf(out, "{\n");
    dump_program(out, "program", prog, sim.env.time_unit(), sim.env.time_precision(), true);
    std::fprintf(out, "}\n");
    std::fclose(out);
    return 0;
}
