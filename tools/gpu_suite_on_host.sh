#!/bin/bash
# The whole `-m gpu` suite against the CPU emulation of the device (tests/native/cuda_on_host): every kernel of
# libcxxdes_b200/csrc compiled by g++, CUDA threads as fibres.  TEST INFRASTRUCTURE -- checks the kernels' logic where no
# GPU is at hand; says nothing about speed.  tests/test_kernels_on_host.py runs the same thing without the tests that
# are sized for the device; this script runs everything (about a quarter of an hour on 8 cores; expect the tests that
# need two kernels at once or link the CUDA build to fail: see DESELECT in tests/test_kernels_on_host.py).
#   tools/gpu_suite_on_host.sh [--asan] [pytest arguments, e.g. tests/test_models_gpu.py -k aloha]
# --asan: the emulated library built with -fsanitize=address,undefined -- a memory and undefined-behaviour check of
# every kernel and of the C ABI's host side (device buffers and shared memory are heap blocks of exactly the size asked
# for); 3-4 times slower.  UBSAN_OPTIONS=halt_on_error=1 turns a report into a failed test.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /dev/null     # (nvcc must not run under the ASan preload)
if [ "$1" = "--asan" ]; then
    shift
    lib=$(python tests/native/cuda_on_host/build.py --asan | tail -1)
    export LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0
else
    lib=$(python tests/native/cuda_on_host/build.py | tail -1)
fi
# (the C++ hosts of examples/ link -lcxxdes_b200: under that name they find the emulated library here)
export CXXDES_B200_LIBRARY="$lib" CUDA_ON_HOST_SMS="${CUDA_ON_HOST_SMS:-2}" LD_LIBRARY_PATH="$(dirname "$lib")/lib"
if [ $# -eq 0 ]; then set -- tests; fi
exec python -m pytest "$@" -m gpu -q -p no:cacheprovider -n "${JOBS:-6}" --timeout 3600 --durations=15
