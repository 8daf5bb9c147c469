#!/bin/bash
# Memory checking without a GPU tool (compute-sanitizer is closed on this pool): the emulator build of the kernel + driver sources
# (tests/emu) compiled with AddressSanitizer.  "Device" memory is malloc'd there, so every out-of-bounds global access of a kernel
# lands in a redzone.  Runs tools/sanitize_smoke.py and, with --tests, the emulator half of the parity tests.
#   tools/asan_emu.sh [--tests]
set -e
cd "$(dirname "$0")/.."
OUT=${ASAN_OUT:-/tmp/scb_asan}
mkdir -p "$OUT"
LIB="$OUT/libscb_emu_asan.so"
if [ ! -e "$LIB" ] || [ -n "$(find seamlesscloneoptimization_b200/csrc include tests/emu/emu_cuda.h -newer "$LIB" -type f | head -1)" ]; then
    # one translation unit with every kernel template instantiated: ~20 minutes with -fsanitize=address -g
    g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -DSCB_EMU -x c++ -Itests/emu -Iseamlesscloneoptimization_b200/csrc -ffp-contract=off \
        -shared -fPIC -Wno-unknown-pragmas seamlesscloneoptimization_b200/csrc/scb_api.cu seamlesscloneoptimization_b200/csrc/scb_i8.cu seamlesscloneoptimization_b200/csrc/scb_stamp.cu -o "$LIB" -lpthread
fi
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1   # fibers: swapcontext is only partly supported
SCB_LIBRARY="$OUT/libscb_emu_asan.so" python tools/sanitize_smoke.py 2>&1 | grep -v "doesn't fully support makecontext"
if [ "$1" == "--tests" ]; then
    SCB_EMU_LIBRARY="$OUT/libscb_emu_asan.so" python -m pytest tests/test_pipeline.py -x -q -m "not gpu" -k "emu and (transform_length or golden or flags or orientations or batch or sharded or int8 or plan_cache or two_contexts or corner)" 2>&1 | grep -v "doesn't fully support makecontext" | tail -5
fi
