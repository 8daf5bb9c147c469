#!/bin/bash
# Compile-time A/B variants of libscb.so:  tools/build_variant.sh NAME [-DSCB_TRI_UNROLL=16 ...]
# Recompiles csrc/scb_api.cu (the only translation unit that sees the SCB_TRI_* / SCB_RHS2_* macros) with the extra flags and links it
# with the objects of the current main build -> seamlesscloneoptimization_b200/lib/variants/libscb_NAME.so.  Select it with
# SCB_LIBRARY=... (package, bench.py, tools/) or SCB_TEST_LIBRARY=... (pytest's cuda_lib fixture).  Variants are git-ignored artefacts.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
pkg=seamlesscloneoptimization_b200
out=$pkg/lib/variants
mkdir -p $out
python -c "import __graft_entry__ as g; g.build_cuda()"
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c $pkg/csrc/scb_api.cu -o $out/scb_api_$name.o
others=$(ls $pkg/lib/obj/*.o | grep -v '/scb_api\.')
nvcc --shared -gencode arch=compute_100a,code=sm_100a -o $out/libscb_$name.so $out/scb_api_$name.o $others
rm -f $out/scb_api_$name.o
echo "built $out/libscb_$name.so"
