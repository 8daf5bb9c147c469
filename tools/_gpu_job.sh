set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_i_bench.json 2> gpurun_out/r2_i_bench.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:i8_gemm_pkernel -s 6 -c 2 -o gpurun_out/r2_i_i8prof python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_i_ncu.log 2>&1
