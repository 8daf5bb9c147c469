set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline.py -m gpu -q -x -k "banded or batch or golden or int8" 2>&1 | tail -5 > gpurun_out/r2_l_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_l_bench_cfg2.json 2> gpurun_out/r2_l_bench_cfg2.err
SCB_PLAN_CACHE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_l_bench_cfg2_nocache.json 2> gpurun_out/r2_l_bench_cfg2_nocache.err
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_l_bench_cfg4.json 2> gpurun_out/r2_l_bench_cfg4.err
for t in 1 2 4 8; do for l in 4 8; do
SCB_SUBMIT_THREADS=$t SCB_LANES=$l timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_l_bench_cfg3_t${t}_l${l}.json 2> gpurun_out/r2_l_bench_cfg3_t${t}_l${l}.err
done; done
