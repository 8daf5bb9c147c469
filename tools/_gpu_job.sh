set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline.py -m gpu -q -x -k "batch" 2>&1 | tail -5 > gpurun_out/r2_t_pytest_batch.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_t_pytest_gpu.log
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t_bench_cfg3.json 2> gpurun_out/r2_t_bench_cfg3.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t_bench_cfg2.json 2> gpurun_out/r2_t_bench_cfg2.err
