set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_n_pytest_multigpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_n_bench_cfg2.json 2> gpurun_out/r2_n_bench_cfg2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_n_bench_cfg2_reference.json 2> gpurun_out/r2_n_bench_cfg2_reference.err
