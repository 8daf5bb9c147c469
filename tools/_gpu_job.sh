set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_j_pytest_gpu.log
for w in cfg2 cfg1 cfg5 cfg4 cfg3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_j_bench_$w.json 2> gpurun_out/r2_j_bench_$w.err
done
timeout 600 python bench.py > gpurun_out/r2_j_bench_default.json 2> gpurun_out/r2_j_bench_default.err
