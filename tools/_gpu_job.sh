set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_q_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_q_bench_cfg2.json 2> gpurun_out/r2_q_bench_cfg2.err
SCB_I8_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_q_bench_cfg2_nofuse.json 2> gpurun_out/r2_q_bench_cfg2_nofuse.err
for w in cfg1 cfg5 cfg4; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_q_bench_$w.json 2> gpurun_out/r2_q_bench_$w.err
done
