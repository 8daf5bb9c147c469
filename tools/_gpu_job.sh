set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_p_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_p_bench_cfg2.json 2> gpurun_out/r2_p_bench_cfg2.err
SCB_I8_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_p_bench_cfg2_nofuse.json 2> gpurun_out/r2_p_bench_cfg2_nofuse.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r2_p_launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p_ncu.log 2>&1
