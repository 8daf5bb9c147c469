set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_r_smoke.log 2>&1
timeout 600 python -m pytest tests/test_pipeline.py -m gpu -q -x -k "i8 or int8 or golden or full_size_vs" 2>&1 | tail -4 > gpurun_out/r2_r_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_r_bench_cfg2.json 2> gpurun_out/r2_r_bench_cfg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_r_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_r_ncu.log 2>&1
