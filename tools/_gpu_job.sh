set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest_gpu.log
for w in cfg2 cfg1 cfg5 cfg4 cfg3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
PARITY_SEEDS="0 1 2 3" timeout 600 python tools/parity_report.py cfg1 cfg2 cfg5 > gpurun_out/final_parity.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:rhs_fold|i8_gemm_pkernel|tri_solve|i8_digitize|tri_low" -s 12 -c 7 -o gpurun_out/final_prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu2.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/final_smi.txt
