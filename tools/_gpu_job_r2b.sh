# One GPU call, bounded by a deadline: A/B of the opt-in kernel variants with the byte-equality gate (tools/ab_select.py), then the GPU
# test suite, smoke, the bench line and the ncu launch list under the SELECTED switches.  Every step writes into gpurun_out/ as it goes.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
START=$(date +%s)
LIMIT=${JOB_LIMIT:-460}
left() { echo $(( LIMIT - ($(date +%s) - START) )); }
step() {  # step <max seconds> <command...>: skipped when fewer than 20 s remain
  local max=$1; shift
  local l=$(left)
  if [ $l -lt 20 ]; then echo "SKIP (deadline): $*" >> gpurun_out/job.log; return 1; fi
  [ $l -lt $max ] && max=$l
  echo "[$(( $(date +%s) - START )) s] timeout $max $*" >> gpurun_out/job.log
  timeout $max "$@"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
step 420 python tools/ab_select.py --out gpurun_out/ab --workloads cfg2 cfg1 --steps 20 > gpurun_out/ab.log 2>&1
[ -f gpurun_out/ab/selected.env ] && . gpurun_out/ab/selected.env
env | grep '^SCB_' > gpurun_out/selected_env.txt
step 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_selected.log 2>&1
step 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_selected_full.log 2>&1
tail -15 gpurun_out/pytest_gpu_selected_full.log > gpurun_out/pytest_gpu_selected.log
step 150 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2_selected.json 2> gpurun_out/bench_cfg2_selected.err
step 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_selected.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
step 240 ncu --set full --clock-control none --import-source on -k "regex:rhs_fold|i8_gemm_p|tri_solve|i8_digitize|tri_low" -s 12 -c 7 -o gpurun_out/prof_selected python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
step 90 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_selected.json 2> gpurun_out/bench_cfg4_selected.err
step 90 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_selected.json 2> gpurun_out/bench_cfg1_selected.err
step 90 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg5_selected.json 2> gpurun_out/bench_cfg5_selected.err
echo "done at $(( $(date +%s) - START )) s" >> gpurun_out/job.log
