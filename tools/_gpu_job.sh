set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/r2_a_pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/r2_a_pytest_gpu.log
for eng in tri fft; do
  PARITY_SEEDS="0 1 2 3" SCB_ENGINE=$eng python tools/parity_report.py cfg1 cfg2 cfg5 > gpurun_out/r2_a_parity_${eng}.txt 2>&1
done
PARITY_SEEDS="0 1 2 3" SCB_ENGINE=fft SCB_REFINE=0 python tools/parity_report.py cfg1 cfg2 cfg5 > gpurun_out/r2_a_parity_fft_norefine.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_a_bench_cfg2.json 2> gpurun_out/r2_a_bench_cfg2.err
