set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_m_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x -s 2>&1 | tail -30 > gpurun_out/r2_m_pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_m_bench_n2.json 2> gpurun_out/r2_m_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --steps 10 --warmup 3 --no-sharded-graph > gpurun_out/r2_m_bench_cfg4_n2_nograph.json 2> gpurun_out/r2_m_bench_cfg4_n2_nograph.err
tail -5 gpurun_out/r2_m_bench_n2.err
