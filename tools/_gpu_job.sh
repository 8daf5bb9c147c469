set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_o_ngpus.txt
nproc >> gpurun_out/r2_o_ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_o_bench_n8.json 2> gpurun_out/r2_o_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_o_bench_n4.json 2> gpurun_out/r2_o_bench_n4.err
