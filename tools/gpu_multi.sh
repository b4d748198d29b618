set -x
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/dist_gpu_check.py > gpurun_out/m${N}_distcheck.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 300 --warmup 5 --no-e2e > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
