N=$1; TAG=$2
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
