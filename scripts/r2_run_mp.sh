# usage: bash scripts/r2_run_mp.sh N TAG   (under gpurun --gpus N)
N=$1; TAG=$2
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 200 $TR scripts/run_mesh.py 1000000 4000000 16 check 5 > gpurun_out/${TAG}_mesh_cfg3shape.json 2> gpurun_out/${TAG}_mesh_cfg3shape.err
tail -3 gpurun_out/${TAG}_mesh_cfg3shape.err; cat gpurun_out/${TAG}_mesh_cfg3shape.json
timeout 300 $TR scripts/run_mesh.py 16000000 64000000 16 check 7 > gpurun_out/${TAG}_mesh_cfg5.json 2> gpurun_out/${TAG}_mesh_cfg5.err
tail -3 gpurun_out/${TAG}_mesh_cfg5.err; cat gpurun_out/${TAG}_mesh_cfg5.json
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
