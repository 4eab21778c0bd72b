N=$1; TAG=$2
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
SLA_MESH_TIMELINE=1 SLA_TAG=${TAG}_cfg5 timeout 300 $TR scripts/run_mesh.py 16000000 64000000 16 check 7 > gpurun_out/${TAG}_mesh_cfg5.json 2> gpurun_out/${TAG}_mesh_cfg5.err
tail -3 gpurun_out/${TAG}_mesh_cfg5.err; cat gpurun_out/${TAG}_mesh_cfg5.json
cat gpurun_out/${TAG}_cfg5_timeline_rank0.json
