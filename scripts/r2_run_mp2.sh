N=$1; TAG=$2
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
( timeout 300 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q --timeout 100 -k mesh 2>&1 | tail -5 ) > gpurun_out/${TAG}_mesh_tests.log
cat gpurun_out/${TAG}_mesh_tests.log
SLA_MESH_TIMELINE=1 SLA_TAG=${TAG}_cfg3shape timeout 200 $TR scripts/run_mesh.py 1000000 4000000 16 check 5 > gpurun_out/${TAG}_mesh_cfg3shape.json 2> gpurun_out/${TAG}_mesh_cfg3shape.err
tail -3 gpurun_out/${TAG}_mesh_cfg3shape.err; cat gpurun_out/${TAG}_mesh_cfg3shape.json
SLA_MESH_TIMELINE=1 SLA_TAG=${TAG}_cfg5 timeout 300 $TR scripts/run_mesh.py 16000000 64000000 16 check 7 > gpurun_out/${TAG}_mesh_cfg5.json 2> gpurun_out/${TAG}_mesh_cfg5.err
tail -3 gpurun_out/${TAG}_mesh_cfg5.err; cat gpurun_out/${TAG}_mesh_cfg5.json
cat gpurun_out/${TAG}_cfg5_timeline_rank0.json
timeout 120 python scripts/r2_probe.py quick > gpurun_out/${TAG}_probe.json 2>&1; cat gpurun_out/${TAG}_probe.json
