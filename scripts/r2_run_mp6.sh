N=$1; TAG=$2
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","device_ms_per_step","vs_one_gpu","n_gpus")}); print(d.get("one_gpu")); print(d.get("e2e"))
PY
timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 600 gpurun_out/${TAG}_bench_ref.json
