#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 200 python scripts/r2_small_e2e_probe.py > gpurun_out/r2y_small_probe.jsonl 2> gpurun_out/r2y_small_probe.err
cat gpurun_out/r2y_small_probe.jsonl; tail -3 gpurun_out/r2y_small_probe.err
timeout 200 python bench.py --workload cfg1 --steps 200 --warmup 20 > gpurun_out/r2y_bench_cfg1.json 2> gpurun_out/r2y_bench_cfg1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_bench_cfg1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","cpu_baseline")})
PY
