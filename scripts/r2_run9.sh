#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 100 -k "random_solve_small or fixed_cases or doctest or small_instance_path or cfg1" > gpurun_out/r2z_tests.log 2>&1
tail -3 gpurun_out/r2z_tests.log
timeout 200 python scripts/r2_small_e2e_probe.py > gpurun_out/r2z_small_probe.jsonl 2> gpurun_out/r2z_small_probe.err
cat gpurun_out/r2z_small_probe.jsonl; tail -3 gpurun_out/r2z_small_probe.err
timeout 200 python bench.py --workload cfg1 --steps 200 --warmup 20 > gpurun_out/r2z_bench_cfg1.json 2> gpurun_out/r2z_bench_cfg1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_cfg1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","cpu_baseline")})
PY
