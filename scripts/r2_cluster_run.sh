#!/bin/bash
# round 2: first GPU run of the cluster engine -- its parity test, the neighbouring option tests, the probe, a bench line
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 400 --timeout-method=thread \
  -k "cluster_engine or engine_options or bound_pruned or learn_the_graph or seeded_instances" > gpurun_out/r2w_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2w_tests.log
tail -5 gpurun_out/r2w_tests.log
timeout 300 python scripts/r2_cluster_probe.py > gpurun_out/r2w_probe.jsonl 2> gpurun_out/r2w_probe.err
tail -c 3000 gpurun_out/r2w_probe.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench_cfg3.json 2> gpurun_out/r2w_bench.err
tail -c 600 gpurun_out/r2w_bench_cfg3.json
