#!/bin/bash
# Last GPU call of round 2 (9 GPU-minutes were left): the whole GPU suite on the final tree, then smoke, then -- if the
# budget still allows -- the default bench line.  Every step writes its own file so that a cut-off call keeps what ran.
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
P=gpurun_out/r02c
timeout 420 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --durations=12 > ${P}_gpu_tests.log 2>&1
tail -25 ${P}_gpu_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -3 ${P}_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > ${P}_bench_cfg3.json 2> ${P}_bench_cfg3.err
tail -c 600 ${P}_bench_cfg3.json
