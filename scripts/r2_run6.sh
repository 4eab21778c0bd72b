set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
( timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread -k "pruned or mesh or narrow_scan or engine_options or seeded or cfg3 or learn" 2>&1 | tail -8 ) > gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_cfg3.json 2> gpurun_out/r2q_bench_cfg3.err
timeout 200 python scripts/r2_probe.py quick > gpurun_out/r2q_probe.json 2>&1; cat gpurun_out/r2q_probe.json
timeout 200 python scripts/profile_cfg3.py 0 > gpurun_out/r2q_plain_p0.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel' --launch-skip 12 -c 2 -o gpurun_out/r2q_ncu_full_gather_pruned -f python scripts/profile_cfg3.py 0 > gpurun_out/r2q_ncu_full0.log 2>&1
tail -c 300 gpurun_out/r2q_bench_cfg3.json
