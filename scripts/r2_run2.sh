set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2b_mesh_tests.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err
timeout 300 python scripts/r2_probe.py > gpurun_out/r2b_probe.json 2> gpurun_out/r2b_probe.err
timeout 300 python scripts/profile_cfg3.py 0 > gpurun_out/r2b_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel' --launch-skip 10 -c 2 -o gpurun_out/r2b_ncu_gather -f python scripts/profile_cfg3.py 0 > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_mesh_tests.log gpurun_out/r2b_tests.log
cat gpurun_out/r2b_probe.json
