set -x
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q --timeout 120 2>&1 | tail -30 ) > gpurun_out/r2c_mesh_tests.log
cat gpurun_out/r2c_mesh_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err
timeout 200 python scripts/r2_probe.py > gpurun_out/r2c_probe.json 2> gpurun_out/r2c_probe.err
timeout 200 python scripts/profile_cfg3.py 0 > gpurun_out/r2c_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel' --launch-skip 10 -c 2 -o gpurun_out/r2c_ncu_gather -f python scripts/profile_cfg3.py 0 > gpurun_out/r2c_ncu.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_partitioned.py 2>&1 | tail -30 ) > gpurun_out/r2c_tests.log
cat gpurun_out/r2c_tests.log
cat gpurun_out/r2c_probe.json
