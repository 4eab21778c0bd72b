set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
( timeout 400 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q --timeout 100 --timeout-method=thread -k mesh 2>&1 | tail -30 ) > gpurun_out/r2l_mesh_tests.log
cat gpurun_out/r2l_mesh_tests.log
