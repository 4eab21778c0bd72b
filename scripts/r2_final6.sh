#!/bin/bash
# The last GPU seconds of round 2: the GPU suite minus its three slowest tests (which do not go through
# assert_equals_model / check_batch) after `restarts` joined the counters compared with the CPU model.
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
timeout 100 python -m pytest tests -m gpu -q --timeout 60 --timeout-method=thread -k "not feasibility_boundary and not bound_pruned and not cfg5_khosla" > gpurun_out/r02f_gpu_tests_subset.log 2>&1
tail -8 gpurun_out/r02f_gpu_tests_subset.log
