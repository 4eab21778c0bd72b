#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 200 -k "narrow or invalid_large or cfg3 or sign_handling or device_generator or failed_solve" > gpurun_out/r2p_tests.log 2>&1
tail -3 gpurun_out/r2p_tests.log
timeout 200 python scripts/r2_upload_ride_ab.py > gpurun_out/r2p_upload_ride_ab.jsonl 2> gpurun_out/r2p_upload_ride_ab.err
cat gpurun_out/r2p_upload_ride_ab.jsonl; tail -3 gpurun_out/r2p_upload_ride_ab.err
