#!/bin/bash
# round 2: coalesced first-round person pass -- parity tests of the new tree, then the variants (head = before the change)
mkdir -p gpurun_out
set -x
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu --timeout 300 --timeout-method=thread \
  -k "not cfg5" > gpurun_out/r2x_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2x_tests.log
tail -5 gpurun_out/r2x_tests.log
timeout 400 python scripts/r2_variant_assign.py variants/head.so variants/per256.so variants/per512.so variants/per1024.so variants/per2048.so variants/u8per1024.so \
  > gpurun_out/r2x_variants.jsonl 2> gpurun_out/r2x_variants.err
cat gpurun_out/r2x_variants.jsonl
tail -3 gpurun_out/r2x_variants.err
