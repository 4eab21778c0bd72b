set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
timeout 300 python scripts/variant_run.py sparse_linear_assignment_b200/libsla_b200.so variants/key4.so variants/key3.so variants/key6.so > gpurun_out/r2a_variants.jsonl 2> gpurun_out/r2a_variants.err
timeout 300 python scripts/r2_probe.py > gpurun_out/r2a_probe.json 2> gpurun_out/r2a_probe.err
cat gpurun_out/r2a_tests.log
tail -c 1500 gpurun_out/r2a_bench_cfg3.json
cat gpurun_out/r2a_variants.jsonl gpurun_out/r2a_probe.json
