set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/f_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_cfg3.json 2> gpurun_out/f_bench_cfg3.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_ref_cfg3.json 2>/dev/null
for w in cfg1 cfg2 cfg4 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/f_bench_$w.json 2>gpurun_out/f_bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel|assign_wide' --launch-skip 20 -c 2 -o gpurun_out/f_ncu_full_round1 -f python scripts/profile_cfg3.py 1 > gpurun_out/f_ncu_full.log 2>&1
python benchmarks/reference_harness.py > gpurun_out/f_reference_harness.md 2>gpurun_out/f_reference_harness.err
cat gpurun_out/f_tests.log
