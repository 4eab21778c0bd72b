# Round-2 evidence run on ONE B200 (under gpurun): tests, bench lines, ncu launch list and full captures.
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
P=gpurun_out/r02
( timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread 2>&1 | tail -15 ) > ${P}_gpu_tests.log
cat ${P}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; cat ${P}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > ${P}_bench_cfg3.json 2> ${P}_bench_cfg3.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > ${P}_bench_reference_arm_cfg3.json 2>/dev/null
for w in cfg1 cfg2 cfg4 cfg5; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > ${P}_bench_$w.json 2> ${P}_bench_$w.err; done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_plain_bench.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_bench.log 2>&1
timeout 200 python scripts/profile_cfg3.py 1 > ${P}_plain_p1.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel|assign_wide|first_assign' --launch-skip 26 -c 5 -o ${P}_ncu_full_round1_u16 -f python scripts/profile_cfg3.py 1 > ${P}_ncu_full1.log 2>&1
timeout 200 python scripts/profile_cfg3.py 0 > ${P}_plain_p0.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel' --launch-skip 12 -c 2 -o ${P}_ncu_full_gather_pruned -f python scripts/profile_cfg3.py 0 > ${P}_ncu_full0.log 2>&1
timeout 200 python scripts/profile_cfg3.py 1 0 narrow_scan=0 > ${P}_plain_p2.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bid_regular_kernel' --launch-skip 12 -c 1 -o ${P}_ncu_full_round1_f64 -f python scripts/profile_cfg3.py 1 0 narrow_scan=0 > ${P}_ncu_full2.log 2>&1
timeout 500 python benchmarks/reference_harness.py > ${P}_reference_harness.md 2> ${P}_reference_harness.err
tail -c 600 ${P}_bench_cfg3.json
