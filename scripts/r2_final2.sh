# Round-2 closing evidence run on ONE B200 (under gpurun): the whole GPU suite, smoke, the default bench line, the
# reference arm, cfg1 and the reference's criterion groups on the final tree.
set -x
mkdir -p gpurun_out
export SLA_MESH_TIMEOUT_S=10
P=gpurun_out/r02b
( timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread 2>&1 | tail -15 ) > ${P}_gpu_tests.log
cat ${P}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; cat ${P}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > ${P}_bench_cfg3.json 2> ${P}_bench_cfg3.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > ${P}_bench_reference_arm_cfg3.json 2>/dev/null
timeout 200 python bench.py --workload cfg1 --steps 200 --warmup 20 > ${P}_bench_cfg1.json 2> ${P}_bench_cfg1.err
timeout 300 python benchmarks/reference_harness.py > ${P}_reference_harness.md 2> ${P}_reference_harness.err
tail -c 400 ${P}_bench_cfg3.json; tail -c 300 ${P}_bench_cfg1.json
