#!/bin/bash
# The remaining 4 GPU-minutes of round 2: ncu launch list of the default bench command on the final tree (the plain
# command first, as the recipe asks), then the cfg1 line.
set -x
mkdir -p gpurun_out
P=gpurun_out/r02d
timeout 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_bench_plain.json 2> ${P}_bench_plain.err || exit 1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_bench.log 2>&1
wc -l ${P}_launches_bench_cfg3.csv
timeout 60 python bench.py --workload cfg1 --steps 200 --warmup 20 > ${P}_bench_cfg1.json 2> ${P}_bench_cfg1.err
tail -c 300 ${P}_bench_cfg1.json
