#!/bin/bash
# What is left of the round's GPU budget (3.5 minutes): the other bench lines on the final tree.
set -x
mkdir -p gpurun_out
P=gpurun_out/r02e
timeout 40 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_arm_cfg3.json 2> ${P}_ref.err
timeout 40 python bench.py --workload cfg2 --steps 10 --warmup 3 > ${P}_bench_cfg2.json 2> ${P}_cfg2.err
timeout 40 python bench.py --workload cfg4 --steps 5 --warmup 3 > ${P}_bench_cfg4.json 2> ${P}_cfg4.err
timeout 70 python bench.py --workload cfg5 --steps 5 --warmup 3 > ${P}_bench_cfg5.json 2> ${P}_cfg5.err
for f in ${P}_bench_*.json; do echo $f; head -c 300 $f; echo; done
