#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r2_mid_timing.py "$@" >> gpurun_out/r2m_mid_timing.jsonl 2> gpurun_out/r2m_mid_timing.err
tail -4 gpurun_out/r2m_mid_timing.jsonl; tail -3 gpurun_out/r2m_mid_timing.err
