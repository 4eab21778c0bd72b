set -x
mkdir -p gpurun_out
timeout 200 python scripts/profile_mesh_lockstep.py 8000000 32000000 16 2 > gpurun_out/r2i_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mesh_' --launch-skip 0 -c 30 -o gpurun_out/r2i_ncu_mesh -f python scripts/profile_mesh_lockstep.py 8000000 32000000 16 1 > gpurun_out/r2i_ncu.log 2>&1
cat gpurun_out/r2i_plain.log; tail -3 gpurun_out/r2i_ncu.log
