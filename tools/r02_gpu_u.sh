#!/bin/bash
# round 2, GPU call U: ncu --set full of the R1 tick after the rare-path restructuring (source-level), 2^18 games
mkdir -p gpurun_out
python tools/profile_targets_r1.py 262144 8 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/prof_r1t_r02 python tools/profile_targets_r1.py 262144 8 2 > gpurun_out/u_ncu_r1.log 2>&1
tail -n 2 gpurun_out/u_ncu_r1.log
