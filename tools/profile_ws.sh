#!/bin/bash
# Profile capture of the two headline kernels (run on the GPU box from the repo root; outputs under gpurun_out/):
#   1. a plain run that must exit 0 without ncu,
#   2. the launch list (gpu__time_duration per launch) -> profiles/r01_ncu_launches_ws.csv,
#   3. one ncu --set full capture of c3_ws_kernel / warp_u8_ws_kernel (launches 7-8) with source correlation;
#      read it back with `ncu -i gpurun_out/ws_full.ncu-rep --page raw --csv` (summary: profiles/r01_ncu_ws_full_summary.txt)
#      and `--page source --csv` (instruction counts / stall samples per SASS line).
set -x
B="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"c3_ws_kernel|warp_u8_ws_kernel" --launch-skip 6 -c 2 -f -o gpurun_out/ws_full $B > gpurun_out/ncu_full.log 2>&1
