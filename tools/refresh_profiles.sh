#!/bin/bash
# Regenerates the measured evidence of a round on the GPU box (run from the repo root: outputs under gpurun_out/refresh/,
# copied into profiles/ by hand afterwards). Every profiled command first runs once without ncu.
set -x
O=gpurun_out/refresh
mkdir -p $O
python bench.py > $O/bench_final.json 2> $O/bench_final.err || exit 1
B="python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-modes12"
timeout 300 $B > $O/prof_plain.json 2> $O/prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv $B > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"c3_ws_kernel|warp_u8_ws_kernel" --launch-skip 6 -c 2 -f -o $O/ws_full $B > $O/ncu_full.log 2>&1
python tools/fwd_prof.py rot 16 3 > $O/fwd_plain.txt 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_raster_kernel" --launch-skip 1 -c 1 -f -o $O/fwd_raster python tools/fwd_prof.py rot 16 3 > $O/ncu_fwd.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/fwd_launches.csv python tools/fwd_prof.py rot 16 2 > /dev/null 2>&1
python tools/opbench.py 32 > $O/opbench_b32.txt 2>&1
for n in 64 16; do echo "# --- $n x 1080p"; python tools/fwdbench.py $n 2>&1 | grep -v '^{'; done > $O/fwdbench.txt
echo "# --- 37 x 436x1024 (cfg 3 size)" >> $O/fwdbench.txt
python tools/fwdbench.py 37 436 1024 2>&1 | grep -v '^{' >> $O/fwdbench.txt
python tools/rough_flows.py > $O/rough_tma.txt 2>&1
OFK_C3_WS=0 OFK_WARP_WS=0 python tools/rough_flows.py > $O/rough_gather.txt 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt
