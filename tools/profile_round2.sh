#!/bin/bash
# ncu evidence of round 2 (run on a B200 through gpurun; outputs under gpurun_out/, summaries copied to profiles/)
set -u
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bev --no-cold-cli"
$B > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
G="python tools/greedy_perf.py P28K 4800"
$G "" > gpurun_out/r02_plain_greedy.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:greedy_warp -s 2 -c 1 -o gpurun_out/r02_greedy_warp $G "" > gpurun_out/r02_ncu_greedy.log 2>&1
ADSEL_GREEDY_WARP=0 $G "" > gpurun_out/r02_plain_greedy_fast.log 2>&1 || exit 1
ADSEL_GREEDY_WARP=0 ncu --set full --clock-control none --import-source on -k regex:greedy_fast -s 2 -c 1 -o gpurun_out/r02_greedy_fast $G "" > gpurun_out/r02_ncu_greedy_fast.log 2>&1
tail -1 gpurun_out/r02_plain_greedy.log | cut -c1-120
tail -1 gpurun_out/r02_plain_greedy_fast.log | cut -c1-120
