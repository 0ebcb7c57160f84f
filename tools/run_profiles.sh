#!/bin/bash
# Profiling pass of round 2 (one B200): launch list of a bench step and a full ncu capture of the two contraction kernels.
set -o pipefail
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_n65536.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"bwd2_kernel|fwd_kernel" -s 4 -c 2 -f -o gpurun_out/r2_full_two_sided $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
