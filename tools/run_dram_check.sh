#!/bin/bash
# DRAM traffic of the two-sided backward with and without the ring's L2 evict_last policy (ncu, dram metrics only).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
for H in 1 0; do
  CLIPNCE_BWD2_L2HINT=$H $CMD > gpurun_out/plain_$H.log 2>&1 || { echo "plain failed"; exit 1; }
  CLIPNCE_BWD2_L2HINT=$H ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:bwd2_kernel -s 3 -c 1 --csv --log-file gpurun_out/r2_dram_hint$H.csv $CMD > /dev/null 2>&1
  echo "hint=$H"; grep bwd2 gpurun_out/r2_dram_hint$H.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_sided" 2>&1 | tail -2
bash tools/run_bwd2_sweep.sh CLIPNCE_BWD2_L2HINT=1 CLIPNCE_BWD2_L2HINT=0 CLIPNCE_BWD2_L2HINT=1 CLIPNCE_BWD2_L2HINT=0
