#!/bin/bash
# kernel launch list (ncu, durations only) of the tri-modal step, grouped vs three pair steps
for n in 4096 8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tri_launches_n$n.csv \
    python tools/bench_trimodal.py --n $n --steps 1 > gpurun_out/tri_ncu_n$n.log 2>&1
done
