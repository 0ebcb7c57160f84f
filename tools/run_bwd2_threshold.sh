#!/bin/bash
# two-sided backward against the two-sweep backward between 8192 and 16384 rows (where to switch)
for n in 8192 12288; do
  for hook in 0 8192; do
    if [ $hook = 0 ]; then env="CLIPNCE_NO_BWD2=1"; else env="CLIPNCE_BWD2_MIN_N=$hook"; fi
    env $env timeout 200 python bench.py --n $n --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/thr_${n}_$hook.json 2> gpurun_out/thr_${n}_$hook.err
    python - <<PY
import json
try:
    r = json.load(open("gpurun_out/thr_${n}_$hook.json"))
    print("n=$n", "two-sided" if $hook else "two sweeps", "ms/step", round(r["ms_per_step"], 4), "bwd", round(r["roofline"]["ms_per_launch"], 4), "parity", r["parity"]["ok"], r["roofline"]["kernel"][:24])
except Exception as e:
    print("n=$n hook=$hook failed", e); print(open("gpurun_out/thr_${n}_$hook.err").read()[-800:])
PY
  done
done
