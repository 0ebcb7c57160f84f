#!/bin/bash
# one pair at the reference's batch sizes: both backward sides as one grouped sweep (default) against two pair sweeps
for n in 256 1024 4096; do
  for ng in 0 1; do
    CLIPNCE_NO_GROUP=$ng timeout 200 python bench.py --n $n --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/small_${n}_$ng.json 2> gpurun_out/small_${n}_$ng.err
    python - <<PY
import json
try:
    r = json.load(open("gpurun_out/small_${n}_$ng.json"))
    print("n=$n NO_GROUP=$ng ms/step", round(r["ms_per_step"], 4), "eager", round(r["eager_ms_per_step"], 3), "launches", r["gpu_launches"], "parity", r["parity"]["ok"] if r.get("parity") else None)
except Exception as e:
    print("n=$n NO_GROUP=$ng failed", e); print(open("gpurun_out/small_${n}_$ng.err").read()[-1200:])
PY
  done
done
