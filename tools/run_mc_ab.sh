#!/bin/bash
# A/B of the 4-CTA-cluster multicast backward (two-sweep formulation, CLIPNCE_NO_BWD2=1) at the headline shape
for mc in 0 1; do
  CLIPNCE_NO_BWD2=1 CLIPNCE_BWD_MC=$mc timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline \
    > gpurun_out/mc_ab_$mc.json 2> gpurun_out/mc_ab_$mc.err
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/mc_ab_$mc.json"))
    print("MC=$mc ms/step", round(r["ms_per_step"], 3), "bwd ms", round(r["roofline"]["ms_per_launch"], 3), "parity", r["parity"]["ok"], r["parity"]["dA_rel"], r["parity"]["dB_rel"])
except Exception as e:
    print("MC=$mc failed", e); print(open("gpurun_out/mc_ab_$mc.err").read()[-1500:])
PY
done
