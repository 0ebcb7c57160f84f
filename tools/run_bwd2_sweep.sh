#!/bin/bash
# A/B of the two-sided backward on one B200: producer split sweep + the two-sweep path on the same box.
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout 300 "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity 2>>gpurun_out/bwd2_sweep.err | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    r = j['roofline']
    print('ms_per_step %.3f  eager %.3f  bwd_ms %.3f fwd_ms %.3f  step_frac_burst %.3f  launches %d' % (j['ms_per_step'], j['eager_ms_per_step'], r['ms_per_launch'], r['fwd_ms_per_launch'], r['step_frac_of_burst'], j['gpu_launches']))
"; }
for cfg in "$@"; do run "$cfg" env $cfg; done
