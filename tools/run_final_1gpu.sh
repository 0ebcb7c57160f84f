#!/bin/bash
# Final 1-GPU evidence of the round: bench lines (default, s = 100), config 2, tri-modal, ncu launch list + full capture.
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/bench_final.err || tail -5 gpurun_out/bench_final.err
python bench.py --scale 100 --mix 0.1 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_scale100_final.json 2>> gpurun_out/bench_final.err
python bench.py --n 32768 --d 768 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_n32768_d768_final.json 2>> gpurun_out/bench_final.err
python tools/bench_config2.py > gpurun_out/r2_bench_config2_n4096_final.json 2>> gpurun_out/bench_final.err
for n in 1024 4096; do python tools/bench_trimodal.py --n $n 2>> gpurun_out/bench_final.err | tail -1 >> gpurun_out/r2_bench_trimodal_final.jsonl; done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_n65536_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph --scale 100 --mix 0.1"
ncu --set full --clock-control none --import-source on -k regex:"bwd2_kernel|fwd_kernel" -s 4 -c 2 -f -o gpurun_out/r2_full_scale100 $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_full_scale100.ncu-rep --page raw --csv > gpurun_out/r2_full_scale100_raw.csv 2> /dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/*raw.csv
python - <<'PY'
import json
for f in ("r2_bench_1gpu_final", "r2_bench_1gpu_scale100_final", "r2_bench_1gpu_n32768_d768_final"):
    try:
        r = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(r["ms_per_step"], 3), round(r["value"] / 1e6, 3), "e2e", round(r["e2e"]["value"] / 1e6, 3), r["parity"]["ok"], r["clocks"], round(r["roofline"]["ms_per_launch"], 3), round(r["roofline"]["fwd_ms_per_launch"], 3), round(r["roofline"]["step_frac_of_burst"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
