#!/bin/bash
# Ad-hoc bench matrix (run on the GPU box).
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])'
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 0 2>> gpurun_out/bench_err.log | python -c "$show" "bf16x3 greedy pipelined run$i"
D2T_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 0 2>> gpurun_out/bench_err.log | python -c "$show" "bf16x3 greedy pipelined nopdl run$i"
done
tail -5 gpurun_out/bench_err.log
