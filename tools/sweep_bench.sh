#!/bin/bash
# usage: tools/sweep_bench.sh  (on the GPU box)
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for cfg in "greedy bf16x3" "beam bf16x3" "greedy bf16" "beam bf16"; do
set -- $cfg
timeout 300 python bench.py --steps 6 --warmup 3 --cpu-sample 0 --mode $1 --precision $2 --sequential 2>> gpurun_out/bench_err.log | python -c "$show" "$1 $2 sequential"
done
tail -5 gpurun_out/bench_err.log
