#!/bin/bash
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for cfg in "greedy 4 --no-overlap" "greedy 8 --no-overlap" "greedy 4 " "greedy 8 " "beam 4 --no-overlap" "beam 2 --no-overlap" "beam 4 "; do
set -- $cfg
timeout 300 python bench.py --steps 16 --warmup 3 --cpu-sample 0 --mode $1 --decode-merge $2 $3 2>> gpurun_out/bench_err.log | python -c "$show" "$1 merge=$2 $3"
done
tail -5 gpurun_out/bench_err.log
