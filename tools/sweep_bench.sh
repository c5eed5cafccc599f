#!/bin/bash
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for cfg in "greedy 4 116" "greedy 4 124" "greedy 4 140" "greedy 8 124" "greedy 8 132" "greedy 8 140" "beam 4 124" "beam 4 140" "beam 8 132"; do
set -- $cfg
timeout 300 python bench.py --steps 16 --warmup 3 --cpu-sample 0 --mode $1 --decode-merge $2 --encoder-sms $3 2>> gpurun_out/bench_err.log | python -c "$show" "$1 merge=$2 sms=$3"
done
