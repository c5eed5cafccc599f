#!/bin/bash
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
timeout 300 python bench.py --steps 6 --warmup 3 --cpu-sample 0 --head Attnv2 --batch 512 --sequential 2>> gpurun_out/bench_err.log | python -c "$show" "attnv2 B=512 sequential"
timeout 300 python bench.py --steps 6 --warmup 3 --cpu-sample 0 --head Attnv2 --batch 512 2>> gpurun_out/bench_err.log | python -c "$show" "attnv2 B=512 pipelined"
tail -5 gpurun_out/bench_err.log
