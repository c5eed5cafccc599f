#!/bin/bash
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for prec in bf16x3 bf16; do
D2T_TC4=1 timeout 200 python bench.py --steps 8 --warmup 3 --cpu-sample 0 --sequential --precision $prec 2>> gpurun_out/bench_err.log | python -c "$show" "$prec sequential tc4=1"
D2T_TC4=1 timeout 200 python bench.py --steps 8 --warmup 3 --cpu-sample 0 --precision $prec 2>> gpurun_out/bench_err.log | python -c "$show" "$prec pipelined tc4=1"
done
tail -5 gpurun_out/bench_err.log
