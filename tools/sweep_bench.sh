#!/bin/bash
# Ad-hoc bench matrix (run on the GPU box).
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])'
for mode in greedy beam; do
  timeout 300 python bench.py --steps 8 --warmup 3 --cpu-sample 0 --mode $mode 2>> gpurun_out/bench_err.log | python -c "$show" "bf16x3 $mode pipelined"
done
D2T_FUSE_LN=0 timeout 300 python bench.py --steps 8 --warmup 3 --cpu-sample 0 --sequential 2>> gpurun_out/bench_err.log | python -c "$show" "bf16x3 greedy sequential nofuse"
timeout 300 python bench.py --steps 8 --warmup 3 --cpu-sample 0 --sequential 2>> gpurun_out/bench_err.log | python -c "$show" "bf16x3 greedy sequential"
tail -5 gpurun_out/bench_err.log
