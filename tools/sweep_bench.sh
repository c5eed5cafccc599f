#!/bin/bash
# Sweep the encoder SM budget of the pipelined schedule (run on the GPU box).
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"])'
for sms in 88 72 56; do
  python bench.py --steps 10 --warmup 3 --cpu-sample 0 --mode beam --encoder-sms $sms 2>> gpurun_out/bench_err.log | python -c "$show" "beam sms=$sms"
done
tail -5 gpurun_out/bench_err.log
