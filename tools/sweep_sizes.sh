#!/bin/bash
# BASELINE config 5: HybridViT beam-5 throughput over image sizes and batch sizes on one GPU (bench.py defaults otherwise).
show='import json,sys
d=json.load(sys.stdin)
r=d["roofline"] or {}
print(sys.argv[1], "| formulas/s", round(d["value"]), "| e2e", round(d["e2e"]["value"]), "| ms/batch", round(d["ms_per_step"],1), "| enc ms", round(r.get("encode_ms",0),1), "| dec ms", round(r.get("decode_ms",0),1), "| conv frac", round(r.get("frac",0),3), "|", d["clocks"]["reasons"])'
prec=${1:-bf16x3}
for cfg in "64 256 32" "64 256 64" "64 256 128" "64 256 256" "64 256 512" "64 256 1024" "96 384 128" "128 512 64" "160 704 32" "192 896 32"; do
set -- $cfg
steps=8; if [ $3 -ge 512 ]; then steps=4; fi
timeout 400 python bench.py --steps $steps --warmup 3 --cpu-sample 0 --mode beam --precision $prec --height $1 --width $2 --batch $3 2>> gpurun_out/sweep_err.log | python -c "$show" "$prec ${1}x${2} B=$3"
done
tail -3 gpurun_out/sweep_err.log
