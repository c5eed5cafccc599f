#!/bin/bash
# Schedule sweep after the CTA-pair convolution (bench.py main record only) + decode duration vs rows per call.
O=gpurun_out
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
while read -r mode merge sms prec extra; do
  [ -z "$mode" ] && continue
  steps=$((merge * 3)); [ $steps -lt 12 ] && steps=12
  timeout 300 python bench.py --steps $steps --warmup 4 --cpu-sample 0 --records none --mode $mode --precision $prec --decode-merge $merge --encoder-sms $sms $extra 2>> $O/bench_err.log | python -c "$show" "$mode $prec merge=$merge sms=$sms $extra" | tee -a $O/r02c_schedule_sweep2.txt
done <<'LIST'
greedy 4 132 bf16x3
greedy 8 132 bf16x3
greedy 8 148 bf16x3 --no-overlap
greedy 12 148 bf16x3 --no-overlap
greedy 16 148 bf16x3 --no-overlap
beam 4 132 bf16x3
beam 4 148 bf16x3 --no-overlap
beam 8 148 bf16x3 --no-overlap
greedy 4 132 bf16
greedy 8 148 bf16 --no-overlap
greedy 16 148 bf16 --no-overlap
beam 4 132 bf16
beam 4 148 bf16 --no-overlap
beam 8 148 bf16 --no-overlap
LIST
python tools/pipeline_timing.py rows 2>&1 | tee $O/r02c_decode_vs_rows.txt
