#!/bin/bash
# Round-2 (third pass) evidence after the CTA-pair convolution: launch lists, ncu --set full of the pair kernel in both
# tensor-core modes, and a schedule sweep.  Run under gpurun on ONE GPU.
set -u
O=gpurun_out
P="python tools/profile_path.py"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 300 "$@" > /dev/null 2>&1 || { echo "$name plain run failed"; return; }
  timeout 600 $NCU -k regex:$rx -s $skip -c $cnt -o $O/$name "$@" > $O/ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/r02c_ncu_$name.txt 2>&1
  python tools/ncu_hot_lines.py $O/$name.ncu-rep >> $O/r02c_ncu_$name.txt 2>&1
  rm -f $O/$name.ncu-rep $O/ncu_$name.log; }
for prec in bf16x3 bf16; do
  CMD="$P --batch 256 --steps 6 --warm 0 --mode greedy --precision $prec"
  $CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02c_launches_${prec}_greedy_B256.csv $CMD > /dev/null 2>&1
  echo "launches $prec rc=$?"
  python tools/summarize_launches.py $O/r02c_launches_${prec}_greedy_B256.csv > $O/r02c_launches_${prec}_greedy_B256.summary.txt 2>&1
done
cap conv_tc5_pair_bf16 conv_gemm_tc5 8 2 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16
cap conv_tc5_pair_bf16x3 conv_gemm_tc5 8 2 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16x3
show='import json,sys
d=json.load(sys.stdin)
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "enc", round(d["roofline"]["encode_ms"],1), "dec", round(d["roofline"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for cfg in "greedy 4 132 --" "greedy 4 140 --" "greedy 4 148 --" "greedy 4 148 --no-overlap" "greedy 8 148 --no-overlap" "greedy 8 140 --" "beam 4 132 --" "beam 4 148 --no-overlap" "beam 2 148 --no-overlap" "beam 2 132 --"; do
  set -- $cfg
  timeout 300 python bench.py --steps 16 --warmup 4 --cpu-sample 0 --records none --mode $1 --decode-merge $2 --encoder-sms $3 $4 2>> $O/bench_err.log | python -c "$show" "$1 merge=$2 sms=$3 $4" | tee -a $O/r02c_schedule_sweep.txt
done
