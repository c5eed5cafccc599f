#!/bin/bash
# End-of-round evidence (run under gpurun, ONE GPU): launch lists of one encode + decode steps in both tensor-core modes,
# `ncu --set full` of the kernels this round changed, every profiled command first run plainly (exit 0 required).
set -u
O=gpurun_out
P="python tools/profile_path.py"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 300 "$@" > /dev/null 2>&1 || { echo "$name plain run failed"; return; }
  timeout 600 $NCU -k regex:$rx -s $skip -c $cnt -o $O/$name "$@" > $O/ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/r02f_ncu_$name.txt 2>&1
  python tools/ncu_hot_lines.py $O/$name.ncu-rep >> $O/r02f_ncu_$name.txt 2>&1
  rm -f $O/$name.ncu-rep $O/ncu_$name.log; }
for prec in bf16x3 bf16; do
  CMD="$P --batch 256 --steps 6 --warm 0 --mode greedy --precision $prec"
  $CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02f_launches_${prec}_greedy_B256.csv $CMD > /dev/null 2>&1
  echo "launches $prec rc=$?"
  python tools/summarize_launches.py $O/r02f_launches_${prec}_greedy_B256.csv > $O/r02f_launches_${prec}_greedy_B256.summary.txt 2>&1
done
CMD="$P --batch 256 --images 1024 --steps 3 --warm 0 --mode beam"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02f_launches_bf16x3_beam5_5120rows.csv $CMD > /dev/null 2>&1
python tools/summarize_launches.py $O/r02f_launches_bf16x3_beam5_5120rows.csv > $O/r02f_launches_bf16x3_beam5_5120rows.summary.txt 2>&1
cap conv0_direct "conv0_direct" 0 1 $P --batch 256 --steps 2 --warm 0 --mode greedy
cap encoder_attention "encoder_attention" 0 1 $P --batch 256 --steps 2 --warm 0 --mode greedy
cap wide_decode_tc3 "conv_gemm_tc3" 12 3 $P --batch 256 --images 1024 --steps 3 --warm 0 --mode beam
cap attn_greedy2560 decode_attention 400 2 $P --batch 256 --images 2560 --steps 104 --warm 0 --mode greedy
cap lstm_attention "lstm_attention_step" 60 1 $P --head Attnv2 --batch 512 --steps 70 --warm 0 --mode greedy
ls -la $O/r02f_* | awk '{print $5, $9}'
