#!/bin/bash
# End-of-round ncu evidence (run under gpurun, ONE GPU): a launch list of one encode + decode steps, and one
# `ncu --set full` capture per kernel family.  Every profiled command line is first run plainly (exit 0 required).
#   bash tools/ncu_evidence.sh        -> gpurun_out/r02_*.ncu-rep, gpurun_out/r02_launches_*.csv
set -u
O=gpurun_out
P="python tools/profile_path.py"
NCU="ncu --set full --clock-control none --import-source on -f"
run() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && $NCU -k regex:$rx -s $skip -c $cnt -o $O/r02_$name "$@" > $O/ncu_$name.log 2>&1
  echo "$name rc=$?"
  # gpurun copies at most 64 MiB back: keep the text summary (+ the hottest source lines), drop the report
  if [ -f $O/r02_$name.ncu-rep ]; then
    python tools/ncu_summary.py $O/r02_$name.ncu-rep > $O/r02_ncu_$name.txt 2>&1
    python tools/ncu_hot_lines.py $O/r02_$name.ncu-rep >> $O/r02_ncu_$name.txt 2>&1
    rm -f $O/r02_$name.ncu-rep
  fi
  rm -f $O/plain_$name.log $O/ncu_$name.log
}
# launch lists (device time per launch; cold cache, serialised: compare shares)
CMD="$P --batch 256 --steps 6 --warm 0 --mode greedy"
$CMD > $O/plain_launches_greedy.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_bf16x3_greedy_B256.csv $CMD > /dev/null 2>&1
echo "launches greedy rc=$?"
CMD="$P --batch 256 --steps 4 --warm 0 --mode beam"
$CMD > $O/plain_launches_beam.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_bf16x3_beam5_B256.csv $CMD > /dev/null 2>&1
echo "launches beam rc=$?"
# dominant kernel: layer3 3x3 convolutions (launches 20.. of conv_gemm_tc3 are layer3 blocks), both tensor-core modes
run conv_tc3_bf16x3 conv_gemm_tc3 22 2 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16x3
run conv_tc3_bf16 conv_gemm_tc3 22 2 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16
# decode attention on the shapes the schedule runs: 1024 merged greedy rows, 1280 / 5120 beam rows, step 100
run attn_greedy1024 decode_attention 800 2 $P --batch 256 --images 1024 --steps 104 --warm 0 --mode greedy
run attn_beam1280 decode_attention 800 2 $P --batch 256 --steps 104 --warm 0 --mode beam
run attn_beam5120 decode_attention 800 2 $P --batch 256 --images 1024 --steps 104 --warm 0 --mode beam
run beam_step beam_step_kernel 100 1 $P --batch 256 --steps 104 --warm 0 --mode beam
run greedy_pick greedy_pick_kernel 100 1 $P --batch 256 --steps 104 --warm 0 --mode greedy
run decode_gemm "conv_gemm_tc_kernel" 300 4 $P --batch 256 --steps 16 --warm 0 --mode greedy
run layernorm layernorm_kernel 40 2 $P --batch 256 --steps 8 --warm 0 --mode greedy
run encoder_misc "conv0_direct|maxpool2x2|encoder_attention|assemble_tokens" 0 6 $P --batch 256 --steps 2 --warm 0 --mode greedy
# LSTM head (config/train.yaml stack)
run lstm_greedy "lstm_attention_step|lstm_pointwise|lstm_pick" 60 3 $P --head Attnv2 --batch 256 --steps 40 --warm 0 --mode greedy
run lstm_beam "attn_beam_step" 30 1 $P --head Attnv2 --batch 256 --steps 40 --warm 0 --mode beam
# preprocessing
run prep "prep_" 0 4 $P --batch 8 --steps 2 --warm 0 --mode greedy --prep 64
ls -la $O/r02_* | awk '{print $5, $9}'
