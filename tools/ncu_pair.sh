set -u
O=gpurun_out
P="python tools/profile_path.py"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1 rx=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rx -s $skip -c 1 -o $O/$name "$@" > $O/ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/r02b_ncu_$name.txt 2>&1
  python tools/ncu_hot_lines.py $O/$name.ncu-rep >> $O/r02b_ncu_$name.txt 2>&1
  rm -f $O/$name.ncu-rep; }
timeout 120 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16 --opt pair=1 > /dev/null 2>&1 || exit 1
cap conv_tc5_pair_bf16 conv_gemm_tc5 6 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16 --opt pair=1
cap conv_tc3_tma_bf16 conv_gemm_tc3 22 $P --batch 256 --steps 2 --warm 0 --mode greedy --precision bf16
