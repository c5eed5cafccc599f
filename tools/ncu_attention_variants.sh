set -x
for v in "perrow attn_staged=0 attn_image_block=0" "image attn_staged=0 attn_image_block=1" "staged attn_staged=1 attn_image_block=0"; do
  set -- $v
  name=$1; o1=$2; o2=$3
  CMD="python tools/profile_path.py --batch 256 --steps 104 --warm 0 --mode beam --opt $o1 --opt $o2"
  $CMD > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_attention -s 800 -c 2 -o gpurun_out/attn_$name -f $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
done
