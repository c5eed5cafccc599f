#!/bin/bash
# Scaling run on one box: N = 1, 2, 4, 8 ranks back to back (weak scaling, 256 images per rank).
N=${1:-8}
for n in 1 2 4 8; do
  if [ $n -gt $N ]; then break; fi
  for mode in greedy beam; do
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 8 --warmup 3 --cpu-sample 0 --mode $mode > gpurun_out/scale_${mode}_n$n.json 2>> gpurun_out/scale_err.log
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --cpu-sample 0 --mode $mode > gpurun_out/scale_${mode}_n$n.json 2>> gpurun_out/scale_err.log
    fi
    python - <<PY
import json
d=json.load(open("gpurun_out/scale_${mode}_n$n.json"))
print("$mode n=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1))
PY
  done
done
tail -3 gpurun_out/scale_err.log
