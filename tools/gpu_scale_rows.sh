#!/bin/bash
# Strong scaling of ONE 31x512x512 scene split into row bands (bench.py --shard rows) at 1..N GPUs of this box, eager and
# CUDA-graph replay, each run under its own timeout (a hung collective must not hold the box).
# Usage: gpu_scale_rows.sh NMAX ["peer nccl"] ["1 2 4 8"]
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
NMAX=${1:-2}
COMMS=${2:-"peer nccl"}
NLIST=${3:-"1 2 4 8"}
for n in $NLIST; do
  [ $n -gt $NMAX ] && break
  for c in $COMMS; do
  [ $n -eq 1 ] && [ $c != peer ] && continue
  for g in off on; do
    tag=r02_rows_n${n}_${c}_graph_${g}
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --shard rows --steps 20 --warmup 4 --no-cpu-baseline --cuda-graph $g > $O/$tag.json 2> $O/$tag.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus $n --shard rows --comm $c --steps 20 --warmup 4 --no-cpu-baseline --cuda-graph $g > $O/$tag.json 2> $O/$tag.err
    fi
    echo "$tag rc=$? $(python - <<PY
import json
try:
    j = json.loads([l for l in open('$O/$tag.json') if l.startswith('{')][-1])
    print(f"{j['ms_per_step']:.3f} ms/scene  {j['value']:.1f} cubes/s  e2e {j['e2e']['ms_per_step']:.3f} ms  parity {j['config']['output_check']['max_abs_err_over_max_abs_ref']:.2e}")
except Exception as e:
    print('no line:', e)
PY
)"
    tail -2 $O/$tag.err | cut -c1-300
  done
  done
done
