#!/bin/bash
# Round-2 (second half) ncu evidence, run on the GPU box through gpurun: launch list + DRAM traffic of one cube512 step
# after the CTA-pair fused MLP, and full captures of the new MLP kernel + the pair-mode GEMM engine at the level-1 shapes.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r02b_traffic.csv python tools/profile_step.py infer > $O/r02b_traffic.log 2>&1
python tools/ncu_traffic.py $O/r02b_traffic.csv > $O/r02b_traffic_cube512.json
python tools/ncu_summary.py launches $O/r02b_traffic.csv > $O/r02b_launches_cube512_fp32.txt
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"mlp_tc_kernel|gemm_tc_kernel" -s 60 -c 24 \
    -o $O/r02b_hot python tools/profile_step.py infer > $O/r02b_hot.log 2>&1
python tools/ncu_summary.py report $O/r02b_hot.ncu-rep > $O/r02b_hot_kernels_full.txt
ls -la $O/*.ncu-rep
for f in $O/r02b_hot.ncu-rep; do
  if [ $(stat -c %s $f) -gt 30000000 ]; then rm -f $f; fi
done
tail -5 $O/r02b_hot_kernels_full.txt
