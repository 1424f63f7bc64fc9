#!/bin/bash
# Round-2 closing ncu launch list + DRAM traffic of one cube512 step on the final build.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r02d_traffic.csv python tools/profile_step.py infer > $O/r02d_traffic.log 2>&1
python tools/ncu_traffic.py $O/r02d_traffic.csv > $O/r02d_traffic_cube512.json
python tools/ncu_summary.py launches $O/r02d_traffic.csv > $O/r02d_launches_cube512_fp32.txt
head -12 $O/r02d_launches_cube512_fp32.txt
