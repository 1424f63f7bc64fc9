#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun): launch list + DRAM traffic of one cube512 step, full captures of
# the hot kernels at the level-1 C=128 shapes, the metric / degradation kernels.  Text summaries land in gpurun_out/.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r02_traffic.csv python tools/profile_step.py infer > $O/r02_traffic.log 2>&1
python tools/ncu_traffic.py $O/r02_traffic.csv > $O/r02_traffic_cube512.json
python tools/ncu_summary.py launches $O/r02_traffic.csv > $O/r02_launches_cube512_fp32.txt
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_tc_kernel|mlp_tc_kernel|window_attn_mma|dwgram_tma|spectral_finish|local_gate" -s 150 -c 36 \
    -o $O/r02_hot python tools/profile_step.py infer > $O/r02_hot.log 2>&1
python tools/ncu_summary.py report $O/r02_hot.ncu-rep > $O/r02_hot_kernels_full.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"psnr_ssim|degrade" \
    -o $O/r02_metrics python tools/profile_step.py metrics > $O/r02_metrics.log 2>&1
python tools/ncu_summary.py report $O/r02_metrics.ncu-rep > $O/r02_metrics_kernels_full.txt
ls -la $O/*.ncu-rep
# keep the raw reports only if they fit the 64 MiB return channel
for f in $O/r02_hot.ncu-rep $O/r02_metrics.ncu-rep; do
  if [ $(stat -c %s $f) -gt 30000000 ]; then rm -f $f; fi
done
tail -5 $O/r02_hot_kernels_full.txt
