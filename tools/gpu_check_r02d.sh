#!/bin/bash
# round-2 closing pass: the metrics / degradation tests (new: mphsir_sr_degrade), memcheck over the same tests, and the default
# bench line of the final build.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 120 python -m pytest tests/test_metrics_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02d_metrics_tests.txt 2>&1
tail -3 $O/r02d_metrics_tests.txt
MPHSIR_SANITIZE=1 timeout 150 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 --launch-timeout 0 \
  python -m pytest tests/test_metrics_gpu.py -m gpu -q -x -p no:cacheprovider -k "sr_degrade or blur or reference_default" > $O/r02d_memcheck_metrics.txt 2>&1
grep -E "ERROR SUMMARY|passed|failed" $O/r02d_memcheck_metrics.txt | tail -3
timeout 200 python bench.py > $O/r02d_default_n1.json 2> $O/r02d_default_n1.err
python - <<P
import json
d=json.loads(open("gpurun_out/r02d_default_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["train"]["value"], d["train"]["e2e"]["value"], d["config"]["output_check"]["max_abs_err_over_max_abs_ref"])
P
