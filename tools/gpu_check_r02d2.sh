#!/bin/bash
# memcheck over the degradation kernels after padding the blur kernel's weight array (compute-sanitizer flagged the 12 bytes
# an LDS.128 read past kw[20]); then the metrics tests once more without the sanitizer.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
MPHSIR_SANITIZE=1 timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 --launch-timeout 0 \
  python -m pytest tests/test_metrics_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02d_memcheck_metrics.txt 2>&1
grep -E "ERROR SUMMARY|passed|failed" $O/r02d_memcheck_metrics.txt | tail -3
timeout 120 python -m pytest tests/test_metrics_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02d_metrics_tests.txt 2>&1
tail -2 $O/r02d_metrics_tests.txt
