#!/bin/bash
# compute-sanitizer evidence for the warp-specialised kernels (19 warps, five mbarrier pipelines, TMA landing zones
# converted in place): memcheck over the operator-level GPU tests at their small shapes, racecheck (shared-memory hazards)
# and synccheck over the tensor-core operators.  Run on the GPU box through gpurun; summaries land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
export MPHSIR_SANITIZE=1
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, seconds, pytest args...
  local name=$1 tool=$2 secs=$3; shift 3
  echo "=== $tool: pytest $*" > $O/r02_${name}.txt
  timeout $secs $SAN --tool $tool --print-limit 30 --launch-timeout 0 python -m pytest "$@" -q -x -p no:cacheprovider >> $O/r02_${name}.txt 2>&1
  echo "exit code $? (124 = stopped by the $secs s limit)" >> $O/r02_${name}.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|exit code" $O/r02_${name}.txt | tail -6
}
run memcheck_ops memcheck 600 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py tests/test_metrics_gpu.py -k "not scene_shape"
run memcheck_sharded memcheck 300 tests/test_sharded_gpu.py -k "virtual_ranks_match_single_gpu and fp32 and not 4-"
run memcheck_train memcheck 600 tests/test_train_ops_gpu.py
run racecheck_tc racecheck 600 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -k "gemm or mlp or dwgram or conv or window"
run synccheck_tc synccheck 300 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -k "gemm or mlp or dwgram or conv or window"
