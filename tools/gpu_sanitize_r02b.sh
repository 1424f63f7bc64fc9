#!/bin/bash
# compute-sanitizer pass over the kernels changed in the second half of round 2: the CTA-pair fused MLP (cta_group::2, relaxed
# cluster arrives, A operands in tensor memory), the pair-mode GEMM engine and the one-launch local gate.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
export MPHSIR_SANITIZE=1
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, seconds, pytest args...
  local name=$1 tool=$2 secs=$3; shift 3
  echo "=== $tool: pytest $*" > $O/r02b_${name}.txt
  timeout $secs $SAN --tool $tool --print-limit 30 --launch-timeout 0 python -m pytest "$@" -q -x -p no:cacheprovider >> $O/r02b_${name}.txt 2>&1
  echo "exit code $? (124 = stopped by the $secs s limit)" >> $O/r02b_${name}.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|exit code" $O/r02b_${name}.txt | tail -6
}
SEL="fused_mlp or local_gate or gemm_tc_bias or conv"
run memcheck memcheck 500 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -k "$SEL"
run synccheck synccheck 300 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -k "$SEL"
run racecheck racecheck 500 tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -k "fused_mlp or local_gate"
