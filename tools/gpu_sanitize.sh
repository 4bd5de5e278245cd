#!/bin/bash
# GPU box: compute-sanitizer over the kernel tests that exercise every pipeline (mbarrier rings, TMEM double buffer,
# lock-free threshold table, survivor lists, converter warps).  Sizes are the tests' own (small); logs under gpurun_out/.
mkdir -p gpurun_out
SEL='golden or ragged or escalation or tie_probe or streaming'
for TOOL in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 9 \
      python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" -p no:cacheprovider --timeout=240 > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?" | tee -a gpurun_out/sanitize_$TOOL.log
  grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|hazard" gpurun_out/sanitize_$TOOL.log | tail -6
done
