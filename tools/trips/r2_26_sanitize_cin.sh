#!/bin/bash
# round 2: compute-sanitizer (memcheck, synccheck, racecheck) over the CIN kernel tests (tensor-core path included)
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  ( timeout 900 compute-sanitizer --tool $tool --print-limit 12 --error-exitcode 66 \
      python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "cin" ) > gpurun_out/r2_26_sanitize_$tool.log 2>&1
  rc=$?
  errs=$(grep -c "^========= .*\(Invalid\|Race\|hazard\|Barrier error\|Misaligned\|out of bounds\)" gpurun_out/r2_26_sanitize_$tool.log)
  echo "$tool: exit $rc, $(grep -E '[0-9]+ passed|[0-9]+ failed' gpurun_out/r2_26_sanitize_$tool.log | tail -1), error records $errs, $(grep 'ERROR SUMMARY' gpurun_out/r2_26_sanitize_$tool.log | tail -1)"
  grep "^========= .*\(Invalid\|hazard\|Barrier error\|Misaligned\)" -A3 gpurun_out/r2_26_sanitize_$tool.log | head -12 | cut -c1-220
done
