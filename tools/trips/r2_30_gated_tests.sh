#!/bin/bash
# round 2: the two test groups that were behind RPB_EXPERIMENTAL=1 (tcgen05 tower-tail backward, fused core on local shards), un-gated
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_30_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_30_tests.log
grep -E "passed|failed|FAILED|Error|assert|exit" gpurun_out/r2_30_tests.log | tail -12 | cut -c1-300
