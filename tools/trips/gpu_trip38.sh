#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t38_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t38_tests.log
tail -25 gpurun_out/t38_tests.log | cut -c1-400
