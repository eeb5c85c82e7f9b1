#!/bin/bash
# first GPU trip: SIMT-only parity, then the tcgen05 tests in their own process, then micro-benchmarks
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
export RPB_GEMM_IMPL=1
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider \
  -k "not tcgen05 and not DCN and not xDeepFM and not AutoInt and not FiBiNet and not dcn and not xdeepfm and not autoint and not fibinet" \
  > gpurun_out/t1_simt.log 2>&1
echo "simt tests exit $?" | tee -a gpurun_out/t1_simt.log
unset RPB_GEMM_IMPL
timeout 600 python -m pytest tests/test_zz_linear_tc_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/t1_tc.log 2>&1
echo "tc tests exit $?" | tee -a gpurun_out/t1_tc.log
timeout 900 python tools/microbench.py > gpurun_out/t1_micro.log 2>&1
echo "microbench exit $?" | tee -a gpurun_out/t1_micro.log
tail -5 gpurun_out/t1_simt.log; tail -5 gpurun_out/t1_tc.log; tail -40 gpurun_out/t1_micro.log
