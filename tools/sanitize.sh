#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests at small shapes (the repo carries hand-rolled mbarrier
# protocols, TMA stores, cp.async rings and cross-GPU `red`s).  memcheck = out-of-bounds / misaligned global + shared
# accesses; racecheck = shared-memory hazards; synccheck = barrier misuse.  Run on a GPU box:
#     gpurun --timeout 2400 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_<tool>.log and a one-line summary per tool (copied into profiles/r02_sanitize.md).
# The tcgen05 / TMA kernels run under the sanitizer as they are (small shapes: the bounded mbarrier waits of tc_ptx.cuh trap
# after ~2 s, far beyond a sanitized small-shape launch).
mkdir -p gpurun_out
TESTS="tests/test_kernels_gpu.py tests/test_tower_gpu.py"
SEL="gather or fm_standalone or crossnet or cin or autoint_attention or bilinear or sigmoid_bce or one_kernel_forward_equals or fused_head or tower_mlp"
for tool in memcheck racecheck synccheck; do
  ( timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 66 \
      python -m pytest $TESTS -m gpu -x -q -p no:cacheprovider -k "$SEL" ) > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  errs=$(grep -c "^========= .*\(Invalid\|Race\|hazard\|Barrier error\|Misaligned\|out of bounds\)" gpurun_out/sanitize_$tool.log)
  echo "$tool: exit $rc, $(grep -E '[0-9]+ passed|[0-9]+ failed' gpurun_out/sanitize_$tool.log | tail -1), error records $errs, $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
