#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'wgrad_tf32x3|gemm_tf32x3_v2|gather_fwd_tile|rows_zero' --launch-skip 22 --launch-count 11 -o gpurun_out/t15_deepfm -f python tools/profile_all.py --model DeepFM --steps 3 > gpurun_out/t15_ncu.log 2>&1
tail -5 gpurun_out/t15_ncu.log
ls -la gpurun_out/
