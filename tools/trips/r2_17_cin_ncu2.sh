#!/bin/bash
# round 2: ncu --set full of the CIN kernels after the second forward formulation / staged wgrad operands
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'cin_' --launch-skip 12 --launch-count 12 -f -o gpurun_out/r2_17_cin \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_17_ncu.log 2>&1
ls -la gpurun_out/r2_17_cin.ncu-rep
python tools/ncu_table.py gpurun_out/r2_17_cin.ncu-rep 2>&1 | tail -16
