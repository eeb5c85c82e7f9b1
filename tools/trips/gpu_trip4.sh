#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "not FiBiNet and not fibinet" > gpurun_out/t4_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t4_tests.log
timeout 600 python tools/exp_gather.py > gpurun_out/t4_exp_gather.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:gather_fwd_kernel -c 12 --csv --log-file gpurun_out/t4_gather_variants.csv python tools/exp_gather.py > /dev/null 2>&1
tail -12 gpurun_out/t4_tests.log; cat gpurun_out/t4_exp_gather.log; grep gather_fwd gpurun_out/t4_gather_variants.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | head -60
