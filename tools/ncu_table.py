#!/usr/bin/env python
"""Markdown table of the kernels in one or more .ncu-rep files (ncu --set full): time, DRAM bytes, pipes, occupancy."""
import csv
import subprocess
import sys

COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3,
         'msecond': 1e3}

print('| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 hit % | tensor pipe % | fma pipe % | issue active % | '
      'warps active % | regs | grid | dyn smem KB |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for path in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(d, k):
        i = hdr.index(k)
        v = float(d[i].replace(',', '')) if d[i] not in ('', 'n/a') else 0.0
        return v * SCALE.get(units[i].split('/')[0], 1.0)
    for d in data:
        name = d[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('rpb::', '')
        t, rd, wr = val(d, COLS[0]), val(d, COLS[1]), val(d, COLS[2])
        print(f"| `{name}` | {t:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / t * 1e3:.0f} | {val(d, COLS[3]):.0f} | {val(d, COLS[4]):.1f} | "
              f"{val(d, COLS[5]):.1f} | {val(d, COLS[6]):.1f} | {val(d, COLS[7]):.1f} | {int(val(d, COLS[8]))} | {int(val(d, COLS[9]))} | "
              f"{val(d, COLS[10]) * 1e3:.1f} |")
