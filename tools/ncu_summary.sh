#!/bin/bash
# usage: tools/ncu_summary.sh <report.ncu-rep> <out_prefix>   (run in the build container, no GPU needed)
set -e
REP=$1; OUT=$2
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','launch__waves_per_multiprocessor','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__cycles_active.avg','smsp__cycles_active.max','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_not_selected']
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:70])
    for k in keys:
        if k in hdr: print('  %-66s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
" > ${OUT}_metrics.txt
ncu -i $REP --page source --csv 2>/dev/null > /tmp/_src.csv
mkdir -p /tmp/_cub && cd /tmp/_cub && rm -f *.cubin && cuobjdump -xelf all "/root/repo/bridges-with-reinforcement-learning_b200/libbridges_b200.so" >/dev/null 2>&1
nvdisasm -g -c bw_step.sm_100a.cubin > step.sass 2>/dev/null
python /root/repo/tools/ncu_lines.py /tmp/_src.csv /tmp/_cub/step.sass > ${OUT}_lines.txt
