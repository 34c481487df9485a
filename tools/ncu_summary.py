#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (first profiled launch): python tools/ncu_summary.py rep [more keys]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_fmaheavy.avg.pct', 'sm__inst_executed_pipe_fmalite.avg.pct',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct', 'lts__t_bytes.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct',
        '_per_issue_active.ratio', 'sm__cycles_elapsed.avg ', 'smsp__cycles_active.avg ', 'sm__pipe_tensor']
def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== launch:', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, r):
            if any(k in h + ' ' for k in keys):
                print(f'{h} [{u}] = {v}')
if __name__ == '__main__':
    main()
