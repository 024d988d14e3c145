#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu -i X --page raw --csv) into the few numbers DESIGN.md cites."""
import csv, subprocess, sys
KEEP = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.sum',
 'lts__t_sectors_srcunit_tex_op_red.sum','lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum',
 'lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum','lts__t_requests_srcunit_tex_op_red.sum',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','smsp__inst_executed_op_shared_atom.sum','smsp__inst_executed_op_global_red.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum',
 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','launch__grid_size','launch__block_size',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__cycles_active.avg','sm__cycles_elapsed.max']
def main(rep, idx=0):
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2+idx]
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('kernel:', name)
    for i,h in enumerate(hdr):
        if h in KEEP or ('warp_issue_stalled' in h and h.endswith('per_warp_active.pct')):
            try:
                if 'stalled' in h and float(vals[i]) < 2: continue
            except ValueError: pass
            print(f"{h:84s} {units[i]:16s} {vals[i]}")
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else 0)
