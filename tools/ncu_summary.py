"""Summarise an .ncu-rep (raw page CSV) into the few numbers we track.
usage: python tools/ncu_summary.py report.ncu-rep [more metric substrings]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sectors_op_write.sum',
        'lts__t_sectors_op_read.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'smsp__average_warp', 'smsp__average_warps_issue_stalled',
        'dram__cycles_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'sm__inst_executed_pipe']


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:90], 'grid', d.get('Grid Size'),
              'block', d.get('Block Size'))
        for h, u in zip(hdr, units):
            if any(k in h for k in keys):
                v = d[h]
                try:
                    if float(v) == 0 and 'stalled' in h:
                        continue
                except ValueError:
                    pass
                print('  %-88s %-12s %s' % (h, u, v))


if __name__ == '__main__':
    main()
