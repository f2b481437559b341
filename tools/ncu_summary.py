"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text table for profiles/."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'sm__cycles_elapsed.max.per_second', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
def main(rep, out=None):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append('== kernel: %s  (id %s)' % (d.get('Kernel Name', '?'), d.get('ID', '?')))
        for i, h in enumerate(hdr):
            if h in KEYS or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
                try:
                    v = float(r[i])
                    if 'issue_stalled' in h and v < 0.05: continue
                    lines.append('  %-90s %14.6g %s' % (h, v, units[i]))
                except ValueError:
                    lines.append('  %-90s %14s %s' % (h, r[i], units[i]))
    s = '\n'.join(lines) + '\n'
    if out: open(out, 'w').write(s)
    else: print(s)
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
