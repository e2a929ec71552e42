"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_poll_exact_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.avg.per_cycle_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
]
STALLS = 'smsp__average_warps_issue_stalled_'


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ['# ncu --set full --clock-control none summary of %s' % rep]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        lines.append('')
        lines.append('kernel: %s' % name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append('  %-72s %s %s' % (k, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    stalls.append((float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append('  warp stall reasons (avg warps stalled per issue-active cycle):')
        for v, h in stalls[:8]:
            lines.append('    %-40s %.3f' % (h, v))
    with open(out, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
