#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of metrics profiles/ keeps under version control.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/x.txt
"""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'sm__cycles_active.avg']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('# kernel: %s   (source: %s)' % (name, rep))
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or any(h == k or h.startswith(k + '.') and h.count('.') <= k.count('.') + 1 for k in ()):
                print('%-78s %-18s %s' % (h, u, v))
            if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
                try:
                    stalls.append((float(v), h[len('smsp__pcsamp_warps_issue_stalled_'):]))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        print('warp-state samples (pc sampling): ' + ', '.join('%s %.1f%%' % (n, 100 * s / tot) for s, n in sorted(stalls, reverse=True)[:8]))
        print()


if __name__ == '__main__':
    main()
