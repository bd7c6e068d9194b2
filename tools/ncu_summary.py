#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv            > profiles/..._launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof_X.ncu-rep [regex]    > profiles/..._kernel.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '')[:70]
        v = float(d['Metric Value'].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)')
    print('# source:', path, ' launches:', len(data))
    print('%-72s %5s %12s %7s %10s' % ('kernel', 'n', 'total_us', 'share', 'avg_us'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %5d %12.1f %7.3f %10.1f' % (k, v[0], v[1] / 1e3, v[1] / tot, v[1] / v[0] / 1e3))


def kernel(path, pattern=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full --clock-control none; source:', path)
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if pattern and not re.search(pattern, d['Kernel Name']):
            continue
        print('kernel:', d['Kernel Name'], ' grid', d.get('Grid Size'), ' block', d.get('Block Size'))
        for m in METRICS:
            if m in d:
                print('   %-78s %14s %s' % (m, d[m], units[hdr.index(m)]))
        for m in hdr:
            if 'stall' in m and m.endswith('.pct') and m not in METRICS:
                try:
                    if float(d[m]) >= 5.0:
                        print('   %-78s %14s %s' % (m, d[m], '%'))
                except ValueError:
                    pass


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
