#!/usr/bin/env python3
"""Summarise ncu output into profiles/ (tracked): launch lists (CSV from --metrics gpu__time_duration.sum)
and key raw metrics of .ncu-rep captures.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/ncu_summary.py rep gpurun_out/prof_feat.ncu-rep profiles/r01_feature_kernel.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'smsp__cycles_active.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row['Kernel Name'], []).append(float(row['Metric Value'].replace(',', '')))
    def is_ours(k):
        n = k.replace('void ', '')
        return 'nvbx::' in k or n.startswith('k_')
    ours = {k: v for k, v in agg.items() if is_ours(k)}
    tot = sum(sum(v) for v in ours.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list: `{src}`\n\n')
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n')
        f.write('Shares are over this library\'s kernels only (torch input-generation kernels excluded).\n\n')
        f.write('| kernel | launches | mean us | min us | max us | share of step |\n|---|---:|---:|---:|---:|---:|\n')
        for k, v in ours.items():
            name = k.split('(')[0].replace('void ', '')
            f.write(f'| `{name}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {min(v) / 1e3:.2f} | {max(v) / 1e3:.2f} | '
                    f'{100 * sum(v) / tot:.1f}% |\n')
        other = {k: v for k, v in agg.items() if not is_ours(k)}
        if other:
            f.write('\nOther kernels in the capture (input generation by torch): ' +
                    ', '.join(f'`{k.split("(")[0][:60]}` x{len(v)}' for k, v in other.items()) + '\n')
    print(open(dst).read())


def rep(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full capture: `{src}`\n\n')
        for r in rows[2:]:
            f.write(f'## {r[hdr.index("Kernel Name")][:90]}\n\n| metric | value | unit |\n|---|---:|---|\n')
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f'| {k} | {r[i]} | {units[i]} |\n')
            f.write('\n')
    print(open(dst).read()[:3000])


if __name__ == '__main__':
    {'launches': launches, 'rep': rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
