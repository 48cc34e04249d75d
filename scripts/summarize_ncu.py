"""Summaries of ncu outputs brought back in gpurun_out/ -> profiles/ (tracked).
  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_ffma.md "title"
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_prof.md "title"
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'smsp__issue_active.avg.pct_of_peak_sustained_active']


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith('==')]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name', 'gpu__time_duration.sum') != 'gpu__time_duration.sum':
            continue     # captures that also carry dram__bytes_* (scripts/family_traffic.py reads those)
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        if v != v:      # a launch cut off by the end of the capture
            continue
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('pcfd::', '')
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    with open(dst, 'w') as f:
        f.write(f'# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list '
                f'(cold-cache, serialised: compare SHARES, not absolutes).\n\n')
        f.write(f'{sum(v[0] for v in tot.values())} launches, {total / 1e3:.2f} ms summed kernel time\n\n')
        f.write('| kernel | launches | total us | share |\n|---|---:|---:|---:|\n')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {v[1] / total:.3f} |\n')


def full(src, dst, title):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({src}).\n\n')
        for row in rows[2:]:
            f.write(f'## {row[hdr.index("Kernel Name")]}\n\n| metric | value | unit |\n|---|---:|---|\n')
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f'| {m} | {row[i]} | {units[i]} |\n')
            f.write('\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:5])
