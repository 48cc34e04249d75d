"""Reads an .ncu-rep here (no GPU): per kernel the headline counters and the hottest SASS lines by stall samples.
    python scripts/ncu_hot.py gpurun_out/x.ncu-rep [kernel-regex] [top-n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else '.'
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', f'regex:{rx}'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('====', r[idx['Kernel Name']][:90])
    for w in WANT:
        if w in idx:
            print(f'  {w:70s} {r[idx[w]]} {rows[1][idx[w]]}')
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.3:
                print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:30s} {v:.2f}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{rx}'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hs = [i for i, r in enumerate(rows) if r and r[0] == 'Address'] + [len(rows)]
seen = set()
for a, b in zip(hs[:-1], hs[1:]):
    name = rows[a - 1][1][:80] if a > 0 and len(rows[a - 1]) > 1 else '?'
    if name in seen:
        continue
    seen.add(name)
    h = rows[a]
    ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    data = [(int(r[isamp]), int(r[iex]), r[ia].strip()) for r in rows[a + 1:b] if len(r) > isamp and r[isamp].isdigit()]
    tot, totex = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f'==== {name}: {len(data)} SASS lines, {totex} warp instructions, {tot} samples')
    ops = collections.Counter()
    for s, e, t in data:
        ops[(t.split()[1] if t.startswith('@') else t.split()[0])] += e
    print('  opcodes:', ', '.join(f'{o} {c}' for o, c in ops.most_common(14)))
    for i, (s, e, t) in sorted(enumerate(data), key=lambda x: -x[1][0])[:topn]:
        print(f'  {i:5d} samples {s:6d} ({100.0 * s / max(tot, 1):4.1f}%) exec {e:9d}  {t[:90]}')
