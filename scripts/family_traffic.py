"""profiles/traffic.json: measured DRAM bytes per C-ABI call of every jet family, from an ncu launch list taken with
   PCFD_NO_OVERLAP=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv ... bench.py --no-graph
(one stream, so the kernels of a call are consecutive: dW kernel, optional column-sum kernel, finish kernel).
  python scripts/family_traffic.py gpurun_out/traffic_r1.csv profiles/traffic.json"""
import collections
import csv
import json
import re
import sys


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = row['ID']
        d = per.setdefault(key, {'name': row['Kernel Name'], 'bytes': 0.0})
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        unit = row['Metric Unit']
        if row['Metric Name'].startswith('dram__bytes'):
            d['bytes'] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    fam = collections.defaultdict(lambda: {'calls': 0, 'bytes': 0.0})
    last = None
    for d in per.values():
        n = d['name']
        m = re.search(r'(ws_fwd|ws_dx|ws_dw|jet_fwd|jet_dx|jet_dw|thin_n_fwd|thin_k_fwd|thin_n_dx|small_rows_dw|small_rows)_?kernel<(\d+)', n)
        m1 = re.search(r'(ws_fwd1|ws_dw1)_kernel<', n)      # value-only kernels with the operand in tensor memory
        if m1:
            last = 'jet_fwd_cj1' if m1.group(1) == 'ws_fwd1' else 'jet_dw_cj1'
            fam[last]['calls'] += 1
            fam[last]['bytes'] += d['bytes']
        elif m:
            kind, cj = m.group(1), int(m.group(2))
            if kind == 'small_rows':
                cj, p = 1, ('dx' if '<1>' in n or '<true>' in n else 'fwd')
            elif kind == 'small_rows_dw':
                cj, p = 1, 'dw'
            else:
                p = 'fwd' if 'fwd' in kind else ('dx' if 'dx' in kind else 'dw')
            last = f'jet_{p}_cj{cj}'
            fam[last]['calls'] += 1
            fam[last]['bytes'] += d['bytes']
        elif 'small_rows_dw_kernel' in n:
            last = 'jet_dw_cj1'
            fam[last]['calls'] += 1
            fam[last]['bytes'] += d['bytes']
        elif ('colsum_kernel' in n or 'dw_finish_kernel' in n) and last is not None and '_dw_' in last:
            fam[last]['bytes'] += d['bytes']
        else:
            last = None if not ('colsum' in n or 'dw_finish' in n) else last
    out = {k: v['bytes'] / v['calls'] for k, v in fam.items() if v['calls']}
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps({k: round(v / 1e6, 2) for k, v in out.items()}), '(MB per call)')


if __name__ == '__main__':
    main(*sys.argv[1:3])
