"""Reads a scripts/timeline.py CSV and prints, for the last complete step: the kernels in start order (stream, start, duration),
time per kernel name, the union of busy time, and the gaps where no kernel runs."""
import csv
import re
import sys
from collections import defaultdict


def short(n):
    n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '').replace('pcfd::', '').replace('ws::', '')
    return n[:60]


def main(path, verbose=True):
    rows = [(float(r['start_us']), float(r['dur_us']), r['stream'], short(r['name'])) for r in csv.DictReader(open(path))]
    # steps start with the input copy (gather_blocks_multi_kernel)
    starts = [i for i, r in enumerate(rows) if r[3].startswith('gather_blocks_multi')]
    if len(starts) < 2:
        starts = [0, len(rows)]
    a, b = starts[-2], starts[-1]
    step = rows[a:b]
    t0 = step[0][0]
    end = max(s + d for s, d, _, _ in step)
    print(f'step: {len(step)} kernels, {end - t0:.1f} us from first start to last end; next step starts at {rows[b][0] - t0:.1f} us' if b < len(rows) else '')
    if verbose:
        for s, d, st, n in step:
            print(f'{s - t0:9.1f} {d:8.1f}  s{st:>3}  {n}')
    by = defaultdict(lambda: [0, 0.0])
    for s, d, st, n in step:
        by[n][0] += 1
        by[n][1] += d
    print('\nper kernel:')
    tot = 0.0
    for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f'{d:9.1f} us  x{c:<3} {n}')
        tot += d
    print(f'{tot:9.1f} us  summed kernel time')
    iv = sorted((s, s + d) for s, d, _, _ in step)
    busy, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((cur_e - t0, s - cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f'busy union {busy:.1f} us, idle inside the step {sum(g for _, g in gaps):.1f} us in {len(gaps)} gaps; largest:')
    for at, g in sorted(gaps, key=lambda x: -x[1])[:8]:
        print(f'   {g:7.1f} us at {at:9.1f}')


if __name__ == '__main__':
    main(sys.argv[1], verbose='-q' not in sys.argv)
