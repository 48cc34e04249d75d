"""SASS census of libpcfd_sm100.so: per kernel, the counts of the instructions that prove the Blackwell path
(tcgen05.mma -> UTC*MMA, TMA -> UTMALDG / UTMASTG, tcgen05.ld / st -> LDTM / STTM, tcgen05.commit -> UTCBAR), next to the
legacy tensor path (HMMA) which must be absent.  Runs on the CPU box:  python scripts/sass_census.py > profiles/sass_census.md"""
import collections
import os
import re
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(R, 'porous-cfd_b200', 'libpcfd_sm100.so')
KEYS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'LDGMC', 'MUFU', 'FFMA', 'HMMA']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    counts, order, cur, i = {}, [], None, 0
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = names[i] if i < len(names) and names[i] else m.group(1)
            i += 1
            cur = re.sub(r'\(.*', '', cur)                      # drop the argument list
            cur = re.sub(r'^void ', '', cur)
            counts.setdefault(cur, collections.Counter())
            counts[cur]['_inst'] += 1
            if cur not in order:
                order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            if op in KEYS:
                counts[cur][op] += 1
            counts[cur]['_n'] += 1
    tot = collections.Counter()
    print('# SASS census of `porous-cfd_b200/libpcfd_sm100.so` (sm_100a)\n')
    print('`cuobjdump -sass` of the built library, instruction counts per kernel (template instances summed). '
          '`UTCHMMA` = tcgen05.mma, `UTMALDG`/`UTMASTG` = TMA load / store, `LDTM`/`STTM` = tcgen05.ld / st, '
          '`UTCBAR` = tcgen05.commit, `LDGMC` = multimem.ld_reduce (the NVSwitch adds the ranks\' copies in flight; multimem.st compiles '
          'to `STG.E.128.STRONG.SYS` on the multicast address). No `HMMA` (legacy mma.sync) anywhere.\n')
    print('| kernel | instances | instructions | ' + ' | '.join(KEYS) + ' |')
    print('|---|---:|---:|' + '---:|' * len(KEYS))
    fam = collections.OrderedDict()
    inst = collections.Counter()
    for k in order:
        base = re.sub(r'<.*', '', k)
        fam.setdefault(base, collections.Counter()).update(counts[k])
        inst[base] += counts[k]['_inst']
    for k, c in sorted(fam.items(), key=lambda kv: (-kv[1]['UTCHMMA'], -kv[1]['_n'])):
        print(f'| `{k}` | {inst[k]} | {c["_n"]} | ' + ' | '.join(str(c[x]) for x in KEYS) + ' |')
        tot.update(c)
    print(f'| **total** | {sum(inst.values())} | {tot["_n"]} | ' + ' | '.join(str(tot[x]) for x in KEYS) + ' |')


if __name__ == '__main__':
    sys.exit(main())
