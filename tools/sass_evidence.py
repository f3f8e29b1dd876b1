#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove what a kernel uses (cuobjdump -sass on the built library):

    python tools/sass_evidence.py > profiles/r2_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'a-recsys_b200', 'lib', 'libarx_b200.so')
COLS = [('UTCHMMA', r'\bUTCHMMA'), ('UTMALDG', r'\bUTMALDG'), ('LDTM', r'\bLDTM'), ('UTCBAR', r'\bUTCBAR'),
        ('SYNCS', r'\bSYNCS'), ('LDG.E.128', r'\bLDG\.E\.128|\bLDG\.E\.LTC\S*\.128|\bLDG\.E\.[A-Z.]*128'),
        ('LDGSTS', r'\bLDGSTS'), ('ATOMG', r'\bATOMG'), ('RED', r'\bRED[G.]'), ('REDG.F32x4', r'REDG\.E\.ADD\.F32x4'),
        ('.SYS', r'\.SYS\b'), ('cluster (UCGABAR, STAS, MAPA)', r'\bUCGABAR|\bMAPA\b|\bSTAS\b|\bSTAS\.')]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(anonymous namespace\)::', '', cur)
            cur = re.sub(r'\(.*$', '', cur)
            order.append(cur)
            continue
        if cur is None:
            continue
        for name, pat in COLS:
            if re.search(pat, line):
                counts[cur][name] += 1
    print('# SASS evidence, round 2 (cuobjdump -sass a-recsys_b200/lib/libarx_b200.so, sm_100a)\n')
    print('UTCHMMA = tcgen05.mma, UTMALDG = TMA bulk tensor load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,')
    print('SYNCS = mbarrier ops, LDG.E.128 = 128-bit global loads, REDG.F32x4 = red.global.add.v4.f32 (split accumulation / NVLink')
    print('peer pushes), .SYS = system-scope accesses (peer memory), last column = cluster barrier + st.async into distributed shared memory (the LSTM hidden-state exchange).\n')
    print('| kernel | ' + ' | '.join(n for n, _ in COLS) + ' |')
    print('|---|' + '---:|' * len(COLS))
    keep = [k for k in order if any(counts[k].values())]
    for k in sorted(set(keep), key=lambda k: (-counts[k]['UTCHMMA'], -counts[k]['REDG.F32x4'], k)):
        print('| `%s` | ' % k[:70] + ' | '.join(str(counts[k][n]) for n, _ in COLS) + ' |')


if __name__ == '__main__':
    main()
