#!/usr/bin/env python
"""Hot source lines of one kernel from an .ncu-rep captured with --import-source on:
    python tools/ncu_source_hot.py REP KERNEL_ID_FILTER [top]
Aggregates the warp-stall samples of `ncu --page source --print-source cuda,sass --csv` per CUDA source line."""
import csv
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-id', kid],
                         capture_output=True, text=True).stdout
    fname, hdr, agg, total = None, None, {}, 0
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == 'File Path':
            fname = row[1].split('/')[-1]
            continue
        if row[0] == 'Line No':
            hdr = row
            continue
        if hdr is None or len(row) != len(hdr) or row[2] != '-':      # per-line summary rows have Address '-'
            continue
        d = dict(zip(hdr[4:], row[4:]))
        n = int(d.get('# Samples', '0') or 0)
        if n == 0:
            continue
        total += n
        stalls = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
        agg[(fname, int(row[0]))] = (n, row[1].strip()[:90], stalls, int(d.get('Instructions Executed', '0') or 0))
    print('total samples', total)
    for (f, ln), (n, src, stalls, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ' '.join('%s=%d' % (k[6:], v) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print('%5.1f%%  %s:%d  inst=%d  [%s]  %s' % (100.0 * n / max(total, 1), f, ln, inst, st, src))


if __name__ == '__main__':
    main()
