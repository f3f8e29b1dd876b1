#!/usr/bin/env python
"""Turn ncu output into the tracked summaries under profiles/.

  python tools/ncu_report.py launches <launches.csv> <out.md> "<command line that produced it>"
      launches.csv = `ncu --metrics gpu__time_duration.sum --clock-control none ... --csv --log-file launches.csv`
  python tools/ncu_report.py full <report.ncu-rep> <out.md> "<command>" [--traffic profiles/traffic.json]
      report.ncu-rep = `ncu --set full --clock-control none --import-source on ... -o report`
      (read here with `ncu -i report.ncu-rep --page raw --csv`)
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [('gpu__time_duration.sum', 'duration'),
        ('dram__bytes_read.sum', 'dram read'),
        ('dram__bytes_write.sum', 'dram write'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('smsp__issue_active.avg.pct', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor pipe % (inst)'),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (cycles)'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (cycles)'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('launch__registers_per_thread', 'registers / thread'),
        ('launch__grid_size', 'grid'),
        ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
        ('launch__occupancy_limit_registers', 'occupancy limit (regs)'),
        ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem)'),
        ('launch__waves_per_multiprocessor', 'waves / SM'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long scoreboard'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short scoreboard'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg throttle'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio throttle'),
        ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'stall membar')]


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.strip()[:70]


def launches(path, out, cmd):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = {}
    for r in rows[1:]:
        if r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
            continue
        v = float(r[iv].replace(',', ''))
        v = v / 1e3 if r[iu] in ('ns', 'nsecond') else v * 1e3 if r[iu] in ('ms', 'msecond') else v
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu launch list (cold-cache, serialised: compare SHARES, not absolutes)\n\n')
        f.write('command: `%s`\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n' % cmd)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (k, a[0], a[1], 100 * a[1] / tot))
        f.write('| **all** | %d | %.1f | 100%% |\n' % (n, tot))
    print('wrote', out)


def full(rep, out, cmd, traffic=None):
    if rep.endswith('.csv'):          # already exported on the GPU box: `ncu -i rep --page raw --csv > file.csv`
        txt = open(rep, errors='replace').read()
    else:
        txt = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode(errors='replace')
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tj = {}
    with open(out, 'w') as f:
        f.write('# ncu --set full summary\n\ncommand: `%s`\n(one cold-cache launch per section, under the profiler: '
                'the bench numbers come from CUDA events)\n' % cmd)
        for r in data:
            name = short(r[ix['Kernel Name']])
            f.write('\n## `%s`  (grid %s)\n\n| metric | value |\n|---|---|\n' % (name, r[ix['launch__grid_size']] if 'launch__grid_size' in ix else '?'))
            seen = set()
            for k, label in KEYS:
                if k in ix and r[ix[k]] != '' and label not in seen:
                    seen.add(label)
                    f.write('| %s | %s %s |\n' % (label, r[ix[k]], units[ix[k]]))
            if 'dram__bytes_read.sum' in ix:
                def to_bytes(v, u):
                    v = float(v.replace(',', ''))
                    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
                base = name + ' grid ' + r[ix['launch__grid_size']]
                k = 1
                while '%s #%d' % (base, k) in tj:
                    k += 1
                tj.setdefault('%s #%d' % (base, k), {
                    'dram_bytes_read': to_bytes(r[ix['dram__bytes_read.sum']], units[ix['dram__bytes_read.sum']]),
                    'dram_bytes_write': to_bytes(r[ix['dram__bytes_write.sum']], units[ix['dram__bytes_write.sum']]),
                    'duration_us': r[ix['gpu__time_duration.sum']], 'source': out})
    if traffic:
        json.dump(tj, open(traffic, 'w'), indent=1)
    print('wrote', out)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        tr = sys.argv[sys.argv.index('--traffic') + 1] if '--traffic' in sys.argv else None
        full(sys.argv[2], sys.argv[3], sys.argv[4], tr)
