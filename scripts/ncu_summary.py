"""Summarise an .ncu-rep (ncu --set full --import-source on) into a small text file for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt ["title"]"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
lines = ['# ' + title, '']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            lines.append('%-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    lines.append('')
    lines.append('stall reasons (warps per issue-active cycle):')
    st = [(float(d[h]), h) for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    for v, h in sorted(st, reverse=True)[:9]:
        lines.append('  %-28s %.3f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
    lines.append('')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Address' in r and 'Source' in r)
h2 = rows[hi]
isrc, iexec, isamp = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) <= iexec:
        continue
    try:
        n, s = int(r[iexec]), int(r[isamp])
    except ValueError:
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    ops[op] += n
    samp[op] += s
    tot += n
lines.append('SASS opcode mix (share of executed warp instructions / share of stall samples), %d static instructions:' % (len(rows) - hi - 1))
for op, n in ops.most_common(22):
    lines.append('  %-10s %6.2f%%  %5.1f%%' % (op, 100.0 * n / tot, 100.0 * samp[op] / max(1, sum(samp.values()))))
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))
