"""Executed warp instructions and stall samples per SOURCE LINE of an .ncu-rep (compiled with -lineinfo,
captured with --import-source on).  Usage: python scripts/ncu_lines.py rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fpath = None
agg = {}
tot_i = tot_s = 0
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fpath = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ii, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= ii or not r[0].isdigit():
        continue
    try:
        n, s = int(r[ii]), int(r[isamp])
    except ValueError:
        continue
    key = (fpath, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
    tot_i += n
    tot_s += s
print('total warp instructions %d, samples %d' % (tot_i, tot_s))
by_file = {}
for (f, ln, src), (n, s) in agg.items():
    b = by_file.setdefault(f, [0, 0])
    b[0] += n
    b[1] += s
for f, (n, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print('%-14s inst %5.1f%%  stall samples %5.1f%%' % (f, 100. * n / tot_i, 100. * s / max(tot_s, 1)))
print()
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.2f%% %5.2f%%  %s:%d  %s' % (100. * n / tot_i, 100. * s / max(tot_s, 1), f, ln, src))
