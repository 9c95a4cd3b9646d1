#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line:
instructions executed and stall samples, top-N lines, and totals per named line range."""
import csv, sys, collections
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
cur_file = None
with open(path, newline='') as f:
    for r in csv.reader(f):
        if not r: continue
        if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
        if r[0] in ('Function Name', 'Line No', 'Kernel Name'): 
            if r[0] == 'Line No': hdr = r
            continue
        if r[0] == '' : continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        def num(x):
            try: return float(x)
            except ValueError: return 0.0
        rows.append((cur_file, line, r[1].strip(), num(r[4]), num(r[7])))
tot_s = sum(r[3] for r in rows); tot_i = sum(r[4] for r in rows)
print('total samples %d, total warp-instructions %d' % (tot_s, tot_i))
print('--- top lines by stall samples')
for r in sorted(rows, key=lambda r: -r[3])[:topn]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100*r[3]/tot_s, 100*r[4]/tot_i, r[0], r[1], r[2][:110]))
print('--- top lines by instructions')
for r in sorted(rows, key=lambda r: -r[4])[:topn]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100*r[3]/tot_s, 100*r[4]/tot_i, r[0], r[1], r[2][:110]))
if len(sys.argv) > 3:
    # phase table: file:lo-hi=name,...
    phases = []
    for spec in sys.argv[3].split(','):
        rng, name = spec.split('=')
        fn, lh = rng.split(':'); lo, hi = lh.split('-')
        phases.append((fn, int(lo), int(hi), name))
    agg = collections.OrderedDict()
    for r in rows:
        name = 'other:' + r[0]
        for fn, lo, hi, nm in phases:
            if r[0].startswith(fn) and lo <= r[1] <= hi: name = nm; break
        a = agg.setdefault(name, [0.0, 0.0]); a[0] += r[3]; a[1] += r[4]
    print('--- phases')
    for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%5.1f%% smp %5.1f%% inst  %s' % (100*s/tot_s, 100*i/tot_i, k))
