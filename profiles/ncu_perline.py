#!/usr/bin/env python
"""Per-source-line executed warp-instructions per environment-step from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:KERNEL` (one kernel instance)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
n_env = float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = None; lines = collections.OrderedDict(); tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No', 'Kernel Name') or r[0] == '': continue
    try: ln = int(r[0]); n = float(r[7]); smp = float(r[4])
    except ValueError: continue
    lines[(cur, ln)] = (n, smp, r[1].strip()); tot += n
print('total warp-instructions %d, per environment-step %.0f' % (tot, tot / n_env))
for (f, ln), (n, smp, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('%5.1f%% %6.0f/env  %s:%d  %s' % (100 * n / tot, n / n_env, f, ln, src[:100]))
