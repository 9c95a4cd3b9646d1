#!/usr/bin/env python
"""Executed warp-instructions per SASS opcode from `ncu -i X.ncu-rep --page source --csv --print-source sass`
(one kernel instance); second argument: environments in the launch (per-environment column)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n_env = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[start]
src, ie = hdr.index('Source'), hdr.index('Instructions Executed')
cnt = collections.Counter()
for r in rows[start + 1:]:
    if len(r) <= ie:
        continue
    try:
        n = float(r[ie])
    except ValueError:
        continue
    op = r[src].split()
    if not op:
        continue
    o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
    keep2 = ('MUFU', 'F2I', 'I2F', 'F2F', 'UTC', 'UBLKCP', 'SYNCS', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'TCGEN', 'UTMA')
    o = '.'.join(o.split('.')[:2]) if o.startswith(keep2) else o.split('.')[0]
    cnt[o] += n
tot = sum(cnt.values())
print('total warp-instructions %d (%.0f per environment)' % (tot, tot / n_env))
for o, n in cnt.most_common(60):
    print('%-16s %6.2f%%  %10.0f  %8.1f/env' % (o, 100 * n / tot, n, n / n_env))
