#!/usr/bin/env python
"""Which steps are slow?  Times every step of a C2-style run with its own event pair and prints, next to each
step's duration, how many environments left the fast move path (bail-outs) and how many episodes ended in it."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from paintrl_b200.batched_env import BatchedPaintEnv
from paintrl_b200.config import EnvConfig

ap = argparse.ArgumentParser()
ap.add_argument('--envs', type=int, default=4096)
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=300)
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
cfg = EnvConfig(w['extra'], auto_reset=True, seed=1234, **w['kw'])
dev = torch.device('cuda:0')
env = BatchedPaintEnv(args.envs, cfg, device=dev, texture_size=w.get('texture', (240, 240)))
gen = torch.Generator(device=dev); gen.manual_seed(1234)
acts = (torch.randint(0, cfg.discrete_granularity, (args.steps, args.envs), generator=gen, device=dev, dtype=torch.int64) if cfg.action_mode == "discrete" else torch.rand((args.steps, args.envs, cfg.action_dim), generator=gen, device=dev, dtype=torch.float64) * 2 - 1)
env.reset(torch.randint(0, env.n_starts, (args.envs,), generator=gen, device=dev, dtype=torch.int32))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
rows = []
prev = env.stats()
for i in range(args.steps):
    flush.fill_(i & 255)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(acts[i]); e1.record()
    torch.cuda.synchronize()
    st = env.stats()
    rows.append((e0.elapsed_time(e1) * 1e3, st['move_bailouts'] - prev['move_bailouts'], st['episodes_ended'] - prev['episodes_ended'],
                 st['ray_full_scans'] - prev['ray_full_scans']))
    prev = st
r = np.array(rows[10:])
print('steps %d: us median %.1f mean %.1f p90 %.1f' % (len(r), np.median(r[:, 0]), r[:, 0].mean(), np.percentile(r[:, 0], 90)))
for lo, hi in ((0, 0), (1, 1), (2, 3), (4, 10**9)):
    sel = (r[:, 1] >= lo) & (r[:, 1] <= hi)
    if sel.any():
        print('  steps with %d..%d bail-outs: n %4d  median %.1f us  mean %.1f  (episodes ended / step %.1f)' % (lo, min(hi, 999), sel.sum(), np.median(r[sel, 0]), r[sel, 0].mean(), r[sel, 2].mean()))
print('  slowest: ' + '  '.join('%.0fus/b%d/e%d/f%d' % tuple(x) for x in r[np.argsort(-r[:, 0])[:16]]))
print('  fastest: ' + '  '.join('%.0fus/b%d/e%d/f%d' % tuple(x) for x in r[np.argsort(r[:, 0])[:8]]))
print('  total bail-outs %d over %d env-steps; full scans %d' % (r[:, 1].sum(), len(r) * args.envs, r[:, 3].sum()))
import ctypes
from paintrl_b200 import _capi
lib = _capi.lib()
if hasattr(lib, 'paintrl_debug_fast_reasons'):
    buf = (ctypes.c_ulonglong * 8)()
    if lib.paintrl_debug_fast_reasons(buf) == 1:
        print('  rays leaving the fast path by reason [-, outside grid, empty cell, no entering plane, attempts used up, no vertex candidates]:', list(buf)[:6])
