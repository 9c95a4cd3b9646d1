#!/usr/bin/env python
"""Per-warp timeline of one step (globaltimer at the start / end of every environment's move and paint work).

    PAINTRL_TRACE=1 python -m paintrl_b200.build --force     # trace build (diagnosis only, never a bench value)
    python profiles/timeline.py [--envs N] [--workload c2]
    python -m paintrl_b200.build --force                     # back to the product build
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from paintrl_b200 import _capi
from paintrl_b200.batched_env import BatchedPaintEnv
from paintrl_b200.config import EnvConfig

ap = argparse.ArgumentParser()
ap.add_argument('--envs', type=int, default=4096)
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=40)
ap.add_argument('--out', default=None)
ap.add_argument('--uniform', action='store_true', help='every environment at start point 0 with action 0: identical work per warp, what is left is scheduling')
ap.add_argument('--no-flush', action='store_true')
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
cfg = EnvConfig(w['extra'], auto_reset=True, seed=1234, **w['kw'])
dev = torch.device('cuda:0')
env = BatchedPaintEnv(args.envs, cfg, device=dev, texture_size=w.get('texture', (240, 240)))
lib = _capi.lib()
lib.paintrl_debug_trace.restype = ctypes.c_int
lib.paintrl_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
gen = torch.Generator(device=dev); gen.manual_seed(1234)
acts = torch.randint(0, cfg.discrete_granularity, (args.steps, args.envs), generator=gen, device=dev, dtype=torch.int64)
env.reset(torch.randint(0, env.n_starts, (args.envs,), generator=gen, device=dev, dtype=torch.int32))
if args.uniform:
    acts.zero_()
    env.reset(torch.zeros(args.envs, dtype=torch.int32, device=dev))
    args.steps = min(args.steps, 6)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for i in range(args.steps):
    if not args.no_flush:
        flush.fill_(i & 255)
    env.step(acts[i])
torch.cuda.synchronize()
n = min(args.envs, 65536)
tr = np.zeros((n, 8), dtype=np.uint64)
if lib.paintrl_debug_trace(tr.ctypes.data, n) == 0:
    raise SystemExit('not a trace build: PAINTRL_TRACE=1 python -m paintrl_b200.build --force')
t = tr[:, :6].astype(np.int64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3     # microseconds since the first move warp started
names = ['move start', 'move end', 'paint start', 'paint dependency resolved', 'paint inputs loaded', 'paint end']
print('%s, %d envs: last step, microseconds since the first move warp started' % (args.workload, n))
for k, nm in enumerate(names):
    c = t[:, k]
    print('  %-28s min %7.2f  p10 %7.2f  median %7.2f  p90 %7.2f  max %7.2f' % (nm, c.min(), np.percentile(c, 10), np.median(c), np.percentile(c, 90), c.max()))
dm, dp = t[:, 1] - t[:, 0], t[:, 5] - t[:, 4]
for nm, d in (('move duration per warp', dm), ('paint duration per warp (after inputs)', dp)):
    print('  %-40s min %6.2f  p10 %6.2f  median %6.2f  p90 %6.2f  p99 %6.2f  max %6.2f' % (nm, d.min(), np.percentile(d, 10), np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()))
sm_m, sm_p = (tr[:, 6] & 0xffff).astype(int), tr[:, 7].astype(int)
cnt = ((tr[:, 6] >> 16) & 0xffffffff).astype(np.int64)
max_planes, max_verts = ((tr[:, 6] >> 48) & 0xff).astype(int), ((tr[:, 6] >> 56) & 0xff).astype(int)
off, full, verify, attempts = cnt & 0xf, (cnt >> 8) & 0xff, (cnt >> 16) & 0xff, (cnt >> 24) & 0x7f
slow = np.argsort(-dm)[:24]
print('  slowest move warps: duration us / off-part sub-steps / full plane scans (two-kernel step) or cached-pair tests (one-kernel step) / verify passes / cell attempts')
print('   ' + '  '.join('%.1f/%d/%d/%d/%d' % (dm[i], off[i], full[i], verify[i], attempts[i]) for i in slow))
print('  their SMs: ' + ' '.join(str(sm_m[i]) for i in slow) + '   start us: ' + ' '.join('%.1f' % t[i, 0] for i in slow))
print('  their largest plane list / vertex list: ' + ' '.join('%d/%d' % (max_planes[i], max_verts[i]) for i in slow))
for lo_, hi_ in ((0, 17), (17, 33), (33, 65), (65, 129), (129, 256)):
    sel = (max_planes >= lo_) & (max_planes < hi_)
    if sel.any():
        print('  move duration, largest plane list in [%d, %d): n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (lo_, hi_, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
for lo_, hi_ in ((0, 17), (17, 33), (33, 65), (65, 256)):
    sel = (max_verts >= lo_) & (max_verts < hi_)
    if sel.any():
        print('  move duration, largest vertex list in [%d, %d): n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (lo_, hi_, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
pose = env.get_state()['pose'].cpu().numpy()[:n]
pk = env.pack
a0, a1 = pk.axes
r = pk.ranges
n0 = (pose[:, a0] - r[0, 0]) / (r[0, 1] - r[0, 0]); n1 = (pose[:, a1] - r[1, 0]) / (r[1, 1] - r[1, 0])
print('  their poses after the step (fraction of the part extent along axis 0 / axis 1): ' + ' '.join('%.2f/%.2f' % (n0[i], n1[i]) for i in slow))
edge = np.minimum(np.minimum(n0, 1 - n0), np.minimum(n1, 1 - n1))
for lo_, hi_ in ((-9, 0.02), (0.02, 0.05), (0.05, 0.1), (0.1, 0.2), (0.2, 0.6)):
    sel = (edge >= lo_) & (edge < hi_)
    if sel.any():
        print('  move duration, distance to the bounding box edge in [%.2f, %.2f): n %5d  median %6.2f  p90 %6.2f  max %6.2f   paint median %6.2f' % (lo_, hi_, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max(), np.median(dp[sel])))
for k in range(5, 16):
    sel = attempts == k
    if sel.any():
        print('  move duration, %2d cell attempts: n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (k, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
for k in range(0, 6):
    sel = verify == k
    if sel.any():
        print('  move duration, %d verify passes: n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (k, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
for nm, sel in (('no slow path', (full == 0) & (verify == 0)), ('verify only', (full == 0) & (verify > 0)), ('full scans', full > 0)):
    if sel.any():
        print('  move duration, %-12s: n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (nm, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
for k in range(6):
    sel = off == k
    if sel.any():
        print('  move duration, %d off-part sub-steps: n %5d  median %6.2f  p90 %6.2f  max %6.2f' % (k, sel.sum(), np.median(dm[sel]), np.percentile(dm[sel], 90), dm[sel].max()))
print('  move: warps per SM min %d max %d; last move end per SM: min %.2f median %.2f max %.2f' % (
    np.bincount(sm_m).min(), np.bincount(sm_m).max(), *np.percentile([t[sm_m == s, 1].max() for s in np.unique(sm_m)], [0, 50, 100])))
print('  paint: warps per SM min %d max %d; last paint end per SM: min %.2f median %.2f max %.2f' % (
    np.bincount(sm_p).min(), np.bincount(sm_p).max(), *np.percentile([t[sm_p == s, 5].max() for s in np.unique(sm_p)], [0, 50, 100])))
# occupancy over time: warps in flight
for nm, a, b in (('move', t[:, 0], t[:, 1]), ('paint', t[:, 4], t[:, 5])):
    grid = np.linspace(0, t[:, 5].max(), 41)
    inflight = [(int(((a <= g) & (b > g)).sum())) for g in grid]
    print('  %s warps in flight every %.1f us: %s' % (nm, grid[1], ' '.join(str(v) for v in inflight)))
if args.out:
    np.save(args.out, tr)
