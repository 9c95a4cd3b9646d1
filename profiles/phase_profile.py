#!/usr/bin/env python
"""Phase timing of the step kernels (clock64 between marks, averaged per environment-step).

    PAINTRL_PROFILE=1 python -m paintrl_b200.build --force      # instrumented build
    python profiles/phase_profile.py [--envs N] [--workload c2|c3] [--steps K]
    python -m paintrl_b200.build --force                        # back to the product build

The instrumented build is for diagnosis only; its timings are never bench values.
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from paintrl_b200 import _capi
from paintrl_b200.batched_env import BatchedPaintEnv
from paintrl_b200.config import EnvConfig

NAMES = {0: 'move: load + action + first p', 1: 'move: p / end of later sub-steps', 6: 'ray: guess cell, entry, issue blob loads',
         7: 'ray: slab pass (div, reductions)', 8: 'ray: region check / accept', 9: 'ray: verify / full scan / hit',
         10: 'hook: nearest vertex', 11: 'hook: triangles + pick', 12: 'move: pose, quat, centre', 13: 'move: stores',
         15: 'paint: early row ranks + TCP-row prefetch', 16: 'paint: TMA loads + mbarrier wait', 17: 'stamp: shot floats, bbox', 18: 'stamp: row ranges', 19: 'stamp: words (ball tests, bits)',
         20: 'score + termination', 21: 'obs: normalised pose', 22: 'obs: row ranks', 23: 'obs: rows + words (popc)',
         24: 'obs: TCP row', 25: 'obs: reductions + write', 26: 'paint: outputs, reset, stores'}

ap = argparse.ArgumentParser()
ap.add_argument('--envs', type=int, default=4096)
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=50)
ap.add_argument('--no-flush', action='store_true')
ap.add_argument('--per-env', action='store_true', help='last step: phase cycles of the slowest move warps and by slow-path class')
ap.add_argument('--dump-rays', default=None, help='write the rays that left the fast path to this .npy file')
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
cfg = EnvConfig(w['extra'], auto_reset=True, seed=1234, **w['kw'])
dev = torch.device('cuda:0')
env = BatchedPaintEnv(args.envs, cfg, device=dev)
lib = _capi.lib()
lib.paintrl_debug_profile.restype = ctypes.c_int
lib.paintrl_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
gen = torch.Generator(device=dev); gen.manual_seed(1234)
acts = torch.randint(0, cfg.discrete_granularity, (args.steps + 10, args.envs), generator=gen, device=dev, dtype=torch.int64)
env.reset(torch.randint(0, env.n_starts, (args.envs,), generator=gen, device=dev, dtype=torch.int32))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for i in range(10):
    env.step(acts[i])
buf = (ctypes.c_ulonglong * 64)()
if lib.paintrl_debug_profile(buf, 1) == 0:
    raise SystemExit('not an instrumented build: PAINTRL_PROFILE=1 python -m paintrl_b200.build --force')
for i in range(args.steps):
    if not args.no_flush:
        flush.fill_(i & 255)
    env.step(acts[10 + i])
lib.paintrl_debug_profile(buf, 0)
if args.dump_rays:
    import numpy as np
    rays = np.zeros((4096, 8))
    lib.paintrl_debug_rays.restype = ctypes.c_int
    lib.paintrl_debug_rays.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    n_r = lib.paintrl_debug_rays(rays.ctypes.data, 4096, 0)
    np.save(args.dump_rays, rays[:n_r])
    print('dumped %d slow-path rays to %s' % (n_r, args.dump_rays))
n = args.steps * args.envs
tot_m = sum(buf[i] for i in range(0, 16)); tot_p = sum(buf[i] for i in range(16, 32))
print('%s, %d envs, %d steps: cycles per environment-step (leader lane, L2 %s between steps)' % (args.workload, args.envs, args.steps, 'warm' if args.no_flush else 'flushed'))
for i in range(32):
    if buf[i]:
        print('  [%2d] %-44s %9.0f' % (i, NAMES.get(i, '?'), buf[i] / n))
print('  move kernel total %.0f cycles, paint kernel total %.0f cycles' % (tot_m / n, tot_p / n))

if args.per_env:
    import numpy as np
    lib.paintrl_debug_profile_env.restype = ctypes.c_int
    lib.paintrl_debug_profile_env.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rows = np.zeros((min(args.envs, 65536), 32), dtype=np.uint32)
    lib.paintrl_debug_profile_env(rows.ctypes.data, rows.shape[0])
    cnt = rows[:, 14].astype(np.int64)
    off, full, verify = cnt & 0xf, (cnt >> 8) & 0xff, (cnt >> 16) & 0xff
    ph = rows.astype(np.int64)
    ph[:, 14] = 0
    move_total = ph[:, :14].sum(1)
    slots = [k for k in range(14) if ph[:, k].any()]
    print('last step, move phases per environment (cycles): ' + ' '.join('[%d]' % k for k in slots))
    for nm, sel in (('no slow path', (full == 0) & (verify == 0)), ('verify only', (full == 0) & (verify > 0)), ('full scans', full > 0)):
        if sel.any():
            print('  %-12s n %5d  total %7.0f : %s' % (nm, sel.sum(), move_total[sel].mean(), ' '.join('%6.0f' % ph[sel, k].mean() for k in slots)))
    for i in np.argsort(-move_total)[:12]:
        print('  env %5d off %d full %d verify %d total %7d : %s' % (i, off[i], full[i], verify[i], move_total[i], ' '.join('%6d' % ph[i, k] for k in slots)))
