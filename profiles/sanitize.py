#!/usr/bin/env python
"""A small pass through every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 9 python profiles/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize.py

staged / unstaged paint (RGB and HSI), the one-kernel and the two-kernel step, 8- and 32-lane move groups, the forced
hand-over of the move to the paint warp, grid observation, resets, state export / import, the synchronous and the
pipelined host-buffer step, the beam-fan paint method, the policy kernel and a rollout fragment, the rasteriser
(re-textured part and the loader on the synthetic part) and the grid world.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from paintrl_b200.batched_env import BatchedPaintEnv
from paintrl_b200.config import DEFAULT_EXTRA_CONFIG
from paintrl_b200.param_env import BatchedParamTestEnv

dev = torch.device('cuda:0')
rng = np.random.default_rng(0)


def run(n, extra, steps=6, texture=(240, 240), lanes=None, fused=None, bail=None, **kw):
    for key, val in (('PAINTRL_MOVE_LANES', lanes), ('PAINTRL_FUSED', fused), ('PAINTRL_DEBUG_BAIL_MOD', bail)):
        if val is not None:
            os.environ[key] = str(val)
        else:
            os.environ.pop(key, None)
    if isinstance(extra, dict):
        env = BatchedPaintEnv(n, dict(DEFAULT_EXTRA_CONFIG, **extra), device=dev, auto_reset=True, texture_size=texture, **kw)
    else:
        env = BatchedPaintEnv(n, extra, device=dev, texture_size=texture)
    env.reset(rng.integers(0, env.n_starts, size=n).astype(np.int32))
    out = env.host_buffers()
    for t in range(steps):
        if env.cfg.action_mode == 'discrete':
            acts = rng.integers(0, env.cfg.discrete_granularity, size=n)
        else:
            acts = rng.uniform(-1, 1, size=(n, env.action_dim))
        if t % 3 == 1:
            env.step_host(acts, out)
        elif t % 3 == 2:
            env.step_host_submit(np.ascontiguousarray(acts), out, slot=t & 1)
            env.step_host_wait(slot=t & 1)
        else:
            env.step(acts)
    st = env.get_state()
    env.set_state(status=st['status'], pose=st['pose'], quat=st['quat'], scalars=st['scalars'])
    env.reset(rng.integers(0, env.n_starts, size=3).astype(np.int32), env_ids=[0, 1, 2])
    env.job_status()
    torch.cuda.synchronize()
    stats = env.stats()
    env.close()
    return stats


print('door RGB one kernel  ', run(37, {}, fused=1))
print('door RGB two kernels ', run(37, {}, fused=0))
print('door RGB 8 lanes     ', run(37, {}, lanes=8, fused=0))
from paintrl_b200.config import EnvConfig
print('door RGB forced bails', run(24, {}, fused=0, bail=3))
print('door normal paint    ', run(13, EnvConfig(dict(DEFAULT_EXTRA_CONFIG), auto_reset=True, paint_method='normal'), steps=3))
print('sheet HSI normal     ', run(7, EnvConfig(dict(DEFAULT_EXTRA_CONFIG, Part_NO=1, COLOR_MODE='HSI'), auto_reset=True, paint_method='normal'), steps=3))
print('sheet HSI hybrid     ', run(21, dict(Part_NO=1, COLOR_MODE='HSI', TERMINATION_MODE='hybrid', OVERLAP_PENALTY=True)))
print('door grid continuous ', run(19, dict(START_POINT_MODE='all'), action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4))
print('door 640x640 unstaged', run(9, dict(START_POINT_MODE='edge'), steps=4, texture=(640, 640)))
print('sheet HSI 640 unstaged', run(5, dict(Part_NO=1, COLOR_MODE='HSI'), steps=3, texture=(640, 640)))
p = BatchedParamTestEnv(70, 14, 40, False, 'section', auto_reset=True, device=dev)
p.reset()
for t in range(30):
    p.step(rng.choice(4, size=70))
p.tables()
print('grid world           ', p.stats())
p.close()
from paintrl_b200.rollout import MlpPolicy, RolloutWorker
env = BatchedPaintEnv(160, dict(DEFAULT_EXTRA_CONFIG), device=dev, auto_reset=True)
worker = RolloutWorker(env, MlpPolicy(env.obs_dim, 4, device=dev, seed=3), fragment_length=6)
frag, _ = worker.collect()
frag, _ = worker.collect()
torch.cuda.synchronize()
print('rollout + policy     ', float(frag.reward.sum()))
env.close()
from paintrl_b200 import loader
pack = loader.load_part(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'data', 'urdf', 'painting', 'bulge.urdf'),
                        max_points=5200, part_no=100)
print('loader (rasteriser)  ', pack.n_texels)
print('sanitize pass complete')
