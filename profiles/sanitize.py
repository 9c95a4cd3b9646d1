#!/usr/bin/env python
"""A small pass through every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 9 python profiles/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize.py

staged / unstaged paint (RGB and HSI), 8- and 32-lane move groups, grid observation, resets, state
export / import, the host-buffer step, the rasteriser (re-textured part) and the grid world.
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


def run(n, extra, steps=6, texture=(240, 240), lanes=None, **kw):
    if lanes:
        os.environ['PAINTRL_MOVE_LANES'] = str(lanes)
    else:
        os.environ.pop('PAINTRL_MOVE_LANES', None)
    env = BatchedPaintEnv(n, dict(DEFAULT_EXTRA_CONFIG, **extra), device=dev, auto_reset=True, texture_size=texture, **kw)
    env.reset(rng.integers(0, env.n_starts, size=n).astype(np.int32))
    out = env.host_buffers()
    for t in range(steps):
        if env.cfg.action_mode == 'discrete':
            acts = rng.integers(0, env.cfg.discrete_granularity, size=n)
        else:
            acts = rng.uniform(-1, 1, size=(n, env.action_dim))
        if t % 2:
            env.step_host(acts, out)
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


print('door RGB staged      ', run(37, {}))
print('door RGB 8 lanes     ', run(37, {}, lanes=8))
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
print('sanitize pass complete')
