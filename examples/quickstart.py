#!/usr/bin/env python
"""Three ways to drive the engine (needs a CUDA device; run from the repository root).

    python examples/quickstart.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from paintrl_b200 import BatchedPaintEnv, MlpPolicy, RolloutWorker
from paintrl_b200.config import DEFAULT_EXTRA_CONFIG
from PaintRLEnv.robot_gym_env import PaintGymEnv          # the reference's own import path

# 1. the reference's gym.Env, one environment (what zigzag.py / spiral.py / paint_*.py construct)
env = PaintGymEnv('', with_robot=False, renders=False, rollout=True, extra_config=dict(DEFAULT_EXTRA_CONFIG))
obs = env.reset()
total = 0.0
for t in range(20):
    obs, reward, done, info = env.step(1 if obs[-1] < 0.95 else 0)
    total += reward
print('gym.Env: 20 steps, return %.3f, last info %s' % (total, info))
env.texture_image().save('/tmp/paintrl_texture.png') if hasattr(env.texture_image(), 'save') else None
env.close()

# 2. thousands of environments, device tensors in and out
batch = BatchedPaintEnv(4096, dict(DEFAULT_EXTRA_CONFIG), auto_reset=True)
batch.reset(torch.randint(0, batch.n_starts, (4096,), dtype=torch.int32))
for t in range(100):
    obs, actual, done, info = batch.step(torch.randint(0, 4, (4096,), device=batch.device))
print('batched: %s' % batch.stats())

# 3. rollout fragments with a policy on the GPU (paint_ppo.py's sampling shape)
worker = RolloutWorker(batch, MlpPolicy(batch.obs_dim, 4, device=batch.device), fragment_length=100, use_cuda_graph=True)
for it in range(3):
    fragment, stats = worker.collect()
    worker.advance()
print('rollout: %d env-steps per fragment, %d episodes ended in the last one, mean return %.3f' % (
    stats['env_steps'], stats['episodes'], stats['sum_return'] / max(1.0, stats['episodes'])))
batch.close()
