"""Drop-in for the reference module of the same import path.

PaintRL's scripts do `from PaintRLEnv.robot_gym_env import PaintGymEnv` (paint_ppo.py:8,
zigzag.py:2, spiral.py:2); with this repository's root on `sys.path` that import resolves here
and the environment runs on the B200 engine instead of PyBullet + Python loops.
"""
from paintrl_b200.gym_env import PaintGymEnv, PaintVectorEnv, Part_Dict  # noqa: F401
