"""Drop-in for the reference module of the same import path.

PaintRL's hyper-parameter scripts do `from PaintRLEnv.param_test_env import ParamTestEnv`
(param_test_ppo.py:5, param_test_dqn.py, ...); with this repository's root on `sys.path` that import
resolves here and the grid world runs on the B200 engine.
"""
from paintrl_b200.param_env import BatchedParamTestEnv, ParamTestEnv, Visualizer, spiral, zigzag  # noqa: F401

if __name__ == '__main__':
    spiral(20)
