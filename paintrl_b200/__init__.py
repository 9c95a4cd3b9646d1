"""paintrl_b200 -- the B200-native batched paint-simulation step of translearn/PaintRL.

Public names (resolved lazily, so `import paintrl_b200` stays cheap and never needs a GPU):

    BatchedPaintEnv, EnvConfig, PartPack          the batched engine over the C ABI (include/paintrl.h)
    PaintGymEnv, PaintVectorEnv                   the reference's gym.Env / RLlib VectorEnv surfaces
    BatchedParamTestEnv, ParamTestEnv             the grid world of PaintRLEnv/param_test_env.py
    RolloutWorker, MlpPolicy, gae                 on-GPU rollout fragments
"""
_EXPORTS = {
    'BatchedPaintEnv': 'batched_env', 'EnvConfig': 'config', 'PartPack': 'partpack',
    'PaintGymEnv': 'gym_env', 'PaintVectorEnv': 'gym_env',
    'BatchedParamTestEnv': 'param_env', 'ParamTestEnv': 'param_env',
    'RolloutWorker': 'rollout', 'MlpPolicy': 'rollout', 'gae': 'rollout',
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        import importlib
        return getattr(importlib.import_module('.' + _EXPORTS[name], __name__), name)
    raise AttributeError('module %r has no attribute %r' % (__name__, name))
