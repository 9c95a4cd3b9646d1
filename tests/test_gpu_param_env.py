"""GPU parity for the batched grid-world ParamTestEnv (csrc/paintrl_param.cuh through the C ABI):
every golden trace minted from the reference's own module, the reference's zigzag / spiral drivers through
the drop-in import path, and batches of random walks against the CPU restatement (auto-reset included)."""
import numpy as np
import pytest
import torch

from param_golden_util import load

pytestmark = pytest.mark.gpu
CASES = load()


@pytest.mark.parametrize('name', sorted(CASES))
def test_engine_replays_reference_trace(name, cuda_device):
    from paintrl_b200.param_env import BatchedParamTestEnv
    c = CASES[name]
    env = BatchedParamTestEnv(3, c['size'], c['max_len'], c['repeat'], c['mode'], device=cuda_device)
    first = iter(c['first'])
    obs = env.reset().cpu().numpy()
    f = next(first)
    assert all(np.array_equal(obs[k], f) for k in range(3))
    for t, a in enumerate(c['actions']):
        o, actual, done, info = env.step([int(a)] * 3)
        o, actual, done = o.cpu().numpy(), actual.cpu().numpy(), done.cpu().numpy()
        for k in range(3):
            assert np.array_equal(o[k], c['obs'][t]), (name, t)
            assert actual[k] == c['actual'][t] and bool(done[k]) == bool(c['done'][t]), (name, t)
        assert info['reward'].cpu().numpy()[0] == c['reward'][t] and info['penalty'].cpu().numpy()[0] == c['penalty'][t]
        if done[0] and t + 1 < len(c['actions']):
            assert np.array_equal(env.reset().cpu().numpy()[1], next(first))
    w, v = env.tables()
    assert np.array_equal(w[2].cpu().numpy(), c['world'])
    assert np.array_equal(v[2].cpu().numpy(), c['visit'])
    st = env.stats()
    assert st['env_steps'] == 3 * len(c['actions']) and not st['bad_action_seen']
    env.close()


def test_reference_drivers_through_the_drop_in_module(cuda_device, capsys):
    """zigzag() / spiral() of the reference module (param_test_env.py:283-342) against the totals of the
    golden traces, through `PaintRLEnv.param_test_env`."""
    from PaintRLEnv.param_test_env import ParamTestEnv, spiral, zigzag
    z = CASES['section14_zigzag']
    steps, total = zigzag(14, env=ParamTestEnv(14, train_mode=True))
    assert steps == len(z['actions']) and np.isclose(total, z['actual'].sum(), rtol=0, atol=1e-9)
    s = CASES['section22_spiral']
    steps, total = spiral(22, env=ParamTestEnv(22, train_mode=False))          # prints the step log and the tables
    assert steps == len(s['actions']) and np.isclose(total, s['actual'].sum(), rtol=0, atol=1e-9)
    out = capsys.readouterr().out
    assert 'World Table:' in out and 'Visit Table' in out and 'STEP: 1 ' in out
    env = ParamTestEnv(14)
    assert env.observation_space.shape == (6,) and env.action_space.n == 4 and env.EPISODE_MAX_LENGTH == 900
    obs = env.reset()
    assert obs.shape == (6,) and env.get_current_pos() == (1, 1)
    assert env.world[(1, 1)] == 1 and env.world[(0, 0)] == 0 and env.visit_table[(1, 1)] == 1
    with pytest.raises(IndexError):
        env.step(7)
    env.close()


@pytest.mark.parametrize('mode,size,repeat', [('section', 14, False), ('section', 9, True), ('direct', 7, False),
                                              ('grid', 22, False), ('simple', 30, False)])
def test_batch_matches_oracle_with_auto_reset(mode, size, repeat, cuda_device):
    from oracle.param_oracle import ParamOracle
    from paintrl_b200.param_env import BatchedParamTestEnv
    n, steps = 257, 120
    env = BatchedParamTestEnv(n, size, 40, repeat, mode, auto_reset=True, device=cuda_device)
    oras = [ParamOracle(size, 40, repeat, mode) for _ in range(n)]
    rng = np.random.default_rng(7)
    assert np.array_equal(env.reset().cpu().numpy(), np.stack([o.observation() for o in oras]))
    ended = 0
    for t in range(steps):
        acts = rng.choice(4, size=n, p=[0.35, 0.35, 0.15, 0.15])
        o, actual, done, info = env.step(acts)
        o, actual, done, nxt = o.cpu().numpy(), actual.cpu().numpy(), done.cpu().numpy(), info['next_obs'].cpu().numpy()
        for k in range(n):
            ro, ra, rd, ri = oras[k].step(int(acts[k]))
            assert np.array_equal(o[k], ro) and actual[k] == ra and bool(done[k]) == rd, (t, k)
            if rd:
                ended += 1
                assert np.array_equal(nxt[k], oras[k].reset()), (t, k)
            else:
                assert np.array_equal(nxt[k], ro)
    assert ended > n and env.stats()['episodes_ended'] == ended
    env.close()


def test_bad_action_raises_like_the_reference(cuda_device):
    """param_test_env.py:173-174 raises IndexError for an action outside 0..3: so does the batched mirror, before
    any world is stepped, for host lists and for CUDA tensors.  With the check switched off the offending world
    is left untouched but its output row is fresh (current observation, zero reward, done), the flag is
    reported once and then cleared."""
    from paintrl_b200.param_env import BatchedParamTestEnv
    env = BatchedParamTestEnv(4, 8, device=cuda_device)
    first = env.reset().clone()
    before = env.tables()[0].clone()
    for bad in ([0, 9, 1, -1], torch.tensor([0, 4, 1, 2], device=cuda_device)):
        with pytest.raises(IndexError):
            env.step(bad)
    assert torch.equal(env.tables()[0], before) and env.stats()['env_steps'] == 0
    o, actual, done, info = env.step(torch.tensor([0, 9, 1, -1], device=cuda_device), check_actions=False)
    st = env.stats()
    assert st['bad_action_seen'] and st['env_steps'] == 2
    assert not env.stats()['bad_action_seen']                 # cleared once reported
    after = env.tables()[0]
    assert torch.equal(before[1], after[1]) and torch.equal(before[3], after[3]) and not torch.equal(before[0], after[0])
    assert done.cpu().tolist() == [0, 1, 0, 1] and actual.cpu().tolist()[1] == 0.0 and actual.cpu().tolist()[3] == 0.0
    assert torch.equal(o[1], first[1]) and torch.equal(info['next_obs'][3], first[3])
    env.close()


def test_env_ids_are_validated(cuda_device):
    from paintrl_b200.param_env import BatchedParamTestEnv
    env = BatchedParamTestEnv(4, 8, device=cuda_device)
    for bad, exc in (([4], IndexError), ([-1], IndexError), ([1, 1], ValueError), ([], ValueError)):
        with pytest.raises(exc):
            env.reset(env_ids=bad)
        with pytest.raises(exc):
            env.tables(env_ids=bad)
    assert env.reset(env_ids=[3, 0]).shape[0] == 2
    env.close()


def test_spiral_driver_terminates_on_any_world(cuda_device, capsys):
    """The reference's spiral() keeps stepping in its last direction once the legs are used up
    (param_test_env.py:326-340) until the episode ends at a wall; the generator does the same instead of
    spinning without yielding (sizes where the legs run out before the world is consumed)."""
    from paintrl_b200.param_env import ParamTestEnv, drive, spiral_actions
    for size, max_len in ((5, 900), (9, 900), (22, 900)):
        env = ParamTestEnv(size, max_len=max_len)
        steps, _ = drive(env, spiral_actions(size))
        assert 0 < steps <= max(max_len, (size - 2) ** 2)
        env.close()
    gen = spiral_actions(4)            # grid_size - 3 = 1: one-step legs, then none
    next(gen)
    assert [next(gen) for _ in range(8)] == [0, 1, 2, 3, 3, 3, 3, 3]
