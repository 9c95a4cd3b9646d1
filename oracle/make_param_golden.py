"""Mint golden traces of the reference's grid-world ParamTestEnv (PaintRLEnv/param_test_env.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_param_golden        # -> tests/golden/p_param_test_env.npz

The reference module is imported VERBATIM under the gym shim S2 and a `termcolor` stub; every trace is
the output of the reference's own `ParamTestEnv.reset/step`: observations, rewards, done flags, info and
the final world / visit tables, for its four observation modes, its own zigzag / spiral policies
(param_test_env.py:283-342) and seeded random walks (wall hits, repeat-termination on and off).
"""
import contextlib
import importlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get('PAINTRL_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden', 'p_param_test_env.npz')


def load_reference():
    shims = os.path.join(ROOT, 'oracle', 'shims')
    for p in (os.path.join(REFERENCE_ROOT, 'PaintRLEnv'), shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module('param_test_env')


def _from_generator(gen):
    """Adapt an action generator of paintrl_b200.param_env (send the observation, receive the action) to the
    `policy(obs, t)` shape used below."""
    next(gen)
    return lambda obs, t: gen.send(obs)


def zigzag_actions(size, steps):
    """The reference's own column sweep (param_test_env.py:283-314)."""
    from paintrl_b200.param_env import zigzag_actions as gen
    return _from_generator(gen(size))


def spiral_actions(size):
    """The reference's own inward spiral (param_test_env.py:317-342)."""
    from paintrl_b200.param_env import spiral_actions as gen
    return _from_generator(gen(size))


def random_actions(seed, wall_bias):
    rng = np.random.default_rng(seed)

    def policy(obs, t):
        # mostly right / up so the walk covers ground before it hits a wall
        return int(rng.choice(4, p=wall_bias))
    return policy


def run(mod, mode, size, max_len, repeat, policy, max_steps, episodes=1):
    mod.ParamTestEnv.OBS_MODE = mode
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.ParamTestEnv(size, max_len=max_len, train_mode=True, termination_by_repeat=repeat)
    rec = {k: [] for k in ('actions', 'obs', 'actual', 'done', 'reward', 'penalty', 'first')}
    world = visit = None
    for ep in range(episodes):
        obs = env.reset()
        rec['first'].append(np.asarray(obs, dtype=np.float64))
        for t in range(max_steps):
            a = policy(obs, t)
            obs, actual, done, info = env.step(a)
            rec['actions'].append(a)
            rec['obs'].append(np.asarray(obs, dtype=np.float64))
            rec['actual'].append(actual)
            rec['done'].append(done)
            rec['reward'].append(info['reward'])
            rec['penalty'].append(info['penalty'])
            if done:
                break
        world = np.array([[env.world[(i, j)] for j in range(size)] for i in range(size)], dtype=np.int32)
        visit = np.array([[env.visit_table[(i, j)] for j in range(size)] for i in range(size)], dtype=np.int32)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out['world'], out['visit'] = world, visit
    return out


def main():
    mod = load_reference()
    cases = {
        'section14_zigzag': dict(mode='section', size=14, max_len=900, repeat=False, policy=zigzag_actions(14, 0), max_steps=2000),
        'section22_spiral': dict(mode='section', size=22, max_len=900, repeat=False, policy=spiral_actions(22), max_steps=2000),
        'section14_random': dict(mode='section', size=14, max_len=900, repeat=False,
                                 policy=random_actions(1, [0.3, 0.3, 0.2, 0.2]), max_steps=400, episodes=6),
        'section14_repeat': dict(mode='section', size=14, max_len=900, repeat=True,
                                 policy=random_actions(2, [0.4, 0.4, 0.1, 0.1]), max_steps=400, episodes=6),
        'section6_maxlen': dict(mode='section', size=6, max_len=5, repeat=False,
                                policy=random_actions(3, [0.25, 0.25, 0.25, 0.25]), max_steps=100, episodes=8),
        'simple14_random': dict(mode='simple', size=14, max_len=900, repeat=False,
                                policy=random_actions(4, [0.3, 0.3, 0.2, 0.2]), max_steps=300, episodes=4),
        'direct14_zigzag': dict(mode='direct', size=14, max_len=900, repeat=False, policy=zigzag_actions(14, 0), max_steps=60),
        'grid22_zigzag': dict(mode='grid', size=22, max_len=900, repeat=False, policy=zigzag_actions(22, 0), max_steps=500),
        'grid22_random': dict(mode='grid', size=22, max_len=900, repeat=False,
                              policy=random_actions(5, [0.3, 0.3, 0.2, 0.2]), max_steps=300, episodes=4),
    }
    arrays, meta = {}, {}
    for name, c in cases.items():
        out = run(mod, c['mode'], c['size'], c['max_len'], c['repeat'], c['policy'], c['max_steps'], c.get('episodes', 1))
        meta[name] = dict(mode=c['mode'], size=c['size'], max_len=c['max_len'], repeat=c['repeat'], episodes=c.get('episodes', 1))
        for k, v in out.items():
            arrays['%s/%s' % (name, k)] = v
        print('%-18s %4d steps, %d episode ends, return %.1f' % (name, len(out['actions']), int(out['done'].sum()), out['actual'].sum()))
    np.savez_compressed(OUT, meta=json.dumps(meta), **arrays)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
