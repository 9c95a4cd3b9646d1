"""Loader for tests/golden/p_param_test_env.npz (traces of the reference's ParamTestEnv)."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'p_param_test_env.npz')


def load():
    with np.load(PATH, allow_pickle=False) as z:
        meta = json.loads(str(z['meta']))
        cases = {}
        for name, m in meta.items():
            cases[name] = dict(m, **{k.split('/', 1)[1]: z[k] for k in z.files if k.startswith(name + '/')})
    return cases
