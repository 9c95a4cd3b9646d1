"""GPU parity, randomised configurations: seeded draws over every configuration axis of the reference
(part, colour mode, observation mode and granularity, action mode / shape / granularity, termination mode and
its thresholds, penalties, start-point mode, batch size), each stepped against the C oracle with auto-reset."""
import numpy as np
import pytest

from test_gpu_oracle_batch import BASE, _run_case

pytestmark = pytest.mark.gpu


def _draw(rng):
    obs_mode = rng.choice(['section', 'grid', 'simple', 'discrete'])
    obs_grad = int({'section': rng.choice([1, 2, 4, 5, 8, 18, 36]), 'grid': rng.choice([1, 2, 4, 7, 10]),
                    'simple': 4, 'discrete': rng.choice([4, 6, 12])}[obs_mode])
    action_mode = rng.choice(['discrete', 'continuous'])
    kw = dict(obs_mode=str(obs_mode), obs_grad=obs_grad, action_mode=str(action_mode))
    if action_mode == 'continuous':
        kw['action_shape'] = int(rng.choice([1, 2]))
    else:
        kw['discrete_granularity'] = int(rng.choice([2, 3, 4, 8, 20, 36]))
    term = str(rng.choice(['late', 'early', 'hybrid']))
    extra = dict(BASE, Part_NO=int(rng.choice([0, 1, 5, 9])), COLOR_MODE=str(rng.choice(['RGB', 'HSI'])),
                 TERMINATION_MODE=term, START_POINT_MODE=str(rng.choice(['fixed', 'anchor', 'edge', 'all'])),
                 TURNING_PENALTY=bool(rng.integers(0, 2)), OVERLAP_PENALTY=bool(rng.integers(0, 2)),
                 Expected_Episode_Length=int(rng.choice([30, 245, 600])), EPISODE_MAX_LENGTH=int(rng.choice([8, 40, 245])),
                 SWITCH_THRESHOLD=float(rng.choice([0.1, 0.5, 0.9])))
    return extra, kw, int(rng.choice([3, 16, 45])), int(rng.choice([8, 14, 20]))


@pytest.mark.parametrize('seed', range(20))
def test_random_configuration_matches_oracle(seed, cuda_device):
    extra, kw, num_envs, steps = _draw(np.random.default_rng(1000 + seed))
    _run_case(extra, kw, num_envs, steps, cuda_device, status_every=6)
