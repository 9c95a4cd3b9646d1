"""GPU parity through the reference-facing Python surface: the zigzag / spiral baseline scripts'
own control loops (zigzag.py:22-64, spiral.py:22-57) drive the drop-in `PaintGymEnv` and must
retrace the golden episodes minted by running those same loops on the verbatim reference."""
import random

import numpy as np
import pytest

from golden_util import Golden

pytestmark = pytest.mark.gpu


def _make_env(g, cuda_device):
    from PaintRLEnv.robot_gym_env import PaintGymEnv
    PaintGymEnv.DEVICE = cuda_device
    cfg = g.config
    PaintGymEnv.change_action_mode(cfg.get('action_shape', 1), cfg.get('action_mode', 'discrete'),
                                   cfg.get('discrete_granularity', 4))
    PaintGymEnv.change_obs_mode(cfg['obs_mode'], cfg['obs_grad'])
    return PaintGymEnv('unused-urdf-root', with_robot=False, renders=True, render_video=False,
                       rollout=g.rollout, extra_config=cfg['extra_config'])


def test_zigzag_script_loop_retraces_the_golden(cuda_device):
    g = Golden('g5_sheet_zigzag_discrete')
    with _make_env(g, cuda_device) as env:
        # zigzag.py:28-63, verbatim control flow
        horizontal_move, up, done = 0, True, False
        obs = [0] * 5
        t = 0
        while not done:
            current_pos = 0 if obs[-1] == 0 else round(1 / obs[-1])
            if up:
                if current_pos % 22 != 19:
                    action = 1
                elif horizontal_move < 2:
                    action = 0
                    horizontal_move += 1
                else:
                    horizontal_move, up = 0, False
                    continue
            else:
                if current_pos % 22 != 2:
                    action = 3
                elif horizontal_move < 2:
                    action = 0
                    horizontal_move += 1
                else:
                    horizontal_move, up = 0, True
                    continue
            assert action == g.actions(0, t), ('policy diverged at step', t)
            obs, reward, done, info = env.step(action)
            assert isinstance(obs, list) and isinstance(obs[0], np.float64)
            assert np.array_equal(np.asarray(obs), g['obs'][0, t + 1]), t
            assert reward == g['actual'][0, t] and info['reward'] == g['reward'][0, t]
            assert info['penalty'] == g['penalty'][0, t]
            assert done == bool(g['done'][0, t])
            t += 1
        assert t == int(g.lengths[0])
        assert np.array_equal(env.texture_status(), g['status_final'][0])
        assert len(env.replay_buffer) == t


def test_spiral_script_loop_with_robot_reset(cuda_device):
    g = Golden('g7_sheet_spiral_simple')
    with _make_env(g, cuda_device) as env:
        # spiral.py:28-38: centre of the start points, robot.reset(center_point)
        start_points = getattr(env, '_start_points')
        axis_1 = [sp[0][1] for sp in start_points]
        axis_2 = [sp[0][2] for sp in start_points]
        x = min(axis_1) + (max(axis_1) - min(axis_1)) / 2
        y = min(axis_2) + (max(axis_2) - min(axis_2)) / 2
        center_point = [[start_points[0][0][0], x, y], start_points[0][1]]
        assert np.array_equal(np.array(center_point), g['set_pose'][0])
        env.robot.reset(center_point)
        done, direction, strait_counter = False, 0, 1
        current_counter, t = strait_counter, 0
        while not done:
            current_counter -= 1
            obs, reward, done, info = env.step(direction % 4)
            assert np.array_equal(np.asarray(obs), g['obs'][0, t + 1]), t
            assert reward == g['actual'][0, t]
            if current_counter == 0:
                strait_counter += 1
                direction += 1
                current_counter = strait_counter
            t += 1
        assert t == int(g.lengths[0])
        assert np.array_equal(env.texture_status(), g['status_final'][0])


def test_grid_observation_is_an_ndarray_and_spaces_keep_the_reference_quirks(cuda_device):
    from PaintRLEnv.robot_gym_env import PaintGymEnv
    PaintGymEnv.DEVICE = cuda_device
    PaintGymEnv.OBS_GRAD = 4
    PaintGymEnv.change_action_mode(2, 'continuous')
    PaintGymEnv.change_obs_mode('grid', 4)
    assert PaintGymEnv.action_space.shape == (2,) and PaintGymEnv.observation_space.shape == (16,)
    random.seed(7)
    env = PaintGymEnv('x', with_robot=False)
    obs, r, done, info = env.step([0.3, -0.8])
    assert isinstance(obs, np.ndarray) and obs.shape == (16,) and set(info) == {'reward', 'penalty'}
    assert abs(env.robot.get_angle_diff() - np.arctan(abs(-0.8 / 0.3))) < 1e-12
    env.close()
    PaintGymEnv.change_obs_mode('section', 4)
    assert PaintGymEnv.observation_space.shape == (20,)       # robot_gym_env.py:186: 18 + 2 whatever grad is
    PaintGymEnv.change_action_mode(1, 'discrete', 4)
    with pytest.raises(NotImplementedError):
        PaintGymEnv('x', with_robot=True)


def test_vector_env_matches_single_envs(cuda_device):
    from paintrl_b200.gym_env import PaintVectorEnv
    from oracle.oracle import OracleBatch
    n = 48
    vec = PaintVectorEnv(n, None, device=cuda_device, seed=5)
    ora = OracleBatch(vec._engine.pack, vec._cfg, n)
    rng = np.random.RandomState(5)                    # same stream as the vector env's start draws
    first = vec.vector_reset()
    o_first = ora.reset(rng.randint(0, 4, size=n).astype(np.int32))
    assert len(first) == n and np.array_equal(np.stack(first), o_first)
    act_rng = np.random.default_rng(11)
    for t in range(60):
        acts = act_rng.integers(0, 4, size=n)
        obs, rew, done, infos = vec.vector_step(acts)
        o_obs, o_r, o_p, o_a, o_d = ora.step(acts)
        assert np.array_equal(np.stack(obs), o_obs) and rew == o_a.tolist() and done == [bool(d) for d in o_d]
        assert infos[3] == {'reward': float(o_r[3]), 'penalty': float(o_p[3])}
        for i in np.flatnonzero(o_d):                 # RLlib's sampler: reset_at(i) for every done env
            start = rng.randint(0, 4, size=1).astype(np.int32)
            assert np.array_equal(vec.reset_at(int(i)), ora.reset(start, env_ids=[int(i)])[0])
    assert vec.get_unwrapped() == []
    vec.close()
    ora.close()


def test_texture_image_follows_the_painted_status(cuda_device):
    """`texture_image()` (get_texture_image, bullet_paint_wrapper.py:737-738): painted front texels show the
    paint colour, everything else the labelled initial texture."""
    import numpy as np
    from PaintRLEnv.robot_gym_env import PaintGymEnv
    from paintrl_b200.config import DEFAULT_EXTRA_CONFIG
    env = PaintGymEnv('', with_robot=False, renders=False, rollout=True, extra_config=dict(DEFAULT_EXTRA_CONFIG))
    env.reset()
    for a in (1, 1, 0, 3):
        env.step(a)
    status = env.texture_status()
    img = np.asarray(env.texture_image())
    pack = env._engine.pack
    assert img.shape == (240, 240, 3) and (status == 255).sum() > 0
    flat = img.reshape(-1)
    off = pack.texel_offsets()
    painted = status == 255
    assert (flat[off[painted]] == 255).all() and (flat[off[painted] + 1] == 0).all() and (flat[off[painted] + 2] == 0).all()
    assert (flat[off[~painted]] == 191).all()
    env.close()
