"""The reference's Python surfaces over the batched CUDA engine.

`PaintGymEnv`    -- drop-in for PaintRLEnv/robot_gym_env.py:120-422 (gym.Env, one environment):
                    same constructor, class attributes, classmethods (with their quirks), spaces,
                    `extra_config` keys, `step`/`reset` return conventions, and the attributes the
                    reference's scripts reach into (`_start_points`, `robot.reset(pose)`,
                    `robot.get_angle_diff()`, `replay_buffer`; zigzag.py:26, spiral.py:26-38,
                    robot_gym_env.py:425-449).
`PaintVectorEnv` -- the RLlib `VectorEnv` shape (vector_reset / reset_at / vector_step /
                    get_unwrapped) for thousands of environments on one GPU.

Both call `BatchedPaintEnv` -> the C ABI (include/paintrl.h).  PyBullet is not needed at step
time; the part tables come from the frozen part packs.  `gym` is used when importable, otherwise
a minimal stand-in for `gym.Env` / `spaces.Box` / `spaces.Discrete` (this image has neither gym
nor gymnasium).  There is no CPU fallback: constructing an environment without a CUDA device raises.
"""
import random

import numpy as np

from .config import DEFAULT_EXTRA_CONFIG, EnvConfig
from .partpack import PART_DICT

try:                                    # pragma: no cover - gym is absent in the build image
    import gym
    from gym import spaces
    from gym.utils import seeding
    _GymEnv = gym.Env
except Exception:                       # noqa: BLE001 - any import problem means "no gym"
    gym = None

    class _GymEnv(object):
        metadata = {}
        reward_range = (-float('inf'), float('inf'))

    class _Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float64):
            if shape is None:
                low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
                shape = low.shape
            else:
                low, high = np.full(shape, low, dtype=dtype), np.full(shape, high, dtype=dtype)
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return 'Box%s' % (self.shape,)

    class _Discrete(object):
        def __init__(self, n):
            self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return 'Discrete(%d)' % self.n

    class spaces(object):               # noqa: N801 - mirrors `from gym import spaces`
        Box = _Box
        Discrete = _Discrete

    class seeding(object):              # noqa: N801
        @staticmethod
        def np_random(seed=None):
            seed = int(np.random.SeedSequence(seed).entropy % (2 ** 32)) if seed is None else int(seed)
            return np.random.RandomState(seed), seed

Part_Dict = PART_DICT


def _obs_space(mode, grad):
    # robot_gym_env.py:166-173
    if mode == 'section':
        return spaces.Box(low=0.0, high=1.0, shape=(grad + 2,), dtype=np.float64)
    if mode == 'grid':
        return spaces.Box(low=0.0, high=1.0, shape=(grad ** 2,), dtype=np.float64)
    if mode == 'simple':
        return spaces.Box(low=0.0, high=1.0, shape=(2,), dtype=np.float64)
    return spaces.Box(low=0.0, high=1.0, shape=(grad + 1,), dtype=np.float64)


class _RobotView(object):
    """What the scripts touch on `env.robot` (robot.py:366-381)."""

    def __init__(self, env):
        self._env = env

    def reset(self, pose):
        # Robot.reset(pose): pose = [position, normal] (robot.py:366-372); also draws the ten
        # random.choice values of _reset_termination_variables (robot.py:214, 8-11)
        pos, normal = pose
        self._env._engine.set_pose(np.asarray(pos, dtype=np.float64)[None, :],
                                   np.asarray(normal, dtype=np.float64)[None, :])
        _draw_tmp_dir_name()

    def get_angle_diff(self):
        return float(self._env._engine.get_state(status=False)['angle_diff'][0])

    def termination_request(self):
        return bool(self._env._engine.get_state(status=False)['terminate'][0])

    def get_observation(self):
        st = self._env._engine.get_state(status=False)
        return tuple(st['pose'][0].tolist()), tuple(st['quat'][0].tolist())


def _draw_tmp_dir_name():
    import string
    for _ in range(10):                 # robot.py:8-11, 214: keeps the global `random` stream in step
        random.choice(string.ascii_letters)


class PaintGymEnv(_GymEnv):
    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 30}
    reward_range = (-1e3, 1e3)

    # robot_gym_env.py:126-132
    ACTION_SHAPE = 1
    ACTION_MODE = 'discrete'
    DISCRETE_GRANULARITY = 4
    OBS_MODE = 'section'
    OBS_GRAD = 4
    # Robot.PAINT_METHOD (robot.py:172): 'fast' -- the reference's setting -- or 'normal' (beam fan per shot)
    PAINT_METHOD = 'fast'
    EXTRA_CONFIG = dict(DEFAULT_EXTRA_CONFIG)

    action_space = spaces.Discrete(DISCRETE_GRANULARITY)
    observation_space = _obs_space(OBS_MODE, OBS_GRAD)

    # B200 build only: which CUDA device the single-environment engine lives on
    DEVICE = None

    @classmethod
    def change_obs_mode(cls, mode='section', grad=5):
        """robot_gym_env.py:176-193, quirks included: the section space is sized 18 + 2 whatever
        `grad` is, and the grid / discrete spaces use the OBS_GRAD in force *before* this call.
        The observation actually returned by step()/reset() always has the true length."""
        cls.OBS_MODE = mode
        if mode == 'section':
            cls.observation_space = spaces.Box(low=0.0, high=1.0, shape=(18 + 2,), dtype=np.float64)
        elif mode == 'grid':
            cls.observation_space = spaces.Box(low=0.0, high=1.0, shape=(cls.OBS_GRAD ** 2,), dtype=np.float64)
        elif mode == 'simple':
            cls.observation_space = spaces.Box(low=0.0, high=1.0, shape=(2,), dtype=np.float64)
        else:
            cls.observation_space = spaces.Box(low=0.0, high=1.0, shape=(cls.OBS_GRAD + 1,), dtype=np.float64)
        cls.OBS_GRAD = grad

    @classmethod
    def change_action_mode(cls, shape=2, mode='continuous', discrete_granularity=20):
        """robot_gym_env.py:195-205 (1-D continuous keeps the reference's Box(-1, -1))."""
        cls.ACTION_SHAPE = shape
        cls.ACTION_MODE = mode
        if mode == 'continuous':
            if shape == 1:
                cls.action_space = spaces.Box(np.array(-1,), np.array(-1,), dtype=np.float64)
            else:
                cls.action_space = spaces.Box(np.array((-1, -1)), np.array((1, 1)), dtype=np.float64)
        else:
            cls.action_space = spaces.Discrete(discrete_granularity)

    def __init__(self, urdf_root, with_robot=True, renders=False, render_video=False, rollout=False,
                 extra_config=None):
        if extra_config is None:
            extra_config = self.EXTRA_CONFIG
        if with_robot:
            # robot.py:220-233, 331-345: KUKA IK needs PyBullet; every script passes with_robot=False
            raise NotImplementedError('with_robot=True (KUKA IK through PyBullet) is outside the B200 paint path; '
                                      'construct with with_robot=False like paint_ppo.py:87 / zigzag.py:26')
        self._urdf_root = urdf_root
        self._with_robot = with_robot
        self._renders = renders          # accepted for signature parity; nothing is drawn
        self._render_video = render_video
        self._rollout = rollout
        granularity = self.action_space.n if self.ACTION_MODE != 'continuous' else self.DISCRETE_GRANULARITY
        self._cfg = EnvConfig(extra_config, action_mode=self.ACTION_MODE, action_shape=self.ACTION_SHAPE,
                              discrete_granularity=granularity, obs_mode=self.OBS_MODE, obs_grad=self.OBS_GRAD,
                              paint_method=self.PAINT_METHOD)
        self._setup_extra_config(extra_config)
        from .batched_env import BatchedPaintEnv      # raises without CUDA: no CPU fallback
        self._engine = BatchedPaintEnv(1, self._cfg, device=self.DEVICE, urdf_root=urdf_root)
        self._host = self._engine.host_buffers(pinned=True)
        starts = self._engine.pack.start_points(self.START_POINT_MODE)
        self._start_points = [[list(map(float, s[0])), list(map(float, s[1]))] for s in starts]
        self.robot = _RobotView(self)
        self.replay_buffer = []
        self.reset()                                   # robot_gym_env.py:287

    def _setup_extra_config(self, config):
        # robot_gym_env.py:240-252
        self.RENDER_WIDTH = config['RENDER_WIDTH']
        self.RENDER_HEIGHT = config['RENDER_HEIGHT']
        self._part_name = Part_Dict[config['Part_NO']][0]
        self._max_possible_point = Part_Dict[config['Part_NO']][1]
        self.Expected_Episode_Length = config['Expected_Episode_Length']
        self.EPISODE_MAX_LENGTH = config['EPISODE_MAX_LENGTH']
        self.TERMINATION_MODE = config['TERMINATION_MODE']
        self.SWITCH_THRESHOLD = config['SWITCH_THRESHOLD']
        self.START_POINT_MODE = config['START_POINT_MODE']
        self.TURNING_PENALTY = config['TURNING_PENALTY']
        self.OVERLAP_PENALTY = config['OVERLAP_PENALTY']
        self.COLOR_MODE = config['COLOR_MODE']

    def __enter__(self):
        self.reset()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def _format_obs(self, row):
        # robot_gym_env.py:306-319: grid -> ndarray, everything else -> list of np.float64
        if self.OBS_MODE == 'grid':
            return np.array(row, dtype=np.float64)
        return [np.float64(v) for v in row]

    def step(self, action):
        if self.ACTION_MODE == 'continuous':
            a = np.asarray(action, dtype=np.float64).reshape(1, -1)
        else:
            a = np.asarray([int(action)], dtype=np.int64)
        out = self._engine.step_host(a, self._host)
        done = bool(out['done'][0])
        if self._renders and self._rollout:              # robot_gym_env.py:363-367
            self.replay_buffer.append(action)
            if done:
                print(self.replay_buffer)
        info = {'reward': float(out['reward'][0]), 'penalty': float(out['penalty'][0])}
        return self._format_obs(out['obs'][0]), float(out['actual'][0]), done, info

    def reset(self):
        if self._rollout:
            index = 0
            self.replay_buffer = []
        else:
            random.randint(0, 7)                         # painted_mode draw, unused (:378)
            index = random.randint(0, len(self._start_points) - 1)
        obs = self._engine.reset(index)
        _draw_tmp_dir_name()
        return self._format_obs(obs[0].cpu().numpy())

    def render(self, mode='human'):
        if mode == 'human':
            raise Exception('please set render parameter to true to see the result')
        raise NotImplementedError('camera rendering needs PyBullet (robot_gym_env.py:389-415); use texture_image()')

    def texture_status(self):
        """First-channel value of every front texel (get_texture_image's R plane restricted to
        profile[front], bullet_paint_wrapper.py:737-738), in part-pack order."""
        return self._engine.get_state()['status'][0].cpu().numpy()

    def texture_image(self):
        """`get_texture_image` (bullet_paint_wrapper.py:737-738): the part's texture as the simulation has
        painted it so far, a PIL image when Pillow is importable, else the uint8 array [W, H, 3]."""
        arr = self._engine.pack.texture_image(self.texture_status(), self._engine.cfg.color_mode)
        try:
            from PIL import Image
            return Image.fromarray(arr, 'RGB')
        except Exception:                       # noqa: BLE001 - Pillow is optional
            return arr

    def close(self):
        if getattr(self, '_engine', None) is not None:
            self._engine.close()
            self._engine = None

    def seed(self, seed=None):
        _, seed = seeding.np_random(seed)
        return seed


class PaintVectorEnv(object):
    """RLlib `VectorEnv`-shaped view of `num_envs` environments on one GPU.

    RLlib (absent in this image) calls `vector_reset()`, then repeatedly `vector_step(actions)`
    and `reset_at(i)` for every environment that reported done -- the gym contract: the
    observation returned with done=True is the terminal one (robot_gym_env.py:358), the first
    observation of the next episode is what reset returns (robot_gym_env.py:387).
    `step_arrays` / `reset_done` are the batched equivalents that avoid per-environment Python.
    """

    def __init__(self, num_envs, extra_config=None, action_mode='discrete', action_shape=1,
                 discrete_granularity=4, obs_mode='section', obs_grad=4, device=None, rollout=False, seed=0,
                 paint_method='fast', urdf_root=None):
        from .batched_env import BatchedPaintEnv
        self._cfg = EnvConfig(extra_config, action_mode=action_mode, action_shape=action_shape,
                              discrete_granularity=discrete_granularity, obs_mode=obs_mode, obs_grad=obs_grad,
                              auto_reset=False, seed=seed, paint_method=paint_method)
        self._engine = BatchedPaintEnv(num_envs, self._cfg, device=device, urdf_root=urdf_root)
        self.num_envs = int(num_envs)
        self._rollout = rollout
        self._rng = np.random.RandomState(seed)
        if action_mode == 'continuous':
            self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(action_shape,), dtype=np.float64)
        else:
            self.action_space = spaces.Discrete(discrete_granularity)
        self.observation_space = _obs_space(obs_mode, obs_grad)
        self._host = self._engine.host_buffers(pinned=True)

    def _draw_starts(self, n):
        if self._rollout:
            return np.zeros(n, dtype=np.int32)
        return self._rng.randint(0, self._engine.n_starts, size=n).astype(np.int32)

    # ---- RLlib VectorEnv surface
    def vector_reset(self):
        obs = self._engine.reset(self._draw_starts(self.num_envs)).cpu().numpy()
        return [obs[i] for i in range(self.num_envs)]

    def reset_at(self, index):
        obs = self._engine.reset(self._draw_starts(1), env_ids=[int(index)])
        return obs[0].cpu().numpy()

    def vector_step(self, actions):
        obs, actual, done, reward, penalty = self.step_arrays(actions)
        # the VectorEnv contract wants Python lists; build them with bulk conversions (tolist / row views), not one
        # NumPy scalar access per environment -- at thousands of environments that loop, not the engine, set the pace
        infos = [{'reward': r, 'penalty': p} for r, p in zip(reward.tolist(), penalty.tolist())]
        return list(obs), actual.tolist(), done.tolist(), infos

    def get_unwrapped(self):
        return []

    # ---- batched equivalents
    def step_arrays(self, actions):
        """One `paintrl_step_host` for all environments; returns NumPy copies of
        (obs[B, D], actual[B], done[B], reward[B], penalty[B])."""
        out = self._engine.step_host(np.asarray(actions), self._host)
        return (out['obs'].copy(), out['actual'].copy(), out['done'].astype(bool), out['reward'].copy(),
                out['penalty'].copy())

    def reset_done(self, done):
        """Reset every environment flagged in `done` with one launch; returns (indices, first obs)."""
        ids = np.flatnonzero(np.asarray(done)).astype(np.int32)
        if ids.size == 0:
            return ids, np.zeros((0, self._engine.obs_dim))
        obs = self._engine.reset(self._draw_starts(ids.size), env_ids=ids)
        return ids, obs.cpu().numpy()

    def close(self):
        self._engine.close()
