"""Runs PaintRL's OWN sources, verbatim, from /root/reference under shims S1-S5.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing here is imported by the product; it is
used in the build container to (a) mint the golden vectors and part packs under tests/golden and
paintrl_b200/data (oracle/make_golden.py) and (b) cross-check the C restatement
(oracle/paint_oracle.c).  /root/reference does not exist on the GPU box, so nothing under
`-m gpu`, `smoke()` or `bench.py` imports this module.

Shims:
  S1  oracle/shims/pybullet.py        exact ray-vs-hull + scalar multiplyTransforms, no-op GUI
  S2  oracle/shims/gym, pybullet_data minimal Env/spaces/seeding/error/logger
  S3  WritableDataKDTree              `cKDTree.data` became read-only in modern SciPy;
                                      bullet_paint_wrapper.py:944-947 writes into it
  S4  headless policy                 construct with renders=True (renders=False dies at
                                      bullet_paint_wrapper.py:597), silence prints, no-op
                                      Part.write_text_info (bullet_paint_wrapper.py:327 breaks on
                                      1-D continuous actions)
  S5  HSI texels as Python ints       NumPy-1 promotion `uint8 - int -> int64`
                                      (bullet_paint_wrapper.py:406-417) instead of NumPy-2 wrap
"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get('PAINTRL_REFERENCE', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')
_REF_MODULES = ('bullet_paint_wrapper', 'robot', 'robot_gym_env', 'video_renderer')

DEFAULT_EXTRA_CONFIG = {
    'RENDER_HEIGHT': 720, 'RENDER_WIDTH': 960, 'Part_NO': 0,
    'Expected_Episode_Length': 245, 'EPISODE_MAX_LENGTH': 245,
    'TERMINATION_MODE': 'late', 'SWITCH_THRESHOLD': 0.9, 'START_POINT_MODE': 'anchor',
    'TURNING_PENALTY': False, 'OVERLAP_PENALTY': False, 'COLOR_MODE': 'RGB',
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'PaintRLEnv', 'robot_gym_env.py'))


class WritableDataKDTree(object):
    """S3: a cKDTree whose `.data` is a writable *copy*; queries see the original points."""

    def __init__(self, data, *args, **kwargs):
        from scipy.spatial import cKDTree
        self._tree = cKDTree(data, *args, **kwargs)
        self.data = np.array(self._tree.data, dtype=np.float64, copy=True)

    @property
    def tree_points(self):
        return self._tree.data

    def query(self, *args, **kwargs):
        return self._tree.query(*args, **kwargs)

    def query_ball_point(self, *args, **kwargs):
        return self._tree.query_ball_point(*args, **kwargs)


def load_reference_modules():
    """Import fresh copies of the reference modules (their class attributes are process-global
    configuration, robot_gym_env.py:126-205) with the shim directories on sys.path."""
    if not reference_available():
        raise RuntimeError('reference sources not found under %s' % REFERENCE_ROOT)
    ref_dir = os.path.join(REFERENCE_ROOT, 'PaintRLEnv')
    for path in (ref_dir, _SHIMS):
        if path in sys.path:
            sys.path.remove(path)
    sys.path.insert(0, ref_dir)
    sys.path.insert(0, _SHIMS)
    for name in _REF_MODULES + ('pybullet', 'pybullet_data', 'gym', 'gym.spaces', 'gym.error',
                                'gym.logger', 'gym.utils', 'gym.utils.seeding'):
        sys.modules.pop(name, None)
    bpw = importlib.import_module('bullet_paint_wrapper')
    assert bpw.__file__.startswith(REFERENCE_ROOT), bpw.__file__
    bpw.cKDTree = WritableDataKDTree                      # S3
    bpw.Part.write_text_info = lambda self, *a, **k: None  # S4
    rob = importlib.import_module('robot')
    rge = importlib.import_module('robot_gym_env')
    assert rge.__file__.startswith(REFERENCE_ROOT), rge.__file__
    return bpw, rob, rge


class ReferenceEnv(object):
    """One verbatim `PaintGymEnv` plus the hooks the golden generator needs."""

    def __init__(self, extra_config=None, action_mode='discrete', action_shape=1,
                 discrete_granularity=4, obs_mode='section', obs_grad=4, rollout=False,
                 quiet=True, paint_method='fast', urdf_root=None, extra_parts=None):
        self.bpw, self.rob, self.rge = load_reference_modules()
        self.rob.Robot.PAINT_METHOD = paint_method          # robot.py:172 (a class constant edited in source upstream)
        if extra_parts:
            # parts outside the reference's table (the repository's synthetic test part): Part_NO -> [urdf, max points],
            # found under `urdf_root`/urdf/painting like the reference's own (robot_gym_env.py:106-117, 273)
            self.rge.Part_Dict.update(extra_parts)
        cfg = dict(DEFAULT_EXTRA_CONFIG)
        if extra_config:
            cfg.update(extra_config)
        self.extra_config = cfg
        cls = self.rge.PaintGymEnv
        # Same effect as editing the class attributes by hand (robot_gym_env.py:126-132); the
        # classmethods are not used here because change_obs_mode sizes the spaces with stale
        # values (robot_gym_env.py:185-193) -- that quirk is mirrored in the product separately.
        cls.ACTION_MODE = action_mode
        cls.ACTION_SHAPE = action_shape
        cls.DISCRETE_GRANULARITY = discrete_granularity
        cls.OBS_MODE = obs_mode
        cls.OBS_GRAD = obs_grad
        spaces = self.rge.spaces
        if action_mode == 'continuous':
            if action_shape == 2:
                cls.action_space = spaces.Box(np.array((-1, -1)), np.array((1, 1)), dtype=np.float64)
            else:
                cls.action_space = spaces.Box(low=-1.0, high=1.0, shape=(1,), dtype=np.float64)
        else:
            cls.action_space = spaces.Discrete(discrete_granularity)
        self.quiet = quiet
        self.randint_log = []
        real_randint = self.rge.randint

        def logged_randint(a, b):
            if self._forced_start is not None and (a, b) == (0, len(self.env._start_points) - 1):
                v = self._forced_start
            else:
                v = real_randint(a, b)
            self.randint_log.append((a, b, v))
            return v

        self._forced_start = None
        self.rge.randint = logged_randint
        with self._silence():
            self.env = cls(urdf_root or os.path.join(REFERENCE_ROOT, 'PaintRLEnv'), with_robot=False,
                           renders=True, render_video=False, rollout=rollout, extra_config=cfg)
        self.part = self.bpw._urdf_cache[self.env._part_id]
        if cfg['COLOR_MODE'] != 'RGB':
            # S5: plain Python ints so `texels[t] -= q` may go negative like under NumPy 1.x
            self.part.texels = [int(v) for v in self.part.texels]
            self.part.init_texture = list(self.part.texels)

    def _silence(self):
        return contextlib.redirect_stdout(io.StringIO()) if self.quiet else contextlib.nullcontext()

    def reset(self, start_index=None):
        """reset(); with `start_index` the second randint of robot_gym_env.py:381 is forced."""
        self._forced_start = start_index
        with self._silence():
            obs = self.env.reset()
        self._forced_start = None
        return obs

    def step(self, action):
        with self._silence():
            return self.env.step(action)

    # ---- state probes -------------------------------------------------------------------
    def front_status(self):
        """First-channel value of every front texel, in `part.profile[front]` order."""
        part = self.part
        return np.array([int(part.texels[part.get_texel(i, j)]) for (i, j) in part.profile[part.side]],
                        dtype=np.int64)

    def pose(self):
        return (np.array(self.env.robot._pose, dtype=np.float64),
                np.array(self.env.robot._orn, dtype=np.float64))
