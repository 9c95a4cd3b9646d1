"""Environment configuration: PaintGymEnv's class attributes + `extra_config`.

Mirrors PaintRLEnv/robot_gym_env.py:126-157 (defaults) and :240-252 (`_setup_extra_config`
reads every key with `[]`, so a missing key is a KeyError -- same here).
"""
import ctypes
import math

import numpy as np

from . import _capi

PAINT_RADIUS = 0.051          # bullet_paint_wrapper.py:42
STEP_SIZE = PAINT_RADIUS      # bullet_paint_wrapper.py:43

# robot_gym_env.py:134-157
DEFAULT_EXTRA_CONFIG = {
    'RENDER_HEIGHT': 720,
    'RENDER_WIDTH': 960,
    'Part_NO': 0,
    'Expected_Episode_Length': 245,
    'EPISODE_MAX_LENGTH': 245,
    'TERMINATION_MODE': 'late',
    'SWITCH_THRESHOLD': 0.9,
    'START_POINT_MODE': 'anchor',
    'TURNING_PENALTY': False,
    'OVERLAP_PENALTY': False,
    'COLOR_MODE': 'RGB',
}

_OBS_MODES = {'section': 0, 'grid': 1, 'simple': 2, 'discrete': 3}
_TERM_MODES = {'late': 0, 'early': 1, 'hybrid': 2}


def obs_dim(obs_mode, obs_grad):
    """observation_space.shape[0] (robot_gym_env.py:166-173)."""
    if obs_mode == 'section':
        return obs_grad + 2
    if obs_mode == 'grid':
        return obs_grad ** 2
    if obs_mode == 'simple':
        return 2
    return obs_grad + 1


def direction_normalize(action):
    """robot.py:151-160 with NumPy, exactly as the reference evaluates it."""
    if len(action) == 1:
        phi = (action[0] + 1) * np.pi
        return 1 * np.cos(phi), 1 * np.sin(phi)
    rho, phi = np.sqrt(action[0] ** 2 + action[1] ** 2), np.arctan2(action[1], action[0])
    x, y = abs(action[0]), abs(action[1])
    if x == 0 and y == 0:
        return x, y
    m = max(x, y)
    return m * np.cos(phi), m * np.sin(phi)


def turning_angle(delta_axis1, delta_axis2):
    """robot.py:352-358."""
    if delta_axis1 != 0:
        return math.atan(abs(delta_axis2 / delta_axis1))
    return math.pi / 2


def discrete_table(n):
    """(u1, u2, turning angle) for each discrete action a in 0..n (n + 1 rows).

    robot_gym_env.py:342-347 (`[2 * (a - n / 2) / n]`), robot.py:390-397 (clip, normalise, scale)
    and robot.py:352-358, evaluated on the host with the same NumPy / libm calls as the
    reference so that discrete directions carry the reference's own 1e-16 residues.  Actions outside
    0..n-1 are not rejected by the reference but clipped (robot.py:390-393): a < 0 acts like 0
    (both give -1) and a >= n like the extra row n (clipped to +1).
    """
    table = np.zeros((n + 1, 3), dtype=np.float64)
    for a in range(n + 1):
        act = a - n / 2
        act = [2 * act / n]
        act = [min(1, max(-1, v)) for v in act]
        u1, u2 = direction_normalize(act)
        table[a] = (u1, u2, turning_angle(u1 * STEP_SIZE, u2 * STEP_SIZE))
    return table


def uniform_paint_plain(point_density):
    """Robot._paint_plain for RGB: `_get_uniformed_plain(density)` (robot.py:14-36) -- the grid points of pitch
    1.8 / sqrt(density) inside a disc of radius 0.1 at z = 0.2 in the TCP frame.  The two running coordinates are
    accumulated with repeated float additions exactly like the reference's while loops, so the table is bit-identical."""
    projection_distance = 0.2
    ratio = projection_distance / 0.5
    radius = 0.25 * ratio
    resolution = 1.8 / math.sqrt(point_density)
    points = []
    i = -radius
    while i <= radius:
        j = -radius
        while j <= radius:
            if math.sqrt(math.pow(i, 2) + math.pow(j, 2)) <= radius:
                points.append((i, j, projection_distance))
            j += resolution
        i += resolution
    return np.array(points, dtype=np.float64).reshape(-1, 3)


def beta_paint_plain(beta, point_density, rng=None):
    """Robot._paint_plain for HSI: `_get_beta_plain(beta, density)` (robot.py:39-69) -- rings of width
    1.8 / sqrt(density) up to radius 0.1, ring i holding round(450 * w_i / sum w) points with
    w_i = (1 - (i / circles)^2)^(beta - 1), each at a radius drawn uniformly inside its ring (the reference draws from
    the process-global `random.uniform`; pass `rng` -- anything with `.uniform(a, b)` -- for a reproducible table)."""
    import random as _random
    rng = rng or _random
    radius = 0.25 * (0.2 / 0.5)
    resolution = 1.8 / math.sqrt(point_density)
    circles = math.ceil(radius / resolution)
    weights = {i: (1 - (i / circles) ** 2) ** (beta - 1) for i in range(1, circles + 1)}
    total = sum(weights.values())
    counts = {i: round(450 * w / total) for i, w in weights.items()}
    points = []
    for i in range(1, circles + 1):
        lower, upper = (i - 1) * resolution, i * resolution
        step = 2 * math.pi / counts[i] if counts[i] else 0
        for j in range(counts[i]):
            r = rng.uniform(lower, upper)
            theta = j * step
            points.append((r * np.cos(theta), r * np.sin(theta), 0.2))
    return np.array(points, dtype=np.float64).reshape(-1, 3)


class EnvConfig(object):
    """One immutable configuration of the batched environment."""

    def __init__(self, extra_config=None, action_mode='discrete', action_shape=1,
                 discrete_granularity=4, obs_mode='section', obs_grad=4, auto_reset=False, seed=0,
                 max_possible_point=None, paint_method='fast', beam_plain=None):
        cfg = DEFAULT_EXTRA_CONFIG if extra_config is None else extra_config
        # every key is required, like robot_gym_env.py:240-252
        self.render_width = cfg['RENDER_WIDTH']
        self.render_height = cfg['RENDER_HEIGHT']
        self.part_no = cfg['Part_NO']
        self.expected_episode_length = cfg['Expected_Episode_Length']
        self.episode_max_length = cfg['EPISODE_MAX_LENGTH']
        self.termination_mode = cfg['TERMINATION_MODE']
        self.switch_threshold = cfg['SWITCH_THRESHOLD']
        self.start_point_mode = cfg['START_POINT_MODE']
        self.turning_penalty = bool(cfg['TURNING_PENALTY'])
        self.overlap_penalty = bool(cfg['OVERLAP_PENALTY'])
        self.color_mode = cfg['COLOR_MODE']
        if action_mode not in ('discrete', 'continuous'):
            raise ValueError('ACTION_MODE %r' % (action_mode,))
        if action_mode == 'continuous' and action_shape not in (1, 2):
            raise ValueError('ACTION_SHAPE %r' % (action_shape,))
        if obs_mode not in _OBS_MODES:
            # robot_gym_env.py:172-173 treats every other value like 'discrete' for the space but
            # _augmented_observation (:306-319) then falls through to section + pose; refuse.
            raise ValueError('OBS_MODE %r' % (obs_mode,))
        self.action_mode = action_mode
        self.action_shape = action_shape if action_mode == 'continuous' else 1
        self.discrete_granularity = int(discrete_granularity)
        self.obs_mode = obs_mode
        self.obs_grad = int(obs_grad)
        self.auto_reset = bool(auto_reset)
        self.seed = int(seed)
        self.max_possible_point = max_possible_point
        # Robot.PAINT_METHOD (robot.py:172): 'fast' (ball query per shot) or 'normal' (beam fan per shot).  `beam_plain`
        # [n, 3] is Robot._paint_plain; None derives it from the part's texel density at engine construction
        # (robot_gym_env.py:284-285: uniform disc for RGB, beta rings -- seeded with `seed` -- for HSI)
        if paint_method not in ('fast', 'normal'):
            raise ValueError('PAINT_METHOD %r' % (paint_method,))
        self.paint_method = paint_method
        self.beam_plain = None if beam_plain is None else np.ascontiguousarray(beam_plain, dtype=np.float64).reshape(-1, 3)
        self.extra_config = dict(cfg)

    def paint_plain(self, density):
        """Robot.set_up_paint_params (robot.py:244-249): the beam table for this colour mode."""
        if self.beam_plain is not None:
            return self.beam_plain
        if self.color_mode == 'RGB':
            return uniform_paint_plain(density)
        import random as _random
        return beta_paint_plain(2, density, rng=_random.Random(self.seed))

    @property
    def obs_dim(self):
        return obs_dim(self.obs_mode, self.obs_grad)

    @property
    def action_dim(self):
        return self.action_shape if self.action_mode == 'continuous' else 1

    def to_c(self, max_possible_point, density=None):
        """Build the `PaintrlConfig` struct; returns (struct, keepalive).  `density`: the part's texel density
        (Part.get_side_density, bullet_paint_wrapper.py:830-839), needed to derive the beam table in normal mode."""
        c = _capi.PaintrlConfig()
        c.abi_version = _capi.PAINTRL_ABI_VERSION
        c.action_mode = 0 if self.action_mode == 'discrete' else 1
        c.action_shape = self.action_shape
        c.discrete_granularity = self.discrete_granularity
        table = discrete_table(self.discrete_granularity)
        c.discrete_table = table.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        c.obs_mode = _OBS_MODES[self.obs_mode]
        c.obs_grad = self.obs_grad
        c.color_mode = 0 if self.color_mode == 'RGB' else 1
        c.termination_mode = _TERM_MODES[self.termination_mode]
        c.switch_threshold = float(self.switch_threshold)
        c.expected_episode_length = int(self.expected_episode_length)
        c.episode_max_length = int(self.episode_max_length)
        c.turning_penalty = int(self.turning_penalty)
        c.overlap_penalty = int(self.overlap_penalty)
        mpp = self.max_possible_point if self.max_possible_point is not None else max_possible_point
        c.max_possible_point = float(mpp)
        c.auto_reset = int(self.auto_reset)
        c.seed = self.seed
        keep = [table]
        c.paint_method = 0
        c.n_beams = 0
        c.beam_plain = None
        if self.paint_method == 'normal':
            if self.beam_plain is None and not density:
                raise ValueError('PAINT_METHOD normal needs a beam table or the part\'s texel density')
            plain = np.ascontiguousarray(self.paint_plain(density), dtype=np.float64)
            c.paint_method = 1
            c.n_beams = int(plain.shape[0])
            c.beam_plain = plain.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            keep.append(plain)
        return c, keep
