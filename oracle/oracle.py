"""Python face of the C restatement (oracle/paint_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product (paintrl_b200/) never does.

`OracleBatch` holds N independent single-environment oracles that share one part pack and one
configuration, and mirrors the call sequence of the reference:

    reset(start_index)  PaintGymEnv.reset          robot_gym_env.py:370-387
    step(actions)       PaintGymEnv.step           robot_gym_env.py:349-368
    set_pose(pos, n)    Robot.reset(pose)          robot.py:366-372 (spiral.py:28-38)

The action -> unit-direction part of the step (robot_gym_env.py:342-347, robot.py:390-395,
151-160) is evaluated here with NumPy, exactly as the reference does, and handed to C.
"""
import ctypes
import os
import subprocess

import numpy as np

from paintrl_b200 import _capi
from paintrl_b200.config import EnvConfig, direction_normalize

_DIR = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_DIR, 'paint_oracle.c')
LIB = os.path.join(_DIR, '_build', 'libpaint_oracle.so')


def build(force=False):
    """gcc the C restatement into oracle/_build/ (strict IEEE: no FMA contraction)."""
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-std=gnu11', '-ffp-contract=off', '-fno-fast-math',
           '-fopenmp', SRC, '-o', LIB, '-lm']
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        h = ctypes.CDLL(LIB)
        vp, dp, i32 = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int32
        h.oracle_create.restype = vp
        h.oracle_create.argtypes = [ctypes.POINTER(_capi.PaintrlPartPack), ctypes.POINTER(_capi.PaintrlConfig)]
        h.oracle_destroy.argtypes = [vp]
        h.oracle_obs_dim.restype = i32
        h.oracle_obs_dim.argtypes = [vp]
        h.oracle_grid_cells.restype = ctypes.POINTER(i32)
        h.oracle_grid_cells.argtypes = [vp]
        h.oracle_set_pose.argtypes = [vp, dp, dp, dp]
        h.oracle_reset.argtypes = [vp, i32, dp]
        h.oracle_step.argtypes = [vp, ctypes.c_double, ctypes.c_double, dp, dp, dp, dp,
                                  ctypes.POINTER(ctypes.c_uint8)]
        h.oracle_reset_batch.argtypes = [vp, i32, vp, vp, i32]
        h.oracle_step_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32]
        h.oracle_status.restype = ctypes.POINTER(ctypes.c_int16)
        h.oracle_status.argtypes = [vp]
        h.oracle_get_pose.argtypes = [vp, dp, dp]
        h.oracle_get_scalars.argtypes = [vp, dp]
        h.oracle_set_state.argtypes = [vp, vp, vp, vp, vp]
        h.oracle_ray_test.restype = i32
        h.oracle_ray_test.argtypes = [vp, dp, dp, dp]
        h.oracle_ball_query.restype = i32
        h.oracle_ball_query.argtypes = [vp, dp, vp]
        h.oracle_nearest_vertex.restype = i32
        h.oracle_nearest_vertex.argtypes = [vp, dp]
        h.oracle_rasterize.restype = i32
        h.oracle_rasterize.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp]
        _lib = h
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def action_directions(cfg, actions):
    """[n, 2] unit directions for a batch of raw actions, by the reference's own formulas."""
    n = len(actions)
    out = np.zeros((n, 2), dtype=np.float64)
    for i in range(n):
        if cfg.action_mode == 'continuous':
            act = list(np.atleast_1d(actions[i]))               # robot_gym_env.py:343-344
        else:
            a = actions[i] - cfg.discrete_granularity / 2       # robot_gym_env.py:346-347
            act = [2 * a / cfg.discrete_granularity]
        for k, a in enumerate(act):                             # robot.py:390-393
            if not -1 <= a <= 1:
                act[k] = -1 if a < -1 else 1
        out[i] = direction_normalize(act)                       # robot.py:395
    return out


class OracleBatch(object):
    def __init__(self, pack, cfg, num_envs, threads=None):
        assert isinstance(cfg, EnvConfig)
        self.pack, self.cfg, self.num_envs = pack, cfg, int(num_envs)
        self.threads = int(threads or os.cpu_count() or 1)
        self._h = lib()
        self._cpack, self._keep_pack = pack.to_c(cfg.start_point_mode, cfg.color_mode, with_nn_rep=cfg.paint_method == 'normal')
        self._ccfg, self._keep_cfg = cfg.to_c(pack.max_points, density=pack.meta.get('density'))
        self.envs = []
        for _ in range(self.num_envs):
            e = self._h.oracle_create(ctypes.byref(self._cpack), ctypes.byref(self._ccfg))
            if not e:
                raise RuntimeError('oracle_create failed (unsupported observation layout)')
            self.envs.append(e)
        self._env_array = (ctypes.c_void_p * self.num_envs)(*self.envs)
        self.obs_dim = int(self._h.oracle_obs_dim(self.envs[0]))
        self.n_texels = pack.n_texels

    def close(self):
        for e in self.envs:
            self._h.oracle_destroy(e)
        self.envs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference call mirror
    def reset(self, start_index, env_ids=None):
        ids = range(self.num_envs) if env_ids is None else env_ids
        start_index = np.broadcast_to(np.asarray(start_index, dtype=np.int32), (len(ids),))
        obs = np.zeros((len(ids), self.obs_dim), dtype=np.float64)
        for k, i in enumerate(ids):
            self._h.oracle_reset(self.envs[i], int(start_index[k]), _dp(obs[k]))
        return obs

    def set_pose(self, pos, normal, env_ids=None):
        ids = range(self.num_envs) if env_ids is None else env_ids
        pos = np.ascontiguousarray(np.broadcast_to(pos, (len(ids), 3)), dtype=np.float64)
        normal = np.ascontiguousarray(np.broadcast_to(normal, (len(ids), 3)), dtype=np.float64)
        obs = np.zeros((len(ids), self.obs_dim), dtype=np.float64)
        for k, i in enumerate(ids):
            self._h.oracle_set_pose(self.envs[i], _dp(pos[k]), _dp(normal[k]), _dp(obs[k]))
        return obs

    def step(self, actions):
        dirs = np.ascontiguousarray(action_directions(self.cfg, actions))
        return self.step_directions(dirs)

    def step_directions(self, dirs):
        n = self.num_envs
        obs = np.zeros((n, self.obs_dim), dtype=np.float64)
        reward = np.zeros(n)
        penalty = np.zeros(n)
        actual = np.zeros(n)
        done = np.zeros(n, dtype=np.uint8)
        self._h.oracle_step_batch(self._env_array, n, dirs.ctypes.data, obs.ctypes.data,
                                  reward.ctypes.data, penalty.ctypes.data, actual.ctypes.data,
                                  done.ctypes.data, self.threads)
        return obs, reward, penalty, actual, done

    # ------------------------------------------------------------------ probes
    def status(self, i=None):
        if i is None:
            return np.stack([self.status(k) for k in range(self.num_envs)])
        ptr = self._h.oracle_status(self.envs[i])
        return np.ctypeslib.as_array(ptr, shape=(self.n_texels,)).copy()

    def pose(self, i):
        pos, quat = np.zeros(3), np.zeros(4)
        self._h.oracle_get_pose(self.envs[i], _dp(pos), _dp(quat))
        return pos, quat

    def scalars(self, i):
        out = np.zeros(12)
        self._h.oracle_get_scalars(self.envs[i], _dp(out))
        keys = ('total_reward', 'total_return', 'step_counter', 'term_counter', 'last_on_part',
                'terminate', 'last_angle', 'angle_diff', 'rate', 'succeeded', 'pixel_counter',
                'anomalies')
        return dict(zip(keys, out))

    def set_state(self, i, status=None, pose=None, quat=None, scalars=None):
        def ptr(a, dt):
            return None if a is None else np.ascontiguousarray(a, dtype=dt).ctypes.data
        self._h.oracle_set_state(self.envs[i], ptr(status, np.int16), ptr(pose, np.float64),
                                 ptr(quat, np.float64), ptr(scalars, np.float64))

    def grid_cells(self):
        ptr = self._h.oracle_grid_cells(self.envs[0])
        return np.ctypeslib.as_array(ptr, shape=(self.n_texels,)).copy()

    def ray_test(self, frm, to):
        frm, to, hit = (np.ascontiguousarray(v, dtype=np.float64) for v in (frm, to, np.zeros(3)))
        ok = self._h.oracle_ray_test(self.envs[0], _dp(frm), _dp(to), _dp(hit))
        return (hit if ok else None)

    def ball_query(self, center):
        center = np.ascontiguousarray(center, dtype=np.float64)
        mask = np.zeros(self.n_texels, dtype=np.uint8)
        self._h.oracle_ball_query(self.envs[0], _dp(center), mask.ctypes.data)
        return np.flatnonzero(mask)

    def nearest_vertex(self, point):
        point = np.ascontiguousarray(point, dtype=np.float64)
        return int(self._h.oracle_nearest_vertex(self.envs[0], _dp(point)))


def rasterize(tri_a, tri_b, tri_c, tri_uv, width, height):
    """Front texels of a part at a texture size, by the reference's rasterisation rule
    (bullet_paint_wrapper.py:191-212, 604-618; oracle_rasterize in paint_oracle.c).
    Returns (ij [N,2] int32 sorted by (i, j), pos [N,3] float64, owner [N] int32 = tri * 4 + kind)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (tri_a, tri_b, tri_c, tri_uv)]
    n_tris = arrs[0].shape[0]
    owner = np.empty(width * height, dtype=np.int32)
    pos = np.zeros((width * height, 3), dtype=np.float64)
    n = lib().oracle_rasterize(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data,
                               n_tris, width, height, owner.ctypes.data, pos.ctypes.data)
    if n < 0:
        raise ValueError('a UV coordinate maps outside the %dx%d texture' % (width, height))
    idx = np.flatnonzero(owner >= 0)
    assert len(idx) == n
    ij = np.stack([idx // height, idx % height], axis=1).astype(np.int32)
    return ij, pos[idx], owner[idx]


def retextured_pack(base_pack, width, height):
    """`PartPack.retextured` with the texels rasterised by this oracle instead of the GPU."""
    a = base_pack.arrays
    ij, pos, _ = rasterize(a['tri_a'], a['tri_b'], a['tri_c'], a['tri_uv'], width, height)
    return base_pack.retextured(width, height, texels=(ij, pos))
