"""The reference's grid-world `ParamTestEnv` (PaintRLEnv/param_test_env.py) over the batched CUDA engine.

`BatchedParamTestEnv` -- thousands of independent size x size worlds on one GPU (C ABI `paintrl_param_*`,
                         kernels in csrc/paintrl_param.cuh), device tensors in and out.
`ParamTestEnv`        -- drop-in for param_test_env.py:96-246 (gym.Env, one world): same constructor, class
                         attributes, `reset` / `step` conventions (`np.append(obs, [i/size, j/size])`,
                         `info = {'reward', 'penalty'}`), `get_current_pos`, `world` / `visit_table` dicts,
                         the non-train-mode step log and end-of-episode tables, and the module's own
                         `zigzag()` / `spiral()` drivers (param_test_env.py:283-342).
There is no CPU fallback: constructing either without a CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from .gym_env import _GymEnv, seeding, spaces

OBS_MODES = {'section': 0, 'simple': 1, 'direct': 2, 'grid': 3}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedParamTestEnv(object):
    def __init__(self, num_envs, size, max_len=900, termination_by_repeat=False, obs_mode='section',
                 auto_reset=False, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('paintrl_b200 needs a CUDA device: there is no CPU fallback')
        if obs_mode not in OBS_MODES:
            obs_mode = 'simple'                 # param_test_env.py:141-142: anything else observes nothing
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.num_envs, self.size, self.obs_mode = int(num_envs), int(size), obs_mode
        self.auto_reset = bool(auto_reset)
        self._lib = _capi.lib()
        cfg = _capi.PaintrlParamConfig(_capi.PAINTRL_ABI_VERSION, self.size, int(max_len), int(bool(termination_by_repeat)),
                                       OBS_MODES[obs_mode], int(self.auto_reset))
        handle = ctypes.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(index):
            _capi.check(self._lib.paintrl_param_create(ctypes.byref(cfg), self.num_envs, index, ctypes.byref(handle)))
        self._h = handle
        self.obs_dim = int(self._lib.paintrl_param_obs_dim(self._h))
        B, f64, dev = self.num_envs, torch.float64, self.device
        self.obs = torch.zeros(B, self.obs_dim, dtype=f64, device=dev)
        self.next_obs = torch.zeros(B, self.obs_dim, dtype=f64, device=dev)
        self.reward = torch.zeros(B, dtype=f64, device=dev)
        self.penalty = torch.zeros(B, dtype=f64, device=dev)
        self.actual = torch.zeros(B, dtype=f64, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.paintrl_param_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ids(self, env_ids):
        """Validated device list of world indices (in range, unique), like BatchedPaintEnv._ids."""
        if env_ids is None:
            return None, self.num_envs
        host = np.asarray(env_ids.cpu() if isinstance(env_ids, torch.Tensor) else env_ids).reshape(-1)
        if host.size == 0 or host.dtype.kind not in 'iu':
            raise ValueError('env_ids must be a non-empty list of integers')
        if int(host.min()) < 0 or int(host.max()) >= self.num_envs:
            raise IndexError('env_ids out of range [0, %d)' % self.num_envs)
        if np.unique(host).size != host.size:
            raise ValueError('env_ids must be unique')
        ids = torch.as_tensor(host.astype(np.int32), device=self.device).contiguous()
        return ids, int(ids.numel())

    def reset(self, env_ids=None):
        """ParamTestEnv.reset (param_test_env.py:150-160): first observations [n, obs_dim]."""
        ids, n = self._ids(env_ids)
        out = self.obs if ids is None else torch.zeros(n, self.obs_dim, dtype=torch.float64, device=self.device)
        _capi.check(self._lib.paintrl_param_reset(self._h, _ptr(ids), n, _ptr(out), self._stream()))
        if ids is None:
            self.next_obs.copy_(out)
        return out

    def step(self, actions, check_actions=True):
        """ParamTestEnv.step for every world (param_test_env.py:218-240): (obs, actual_reward, done, info).
        An action outside 0..3 raises IndexError BEFORE anything is stepped, like the reference (:173-174); host
        inputs are checked on the host, a CUDA tensor with one reduction (`check_actions=False` skips that
        synchronisation: an offending world then reports done with zero reward and `stats()['bad_action_seen']`)."""
        if check_actions and not (isinstance(actions, torch.Tensor) and actions.is_cuda):
            host = np.asarray(actions.cpu() if isinstance(actions, torch.Tensor) else actions).reshape(-1)
            if host.size and (int(host.min()) < 0 or int(host.max()) > 3):
                raise IndexError('No such action!')
        a = torch.as_tensor(actions, device=self.device).to(torch.int64).contiguous()
        if a.numel() != self.num_envs:
            raise ValueError('expected %d actions' % self.num_envs)
        if check_actions and isinstance(actions, torch.Tensor) and actions.is_cuda and bool(((a < 0) | (a > 3)).any()):
            raise IndexError('No such action!')
        _capi.check(self._lib.paintrl_param_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.penalty),
                                                 _ptr(self.actual), _ptr(self.done), _ptr(self.next_obs), self._stream()))
        return self.obs, self.actual, self.done, {'reward': self.reward, 'penalty': self.penalty, 'next_obs': self.next_obs}

    def tables(self, env_ids=None):
        """(world, visit_table) as int32 [n, size, size] (param_test_env.py:113-131)."""
        ids, n = self._ids(env_ids)
        w = torch.zeros(n, self.size, self.size, dtype=torch.int32, device=self.device)
        v = torch.zeros_like(w)
        _capi.check(self._lib.paintrl_param_tables(self._h, _ptr(ids), n, _ptr(w), _ptr(v), self._stream()))
        return w, v

    def stats(self):
        vals = [ctypes.c_uint64(0) for _ in range(3)]
        bad = ctypes.c_int32(0)
        _capi.check(self._lib.paintrl_param_stats(self._h, ctypes.byref(vals[0]), ctypes.byref(vals[1]), ctypes.byref(vals[2]),
                                                  ctypes.byref(bad)))
        return {'env_steps': vals[0].value, 'episodes_ended': vals[1].value, 'kernel_launches': vals[2].value,
                'bad_action_seen': bool(bad.value)}


class Visualizer(object):
    """Text dump of a size x size table (the role of param_test_env.py:252-280): interior cells whose value is in
    `marks` get a leading `*` (the reference colours them red through termcolor)."""

    def __init__(self, size):
        self._size = size

    def _dump(self, title, table, marks):
        n = self._size
        print(title)
        print('|'.join('%3s' % c for c in range(n)))
        for i in range(n):
            row = []
            for j in range(n):
                v = table[(i, j)]
                interior = 0 < i < n - 1 and 0 < j < n - 1
                row.append('%3s' % (('*%s' % v) if interior and v in marks else v))
            print('|'.join(row))

    def print_visit_table(self, table):
        self._dump('Visit Table: count of visit in each state', table, set(range(20)) - {1})

    def print_world_table(self, table):
        self._dump('World Table:', table, {1})


class ParamTestEnv(_GymEnv):
    """Drop-in for PaintRLEnv/param_test_env.py:96-246."""
    reward_range = (-1e3, 1e3)
    action_space = spaces.Discrete(4)
    OBS_MODE = 'section'
    observation_space = spaces.Box(low=0.0, high=1.0, shape=(6,), dtype=np.float64)

    def __init__(self, size, max_len=900, train_mode=True, termination_by_repeat=False):
        self.size = size
        self.EPISODE_MAX_LENGTH = max(max_len, (self.size - 2) ** 2)
        self._mode = train_mode
        self.repeat_termination = termination_by_repeat
        self.init_reward_counter = (self.size - 2) ** 2
        self.ACTION_DEF = {0: '\U0001f806', 1: '\U0001f805', 2: '\U0001f804', 3: '\U0001f807'}
        self._engine = BatchedParamTestEnv(1, size, max_len, termination_by_repeat, self.OBS_MODE)
        self._visualizer = Visualizer(self.size)
        self._i = self._j = 1
        self._step_counter = 0

    # ---- tables as the reference exposes them
    def _dict(self, which):
        t = self._engine.tables()[which][0].cpu().numpy()
        return {(i, j): int(t[i, j]) for i in range(self.size) for j in range(self.size)}

    @property
    def world(self):
        return self._dict(0)

    @property
    def visit_table(self):
        return self._dict(1)

    def get_current_pos(self):
        return self._i, self._j

    def _track(self, obs):
        self._i, self._j = int(round(obs[-2] * self.size)), int(round(obs[-1] * self.size))

    def reset(self):
        obs = self._engine.reset()[0].cpu().numpy().copy()
        self._step_counter = 0
        self._track(obs)
        return obs

    def step(self, action):
        if action not in (0, 1, 2, 3):
            raise IndexError('No such action!')                       # param_test_env.py:173
        o, actual, done, info = self._engine.step([int(action)])
        host = torch.cat([o[0], actual, info['reward'], info['penalty'], done.to(torch.float64)]).cpu().numpy()
        n = self._engine.obs_dim
        observation = host[:n].copy()
        actual_reward, reward, penalty, done = float(host[n]), host[n + 1], float(host[n + 2]), bool(host[n + 3])
        reward = int(reward)
        self._step_counter += 1
        self._track(observation)
        if not self._mode:                                             # param_test_env.py:228-239
            x, y = round(observation[-2] * self.size), round(observation[-1] * self.size)
            print('STEP: {0} ACTION: {1} OBS: [{2}, {3}], REWARD: {4}'.format(self._step_counter, self.ACTION_DEF[action],
                                                                              int(x), int(y), actual_reward))
            if done:
                self._visualizer.print_world_table(self.world)
                self._visualizer.print_visit_table(self.visit_table)
        return observation, actual_reward, done, {'reward': reward, 'penalty': penalty}

    def render(self, mode='human'):
        pass

    def close(self):
        if getattr(self, '_engine', None) is not None:
            self._engine.close()
            self._engine = None

    def seed(self, seed=None):
        _, seed = seeding.np_random(seed)
        return seed


def zigzag_actions(grid_size):
    """The column sweep of param_test_env.py:283-314 as a generator: send it the latest observation, receive the
    next action (1 = +j until the last interior row, one step of 0 = +i, then 3 = -j back down, ...)."""
    heading, obs = 1, (yield None)
    while True:
        row = round(grid_size * obs[-1]) % grid_size
        at_end = row == (grid_size - 2 if heading == 1 else 1)
        if at_end:
            obs = yield 0                      # one column over, then turn around
            heading = 3 if heading == 1 else 1
        else:
            obs = yield heading


def spiral_actions(grid_size):
    """The inward spiral of param_test_env.py:317-342 as a generator of actions (it ignores the observation):
    three legs of grid_size - 3 steps, then legs shrinking by one every second turn."""
    direction, leg, legs_left = 0, grid_size - 3, 3
    yield None
    while leg > 0:
        for _ in range(leg):
            yield direction % 4
        direction += 1
        legs_left -= 1
        if legs_left <= 0:
            legs_left, leg = 2, leg - 1
    # the legs are used up: the reference's loop (param_test_env.py:326-340) never sees `current_counter == 0` again
    # and keeps stepping in the direction it has just turned to until the episode ends (at a wall)
    while True:
        yield direction % 4


def drive(env, actions):
    """Run one episode of `env` under an action generator; returns (steps, total return)."""
    obs = env.reset()
    next(actions)
    steps, total, done = 0, 0, False
    while not done:
        obs, reward, done, _ = env.step(actions.send(obs))
        steps += 1
        total += reward
    print('In {0} steps get {1} rewards'.format(steps, total))
    return steps, total


def zigzag(grid_size=22, env=None):
    """`zigzag()` of the reference module: sweep the grid column by column."""
    return drive(env or ParamTestEnv(grid_size, train_mode=False), zigzag_actions(grid_size))


def spiral(grid_size=22, env=None):
    """`spiral()` of the reference module: spiral inwards."""
    return drive(env or ParamTestEnv(grid_size, train_mode=False), spiral_actions(grid_size))
