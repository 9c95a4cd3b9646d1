"""CPU restatement of the reference's grid-world ParamTestEnv.  TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module.  It restates PaintRLEnv/param_test_env.py line for line on flat
NumPy tables (no dicts), one environment per object, and is pinned by tests/test_param_oracle.py against
traces minted from the reference's own module (oracle/make_param_golden.py ->
tests/golden/p_param_test_env.npz).  Pure-Python loops: the grid is at most a few hundred cells.
"""
import numpy as np

OBS_MODES = ('section', 'simple', 'direct', 'grid')


def obs_dim(mode, size):
    return {'section': 6, 'simple': 2, 'direct': size * size + 2, 'grid': 102}[mode]


class ParamOracle(object):
    def __init__(self, size, max_len=900, termination_by_repeat=False, obs_mode='section'):
        self.size = int(size)
        self.episode_max_length = max(int(max_len), (self.size - 2) ** 2)        # param_test_env.py:112
        self.repeat_termination = bool(termination_by_repeat)
        self.obs_mode = obs_mode
        s = self.size
        self.init_world = np.zeros((s, s), dtype=np.int32)                        # :122-130
        self.init_world[1:s - 1, 1:s - 1] = 1
        self.init_reward_counter = int(self.init_world.sum())
        self.reset()

    def reset(self):                                                              # :150-160
        self.i = self.j = 1
        self.visit = np.zeros((self.size, self.size), dtype=np.int32)
        self.visit[1, 1] += 1
        self.violated_wall = False
        self.repeat_visit = False
        self.reward_counter = self.init_reward_counter
        self.step_counter = 0
        self.world = self.init_world.copy()
        return self.observation()

    def _immediate(self):                                                         # :206-211
        if self.world[self.i, self.j] > 0:
            self.world[self.i, self.j] -= 1
            self.reward_counter -= 1
            return 1
        return 0

    def step(self, action):                                                       # :218-240, 162-184
        immediate = self._immediate()
        self.step_counter += 1
        if action == 0:
            self.i += 1
        elif action == 1:
            self.j += 1
        elif action == 2:
            self.i -= 1
        elif action == 3:
            self.j -= 1
        else:
            raise IndexError('No such action!')
        s = self.size
        if self.i < 0 or self.i >= s or self.j < 0 or self.j >= s:
            self.i = min(max(self.i, 0), s - 1)
            self.j = min(max(self.j, 0), s - 1)
            self.violated_wall = True
        else:
            self.visit[self.i, self.j] += 1
            if self.visit[self.i, self.j] > 1:
                self.repeat_visit = True
        reward = 0 if self.violated_wall else self._immediate()                   # :213-216
        reward += immediate
        penalty = 0.2
        done = (self.violated_wall or self.reward_counter <= 0 or
                self.step_counter >= self.episode_max_length - 1 or
                (self.repeat_visit and self.repeat_termination))                  # :192-197
        return self.observation(), reward - penalty, done, {'reward': reward, 'penalty': penalty}

    def observation(self):                                                        # :199-204
        s = self.size
        if self.obs_mode == 'section':                                            # :66-93
            cnt = [0, 0, 0, 0]
            mx = [0, 0, 0, 0]
            x, y = self.i, self.j
            for i in range(s):
                for j in range(s):
                    k = -1
                    if 0 < i <= x:
                        if 0 < j <= y:
                            k = 0
                        elif y < j < s - 1:
                            k = 1
                    elif x < i < s - 1:
                        if 0 < j <= y:
                            k = 2
                        elif y < j < s - 1:
                            k = 3
                    if k >= 0:
                        cnt[k] += int(self.world[i, j])
                        mx[k] += 1
            obs = [0 if mx[k] == 0 else cnt[k] / mx[k] for k in range(4)]
        elif self.obs_mode == 'simple':                                           # :18-21
            obs = []
        elif self.obs_mode == 'direct':                                           # :24-30
            obs = list(self.world.astype(np.float64).reshape(-1))
        elif self.obs_mode == 'grid':                                             # :50-63
            max_counter = int(self.init_reward_counter / 100)
            g = np.zeros((10, 10), dtype=np.float64)
            for i in range(s):
                for j in range(s):
                    if i in (0, s - 1) or j in (0, s - 1):
                        continue
                    g[int(i / 2 + 0.5) - 1][int(j / 2 + 0.5) - 1] += self.world[i, j] / max_counter
            obs = list(g.reshape(-1))
        else:
            raise ValueError(self.obs_mode)
        return np.append(np.asarray(obs, dtype=np.float64), [self.i / s, self.j / s])
