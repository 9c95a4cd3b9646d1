"""Shim S2 -- the sliver of `gym` PaintRL touches (robot_gym_env.py:8-10,120-173,420-422;
video_renderer.py:14).  TEST INFRASTRUCTURE ONLY: lets the verbatim reference import here."""
from . import spaces, error, logger, utils  # noqa: F401


class Env(object):
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        return None

    def seed(self, seed=None):
        return None

    @property
    def unwrapped(self):
        return self
