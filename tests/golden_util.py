"""Helpers shared by the parity tests: load a golden trace (tests/golden/*.npz, minted by
oracle/make_golden.py from the verbatim reference) and rebuild its configuration."""
import glob
import json
import os

import numpy as np

from paintrl_b200.config import EnvConfig
from paintrl_b200.partpack import PartPack

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, 'g*.npz')))


class Golden(object):
    def __init__(self, name):
        self.name = name
        with np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False) as z:
            self.data = {k: z[k] for k in z.files}
        self.config = json.loads(str(self.data.pop('config')))
        extra = self.config['extra_config']
        self.cfg = EnvConfig(extra, action_mode=self.config.get('action_mode', 'discrete'),
                             action_shape=self.config.get('action_shape', 1),
                             discrete_granularity=self.config.get('discrete_granularity', 4),
                             obs_mode=self.config['obs_mode'], obs_grad=self.config['obs_grad'],
                             paint_method=self.config.get('paint_method', 'fast'), beam_plain=self.data.get('beam_plain'))
        self.rollout = bool(self.config.get('rollout', False))
        self.pack = PartPack.for_part(extra['Part_NO'])
        self.lengths = self.data['lengths']
        self.n_episodes = len(self.lengths)

    def start_index(self, e):
        return 0 if self.rollout else int(self.data['start_index'][e])

    def actions(self, e, t):
        a = self.data['actions'][e, t]
        if self.cfg.action_mode == 'discrete':
            return int(a[0])
        return a[:self.cfg.action_dim].copy()

    def __getitem__(self, key):
        return self.data[key]
