"""Replay of a sampled SUBSET of a large batch through the C oracle.  TEST INFRASTRUCTURE ONLY.

Environments are independent, so a run of 4096 ... 65536 environments on the
GPU can be checked at full size by logging the per-step results of a few hundred of its environment
indices on the device (`SubsetRecorder`, a handful of index_select calls outside any timed region)
and stepping the C restatement (oracle/paint_oracle.c) through the same actions, start points and
auto-resets afterwards (`replay_subset`).  Used by tests/test_gpu_baseline_sizes.py and by bench.py's
`parity_check` (the checker, never the thing measured).

Reference semantics replayed: PaintGymEnv.step / reset, robot_gym_env.py:349-387.
"""
import numpy as np

from .oracle import OracleBatch

_M64 = (1 << 64) - 1


def splitmix64(x):
    """The engine's start-index stream (paintrl_device.cuh splitmix64), on Python ints."""
    x = (int(x) + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def auto_start_index(seed, env, episode, n_starts):
    """paintrl_kernels.cuh auto_start_index: the start point the engine draws when environment `env`
    auto-resets for the `episode`-th time (no counterpart in the reference, which draws from the
    process-global `random`, robot_gym_env.py:381)."""
    seed, env, episode = int(seed), int(env), int(episode)
    return int(splitmix64((seed & _M64) ^ splitmix64(((env & 0xffffffff) << 32) | (episode & 0xffffffff))) % int(n_starts))


def sample_env_ids(num_envs, count, seed=0, tail=64):
    """A spread of environment indices: an even stride over the batch, the last `tail` indices (the
    final, partial wave of both kernels) and a few random ones."""
    count = min(count, num_envs)
    tail = min(tail, count // 4)
    ids = set(np.linspace(0, num_envs - 1, num=max(1, count // 2), dtype=np.int64).tolist())
    ids.update(range(num_envs - tail, num_envs))
    rng = np.random.default_rng(seed)
    while len(ids) < count:
        ids.add(int(rng.integers(0, num_envs)))
    return np.array(sorted(ids), dtype=np.int64)


class SubsetRecorder(object):
    """Device-side log of the listed environments of a `BatchedPaintEnv` run."""

    def __init__(self, env, env_ids, capacity):
        import torch
        self.env = env
        self.env_ids = np.asarray(env_ids, dtype=np.int64)
        self.ids = torch.as_tensor(self.env_ids, device=env.device)
        n, od, dev, f64 = len(self.env_ids), env.obs_dim, env.device, torch.float64
        self.discrete = env.cfg.action_mode == 'discrete'
        self.actions = torch.zeros((capacity, n) if self.discrete else (capacity, n, env.action_dim),
                                   dtype=torch.int64 if self.discrete else f64, device=dev)
        self.obs = torch.zeros(capacity, n, od, dtype=f64, device=dev)
        self.next_obs = torch.zeros(capacity, n, od, dtype=f64, device=dev)
        self.reward = torch.zeros(capacity, n, dtype=f64, device=dev)
        self.penalty = torch.zeros(capacity, n, dtype=f64, device=dev)
        self.actual = torch.zeros(capacity, n, dtype=f64, device=dev)
        self.done = torch.zeros(capacity, n, dtype=torch.uint8, device=dev)
        self.status = {}          # step index -> int16 [n, n_texels] (post auto-reset planes)
        self.t = 0

    def record(self, actions, status=False):
        """Call after `env.step(actions)` (actions: the full batch's device tensor)."""
        import torch
        t, env, ids = self.t, self.env, self.ids
        a = actions.reshape(env.num_envs, -1) if not self.discrete else actions.reshape(env.num_envs)
        torch.index_select(a, 0, ids, out=self.actions[t])
        torch.index_select(env.obs, 0, ids, out=self.obs[t])
        torch.index_select(env.next_obs if env.cfg.auto_reset else env.obs, 0, ids, out=self.next_obs[t])
        torch.index_select(env.reward, 0, ids, out=self.reward[t])
        torch.index_select(env.penalty, 0, ids, out=self.penalty[t])
        torch.index_select(env.actual, 0, ids, out=self.actual[t])
        torch.index_select(env.done, 0, ids, out=self.done[t])
        if status:
            self.status[t] = env.get_state(env_ids=self.env_ids.astype(np.int32))['status'].cpu().numpy()
        self.t += 1

    def host(self):
        T = self.t
        return {k: getattr(self, k)[:T].cpu().numpy() for k in ('actions', 'obs', 'next_obs', 'reward', 'penalty', 'actual', 'done')}


def replay_subset(pack, cfg, env_ids, start_index, log, status=None, reset_start_index=None, episode0=1, threads=None):
    """Step the oracle through the logged run of the listed environments and compare.

    env_ids     : [n] indices of the environments within their engine (for the seeded start-index stream)
    start_index : [n] start points of the initial reset
    log         : dict from SubsetRecorder.host(): actions [T, n(, A)], obs / next_obs [T, n, od], reward /
                  penalty / actual [T, n], done [T, n]
    status      : {t: int16 [n, n_texels]} planes read back AFTER step t (auto-resets applied)
    reset_start_index : [T, n] explicit start indices of auto-resets, or None for the engine's seeded stream
    episode0    : the engine's per-environment episode counter after the initial reset (1 for a fresh engine)
    Returns {'envs', 'steps', 'ok', 'exact', 'episodes', 'planes_checked', 'mismatch'}.
    """
    env_ids = np.asarray(env_ids, dtype=np.int64)
    n, T = len(env_ids), log['done'].shape[0]
    ora = OracleBatch(pack, cfg, n, threads=threads)
    n_starts = pack.start_points(cfg.start_point_mode).shape[0]
    exact = cfg.color_mode == 'RGB' and cfg.action_mode == 'discrete'
    tol = dict(rtol=1e-5, atol=1e-12)

    def same(a, b):
        return np.array_equal(a, b) if exact else np.allclose(a, b, **tol)

    out = {'envs': int(n), 'steps': int(T), 'ok': True, 'exact': bool(exact), 'episodes': 0, 'planes_checked': 0, 'mismatch': None}

    def fail(what, t):
        out['ok'] = False
        out['mismatch'] = '%s at step %d' % (what, t)
        ora.close()
        return out

    ora.reset(np.asarray(start_index, dtype=np.int32))
    episode = np.full(n, int(episode0), dtype=np.int64)
    for t in range(T):
        o_obs, o_rew, o_pen, o_act, o_done = ora.step(log['actions'][t])
        if not np.array_equal(o_done, log['done'][t]):
            bad = np.flatnonzero(o_done != log['done'][t])
            return fail('done flags differ (envs %s)' % env_ids[bad][:8].tolist(), t)
        for key, val in (('obs', o_obs), ('reward', o_rew), ('penalty', o_pen), ('actual', o_act)):
            if not same(log[key][t], val):
                return fail(key + ' differs', t)
        ids = np.flatnonzero(o_done)
        if len(ids) and cfg.auto_reset:
            if reset_start_index is not None:
                nxt = np.asarray(reset_start_index[t], dtype=np.int32)[ids]
            else:
                nxt = np.array([auto_start_index(cfg.seed, int(env_ids[i]), int(episode[i]), n_starts) for i in ids], dtype=np.int32)
            first = ora.reset(nxt, env_ids=list(ids))
            episode[ids] += 1
            out['episodes'] += len(ids)
            if not same(log['next_obs'][t][ids], first):
                return fail('first observation after auto-reset differs', t)
        keep = np.ones(n, dtype=bool)
        keep[ids] = False
        if not same(log['next_obs'][t][keep], o_obs[keep]):
            return fail('next_obs differs', t)
        if status is not None and t in status:
            for k in range(n):
                if not cfg.auto_reset and o_done[k]:
                    continue
                if not np.array_equal(status[t][k], ora.status(k)):
                    return fail('status plane of env %d differs' % env_ids[k], t)
            out['planes_checked'] += n
    ora.close()
    return out
