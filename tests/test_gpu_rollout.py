"""GPU: the on-device rollout loop (BASELINE config C5 shape) -- a fragment collected through
`step_into` is replayed action by action through a second engine and through the C oracle."""
import numpy as np
import pytest
import torch

from paintrl_b200.config import EnvConfig
from paintrl_b200.partpack import PartPack
from test_gpu_oracle_batch import BASE

pytestmark = pytest.mark.gpu


def test_fragment_replays_through_the_oracle(cuda_device):
    from oracle.oracle import OracleBatch
    from paintrl_b200.batched_env import BatchedPaintEnv
    from paintrl_b200.rollout import MlpPolicy, RolloutWorker, gae, iteration_stats
    n, T = 96, 40
    cfg = EnvConfig(dict(BASE), auto_reset=True, seed=11)
    pack = PartPack.for_part(0)
    env = BatchedPaintEnv(n, cfg, device=cuda_device, pack=pack)
    policy = MlpPolicy(env.obs_dim, cfg.discrete_granularity, device=cuda_device, seed=2)
    worker = RolloutWorker(env, policy, fragment_length=T)
    start = (np.arange(n) % env.n_starts).astype(np.int32)
    worker.start(start)
    ora = OracleBatch(pack, cfg, n)
    probe = OracleBatch(pack, cfg, 1)
    reset_obs = [probe.reset(np.array([i], np.int32))[0].copy() for i in range(env.n_starts)]
    probe.close()
    assert np.array_equal(worker.frag.obs[0].cpu().numpy(), ora.reset(start))
    totals = {'episodes': 0.0, 'sum_reward': 0.0, 'sum_penalty': 0.0, 'new_texels': 0.0, 'max_episode_len': 0.0}
    ep_r, ep_p, ep_l = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int64)
    exp = dict(totals)
    for it in range(2):
        f, stats = worker.collect()
        acts = f.actions.cpu().numpy()
        for t in range(T):
            o, r, p, a, d = ora.step(acts[t])
            assert np.array_equal(f.term_obs[t].cpu().numpy(), o), (it, t)
            assert np.array_equal(f.reward[t].cpu().numpy(), r) and np.array_equal(f.penalty[t].cpu().numpy(), p)
            assert np.array_equal(f.actual[t].cpu().numpy(), a) and np.array_equal(f.done[t].cpu().numpy(), d)
            ep_r += r; ep_p += p; ep_l += 1
            ids = np.flatnonzero(d)
            nxt = f.obs[t + 1].cpu().numpy()
            keep = np.flatnonzero(d == 0)
            assert np.array_equal(nxt[keep], o[keep])
            if len(ids):
                exp['episodes'] += len(ids)
                exp['sum_reward'] += ep_r[ids].sum(); exp['sum_penalty'] += ep_p[ids].sum()
                exp['max_episode_len'] = max(exp['max_episode_len'], float(ep_l[ids].max()))
                ep_r[ids] = 0; ep_p[ids] = 0; ep_l[ids] = 0
                # the engine drew the new episodes' start points from its own seeded stream: follow it by
                # matching the first observation of the new episode against each start point's
                for e in ids:
                    idx = [i for i in range(env.n_starts) if np.array_equal(reset_obs[i], nxt[e])]
                    assert len(idx) == 1, (it, t, e)
                    assert np.array_equal(ora.reset(np.array(idx, np.int32), env_ids=[int(e)])[0], nxt[e])
        for k in totals:
            totals[k] = max(totals[k], stats[k]) if k.startswith('max') else totals[k] + stats[k]
        adv, target = gae(f)
        assert adv.shape == (T, n) and bool(torch.isfinite(adv).all()) and bool(torch.isfinite(target).all())
        assert iteration_stats(stats, device=cuda_device)['env_steps'] == T * n
        worker.advance()
    assert totals['episodes'] == exp['episodes'] and totals['max_episode_len'] == exp['max_episode_len']
    assert np.isclose(totals['sum_reward'], exp['sum_reward'], rtol=1e-12)
    assert np.isclose(totals['sum_penalty'], exp['sum_penalty'], rtol=1e-12)
    assert totals['new_texels'] > 0
    env.close()
    ora.close()


def test_cuda_graph_fragment_equals_eager_fragment(cuda_device):
    """The same seeds through the eager loop and through the captured CUDA graph give identical fragments
    (the per-environment move -> paint hand-off carries no per-launch argument, so a captured step replays)."""
    from paintrl_b200.batched_env import BatchedPaintEnv
    from paintrl_b200.rollout import MlpPolicy, RolloutWorker
    n, T = 128, 12
    frags = []
    for use_graph in (False, True):
        cfg = EnvConfig(dict(BASE), auto_reset=True, seed=3)
        env = BatchedPaintEnv(n, cfg, device=cuda_device)
        worker = RolloutWorker(env, MlpPolicy(env.obs_dim, 4, device=cuda_device, seed=5), fragment_length=T,
                               use_cuda_graph=use_graph)
        worker.start((np.arange(n) % env.n_starts).astype(np.int32))
        out = []
        for it in range(4):
            f, stats = worker.collect()
            out.append({k: getattr(f, k).clone() for k in ('obs', 'actions', 'reward', 'done', 'term_obs', 'logp')})
            out[-1]['stats'] = stats
            worker.advance()
        if use_graph:
            assert worker._graph is not None, worker.graph_error
        frags.append(out)
        env.close()
    for a, b in zip(*frags):
        for k in ('obs', 'actions', 'reward', 'done', 'term_obs', 'logp'):
            assert torch.equal(a[k], b[k]), k
        assert a['stats'] == b['stats']
