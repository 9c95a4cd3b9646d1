"""CPU checks of the subset-replay checker (oracle/replay.py) that the large-batch GPU parity tests and
bench.py's `parity_check` rely on: a log produced by one oracle batch replays clean through another,
a corrupted log is caught, and the Python start-index stream equals the engine's splitmix64 stream
(known answers computed from the C definition)."""
import numpy as np

from paintrl_b200.config import EnvConfig
from paintrl_b200.partpack import PartPack

BASE = {'RENDER_HEIGHT': 720, 'RENDER_WIDTH': 960, 'Part_NO': 0, 'Expected_Episode_Length': 245,
        'EPISODE_MAX_LENGTH': 245, 'TERMINATION_MODE': 'late', 'SWITCH_THRESHOLD': 0.9,
        'START_POINT_MODE': 'anchor', 'TURNING_PENALTY': False, 'OVERLAP_PENALTY': False,
        'COLOR_MODE': 'RGB'}


def _log(pack, cfg, n, steps, seed):
    """A full-batch log produced the way the engine would: seeded auto-reset start indices."""
    from oracle.oracle import OracleBatch
    from oracle.replay import auto_start_index
    ora = OracleBatch(pack, cfg, n)
    rng = np.random.default_rng(seed)
    start = rng.integers(0, 4, size=n).astype(np.int32)
    ora.reset(start)
    episode = np.ones(n, dtype=np.int64)
    keys = ('actions', 'obs', 'next_obs', 'reward', 'penalty', 'actual', 'done')
    log = {k: [] for k in keys}
    status = {}
    for t in range(steps):
        acts = rng.integers(0, 4, size=n)
        obs, rew, pen, act, done = ora.step(acts)
        nxt = obs.copy()
        for i in np.flatnonzero(done):
            idx = auto_start_index(cfg.seed, i, int(episode[i]), 4)
            nxt[i] = ora.reset(np.array([idx], dtype=np.int32), env_ids=[i])[0]
            episode[i] += 1
        for k, v in zip(keys, (acts, obs, nxt, rew, pen, act, done)):
            log[k].append(v)
        if t % 10 == 9:
            status[t] = ora.status()
    ora.close()
    return start, {k: np.stack(v) for k, v in log.items()}, status


def test_subset_of_a_batch_replays_clean_and_corruption_is_caught():
    from oracle.replay import replay_subset, sample_env_ids
    cfg = EnvConfig(dict(BASE, EPISODE_MAX_LENGTH=12), auto_reset=True, seed=99)
    pack = PartPack.for_part(0)
    n, steps = 48, 30
    start, log, status = _log(pack, cfg, n, steps, 3)
    ids = sample_env_ids(n, 16, seed=1, tail=4)
    assert len(ids) == 16 and ids[-1] == n - 1 and len(set(ids.tolist())) == 16
    sub = {k: v[:, ids] for k, v in log.items()}
    st = {t: s[ids] for t, s in status.items()}
    res = replay_subset(pack, cfg, ids, start[ids], sub, status=st)
    assert res['ok'] and res['exact'] and res['episodes'] >= 16 and res['planes_checked'] == 3 * 16, res
    bad = {k: v.copy() for k, v in sub.items()}
    bad['obs'][17, 5, 2] += 1e-9
    res = replay_subset(pack, cfg, ids, start[ids], bad, status=st)
    assert not res['ok'] and 'step 17' in res['mismatch'], res
    bad_st = {t: s.copy() for t, s in st.items()}
    bad_st[19][3, 100] ^= 64
    res = replay_subset(pack, cfg, ids, start[ids], sub, status=bad_st)
    assert not res['ok'] and 'status plane' in res['mismatch'], res


def test_start_index_stream_known_answers():
    """splitmix64 known answers (the published test vector of the generator seeded with 0: first outputs
    0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4) and range of the derived start indices."""
    from oracle.replay import auto_start_index, splitmix64
    assert splitmix64(0) == 0xE220A8397B1DCDAF
    assert splitmix64(0x9E3779B97F4A7C15) == 0x6E789E6AA1B965F4
    draws = [auto_start_index(1234, e, k, 4) for e in range(64) for k in range(1, 5)]
    assert set(draws) == {0, 1, 2, 3}
