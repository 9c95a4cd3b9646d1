"""GPU parity of the normal paint method (Robot.PAINT_METHOD = 'normal', robot.py:172, 414-417: a fan of 124-450 rays per
shot, the texel nearest to every hit; bullet_paint_wrapper.py:562-566): golden traces minted from the reference with
that setting, and batches against the C oracle with auto-reset.  RGB bit-exact; HSI rewards within 1e-5 (integer
thickness planes bit-exact)."""
import numpy as np
import pytest
import torch

from golden_util import Golden
from paintrl_b200.config import EnvConfig
from paintrl_b200.partpack import PartPack
from test_gpu_oracle_batch import BASE

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['g11_door_normal_rgb', 'g12_sheet_normal_hsi'])
def test_normal_paint_golden(cuda_device, name):
    from paintrl_b200.batched_env import BatchedPaintEnv
    g = Golden(name)
    hsi = g.cfg.color_mode != 'RGB'
    tol = dict(rtol=1e-5, atol=1e-12)
    env = BatchedPaintEnv(g.n_episodes, g.cfg, device=cuda_device, pack=g.pack)
    start = np.array([g.start_index(e) for e in range(g.n_episodes)], dtype=np.int32)
    obs = env.reset(start).cpu().numpy()
    for e in range(g.n_episodes):
        assert np.array_equal(obs[e], g['obs'][e, 0])
    T = int(g.lengths.max())
    for t in range(T):
        if g.cfg.action_mode == 'discrete':
            acts = np.array([g.actions(e, min(t, g.lengths[e] - 1)) for e in range(g.n_episodes)])
        else:
            acts = np.stack([g.actions(e, min(t, g.lengths[e] - 1)) for e in range(g.n_episodes)])
        o, actual, done, info = env.step(acts)
        o, actual, done = o.cpu().numpy(), actual.cpu().numpy(), done.cpu().numpy()
        status = env.get_state()['status'].cpu().numpy()
        for e in range(g.n_episodes):
            if t >= g.lengths[e]:
                continue
            ctx = (name, e, t)
            assert int(done[e]) == int(g['done'][e, t]), ctx
            assert int(np.count_nonzero(status[e] != g.pack.status_init(g.cfg.color_mode))) == int(g['painted'][e, t]), ctx
            if hsi or g.cfg.action_mode == 'continuous':
                assert np.allclose(o[e], g['obs'][e, t + 1], **tol), ctx
                assert np.allclose(actual[e], g['actual'][e, t], **tol), ctx
            else:
                assert np.array_equal(o[e], g['obs'][e, t + 1]), ctx
                assert actual[e] == g['actual'][e, t], ctx
            if t == g.lengths[e] - 1:
                assert np.array_equal(status[e], g['status_final'][e]), ctx
    env.close()


@pytest.mark.parametrize('case', ['door_rgb', 'sheet_hsi', 'test_part_hsi_grid'])
def test_normal_paint_batch_matches_oracle(cuda_device, case):
    from oracle.oracle import OracleBatch
    from paintrl_b200.batched_env import BatchedPaintEnv
    extra, kw, n, steps = {
        'door_rgb': (dict(BASE, START_POINT_MODE='edge', OVERLAP_PENALTY=True), dict(), 48, 30),
        'sheet_hsi': (dict(BASE, Part_NO=1, COLOR_MODE='HSI', OVERLAP_PENALTY=True, TURNING_PENALTY=True), dict(), 32, 24),
        'test_part_hsi_grid': (dict(BASE, Part_NO=9, COLOR_MODE='HSI', START_POINT_MODE='all', OVERLAP_PENALTY=True),
                             dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4), 32, 16),
    }[case]
    cfg = EnvConfig(extra, auto_reset=True, paint_method='normal', seed=4, **kw)
    pack = PartPack.for_part(cfg.part_no)
    env = BatchedPaintEnv(n, cfg, device=cuda_device, pack=pack)
    ora = OracleBatch(pack, cfg, n)
    rng = np.random.default_rng(99)
    start = rng.integers(0, env.n_starts, size=n).astype(np.int32)
    assert np.array_equal(env.reset(start).cpu().numpy(), ora.reset(start))
    exact = cfg.color_mode == 'RGB' and cfg.action_mode == 'discrete'
    tol = dict(rtol=1e-5, atol=1e-12)
    for t in range(steps):
        acts = rng.integers(0, 4, size=n) if cfg.action_mode == 'discrete' else rng.uniform(-1, 1, size=(n, 2))
        if t < 6 and cfg.action_mode == 'discrete':
            acts[:] = 1 if t % 2 == 0 else 3          # back and forth: overlapping shots (repeat coats, valid-pixel logic)
        nxt = rng.integers(0, env.n_starts, size=n).astype(np.int32)
        o_g, a_g, d_g, info = env.step(acts, reset_start_index=nxt)
        o_o, r_o, p_o, a_o, d_o = ora.step(acts)
        assert np.array_equal(d_g.cpu().numpy(), d_o), t
        same = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, **tol))
        assert same(o_g.cpu().numpy(), o_o), t
        assert same(info['reward'].cpu().numpy(), r_o) and same(info['penalty'].cpu().numpy(), p_o) and same(a_g.cpu().numpy(), a_o), t
        status = env.get_state()['status'].cpu().numpy()
        for e in range(n):
            if not d_o[e]:
                assert np.array_equal(status[e], ora.status(e)), (t, e)
        ids = np.flatnonzero(d_o)
        if len(ids):
            ro = ora.reset(nxt[ids], env_ids=list(ids))
            assert same(info['next_obs'].cpu().numpy()[ids], ro), t
    env.close()
    ora.close()


def test_normal_paint_needs_a_small_texture_and_a_beam_table(cuda_device):
    from paintrl_b200 import _capi
    from paintrl_b200.batched_env import BatchedPaintEnv
    cfg = EnvConfig(dict(BASE), paint_method='normal', beam_plain=np.zeros((600, 3)))
    with pytest.raises(_capi.PaintrlError):
        BatchedPaintEnv(4, cfg, device=cuda_device)
    with pytest.raises(ValueError):
        EnvConfig(dict(BASE), paint_method='slow')


def test_normal_paint_refuses_textures_beyond_the_staged_plane(cuda_device):
    """The beam-fan kernel keeps the shot / union masks of the whole bit-plane in shared memory: parts with more than
    16384 front texels (door_rr: 17891) are refused at construction, not mis-painted."""
    from paintrl_b200 import _capi
    from paintrl_b200.batched_env import BatchedPaintEnv
    with pytest.raises(_capi.PaintrlError):
        BatchedPaintEnv(4, EnvConfig(dict(BASE, Part_NO=5), paint_method='normal'), device=cuda_device)
