"""GPU parity, gate 2: batches of environments stepped by the CUDA engine and by the C
restatement (oracle/paint_oracle.c, itself pinned to the golden traces) on the same seeded
actions and start points, including same-step auto-reset."""
import numpy as np
import pytest
import torch

from paintrl_b200.config import EnvConfig
from paintrl_b200.partpack import PartPack

pytestmark = pytest.mark.gpu

BASE = {'RENDER_HEIGHT': 720, 'RENDER_WIDTH': 960, 'Part_NO': 0, 'Expected_Episode_Length': 245,
        'EPISODE_MAX_LENGTH': 245, 'TERMINATION_MODE': 'late', 'SWITCH_THRESHOLD': 0.9,
        'START_POINT_MODE': 'anchor', 'TURNING_PENALTY': False, 'OVERLAP_PENALTY': False,
        'COLOR_MODE': 'RGB'}

CASES = {
    # BASELINE config C2 semantics (door, RGB, anchor, discrete-4, section-4, late)
    'c2_door': (dict(BASE), dict(), 192, 70),
    # C3: sheet, HSI, penalties, hybrid
    'c3_sheet_hsi': (dict(BASE, Part_NO=1, COLOR_MODE='HSI', TURNING_PENALTY=True, OVERLAP_PENALTY=True,
                          TERMINATION_MODE='hybrid'), dict(), 96, 40),
    # sheet HSI late: long overlapping walks (thickness saturation)
    'sheet_hsi_late': (dict(BASE, Part_NO=1, COLOR_MODE='HSI', OVERLAP_PENALTY=True), dict(), 64, 120),
    # C4 at 240x240: continuous-2, grid-4, all start points
    'c4_door_grid': (dict(BASE, START_POINT_MODE='all'),
                     dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4), 96, 50),
    'door_discrete_obs': (dict(BASE, START_POINT_MODE='edge', TERMINATION_MODE='early',
                               Expected_Episode_Length=600),
                          dict(obs_mode='discrete', obs_grad=4, discrete_granularity=8), 64, 50),
    'sheet_simple': (dict(BASE, Part_NO=1, START_POINT_MODE='all'), dict(obs_mode='simple'), 64, 60),
    # the other two parts of Part_Dict with a usable max-points entry (robot_gym_env.py:106-117)
    'door_rr_hsi': (dict(BASE, Part_NO=5, COLOR_MODE='HSI', START_POINT_MODE='edge', OVERLAP_PENALTY=True), dict(), 64, 50),
    'test_part_grid': (dict(BASE, Part_NO=9, START_POINT_MODE='all'),
                       dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4), 64, 40),
}


def _run_case(extra, kw, num_envs, steps, device, auto_reset=True, texture=None, status_every=7, pack=None):
    from oracle.oracle import OracleBatch, action_directions, retextured_pack
    from paintrl_b200.batched_env import BatchedPaintEnv
    cfg = EnvConfig(extra, auto_reset=auto_reset, **kw)
    pack = pack if pack is not None else PartPack.for_part(cfg.part_no)
    if texture is None:
        env = BatchedPaintEnv(num_envs, cfg, device=device, pack=pack)
        ora = OracleBatch(pack, cfg, num_envs)
    else:
        # the engine rasterises its texels on the GPU, the oracle with its own C restatement
        env = BatchedPaintEnv(num_envs, cfg, device=device, texture_size=texture)
        opack = retextured_pack(pack, *texture)
        assert np.array_equal(env.pack.front_ij, opack.front_ij) and np.array_equal(env.pack.front_pos, opack.front_pos)
        assert env.pack.max_points == opack.max_points
        ora = OracleBatch(opack, cfg, num_envs)
    rng = np.random.default_rng(20261017)
    n_starts = env.n_starts
    start = rng.integers(0, n_starts, size=num_envs).astype(np.int32)
    obs_g = env.reset(start).cpu().numpy()
    obs_o = ora.reset(start)
    continuous = cfg.action_mode == 'continuous'
    tol = dict(rtol=1e-5, atol=1e-12)
    assert np.array_equal(obs_g, obs_o)
    episodes = 0
    for t in range(steps):
        if continuous:
            acts = rng.uniform(-1, 1, size=(num_envs, cfg.action_dim))
            # the oracle takes the unit directions NumPy computes (as the reference does); the
            # engine computes them on the device from the raw actions
        else:
            acts = rng.integers(0, cfg.discrete_granularity, size=num_envs)
        nxt = rng.integers(0, n_starts, size=num_envs).astype(np.int32)
        o_g, a_g, d_g, info = env.step(acts, reset_start_index=nxt)
        o_o, r_o, p_o, a_o, d_o = ora.step(acts)
        o_g, a_g, d_g = o_g.cpu().numpy(), a_g.cpu().numpy(), d_g.cpu().numpy()
        ctx = ('step', t)
        assert np.array_equal(d_g, d_o), ctx
        if continuous or cfg.color_mode != 'RGB':
            assert np.allclose(o_g, o_o, **tol), ctx
            assert np.allclose(info['reward'].cpu().numpy(), r_o, **tol), ctx
            assert np.allclose(info['penalty'].cpu().numpy(), p_o, **tol), ctx
            assert np.allclose(a_g, a_o, **tol), ctx
        else:
            assert np.array_equal(o_g, o_o), ctx
            assert np.array_equal(info['reward'].cpu().numpy(), r_o), ctx
            assert np.array_equal(info['penalty'].cpu().numpy(), p_o), ctx
            assert np.array_equal(a_g, a_o), ctx
        # status planes: bit-exact, before the oracle side applies the auto-reset
        if t % status_every == 0 or t == steps - 1:
            pre_reset = env.get_state()['status'].cpu().numpy()
            for e in range(num_envs):
                if not d_o[e] or not auto_reset:
                    assert np.array_equal(pre_reset[e], ora.status(e)), ctx + (e,)
        done_ids = np.flatnonzero(d_o)
        if auto_reset and len(done_ids):
            episodes += len(done_ids)
            ro = ora.reset(nxt[done_ids], env_ids=list(done_ids))
            ng = info['next_obs'].cpu().numpy()
            assert np.allclose(ng[done_ids], ro, **tol) if continuous else np.array_equal(ng[done_ids], ro), ctx
    env.close()
    ora.close()
    return episodes


@pytest.mark.parametrize('case', sorted(CASES))
def test_batch_matches_oracle(case, cuda_device):
    extra, kw, num_envs, steps = CASES[case]
    _run_case(extra, kw, num_envs, steps, cuda_device)


def test_host_step_matches_device_step(cuda_device):
    from paintrl_b200.batched_env import BatchedPaintEnv
    a = BatchedPaintEnv(32, dict(BASE), device=cuda_device)
    b = BatchedPaintEnv(32, dict(BASE), device=cuda_device)
    start = np.arange(32, dtype=np.int32) % 4
    a.reset(start)
    b.reset(start)
    rng = np.random.default_rng(3)
    out = b.host_buffers()
    for _ in range(20):
        acts = rng.integers(0, 4, size=32)
        o, actual, done, info = a.step(acts)
        b.step_host(acts, out)
        assert np.array_equal(o.cpu().numpy(), out['obs'])
        assert np.array_equal(actual.cpu().numpy(), out['actual'])
        assert np.array_equal(done.cpu().numpy(), out['done'])
        assert np.array_equal(info['reward'].cpu().numpy(), out['reward'])
        assert np.array_equal(info['penalty'].cpu().numpy(), out['penalty'])
        assert np.array_equal(info['next_obs'].cpu().numpy(), out['next_obs'])
    a.close()
    b.close()


@pytest.mark.parametrize('auto_reset', [True, False])
def test_host_step_scattered_buffers(cuda_device, auto_reset):
    """paintrl_step_host merges device->host copies when the host buffers are carved from one
    allocation (host_buffers); separately allocated, unpinned buffers must give the same results."""
    from paintrl_b200.batched_env import BatchedPaintEnv
    n = 48
    cfg = dict(BASE)
    a = BatchedPaintEnv(n, cfg, device=cuda_device, auto_reset=auto_reset, seed=5)
    b = BatchedPaintEnv(n, cfg, device=cuda_device, auto_reset=auto_reset, seed=5)
    start = np.arange(n, dtype=np.int32) % 4
    a.reset(start)
    b.reset(start)
    rng = np.random.default_rng(4)
    one = a.host_buffers(pinned=True)
    od = a.obs_dim
    scattered = {'obs': np.zeros((n, od)), 'next_obs': np.full((n, od), -7.0), 'done': np.zeros(n, np.uint8),
                 'actual': np.zeros(n), 'penalty': np.zeros(n), 'reward': np.zeros(n)}
    for _ in range(30):
        acts = rng.integers(0, 4, size=n)
        a.step_host(acts, one)
        b.step_host(acts, scattered)
        for k in ('obs', 'reward', 'penalty', 'actual', 'done', 'next_obs'):
            assert np.array_equal(one[k], scattered[k]), k
    no_next = {k: v for k, v in a.host_buffers(pinned=False, next_obs=False).items()}
    assert 'next_obs' not in no_next
    acts = rng.integers(0, 4, size=n)
    a.step_host(acts, no_next)
    b.step_host(acts, scattered)
    for k in ('obs', 'reward', 'penalty', 'actual', 'done'):
        assert np.array_equal(no_next[k], scattered[k]), k
    a.close()
    b.close()


def test_host_steps_interleaved_with_device_steps_and_resets(cuda_device):
    """Repeated step_host calls with the same pinned buffers, interleaved with resets, device-side steps and a
    change of buffers, against the device-path engine."""
    from paintrl_b200.batched_env import BatchedPaintEnv
    n = 96
    a = BatchedPaintEnv(n, dict(BASE), device=cuda_device, auto_reset=True, seed=9)      # host-buffer path
    b = BatchedPaintEnv(n, dict(BASE), device=cuda_device, auto_reset=True, seed=9)      # device path (reference)
    start = (np.arange(n) % 4).astype(np.int32)
    a.reset(start)
    b.reset(start)
    rng = np.random.default_rng(12)
    out = a.host_buffers(pinned=True)
    acts_pinned = torch.zeros(n, dtype=torch.int64, pin_memory=True).numpy()
    for t in range(60):
        acts = rng.integers(0, 4, size=n)
        if t == 25:                        # a reset in between
            a.reset(start)
            b.reset(start)
        if t == 40:                        # new buffers
            out = a.host_buffers(pinned=True)
        if t in (10, 11):                  # device-side steps in between
            a.step(acts)
            b.step(acts)
            continue
        acts_pinned[:] = acts
        a.step_host(acts_pinned, out)
        o, actual, done, info = b.step(acts)
        assert np.array_equal(o.cpu().numpy(), out['obs']), t
        assert np.array_equal(actual.cpu().numpy(), out['actual']) and np.array_equal(done.cpu().numpy(), out['done']), t
        assert np.array_equal(info['next_obs'].cpu().numpy(), out['next_obs']), t
    assert np.array_equal(a.get_state()['status'].cpu().numpy(), b.get_state()['status'].cpu().numpy())
    a.close()
    b.close()


@pytest.mark.parametrize('continuous', [False, True])
def test_out_of_range_actions_are_clipped_like_the_reference(cuda_device, continuous):
    """robot.py:390-393 clips instead of rejecting: discrete actions outside 0..n-1, continuous components
    beyond [-1, 1], infinities and NaN (which the reference's comparison chain sends to +1)."""
    from oracle.oracle import OracleBatch
    from paintrl_b200.batched_env import BatchedPaintEnv
    kw = dict(action_mode='continuous', action_shape=2) if continuous else dict(discrete_granularity=4)
    cfg = EnvConfig(dict(BASE), auto_reset=False, **kw)
    pack = PartPack.for_part(0)
    n = 64
    env = BatchedPaintEnv(n, cfg, device=cuda_device, pack=pack)
    ora = OracleBatch(pack, cfg, n)
    start = (np.arange(n) % 4).astype(np.int32)
    assert np.array_equal(env.reset(start).cpu().numpy(), ora.reset(start))
    rng = np.random.default_rng(5)
    for t in range(12):
        if continuous:
            acts = rng.uniform(-3, 3, size=(n, 2))
            acts[::7, 0] = np.nan
            acts[3::11, 1] = np.inf
            acts[5::13, 0] = -np.inf
        else:
            acts = rng.integers(-3, 9, size=n)
        o_g, a_g, d_g, info = env.step(acts)
        o_o, r_o, p_o, a_o, d_o = ora.step(acts)
        assert np.array_equal(d_g.cpu().numpy(), d_o), t
        if continuous:
            assert np.allclose(o_g.cpu().numpy(), o_o, rtol=1e-5, atol=1e-12) and np.allclose(a_g.cpu().numpy(), a_o, rtol=1e-5, atol=1e-12)
        else:
            assert np.array_equal(o_g.cpu().numpy(), o_o) and np.array_equal(a_g.cpu().numpy(), a_o), t
    assert np.array_equal(env.get_state()['status'].cpu().numpy(), ora.status())
    env.close()
    ora.close()


@pytest.mark.parametrize('num_envs', [1, 5, 33, 257])
def test_ragged_batch_sizes(cuda_device, num_envs):
    """Batch sizes that fill neither a CTA nor a warp group, down to a single environment."""
    _run_case(dict(BASE, START_POINT_MODE='edge'), dict(), num_envs, 12, cuda_device, status_every=5)


def test_subset_reset_leaves_the_other_environments_alone(cuda_device):
    """reset_at / paintrl_reset with env_ids (the RLlib VectorEnv contract): only the listed environments
    start over; set_pose (spiral.py's robot.reset(pose)) moves without clearing the paint."""
    from oracle.oracle import OracleBatch
    from paintrl_b200.batched_env import BatchedPaintEnv
    cfg = EnvConfig(dict(BASE), auto_reset=False)
    pack = PartPack.for_part(0)
    n = 40
    env = BatchedPaintEnv(n, cfg, device=cuda_device, pack=pack)
    ora = OracleBatch(pack, cfg, n)
    start = (np.arange(n) % 4).astype(np.int32)
    env.reset(start)
    ora.reset(start)
    rng = np.random.default_rng(8)
    for t in range(20):
        acts = rng.integers(0, 4, size=n)
        env.step(acts)
        ora.step(acts)
        if t % 5 == 4:
            ids = rng.choice(n, size=7, replace=False).astype(np.int32)
            st = rng.integers(0, 4, size=7).astype(np.int32)
            assert np.array_equal(env.reset(st, env_ids=ids).cpu().numpy(), ora.reset(st, env_ids=list(ids)))
        if t == 12:
            pts = pack.start_points('all')
            pos, nrm = pts[100, 0], pts[100, 1]
            assert np.array_equal(env.set_pose(pos, nrm, env_ids=[3, 9]).cpu().numpy(), ora.set_pose(pos, nrm, env_ids=[3, 9]))
    assert np.array_equal(env.get_state()['status'].cpu().numpy(), ora.status())
    env.close()
    ora.close()


@pytest.mark.parametrize('auto_reset', [True, False])
def test_pipelined_host_steps_match_synchronous_ones(cuda_device, auto_reset):
    """paintrl_step_host_submit / _wait on two staging slots: step t + 1 is submitted before step t is waited for
    (its copy-in and kernels overlap step t's copy-out); every result equals the synchronous paintrl_step_host's."""
    from paintrl_b200 import _capi
    from paintrl_b200.batched_env import BatchedPaintEnv
    n, T = 200, 60
    cfg = dict(BASE, START_POINT_MODE='edge', OVERLAP_PENALTY=True)
    a = BatchedPaintEnv(n, cfg, device=cuda_device, auto_reset=auto_reset, seed=21)
    b = BatchedPaintEnv(n, cfg, device=cuda_device, auto_reset=auto_reset, seed=21)
    start = (np.arange(n) % a.n_starts).astype(np.int32)
    a.reset(start)
    b.reset(start)
    rng = np.random.default_rng(31)
    acts = torch.from_numpy(rng.integers(0, 4, size=(T, n))).pin_memory().numpy()
    outs = [a.host_buffers(pinned=True), a.host_buffers(pinned=True)]
    ref = b.host_buffers(pinned=True)
    with pytest.raises(_capi.PaintrlError):
        a.step_host_wait(0)                          # nothing submitted yet
    a.step_host_submit(acts[0], outs[0], slot=0)
    for t in range(T):
        if t + 1 < T:
            a.step_host_submit(acts[t + 1], outs[(t + 1) & 1], slot=(t + 1) & 1)
        got = a.step_host_wait(slot=t & 1)
        b.step_host(acts[t], ref)
        for k in ('obs', 'reward', 'penalty', 'actual', 'done', 'next_obs'):
            assert np.array_equal(got[k], ref[k]), (t, k)
    # a slot submitted twice without a wait in between drains its first copy-out before it is reused
    a.step_host_submit(acts[0], outs[0], slot=0)
    a.step_host_submit(acts[1], outs[0], slot=0)
    got = a.step_host_wait(0)
    b.step_host(acts[0], ref)
    b.step_host(acts[1], ref)
    for k in ('obs', 'actual', 'done'):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(a.get_state()['status'].cpu().numpy(), b.get_state()['status'].cpu().numpy())
    a.close()
    b.close()
