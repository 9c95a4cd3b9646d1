"""GPU parity at the BASELINE.json batch sizes.

The engine steps the FULL batch of each configuration (C2 4096, C3 16384, C4 8192 at 2048x2048, C5 65536
environments); a few hundred of its environment indices are logged on the device and replayed through the
C oracle afterwards (oracle/replay.py): done flags, observations, rewards, penalties and the first
observation after every auto-reset bit-exact (RGB / discrete) or within 1e-5 (HSI, continuous), status
planes bit-exact at three checkpoints.  These are the regimes the small-batch tests cannot reach: more
than one wave of the move grid with paint CTAs co-resident through the per-environment hand-off flags,
the 8-lane move kernel the engine selects from 16384 environments on, and the generic-axes kernels.

Reference semantics: PaintGymEnv.step / reset, robot_gym_env.py:349-387.
"""
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
C3 = dict(BASE, Part_NO=1, COLOR_MODE='HSI', TURNING_PENALTY=True, OVERLAP_PENALTY=True, TERMINATION_MODE='hybrid')


def _run(extra, kw, num_envs, steps, sample, device, seeded_resets=True, texture=None, pack=None, checkpoints=3):
    from oracle.oracle import retextured_pack
    from oracle.replay import SubsetRecorder, replay_subset, sample_env_ids
    from paintrl_b200.batched_env import BatchedPaintEnv
    cfg = EnvConfig(extra, auto_reset=True, seed=77, **kw)
    if pack is None:
        pack = PartPack.for_part(cfg.part_no)
    if texture is None:
        env = BatchedPaintEnv(num_envs, cfg, device=device, pack=pack)
        opack = pack
    else:
        env = BatchedPaintEnv(num_envs, cfg, device=device, texture_size=texture)
        opack = retextured_pack(pack, *texture)       # the oracle rasterises its own texels
        assert np.array_equal(env.pack.front_ij, opack.front_ij) and np.array_equal(env.pack.front_pos, opack.front_pos)
    gen = torch.Generator(device=device)
    gen.manual_seed(1234)
    if cfg.action_mode == 'discrete':
        actions = torch.randint(0, cfg.discrete_granularity, (steps, num_envs), generator=gen, device=device, dtype=torch.int64)
    else:
        actions = torch.rand((steps, num_envs, cfg.action_dim), generator=gen, device=device, dtype=torch.float64) * 2 - 1
    start = torch.randint(0, env.n_starts, (num_envs,), generator=gen, device=device, dtype=torch.int32)
    nxt = None if seeded_resets else torch.randint(0, env.n_starts, (steps, num_envs), generator=gen, device=device,
                                                   dtype=torch.int32)
    ids = sample_env_ids(num_envs, sample, seed=5)
    rec = SubsetRecorder(env, ids, steps)
    env.reset(start)
    marks = set(np.linspace(steps // 3, steps - 1, num=checkpoints, dtype=int).tolist())
    for t in range(steps):
        env.step(actions[t], reset_start_index=None if nxt is None else nxt[t])
        rec.record(actions[t], status=t in marks)
    stats = env.stats()
    res = replay_subset(opack, cfg, ids, start.cpu().numpy()[ids], rec.host(), status=rec.status,
                        reset_start_index=None if nxt is None else nxt.cpu().numpy()[:, ids])
    env.close()
    assert res['ok'], res
    assert res['planes_checked'] == len(marks) * len(ids)
    assert stats['env_steps'] == num_envs * steps
    return res, stats


def test_c2_door_4096_envs_full_episode(cuda_device):
    """BASELINE configs[1]: 4096 environments, 245 steps, default 32-lane move kernel (two waves)."""
    res, stats = _run(dict(BASE), {}, 4096, 245, 256, cuda_device)
    assert res['exact'] and res['episodes'] > 0


def test_c3_sheet_hsi_16384_envs(cuda_device):
    """BASELINE configs[2]: 16384 environments (the 8-lane move kernel's regime), HSI, penalties, hybrid
    termination (which ends every random-action episode at its first step: reset + one stamp)."""
    res, _ = _run(C3, {}, 16384, 60, 256, cuda_device)
    assert res['episodes'] > 0


def test_c3_sheet_hsi_16384_envs_late_termination(cuda_device):
    """The same batch with late termination: episodes last, so the HSI thickness update, saturation and the
    overlap bookkeeping run in steady state; explicit start indices for the auto-resets."""
    res, _ = _run(dict(C3, TERMINATION_MODE='late'), {}, 16384, 90, 192, cuda_device, seeded_resets=False)


def test_c4_door_2048_texture_8192_envs(cuda_device):
    """BASELINE configs[3] at full size: 2048x2048 texture (plane in global memory), continuous 2-D actions,
    grid observation, every start point."""
    res, _ = _run(dict(BASE, START_POINT_MODE='all'), dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4),
                  8192, 12, 24, cuda_device, texture=(2048, 2048), checkpoints=2)


def test_c5_door_65536_envs(cuda_device):
    """BASELINE configs[4] env side on one GPU: 65536 environments (16 waves of the 8-lane move kernel)."""
    res, stats = _run(dict(BASE), {}, 65536, 40, 320, cuda_device)
    assert res['exact']


@pytest.mark.parametrize('frame', ['axes02', 'axes01'])
@pytest.mark.parametrize('color', ['RGB', 'HSI'])
def test_principal_axes_other_than_1_2(cuda_device, frame, color):
    """Synthetic frames of the door whose principal axes are (0, 2) / (0, 1): the AX12 = false instantiations
    of both kernels (paintrl_capi.cu dispatch), 32- and 8-lane."""
    from pack_util import permuted_pack
    pack = permuted_pack(PartPack.for_part(0), frame)
    extra = dict(BASE, COLOR_MODE=color, START_POINT_MODE='all', OVERLAP_PENALTY=True)
    res, _ = _run(extra, {}, 512, 60, 128, cuda_device, pack=pack)
    assert res['episodes'] >= 0


def test_principal_axes_other_than_1_2_continuous_grid(cuda_device):
    from pack_util import permuted_pack
    pack = permuted_pack(PartPack.for_part(0), 'axes02')
    res, _ = _run(dict(BASE, START_POINT_MODE='edge'), dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4),
                  256, 40, 96, cuda_device, pack=pack, seeded_resets=False)
