"""GPU parity for the load-time texel rasterisation (paintrl_rasterize_texels) and for steps on
re-textured parts (BASELINE config C4: 2048x2048 synthetic texture).

Pins: (1) at 240x240 the GPU rasteriser reproduces the texel set and the FP64 positions of the part
packs minted from the reference's own Part.preprocess (oracle/make_golden.py) bit for bit;
(2) at other sizes it equals the C restatement oracle_rasterize (itself pinned the same way by
tests/test_raster_oracle.py); (3) the step on a re-textured part equals the oracle's."""
import numpy as np
import pytest

from paintrl_b200.partpack import PartPack
from test_gpu_oracle_batch import BASE, _run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('part_no', [0, 1, 5, 9])
def test_rasteriser_reproduces_reference_packs(part_no, cuda_device):
    pack = PartPack.for_part(part_no)
    ij, pos = pack.rasterize(240, 240, device=0)
    assert ij.shape[0] == pack.n_texels
    order = np.lexsort((pack.front_ij[:, 1], pack.front_ij[:, 0]))      # the pack keeps Python's set order
    assert np.array_equal(ij, pack.front_ij[order])
    assert np.array_equal(pos, pack.front_pos[order])


@pytest.mark.parametrize('part_no,size', [(0, (2048, 2048)), (0, (517, 301)), (1, (1024, 1024)), (1, (64, 48)), (0, (1, 1))])
def test_rasteriser_matches_oracle(part_no, size, cuda_device):
    from oracle.oracle import rasterize
    pack = PartPack.for_part(part_no)
    a = pack.arrays
    o_ij, o_pos, _ = rasterize(a['tri_a'], a['tri_b'], a['tri_c'], a['tri_uv'], *size)
    g_ij, g_pos = pack.rasterize(*size, device=0)
    assert np.array_equal(g_ij, o_ij)
    assert np.array_equal(g_pos, o_pos)


def test_rasteriser_rejects_bad_arguments(cuda_device):
    from paintrl_b200 import _capi
    pack = PartPack.for_part(0)
    with pytest.raises(_capi.PaintrlError):
        pack.rasterize(0, 240)
    bare = PartPack(pack.meta, {k: v for k, v in pack.arrays.items() if k != 'tri_uv'})
    with pytest.raises(ValueError):
        bare.rasterize(240, 240)


def test_c4_door_2048_matches_oracle(cuda_device):
    """BASELINE configs[3] at full texture size, few environments (the brute-force oracle scans
    703 k texels per shot): continuous 2-D actions, grid-4 observation, every start point."""
    _run_case(dict(BASE, START_POINT_MODE='all'),
              dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4), 24, 14, cuda_device,
              texture=(2048, 2048), status_every=13)


def test_door_512_section_matches_oracle(cuda_device):
    _run_case(dict(BASE, START_POINT_MODE='edge'), dict(), 48, 40, cuda_device, texture=(512, 512))


def test_sheet_hsi_480_matches_oracle(cuda_device):
    _run_case(dict(BASE, Part_NO=1, COLOR_MODE='HSI', TURNING_PENALTY=True, OVERLAP_PENALTY=True, TERMINATION_MODE='hybrid'),
              dict(), 32, 30, cuda_device, texture=(480, 480))
