"""CPU: the C restatement of the load-time texel rasterisation (oracle_rasterize) against the part
packs minted from the reference's own Part.preprocess -- same texel set, bit-identical positions."""
import numpy as np
import pytest

from paintrl_b200.partpack import PartPack


@pytest.mark.parametrize('part_no', [0, 1, 5, 9])
def test_oracle_rasteriser_reproduces_reference_packs(part_no):
    from oracle.oracle import rasterize
    pack = PartPack.for_part(part_no)
    a = pack.arrays
    assert np.array_equal(a['tri_b'] - a['tri_a'], a['tri_v0']) and np.array_equal(a['tri_c'] - a['tri_a'], a['tri_v1'])
    ij, pos, owner = rasterize(a['tri_a'], a['tri_b'], a['tri_c'], a['tri_uv'], pack.width, pack.height)
    order = np.lexsort((pack.front_ij[:, 1], pack.front_ij[:, 0]))
    assert np.array_equal(ij, pack.front_ij[order])
    assert np.array_equal(pos, pack.front_pos[order])
    assert owner.min() >= 0 and (owner >> 2).max() < a['tri_a'].shape[0]


def test_retextured_pack_properties():
    from oracle.oracle import retextured_pack
    pack = PartPack.for_part(0)
    big = retextured_pack(pack, 480, 480)
    assert big.width == 480 and big.max_points == pack.max_points * 4
    assert 3.7 * pack.n_texels < big.n_texels < 4.3 * pack.n_texels
    corners = np.concatenate([pack.arrays[k] for k in ('tri_a', 'tri_b', 'tri_c')])
    lo, hi = corners.min(0) - 1e-9, corners.max(0) + 1e-9
    assert (big.front_pos >= lo).all() and (big.front_pos <= hi).all()
    assert np.array_equal(big.start_points('all'), pack.start_points('all'))
    assert big.status_init('RGB') == 191 and big.status_init('HSI') == 255
    assert len(np.unique(big.front_ij[:, 0].astype(np.int64) * 480 + big.front_ij[:, 1])) == big.n_texels
    tiny = retextured_pack(pack, 1, 1)
    assert tiny.n_texels == 1
