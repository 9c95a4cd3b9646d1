"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/paintrl.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'paintrl.h')


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(paintrl_[a-z_0-9]+)\s*\(', text)
    return sorted(set(names))


def _lib():
    from paintrl_b200 import build
    build.build()
    from paintrl_b200 import _capi
    return _capi.lib()


def test_header_declares_the_documented_entry_points():
    names = _declared_functions()
    for required in ('paintrl_create', 'paintrl_destroy', 'paintrl_reset', 'paintrl_step', 'paintrl_step_host',
                     'paintrl_set_pose', 'paintrl_get_state', 'paintrl_set_state', 'paintrl_job_status',
                     'paintrl_last_error', 'paintrl_abi_version'):
        assert required in names


def test_library_exports_every_declared_symbol():
    lib = _lib()
    from paintrl_b200 import _capi
    declared = _declared_functions()
    for name in declared:
        assert hasattr(lib, name), 'libpaintrl_b200.so does not export %s' % name
    # the ctypes binding covers exactly the header
    assert sorted(_capi.SIGNATURES) == declared
    assert lib.paintrl_abi_version() == _capi.PAINTRL_ABI_VERSION


def test_struct_layouts_match_the_header_field_order():
    from paintrl_b200 import _capi
    text = open(HEADER).read()

    def fields(struct):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (struct, struct), text, flags=re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        out = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                out.append(re.findall(r'[A-Za-z_0-9]+', part)[-1])
        return out
    assert fields('PaintrlPartPack') == [f[0] for f in _capi.PaintrlPartPack._fields_]
    assert fields('PaintrlConfig') == [f[0] for f in _capi.PaintrlConfig._fields_]
    assert fields('PaintrlStats') == [f[0] for f in _capi.PaintrlStats._fields_]


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    lib = _lib()
    from paintrl_b200.config import EnvConfig
    from paintrl_b200.partpack import PartPack
    cfg = EnvConfig(None)
    pack = PartPack.for_part(0)
    cpack, keep1 = pack.to_c(cfg.start_point_mode, cfg.color_mode)
    ccfg, keep2 = cfg.to_c(pack.max_points)
    handle = ctypes.c_void_p()
    rc = lib.paintrl_create(ctypes.byref(cpack), ctypes.byref(ccfg), 4, 0, ctypes.byref(handle))
    assert rc == -2 and not handle.value           # PAINTRL_E_CUDA
    assert b'no CPU fallback' in lib.paintrl_last_error()
    # and the Python surface refuses as well
    from paintrl_b200.batched_env import BatchedPaintEnv
    with pytest.raises(RuntimeError):
        BatchedPaintEnv(4)


def test_create_rejects_bad_arguments():
    lib = _lib()
    handle = ctypes.c_void_p()
    assert lib.paintrl_create(None, None, 1, 0, ctypes.byref(handle)) == -1
    assert b'null' in lib.paintrl_last_error()
    from paintrl_b200 import _capi
    p, c = _capi.PaintrlPartPack(), _capi.PaintrlConfig()
    assert lib.paintrl_create(ctypes.byref(p), ctypes.byref(c), 1, 0, ctypes.byref(handle)) == -1
    assert b'ABI' in lib.paintrl_last_error()
    assert lib.paintrl_step(None, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.paintrl_num_envs(None) == 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under paintrl_b200/ or PaintRLEnv/ may import it."""
    bad = []
    for top in ('paintrl_b200', 'PaintRLEnv'):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dirpath, f)).read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M) or 'paint_oracle' in src:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_partpacks_are_consistent():
    from paintrl_b200.partpack import PartPack
    for part_no, n_front, n_planes in ((0, 9663, 734), (1, 14482, 10)):
        pk = PartPack.for_part(part_no)
        assert pk.n_texels == n_front and pk.planes_n.shape == (n_planes, 3)
        assert pk.front_ij.shape == (n_front, 2) and pk.front_pos.shape == (n_front, 3)
        assert pk.vtri_start[-1] == len(pk.vtri_idx) and pk.vtri_idx.max() < pk.tri_a.shape[0]
        assert pk.status_init('RGB') == 191 and pk.status_init('HSI') == 255
        assert set(pk.starts) == {'fixed', 'anchor', 'edge', 'all'}
        assert pk.starts['fixed'].shape == (1, 2, 3) and pk.starts['anchor'].shape == (4, 2, 3)
        # unit normals of the collision planes, unit start normals
        assert np.allclose(np.linalg.norm(pk.planes_n, axis=1), 1.0)
        assert np.allclose(np.linalg.norm(pk.starts['all'][:, 1, :], axis=1), 1.0)
