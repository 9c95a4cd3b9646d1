"""Constant per-part tables ("part pack") consumed by the batched paint step.

The reference derives these once per environment at load time
(PaintRLEnv/bullet_paint_wrapper.py:622-648 texel tables, 599-620 kd-tree inputs, 816-832 /
922-963 silhouette table, 740-809 start points).  A pack is the
frozen output of that preprocessing for one (part, texture size); it is stored as an `.npz`
under `paintrl_b200/data/partpacks/` and handed to the C ABI as a `PaintrlPartPack`.
"""
import ctypes
import json
import os

import numpy as np

from . import _capi

PACK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'partpacks')

# robot_gym_env.py:106-117 Part_Dict
PART_DICT = {
    0: ['door_test.urdf', 9148],
    1: ['square.urdf', 14350],
    2: ['door_lf.urdf', 0],
    3: ['door_lr.urdf', 0],
    4: ['door_rf.urdf', 0],
    5: ['door_rr.urdf', 17000],
    6: ['roof.urdf', 0],
    7: ['bonnet.urdf', 0],
    8: ['door_rr_big.urdf', 0],
    9: ['test.urdf', 9148],
}

START_POINT_MODES = ('fixed', 'anchor', 'edge', 'all')
_LOADED = {}      # packs built by loader.load_part in this process: (urdf path, mtime) -> PartPack


class PartPack(object):
    """Host-side (NumPy) view of one part's constant tables."""

    _ARRAYS = ('ranges', 'planes_n', 'planes_off', 'front_ij', 'front_pos', 'vertices',
               'vtri_start', 'vtri_idx', 'tri_a', 'tri_v0', 'tri_v1', 'tri_d00', 'tri_d01',
               'tri_d11', 'tri_inv_denom', 'tri_n', 'grid_lo', 'grid_hi')

    def __init__(self, meta, arrays):
        self.meta = dict(meta)
        self.arrays = arrays
        for name in self._ARRAYS:
            dtype = np.int32 if arrays[name].dtype.kind in 'iu' else np.float64
            setattr(self, name, np.ascontiguousarray(arrays[name], dtype=dtype))
        self.length_width_ratio = float(arrays['length_width_ratio'])
        self.width = int(self.meta['width'])
        self.height = int(self.meta['height'])
        self.axes = tuple(int(a) for a in self.meta['axes'])
        self.max_points = float(self.meta['max_points'])
        self.starts = {m: np.ascontiguousarray(arrays['start_' + m], dtype=np.float64)
                       for m in START_POINT_MODES if 'start_' + m in arrays}

    # ------------------------------------------------------------------------------ io
    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as z:
            arrays = {k: z[k] for k in z.files}
        meta = json.loads(str(arrays.pop('meta')))
        return cls(meta, arrays)

    @classmethod
    def for_part(cls, part_no, width=240, height=240, device=0, urdf_root=None):
        """Pack of `Part_Dict[part_no]` (robot_gym_env.py:106-117) at a texture size.

        Stored packs (`data/partpacks/`: the parts with a usable max-points entry) are read; sizes without a stored
        pack are derived from the part's 240x240 pack by rasterising its front triangles on GPU `device`
        (`retextured`).  A part without a stored pack is built from its URDF by `loader.load_part` -- found under
        `urdf_root`/urdf/painting like the reference finds it (robot_gym_env.py:273; `urdf_root` is PaintGymEnv's
        first argument, or the environment variable PAINTRL_URDF_ROOT)."""
        if part_no not in PART_DICT:
            raise KeyError(part_no)
        name = os.path.splitext(PART_DICT[part_no][0])[0]
        path = os.path.join(PACK_DIR, '%s_%dx%d.npz' % (name, width, height))
        if os.path.isfile(path):
            return cls.load(path)
        base = os.path.join(PACK_DIR, '%s_240x240.npz' % name)
        if os.path.isfile(base):
            return cls.load(base).retextured(width, height, device=device)
        urdf_root = urdf_root or os.environ.get('PAINTRL_URDF_ROOT')
        urdf = os.path.join(urdf_root, 'urdf', 'painting', PART_DICT[part_no][0]) if urdf_root else None
        if urdf is None or not os.path.isfile(urdf):
            raise FileNotFoundError(
                'no stored part pack for Part_NO=%d (%s) and no URDF to build one from (%s): pass urdf_root= / set '
                'PAINTRL_URDF_ROOT to the directory that holds urdf/painting/%s' % (part_no, name, urdf, PART_DICT[part_no][0]))
        from . import loader
        key = (os.path.realpath(urdf), os.path.getmtime(urdf))
        pack = _LOADED.get(key)
        if pack is None:                      # one load per part and process (1-7 s)
            pack = _LOADED[key] = loader.load_part(urdf, device=device)
        if (pack.width, pack.height) != (width, height):
            pack = pack.retextured(width, height, device=device)
        return pack

    def rasterize(self, width, height, device=0):
        """Front texels at a texture size: `(ij [N,2] int32, pos [N,3] float64)` sorted by (i, j), from
        the front triangles' corners and UVs by the reference's rule (Part.preprocess,
        bullet_paint_wrapper.py:604-618, 191-212) -- computed by `paintrl_rasterize_texels` on the GPU."""
        for key in ('tri_b', 'tri_c', 'tri_uv'):
            if key not in self.arrays:
                raise ValueError('this part pack carries no %s: rebuild it with paintrl_b200.loader.load_part' % key)
        lib = _capi.lib()
        tri = [np.ascontiguousarray(self.arrays[k], dtype=np.float64) for k in ('tri_a', 'tri_b', 'tri_c', 'tri_uv')]
        n_tris = tri[0].shape[0]
        n = ctypes.c_int32(0)
        ptrs = [ctypes.c_void_p(a.ctypes.data) for a in tri]
        _capi.check(lib.paintrl_rasterize_texels(*ptrs, n_tris, width, height, device, 0, None, None, ctypes.byref(n)))
        ij = np.zeros((n.value, 2), dtype=np.int32)
        pos = np.zeros((n.value, 3), dtype=np.float64)
        _capi.check(lib.paintrl_rasterize_texels(*ptrs, n_tris, width, height, device, n.value,
                                                 ctypes.c_void_p(ij.ctypes.data), ctypes.c_void_p(pos.ctypes.data),
                                                 ctypes.byref(n)))
        return ij, pos

    def retextured(self, width, height, device=0, texels=None):
        """The same part with a synthetic blank `width` x `height` texture (BASELINE config C4): geometry,
        hull, start points and silhouette table are texture-independent and kept; the front texels are
        re-rasterised; `max_points` (robot_gym_env.py:106-117, a texel count) scales with the texel
        density so that `finished` (robot_gym_env.py:292) keeps its meaning.
        `texels=(ij, pos)` supplies the rasterisation instead (the tests' CPU oracle does)."""
        ij, pos = texels if texels is not None else self.rasterize(width, height, device=device)
        arrays = dict(self.arrays)
        arrays['front_ij'], arrays['front_pos'] = ij, pos
        n = ij.shape[0]
        for key in ('status_init_rgb', 'status_init_hsi'):
            arrays[key] = np.full(n, self.arrays[key][0], dtype=np.int16)
        for key in ('texel_off', 'grid_cells_4', 'grid_cells_10'):
            arrays.pop(key, None)          # texture-size specific; offsets are recomputed on demand
        # synthetic blank texture: everything irrelevant (black, bullet_paint_wrapper.py:583-590) except the front texels
        off = np.minimum((ij[:, 0].astype(np.int64) + ij[:, 1].astype(np.int64) * width) * 3, width * height * 3 - 4)
        for key, init in (('init_texture_rgb', self.arrays['status_init_rgb'][0]), ('init_texture_hsi', self.arrays['status_init_hsi'][0])):
            tex = np.zeros(width * height * 3, dtype=np.uint8)
            for k in range(3):
                tex[off + k] = init
            arrays[key] = tex
        meta = dict(self.meta)
        scale = (width / float(self.width)) * (height / float(self.height))
        meta.update(width=int(width), height=int(height), max_points=float(self.max_points) * scale,
                    density=float(self.meta.get('density', 0.0)) * scale,
                    source='retextured from %s %dx%d by paintrl_rasterize_texels' % (
                        self.meta.get('part_name'), self.width, self.height))
        return PartPack(meta, arrays)

    def reordered_like(self, other):
        """This pack with its front texels in `other`'s order (same texel set required).  The reference keeps the
        texels in the iteration order of a CPython set (bullet_paint_wrapper.py:641); `loader.load_part` and the
        rasteriser emit them sorted by (i, j).  The order carries no meaning for the step; this aligns two packs
        of the same part for table-by-table comparison and for replaying traces recorded in the other order."""
        mine = self.front_ij[:, 0].astype(np.int64) * self.height + self.front_ij[:, 1]
        theirs = other.front_ij[:, 0].astype(np.int64) * other.height + other.front_ij[:, 1]
        if mine.shape != theirs.shape or not np.array_equal(np.sort(mine), np.sort(theirs)):
            raise ValueError('the packs do not hold the same front texels')
        order = np.argsort(mine, kind='stable')[np.searchsorted(np.sort(mine), theirs)]
        arrays = dict(self.arrays)
        for key in ('front_ij', 'front_pos', 'texel_off', 'status_init_rgb', 'status_init_hsi', 'grid_cells_4', 'grid_cells_10'):
            if key in arrays:
                arrays[key] = np.ascontiguousarray(arrays[key][order])
        return PartPack(self.meta, arrays)

    @property
    def n_texels(self):
        return self.front_pos.shape[0]

    def start_points(self, mode):
        """`Part.get_start_points(mode)` (bullet_paint_wrapper.py:749-783): [S,2,3] (pos, normal)."""
        if mode not in self.starts:
            raise ValueError('START_POINT_MODE %r not in %s' % (mode, sorted(self.starts)))
        return self.starts[mode]

    def status_init(self, color_mode):
        """First-channel value of a fresh front texel (bullet_paint_wrapper.py:586)."""
        key = 'status_init_rgb' if color_mode == 'RGB' else 'status_init_hsi'
        values = np.unique(self.arrays[key])
        if values.size != 1:
            raise ValueError('non-uniform initial front colour: %s' % values)
        return int(values[0])

    def init_texture(self, color_mode):
        key = 'init_texture_rgb' if color_mode == 'RGB' else 'init_texture_hsi'
        return np.array(self.arrays[key], dtype=np.uint8)

    def texel_offsets(self):
        """Byte offset of every front texel in the texture plane: `Part.get_texel(i, j)`
        (bullet_paint_wrapper.py:505-506) = min((i + j * W) * 3, len(texels) - 4)."""
        if 'texel_off' in self.arrays:
            return np.asarray(self.arrays['texel_off'], dtype=np.int64)
        ij = self.front_ij.astype(np.int64)
        return np.minimum((ij[:, 0] + ij[:, 1] * self.width) * 3, self.width * self.height * 3 - 4)

    def compose_texture(self, status, color_mode):
        """The reference's whole `Part.texels` array (bullet_paint_wrapper.py:467) rebuilt from the front-texel
        status plane the engine keeps (`paintrl_get_state`): the labelled initial texture with, per front
        texel, the painted colour (255, 0, 0) where the first channel reached 255 (RGB, :358-365), or all three
        channels at the remaining thickness (HSI, :396-417; values below 0 are kept, as under the reference's
        integer texels).  int16 [W * H * 3]; `texture_image` gives the uint8 picture."""
        status = np.asarray(status).reshape(-1)
        if status.shape[0] != self.n_texels:
            raise ValueError('expected %d status values' % self.n_texels)
        tex = self.init_texture(color_mode).astype(np.int16)
        off = self.texel_offsets()
        if color_mode == 'RGB':
            painted = status == 255
            tex[off[painted]] = 255
            tex[off[painted] + 1] = 0
            tex[off[painted] + 2] = 0
        else:
            for k in range(3):
                tex[off + k] = status.astype(np.int16)
        return tex

    def texture_image(self, status, color_mode):
        """`get_texture_image` (bullet_paint_wrapper.py:18-21, 737-738) as a uint8 array [W, H, 3]."""
        tex = self.compose_texture(status, color_mode)
        return (tex & 0xff).astype(np.uint8).reshape(self.width, self.height, 3)

    def nn_representatives(self):
        """For the normal paint method (Part.paint, bullet_paint_wrapper.py:562-566): per front texel, the member of its
        group of texels with the SAME 3-D position that the reference's `cKDTree(pixel_positions).query(point, k=1)`
        reports.  The kd-tree scans a leaf in its own index order and keeps the first of equal distances, so the twin
        it returns is a constant of the tree -- read here from a tree built the way the reference builds it
        (bullet_paint_wrapper.py:619-620, default parameters).  Identity where positions are unique."""
        cached = getattr(self, '_nn_rep', None)
        if cached is not None:
            return cached
        rep = np.arange(self.n_texels, dtype=np.int32)
        uniq, inverse, counts = np.unique(self.front_pos, axis=0, return_inverse=True, return_counts=True)
        groups = np.flatnonzero(counts > 1)
        if groups.size:
            from scipy.spatial import cKDTree
            tree = cKDTree(self.front_pos)
            winners = tree.query(uniq[groups], k=1)[1]
            for g, w in zip(groups, winners):
                rep[np.flatnonzero(inverse.reshape(-1) == g)] = w
        self._nn_rep = rep
        return rep

    # ------------------------------------------------------------------------------ C view
    def to_c(self, start_mode, color_mode, with_nn_rep=False):
        """Build the `PaintrlPartPack` struct; returns (struct, keepalive)."""
        starts = self.start_points(start_mode)
        start_pos = np.ascontiguousarray(starts[:, 0, :])
        start_normal = np.ascontiguousarray(starts[:, 1, :])
        keep = [start_pos, start_normal]

        def dptr(a):
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

        def iptr(a):
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

        p = _capi.PaintrlPartPack()
        p.abi_version = _capi.PAINTRL_ABI_VERSION
        p.width, p.height = self.width, self.height
        p.axis0, p.axis1 = self.axes
        p.n_texels = self.n_texels
        p.texel_pos = dptr(self.front_pos)
        p.texel_ij = iptr(self.front_ij)
        p.n_planes = self.planes_n.shape[0]
        p.plane_n = dptr(self.planes_n)
        p.plane_off = dptr(self.planes_off)
        p.n_vertices = self.vertices.shape[0]
        p.vertices = dptr(self.vertices)
        p.vtri_start = iptr(self.vtri_start)
        p.vtri_idx = iptr(self.vtri_idx)
        p.n_tris = self.tri_a.shape[0]
        p.tri_a, p.tri_v0, p.tri_v1 = dptr(self.tri_a), dptr(self.tri_v0), dptr(self.tri_v1)
        p.tri_d00, p.tri_d01, p.tri_d11 = dptr(self.tri_d00), dptr(self.tri_d01), dptr(self.tri_d11)
        p.tri_inv_denom = dptr(self.tri_inv_denom)
        p.tri_n = dptr(self.tri_n)
        p.range0_min, p.range0_max = float(self.ranges[0, 0]), float(self.ranges[0, 1])
        p.range1_min, p.range1_max = float(self.ranges[1, 0]), float(self.ranges[1, 1])
        p.length_width_ratio = self.length_width_ratio
        p.grid_granularity = self.grid_lo.shape[0]
        p.grid_lo, p.grid_hi = dptr(self.grid_lo), dptr(self.grid_hi)
        p.n_starts = start_pos.shape[0]
        p.start_pos = dptr(start_pos)
        p.start_normal = dptr(start_normal)
        p.status_init = self.status_init(color_mode)
        p.texel_nn_rep = iptr(self.nn_representatives()) if with_nn_rep else None
        return p, keep
