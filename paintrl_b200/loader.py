"""Part loader: the constant tables of a part ("part pack") straight from its URDF / OBJ / MTL / texture.

This is the load-time half of the reference (PaintRLEnv/bullet_paint_wrapper.py: `load_part` :1327-1335 and everything
under it -- OBJ / URDF parsing :1142-1253, principal axes :1294-1300, per-triangle constants :123-146, side
classification :1219-1229, vertex sets :599-620, silhouette table :906-963, hull and neighbour normal correction
:650-698, start points :740-809, labelling :579-597), re-derived here as array code so that the engine no longer depends
on packs minted by running the reference: `load_part(urdf)` produces the same `PartPack` the committed `.npz` files hold
(tests/test_loader.py compares every table bit for bit), and it opens the parts of `Part_Dict` that have no committed
pack (`Part_NO` 2, 3, 4, 6, 7, 8).

What runs where
  * host (NumPy / SciPy, as in the reference: `scipy.spatial.ConvexHull` and `cKDTree` are the reference's own load-time
    dependencies): parsing, triangle constants, sides, vertex sets, hull planes, silhouette scans (the marching rays
    of `_get_exact_boundary`, batched against the hull half-spaces with the slab arithmetic of the step's ray test),
    the two normal-correction passes (the second is sequential by definition: a corrected normal feeds the next
    triangle's average), start points;
  * GPU: the texel rasterisation (`paintrl_rasterize_texels`, csrc/paintrl_raster.cuh) -- or `texels=` supplied by
    the caller.

Arithmetic: `np.dot` / `np.linalg.norm` / `np.cross` are called on the same 3-vectors the reference calls them on
(OpenBLAS ddot contracts the sum into FMAs, so a vectorised einsum would not give the same bits); everything else is
plain FP64 in the reference's operation order.

The front texels come out sorted by (i, j) (the rasteriser's order); the reference's own order is the iteration order
of a CPython `set` of pixel tuples (bullet_paint_wrapper.py:641) and carries no meaning.  `PartPack.reordered_like`
aligns the two for comparisons.
"""
import math
import os
import xml.etree.ElementTree as ElementTree

import numpy as np

from .partpack import PART_DICT, PartPack

PAINT_RADIUS = 0.051                 # bullet_paint_wrapper.py:42
HOOK_DISTANCE = 0.1                  # bullet_paint_wrapper.py:443
GRID_GRANULARITY = 100               # bullet_paint_wrapper.py:447
PARKED = (10.0, 10.0, 10.0)          # Part.IRRELEVANT_POSE, bullet_paint_wrapper.py:445
BASE_POSITION = (-0.4, -0.6, 0.25)   # robot_gym_env.py:275
FRONT, BACK, OTHER = 1, 2, 3         # Side, bullet_paint_wrapper.py:46-50


# ------------------------------------------------------------------------------------------ files
def related_files(urdf_path):
    """(obj, texture) of a painting URDF: the visual mesh and the `map_Kd` of its MTL (bullet_paint_wrapper.py:1153-1176)."""
    root = os.path.dirname(urdf_path)

    def locate(path):
        if os.path.isfile(path):
            return path
        joined = os.path.join(root, path)
        return joined if os.path.isfile(joined) else None

    meshes = ElementTree.parse(urdf_path).getroot().findall('./link/visual/geometry/mesh')
    if not meshes:
        return None, None
    obj_rel = meshes[0].get('filename')
    stem, ext = os.path.splitext(obj_rel)
    mtl = locate(stem + '.mtl')
    if ext != '.obj' or not mtl:
        return None, None
    with open(mtl) as f:
        for line in f:
            if 'map_Kd' in line:
                return locate(obj_rel), locate(line.split(' ')[-1].strip())
    return None, None


def collision_mesh_file(urdf_path):
    meshes = ElementTree.parse(urdf_path).getroot().findall('./link/collision/geometry/mesh')
    return os.path.join(os.path.dirname(urdf_path), meshes[0].get('filename')) if meshes else None


def read_obj(path):
    """Vertices [V, 3], texture coordinates [VT, 2] (v flipped: 1 - v, :1199) and triangular faces as index arrays
    (faces with another vertex count are ignored like the reference does, :1238)."""
    v, vt, fv, ft = [], [], [], []
    with open(path) as f:
        for line in f:
            c = line.split()
            if not c or c[0] == 'vn':
                continue
            if c[0] == 'v':
                v.append([float(x) for x in c[1:]])
            elif c[0] == 'vt':
                vt.append([float(c[1]), 1 - float(c[2])])
            elif c[0] == 'f' and len(c) == 4:
                fv.append([int(x.split('/')[0]) - 1 for x in c[1:]])
                ft.append([int(x.split('/')[1]) - 1 for x in c[1:]])
    return v, vt, np.array(fv, dtype=np.int64).reshape(-1, 3), np.array(ft, dtype=np.int64).reshape(-1, 3)


def to_world(points, base):
    """`multiplyTransforms(base, identity, p, identity)` of shim S1 per point: with the identity quaternion the
    rotation rows are exactly the unit vectors, and ((1 * x + 0 * y) + 0 * z) + b leaves one rounding: x + b."""
    out = []
    for p in points:
        x, y, z = float(p[0]), float(p[1]), float(p[2])
        out.append([((1.0 * x + 0.0 * y) + 0.0 * z) + base[0], ((0.0 * x + 1.0 * y) + 0.0 * z) + base[1],
                    ((0.0 * x + 0.0 * y) + 1.0 * z) + base[2]])
    return out


# ------------------------------------------------------------------------------------------ geometry
def principal_axes(world):
    """bullet_paint_wrapper.py:1288-1300: drop the coordinate with the smallest extent."""
    ext = [max(p[k] for p in world) - min(p[k] for p in world) for k in range(3)]
    drop = ext.index(min(ext))
    return [k for k in range(3) if k != drop], drop


def included_angle(a, b):
    """bullet_paint_wrapper.py:1206-1216."""
    if list(a) == list(b):
        return 0
    d = np.dot(a, b)
    d = 1 if d > 1 else (-1 if d < -1 else d)
    return np.arccos(d)


def classify_side(normal, front_normal):
    """bullet_paint_wrapper.py:1219-1229 as `BarycentricInterpolator.set_side` calls it (:219): within 60 degrees of
    the front normal -> front, of its opposite -> back, else other."""
    limit = np.pi / 3
    if -limit <= included_angle(normal, front_normal) <= limit:
        return FRONT
    if -limit <= included_angle([-c for c in normal], front_normal) <= limit:
        return BACK
    return OTHER


class Triangles(object):
    """Per-face constants of `BarycentricInterpolator` (bullet_paint_wrapper.py:123-146, 263-272) for all faces, as arrays."""

    def __init__(self, world, faces, front_normal):
        w = np.asarray(world, dtype=np.float64).reshape(-1, 3)
        f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        n = len(f)
        self.a, self.b, self.c = w[f[:, 0]], w[f[:, 1]], w[f[:, 2]]
        self.v0, self.v1 = self.b - self.a, self.c - self.a            # np.subtract(b, a), np.subtract(c, a)

        def dots(x, y):       # np.dot per pair of 3-vectors: OpenBLAS ddot contracts, a vectorised sum would not
            return np.array([np.dot(x[t], y[t]) for t in range(n)], dtype=np.float64).reshape(n)

        def cross(x, y):      # np.cross, element by element: products and differences rounded one by one
            return np.stack([x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1], x[:, 2] * y[:, 0] - x[:, 0] * y[:, 2],
                             x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0]], axis=1)

        self.d00, self.d01, self.d11 = dots(self.v0, self.v0), dots(self.v0, self.v1), dots(self.v1, self.v1)
        denom = self.d00 * self.d11 - self.d01 * self.d01
        with np.errstate(divide='ignore', invalid='ignore'):
            self.inv_denom = np.where(denom != 0, 1.0 / np.where(denom != 0, denom, 1.0), 0.0)
            nrm = cross(self.v0, self.v1)
            self.area = np.sqrt(dots(nrm, nrm)) / 2                    # np.linalg.norm(x) = sqrt(x.dot(x))
            # face normal from the corner order (:263-272): the same differences, cross product and norm
            self.normal = nrm / np.sqrt(dots(nrm, nrm))[:, None]
        self.center = (self.a + self.b + self.c) / 3
        self.side = np.array([classify_side(list(self.normal[t]), front_normal) for t in range(n)], dtype=np.int64).reshape(n)
        self.area_valid = self.area >= 1e-4          # BarycentricInterpolator.MIN_AREA


def hull_half_spaces(points):
    """Collision shape of shim S1 (oracle/shims/pybullet.py hull_planes: what Bullet builds for a URDF mesh without a
    concave flag is the convex hull of its vertices): outward half-spaces n.x <= off, one per distinct Qhull facet
    plane (coplanar facets merged when their equations agree to 1e-9, first occurrence kept)."""
    from scipy.spatial import ConvexHull
    eq = ConvexHull(np.asarray(points, dtype=np.float64)).equations
    _, first = np.unique(np.round(eq, 9) + 0.0, axis=0, return_index=True)
    eq = eq[np.sort(first)]
    return np.ascontiguousarray(eq[:, :3]), np.ascontiguousarray(-eq[:, 3])


def rays_hit(normals, offsets, frm, to):
    """Shim S1's slab test (FP64, products and sums rounded one by one, left to right) for a batch of rays [R, 3]:
    boolean hit mask.  Same arithmetic as the step's ray test (csrc/paintrl_device.cuh slab_pass)."""
    d = to - frm
    den = (normals[None, :, 0] * d[:, None, 0] + normals[None, :, 1] * d[:, None, 1]) + normals[None, :, 2] * d[:, None, 2]
    num = offsets[None, :] - ((normals[None, :, 0] * frm[:, None, 0] + normals[None, :, 1] * frm[:, None, 1]) + normals[None, :, 2] * frm[:, None, 2])
    parallel_out = ((den == 0.0) & (num < 0.0)).any(axis=1)
    with np.errstate(divide='ignore', invalid='ignore'):
        t = num / den
    t_in = np.where(den < 0.0, t, -np.inf).max(axis=1)
    t_out = np.where(den > 0.0, t, np.inf).min(axis=1)
    return (~parallel_out) & (t_in <= t_out) & (0.0 <= t_in) & (t_in <= 1.0)


# ------------------------------------------------------------------------------------------ texture labelling
def read_texture(path):
    """Size and RGB bytes of the part's texture (`_cache_texture`, bullet_paint_wrapper.py:1313-1316; PIL like the
    reference: the picture's own bytes survive labelling only where a channel-0 value already equals the label's)."""
    from PIL import Image
    with Image.open(path) as img:
        width, height = img.size
        pixels = np.asarray(img.convert('RGB')).ravel().astype(np.uint8)
    return width, height, pixels


def label_texture(pixels, width, height, front_ij, back_ij, color_mode):
    """`Part._label_part` with render on (bullet_paint_wrapper.py:579-592): everything outside the two profiles
    black, the back profile green, the front profile grey (RGB: 0.75 -> 191) or white (HSI).  `init_part` goes
    through `RGBColorHandler.change_pixel` (:358-365), which leaves a texel alone when its first byte already
    equals the label's first byte, and `get_texel` clamps the last pixel's offset to len - 4 (:505-506) where it
    overlaps its neighbour -- both kept.  Pixels are labelled in (i, j) order inside each of the three groups
    (the reference walks the profiles in `set` order; the order only matters when the clamped last pixel and its
    neighbour are in the same profile)."""
    tex = np.array(pixels, dtype=np.uint8, copy=True)
    limit = tex.shape[0] - 4
    in_profile = np.zeros((width, height), dtype=bool)
    for ij in (front_ij, back_ij):
        if ij is not None and len(ij):
            in_profile[ij[:, 0], ij[:, 1]] = True
    rest = np.argwhere(~in_profile)                      # (i, j) with i major: the reference's target_pixels order
    front_color = (191, 191, 191) if color_mode == 'RGB' else (255, 255, 255)
    groups = [(rest, (0, 0, 0))]
    if back_ij is not None and len(back_ij):
        groups.append((np.asarray(back_ij), (0, 255, 0)))
    groups.append((np.asarray(front_ij), front_color))
    for ij, color in groups:
        ij = np.asarray(ij, dtype=np.int64)
        order = np.lexsort((ij[:, 1], ij[:, 0]))
        ij = ij[order]
        off = (ij[:, 0] + ij[:, 1] * width) * 3
        plain = off < limit - 1                          # texels whose three bytes are theirs alone
        o = off[plain]
        todo = o[tex[o] != color[0]]
        for k in range(3):
            tex[todo + k] = color[k]
        for o in off[~plain]:                            # the last two pixels of the plane, one by one
            o = min(int(o), limit)
            if tex[o] != color[0]:
                tex[o], tex[o + 1], tex[o + 2] = color
    return tex


# ------------------------------------------------------------------------------------------ hook points
def bary_coordinate(tri, t, point):
    """`BarycentricInterpolator._get_bary_coordinate` of triangle t (bullet_paint_wrapper.py:154-163)."""
    v2 = np.subtract(point, tri.a[t])
    d20 = np.dot(v2, tri.v0[t])
    d21 = np.dot(v2, tri.v1[t])
    v = (tri.d11[t] * d20 - tri.d01[t] * d21) * tri.inv_denom[t]
    w = (tri.d00[t] * d21 - tri.d01[t] * d20) * tri.inv_denom[t]
    u = 1.0 - v - w
    if tri.inv_denom[t] == 0:
        return -1, -1, -1
    return u, v, w


def closest_triangle(tri, incident, side, point):
    """`Part._get_closest_bary` (bullet_paint_wrapper.py:508-523)."""
    closest_uvw, closest = -1, None
    for t in incident:
        if tri.side[t] != side:
            continue
        u, v, w = bary_coordinate(tri, t, point)
        if 0 <= u <= 1 and 0 <= v <= 1 and 0 <= w <= 1:
            return t
        if closest is None:
            closest = t
        m = min(u, v, w)
        if m >= closest_uvw:
            closest_uvw, closest = m, t
    return closest


def point_along(point, length, normal):
    """`_get_point_along_normal` (bullet_paint_wrapper.py:56-58)."""
    return [a + b for a, b in zip(point, [i * length for i in normal])]


# ------------------------------------------------------------------------------------------ silhouette table
def march_host(planes_n, planes_off, points, is_min, proof, npa, steps_range, chunk=64):
    """`Part._get_exact_boundary` (bullet_paint_wrapper.py:906-920) for a batch of scans, NumPy: (boundary [n], found [n]).
    Same arithmetic as `paintrl_silhouette_march` (csrc/paintrl_raster.cuh), which the loader uses on the GPU."""
    n = len(points)
    boundary, found = np.zeros(n), np.zeros(n, dtype=bool)
    for k in range(n):
        base = np.array(points[k], dtype=np.float64)
        step = -1e-3 if is_min[k] else 1e-3
        s_np, e_np = float(base[npa]), float(base[npa])
        i = 0
        while i < steps_range and not found[k]:
            m = min(chunk, steps_range - i)
            frm = np.tile(base, (m, 1))
            to = np.tile(base, (m, 1))
            bounds = np.empty(m)
            for r in range(m):
                bounds[r] = base[proof] + (i + r) * step
                s_np -= 1
                e_np += 1
                frm[r, npa], to[r, npa] = s_np, e_np
            frm[:, proof] = bounds
            to[:, proof] = bounds
            miss = ~rays_hit(planes_n, planes_off, frm, to)
            if miss.any():
                boundary[k], found[k] = bounds[int(np.argmax(miss))], True
            i += m
    return boundary, found


def march_gpu(planes_n, planes_off, points, is_min, proof, npa, steps_range, device=0):
    """The same scans by `paintrl_silhouette_march`: one warp per scan on the GPU."""
    import ctypes
    from . import _capi
    lib = _capi.lib()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    mins = np.ascontiguousarray(is_min, dtype=np.int8)
    pn, po = np.ascontiguousarray(planes_n, dtype=np.float64), np.ascontiguousarray(planes_off, dtype=np.float64)
    boundary, found = np.zeros(len(pts)), np.zeros(len(pts), dtype=np.int8)
    _capi.check(lib.paintrl_silhouette_march(pn.ctypes.data, po.ctypes.data, len(po), pts.ctypes.data, mins.ctypes.data, len(pts),
                                             int(proof), int(npa), int(steps_range), int(device), boundary.ctypes.data, found.ctypes.data))
    return boundary, found.astype(bool)


class Silhouette(object):
    """`Part._set_grid_dict` + `_get_exact_boundary` (bullet_paint_wrapper.py:906-963) for one side: per row of the
    100-row grid along the second principal axis, the extent of the collision hull along the first one, found by
    marching a ray in 1 mm steps outwards from the row's extreme vertices until it misses.

    The reference mutates rows of `cKDTree.data` in the sparse-row branch (:944-947; shim S3 makes that a writable
    copy), which later comparisons of the sorted walk and `ConvHull.separate_by_side` then see: `data` is that copy.

    The walk itself is host work (it is a sequential scan of a sorted list); its scans are queued and marched in
    batches by `march` (the GPU kernel, or `march_host`) -- a batch ends where the sparse-row branch needs the
    previous row's result."""

    def __init__(self, masked_vertices, axes, non_principal, ranges, planes_n, planes_off, march=None):
        self.data = np.array(masked_vertices, dtype=np.float64, copy=True)
        self.ax1, self.ax2 = axes
        self.npa = non_principal
        self.ranges = ranges
        self.n, self.off = planes_n, planes_off
        self.march = march or march_host
        self.scans = 0
        self.batches = 0
        self.sparse_rows = 0
        self._run()

    def _flush(self, known, queue):
        if not queue:
            return
        steps_range = int((self.ranges[0][1] - self.ranges[0][0]) / abs(1e-3))
        boundary, found = self.march(self.n, self.off, np.array([q[2] for q in queue]), [q[1] for q in queue], self.ax1, self.npa,
                                     steps_range)
        self.batches += 1
        for (row, is_min, _), b, ok in zip(queue, boundary, found):
            if not ok:
                raise ValueError('silhouette scan of grid row %d never left the part (bullet_paint_wrapper.py:906-920)' % row)
            lo, hi = known[row]
            known[row] = (np.float64(b), hi) if is_min else (lo, np.float64(b))
        del queue[:]

    def _run(self):
        data, ax1, ax2 = self.data, self.ax1, self.ax2
        order = np.argsort(data[:, ax2], kind='stable')
        order = order[data[order, 0] != PARKED[0]]
        r10, r11 = self.ranges[1]
        step_size = (r11 - r10) / GRID_GRANULARITY
        left = right = int(order[0])
        traverse = 0
        known, queue = {}, []
        for i in range(GRID_GRANULARITY):
            cur = traverse
            step_max = r10 + (i + 1) * step_size
            ahead = np.flatnonzero(data[order[cur:], ax2] >= step_max)
            if ahead.size == 0:
                known[i] = (0, 0)
                continue
            index = cur + int(ahead[0])
            if index - cur <= 1:
                self.sparse_rows += 1
                new2 = step_max + 0.5 * step_size
                if (i - 1) not in known:
                    new1 = data[order[index], ax1]
                else:
                    self._flush(known, queue)           # the previous row's extent is an input here
                    new1 = (known[i - 1][0] + known[i - 1][1]) / 2
                for row in (left, right):
                    data[row, ax2] = new2
                for row in (left, right):
                    data[row, ax1] = new1
            else:
                target = order[cur:index]
                by1 = target[np.argsort(data[target, ax1], kind='stable')]
                left, right = int(by1[0]), int(by1[-1])
            known[i] = (None, None)
            queue.append((i, True, data[left].copy()))
            queue.append((i, False, data[right].copy()))
            self.scans += 2
            traverse = index + 1
        self._flush(known, queue)
        self.lo = np.array([known[i][0] for i in range(GRID_GRANULARITY)], dtype=np.float64)
        self.hi = np.array([known[i][1] for i in range(GRID_GRANULARITY)], dtype=np.float64)


def grid_index(value, ranges):
    """`Part._get_grid_index_2` (bullet_paint_wrapper.py:844-851)."""
    rel = (value - ranges[1][0]) / (ranges[1][1] - ranges[1][0])
    g = int(rel * GRID_GRANULARITY)
    return 0 if g < 0 else (GRID_GRANULARITY - 1 if g > GRID_GRANULARITY - 1 else g)


def grid_indices(values, ranges):
    rel = (np.asarray(values, dtype=np.float64) - ranges[1][0]) / (ranges[1][1] - ranges[1][0])
    g = np.trunc(rel * GRID_GRANULARITY).astype(np.int64)
    return np.clip(g, 0, GRID_GRANULARITY - 1)


def normalized_pose(points, axes, ranges, lo, hi):
    """`Part.get_normalized_pose` (bullet_paint_wrapper.py:965-978) for an array of points [N, 3] -> two arrays."""
    radius = PAINT_RADIUS
    a1, a2 = points[:, axes[0]], points[:, axes[1]]
    n2 = (a2 - ranges[1][0] + radius) / (ranges[1][1] - ranges[1][0] + 2 * radius)
    g = grid_indices(a2, ranges)
    glo, ghi = lo[g], hi[g]
    with np.errstate(divide='ignore', invalid='ignore'):
        n1 = np.where(ghi - glo == 0, 0.0, (a1 - glo + radius) / (ghi - glo + 2 * radius))
    return np.clip(n1, 0.0, 1.0), np.clip(n2, 0.0, 1.0)


# ------------------------------------------------------------------------------------------ normal correction
class Flat(object):
    """2-D `BarycentricInterpolator` (get_2d_bary, bullet_paint_wrapper.py:287-291) of many triangles at once."""

    def __init__(self, a, b, c):
        self.a = a
        self.v0 = b - a
        self.v1 = c - a
        n = len(a)
        self.d00 = np.array([np.dot(self.v0[k], self.v0[k]) for k in range(n)])
        self.d01 = np.array([np.dot(self.v0[k], self.v1[k]) for k in range(n)])
        self.d11 = np.array([np.dot(self.v1[k], self.v1[k]) for k in range(n)])
        denom = self.d00 * self.d11 - self.d01 * self.d01
        with np.errstate(divide='ignore'):
            self.inv = np.where(denom != 0, 1.0 / np.where(denom != 0, denom, 1.0), 0.0)

    def inside_exact(self, k, point):
        """`is_inside_triangle` of triangle k, the reference's own calls (np.dot contracts to an FMA)."""
        v2 = np.subtract(point, self.a[k])
        d20 = np.dot(v2, self.v0[k])
        d21 = np.dot(v2, self.v1[k])
        v = (self.d11[k] * d20 - self.d01[k] * d21) * self.inv[k]
        w = (self.d00[k] * d21 - self.d01[k] * d20) * self.inv[k]
        u = 1.0 - v - w
        if self.inv[k] == 0:
            return False
        return bool(0 <= u <= 1 and 0 <= v <= 1 and 0 <= w <= 1)

    def first_containing(self, points):
        """For every point [N, 2] the first triangle (in order) that contains it, or -1.  A plain FP64 evaluation
        decides every (point, triangle) pair whose barycentric coordinates are clear of 0 and 1 by a margin far above
        the difference between a contracted and a plain dot product; the rest go through `inside_exact`."""
        n, m = len(points), len(self.a)
        out = np.full(n, -1, dtype=np.int64)
        if m == 0:
            return out
        chunk = max(1, (1 << 22) // m)               # about 4 M (point, triangle) pairs at a time
        if n > chunk:
            for at in range(0, n, chunk):
                out[at:at + chunk] = self.first_containing(points[at:at + chunk])
            return out
        v2x = points[:, None, 0] - self.a[None, :, 0]
        v2y = points[:, None, 1] - self.a[None, :, 1]
        d20 = v2x * self.v0[None, :, 0] + v2y * self.v0[None, :, 1]
        d21 = v2x * self.v1[None, :, 0] + v2y * self.v1[None, :, 1]
        inv = self.inv[None, :]
        v = (self.d11[None, :] * d20 - self.d01[None, :] * d21) * inv
        w = (self.d00[None, :] * d21 - self.d01[None, :] * d20) * inv
        u = 1.0 - v - w
        scale = (np.abs(self.d11[None, :] * d20) + np.abs(self.d01[None, :] * d21)
                 + np.abs(self.d00[None, :] * d21) + np.abs(self.d01[None, :] * d20)) * np.abs(inv)
        tol = 1e-9 * (1.0 + scale)
        lo = np.minimum(np.minimum(u, v), w)
        hi = np.maximum(np.maximum(u, v), w)
        surely_in = (lo >= tol) & (hi <= 1.0 - tol) & (inv != 0)
        surely_out = (lo < -tol) | (hi > 1.0 + tol) | (inv == 0)
        maybe = ~surely_out                     # surely_in or undecided
        for p in range(n):
            for k in np.flatnonzero(maybe[p]):
                if surely_in[p, k] or self.inside_exact(k, points[p]):
                    out[p] = k
                    break
        return out


def included_angle_between(a, b):
    """`_get_included_angle` (bullet_paint_wrapper.py:1206-1216) without the equal-lists shortcut, which returns 0
    where the arccos gives 0 or a few 1e-8: below every threshold the callers compare against."""
    d = np.dot(a, b)
    d = 1 if d > 1 else (-1 if d < -1 else d)
    return np.arccos(d)


def correct_with_hull(tri, world, masked_data, side, front_normal, axes, ranges, lo, hi, normals):
    """`Part._correct_bary_normals_with_conv_hull` (bullet_paint_wrapper.py:650-660) with `ConvHull` (:61-104): a
    triangle of the painted side, away from the rim (normalised pose within (0.01, 0.99) both ways), whose normal is
    more than 30 degrees off the normal of the hull facet above its centre takes the facet's normal."""
    from scipy.spatial import ConvexHull
    simplices = ConvexHull(np.asarray(world, dtype=np.float64)).simplices
    keep = [s for s in simplices if int(np.sum(masked_data[s, 0] != PARKED[0])) >= 2]
    if not keep:
        return 0
    hull = Triangles(world, keep, front_normal)
    hull_normals = np.where((hull.side == side)[:, None], hull.normal, -hull.normal)
    pa = np.asarray(world, dtype=np.float64)
    keep = np.asarray(keep)
    flat = Flat(pa[keep[:, 0]][:, list(axes)], pa[keep[:, 1]][:, list(axes)], pa[keep[:, 2]][:, list(axes)])
    mine = np.flatnonzero(tri.side == side)
    n1, n2 = normalized_pose(tri.center[mine], axes, ranges, lo, hi)
    inner = mine[~((n1 <= 0.01) | (n1 >= 0.99) | (n2 <= 0.01) | (n2 >= 0.99))]
    above = flat.first_containing(tri.center[inner][:, list(axes)])
    corrected = 0
    for t, k in zip(inner, above):
        if k >= 0 and included_angle_between(normals[t], hull_normals[k]) > np.pi / 6:
            normals[t] = hull_normals[k]
            corrected += 1
    return corrected


def smooth_with_neighbours(tri, sides, normals):
    """`Part._smooth_bary_normals_with_neighbors` + `_smooth_normal` (bullet_paint_wrapper.py:662-698), side by side
    in profile order and triangle by triangle in file order: a triangle whose normal is more than 10 degrees off one
    of its 4 nearest same-side neighbours' takes the area-weighted mean normal of the triangles whose centres lie
    within the paint radius -- sequential by definition (a corrected normal feeds the triangles after it).  The
    neighbour lists come from the same `cKDTree` calls, whose traversal order fixes the order of the sum."""
    from scipy.spatial import cKDTree
    n = len(tri.side)
    smoothed = {}
    for side in sides:
        smoothed[side] = 0
        centers = np.where((tri.side == side)[:, None], tri.center, np.array(PARKED, dtype=np.float64)[None, :])
        tree = cKDTree(centers)
        mine = np.flatnonzero(tri.side == side)
        near = tree.query(tri.center[mine], k=min(5, n))[1].reshape(len(mine), -1)
        balls = tree.query_ball_point(tri.center[mine], PAINT_RADIUS, return_sorted=False)
        for row, t in enumerate(mine):
            for b in near[row]:
                if b == t or b >= n:
                    continue
                if abs(included_angle_between(normals[b], normals[t])) > np.pi / 18:
                    ball = np.array([k for k in balls[row] if k != t], dtype=np.int64)
                    if ball.size:
                        avg = np.average(tri.area[ball][:, None] * normals[ball], 0)
                        mag2 = sum(c * c for c in avg)
                        if abs(mag2 - 1.0) > 0.00001:
                            mag = np.sqrt(mag2)
                            avg = tuple(c / mag for c in avg)
                        normals[t] = avg
                        smoothed[side] += 1
                    break
    return smoothed


# ------------------------------------------------------------------------------------------ start points
def corner_points_and_ranges(world, axes):
    """`_get_corner_points_ranges` (bullet_paint_wrapper.py:1255-1285): the four extreme vertices along the two
    diagonals, pulled inwards by half a paint radius, and the extents along the principal axes.  Python's sort is
    stable: [0] is the first of equal minima, [-1] the last of equal maxima."""
    shrink = PAINT_RADIUS / 2
    w = np.asarray(world, dtype=np.float64)
    a0, a1 = axes

    def first_min(key):
        return int(np.argmin(key))

    def last_max(key):
        return int(len(key) - 1 - np.argmax(key[::-1]))

    diag, anti = w[:, a0] + w[:, a1], w[:, a0] - w[:, a1]
    points = []
    for idx, s0, s1 in ((first_min(diag), 1, 1), (last_max(diag), -1, -1), (first_min(anti), 1, -1), (last_max(anti), -1, 1)):
        p = list(world[idx])
        p[a0] += s0 * shrink if s0 > 0 else 0
        p[a0] -= shrink if s0 < 0 else 0
        p[a1] += shrink if s1 > 0 else 0
        p[a1] -= shrink if s1 < 0 else 0
        points.append(p)
    ranges = [[float(w[:, a0].min()), float(w[:, a0].max())], [float(w[:, a1].min()), float(w[:, a1].max())]]
    return points, ranges


def start_point_modes(tri, side, anchors, normals, axes, ranges, lo, hi):
    """`Part.get_start_points` for its four modes (bullet_paint_wrapper.py:749-809)."""
    shrink = PAINT_RADIUS / 2
    a0, a1 = axes
    axis2 = [p[0][a1] for p in anchors]
    a2max, a2min = max(axis2), min(axis2)
    extra = []
    for t in np.flatnonzero((tri.side == side) & tri.area_valid):
        center = [float(c) for c in tri.center[t]]
        lo_t, hi_t = lo[grid_index(center[a1], ranges)], hi[grid_index(center[a1], ranges)]
        if center[a0] - lo_t >= shrink and hi_t - center[a0] >= shrink and a2min <= center[a1] <= a2max:
            extra.append([point_along(center, HOOK_DISTANCE, normals[t]), [-c for c in normals[t]]])
    # edge mode (:785-809): per grid row of the hook point, the outermost candidates when they lie within 15 % of
    # the row's extent from its silhouette; the lowest and highest rows whole
    rows = {}
    for point, orn in extra:
        rows.setdefault(grid_index(point[a1], ranges), []).append([point, orn])
    edge = []
    if rows:
        top, bottom = max(rows), min(rows)
        with np.errstate(divide='ignore', invalid='ignore'):
            for g, members in rows.items():
                if g in (top, bottom):
                    edge.extend(members)
                    continue
                ordered = sorted(members, key=lambda v: v[0][a0])
                extent = np.float64(hi[g] - lo[g])
                if (ordered[0][0][a0] - lo[g]) / extent < 0.15:
                    edge.append(ordered[0])
                if (hi[g] - ordered[-1][0][a0]) / extent < 0.15:
                    edge.append(ordered[-1])

    def pack(items):
        return np.array([[list(p), list(n)] for p, n in items], dtype=np.float64).reshape(-1, 2, 3)

    return {'fixed': pack(anchors[:1]), 'anchor': pack(anchors), 'edge': pack(anchors + edge), 'all': pack(anchors + extra)}


# ------------------------------------------------------------------------------------------ the loader
def rasterize_side(tri, uv, faces_of_side, width, height, device):
    """Texels of one side on the GPU (`paintrl_rasterize_texels`): (ij sorted by (i, j), pos)."""
    import ctypes
    from . import _capi
    lib = _capi.lib()
    arrs = [np.ascontiguousarray(x[faces_of_side], dtype=np.float64) for x in (tri.a, tri.b, tri.c, uv)]
    ptrs = [ctypes.c_void_p(a.ctypes.data) for a in arrs]
    n = ctypes.c_int32(0)
    _capi.check(lib.paintrl_rasterize_texels(*ptrs, len(faces_of_side), width, height, device, 0, None, None, ctypes.byref(n)))
    ij = np.zeros((n.value, 2), dtype=np.int32)
    pos = np.zeros((n.value, 3), dtype=np.float64)
    _capi.check(lib.paintrl_rasterize_texels(*ptrs, len(faces_of_side), width, height, device, n.value,
                                             ctypes.c_void_p(ij.ctypes.data), ctypes.c_void_p(pos.ctypes.data), ctypes.byref(n)))
    return ij, pos


def load_part(urdf_path, max_points=None, part_no=None, base_position=BASE_POSITION, texture_size=None, device=0,
              rasterizer=None):
    """`load_part` + `PaintGymEnv._load_environment`'s part half (bullet_paint_wrapper.py:1327-1335,
    robot_gym_env.py:271-279) -> `PartPack` for the painted (front) side.

    `max_points`: `Part_Dict`'s second column (robot_gym_env.py:106-117); looked up from the URDF's file name when
    omitted.  `texture_size=(W, H)` replaces the texture by a blank one of that size (BASELINE config C4).
    `rasterizer(tri_a, tri_b, tri_c, tri_uv, width, height) -> (ij, pos)` replaces the GPU rasteriser and switches the
    silhouette scans to their NumPy form as well (machines without a GPU: the CPU tests)."""
    urdf_path = os.path.abspath(urdf_path)
    name = os.path.splitext(os.path.basename(urdf_path))[0]
    if max_points is None or part_no is None:
        for no, (fname, pts) in PART_DICT.items():
            if fname == os.path.basename(urdf_path):
                part_no = no if part_no is None else part_no
                max_points = pts if max_points is None else max_points
        if max_points is None:
            raise ValueError('max_points: %s is not in Part_Dict (robot_gym_env.py:106-117)' % os.path.basename(urdf_path))
    obj_path, texture_path = related_files(urdf_path)
    if not texture_path:
        raise FileNotFoundError('Make sure that the .obj file is processed by Blender!')     # bullet_paint_wrapper.py:1331
    width, height, pixels = read_texture(texture_path)
    if texture_size is not None:
        width, height = int(texture_size[0]), int(texture_size[1])
        pixels = np.zeros(width * height * 3, dtype=np.uint8)
    base = [float(b) for b in base_position]

    v, vt, fv, ft = read_obj(obj_path)
    world = to_world(v, base)
    axes, non_principal = principal_axes(world)
    front_normal = [1 if k == non_principal else 0 for k in range(3)]
    tri = Triangles(world, fv, front_normal)
    wa = np.asarray(world, dtype=np.float64)
    uv = np.asarray(vt, dtype=np.float64)[ft]                                    # [T, 3, 2]
    side = FRONT
    # profile order of the sides (Part.preprocess :626-635): first appearance in the face list, `other` dropped
    seen = []
    for s in tri.side:
        if s not in seen:
            seen.append(int(s))
    sides = [s for s in seen if s in (FRONT, BACK)]
    if side not in sides:
        raise ValueError('%s has no triangle facing the front normal %s' % (name, front_normal))

    # texels (Part.preprocess :622-648, BarycentricInterpolator.get_uv_pixels :191-212)
    gpu_stages = rasterizer is None       # the product path: rasteriser and silhouette scans on the GPU
    if rasterizer is None:
        texels = {s: rasterize_side(tri, uv, np.flatnonzero(tri.side == s), width, height, device) for s in sides}
    else:
        texels = {}
        for s in sides:
            f = np.flatnonzero(tri.side == s)
            texels[s] = rasterizer(tri.a[f], tri.b[f], tri.c[f], uv[f], width, height)[:2]
    front_ij, front_pos = texels[side]
    back_ij = texels[BACK][0] if BACK in texels else None
    init = {mode: label_texture(pixels, width, height, front_ij, back_ij, mode) for mode in ('RGB', 'HSI')}
    texel_off = np.minimum((front_ij[:, 0].astype(np.int64) + front_ij[:, 1].astype(np.int64) * width) * 3, width * height * 3 - 4)

    # side-masked vertex sets (Part._build_kd_tree :599-620)
    masked = {}
    incident = [[] for _ in range(len(world))]
    for t, face in enumerate(fv):
        for k in face:
            incident[k].append(t)
    for s in sides:
        m = wa.copy()
        for k, faces in enumerate(incident):
            if faces and not any(tri.side[t] == s for t in faces):
                m[k] = PARKED
        masked[s] = m

    # anchors (set_start_points :740-747, before any normal correction) and ranges
    from scipy.spatial import cKDTree
    corners, ranges = corner_points_and_ranges(world, axes)
    normals = tri.normal.copy()                     # [T, 3], corrected in place below
    vertex_tree = cKDTree(masked[side])
    anchors = []
    for point in corners:
        nearest = int(vertex_tree.query(point, k=1)[1])
        t = closest_triangle(tri, incident[nearest], side, point)
        if t is not None:
            anchors.append([point_along(point, HOOK_DISTANCE, normals[t]), [-c for c in normals[t]]])

    # collision hull (shim S1) and the silhouette table (Part.postprocess :816-821)
    collision = collision_mesh_file(urdf_path)
    if collision is None:
        raise ValueError('%s has no collision mesh' % urdf_path)
    planes_n, planes_off = hull_half_spaces(to_world(read_obj(collision)[0], base))
    length_width_ratio = (ranges[0][1] - ranges[0][0]) / (ranges[1][1] - ranges[1][0])
    march = (lambda *a: march_gpu(*a, device=device)) if gpu_stages else march_host
    silhouette = Silhouette(masked[side], axes, non_principal, ranges, planes_n, planes_off, march=march)
    lo, hi = silhouette.lo, silhouette.hi

    # normal correction (:650-698)
    hull_corrected = correct_with_hull(tri, world, silhouette.data, side, front_normal, axes, ranges, lo, hi, normals)
    smoothed = smooth_with_neighbours(tri, sides, normals)

    starts = start_point_modes(tri, side, anchors, normals, axes, ranges, lo, hi)

    # density (Part.get_density :834-839)
    area = 0
    for g in range(GRID_GRANULARITY):
        area += ((ranges[1][1] - ranges[1][0]) / GRID_GRANULARITY) * (hi[g] - lo[g])
    density = len(front_ij) / area

    front = np.flatnonzero(tri.side == side)
    remap = -np.ones(len(fv), dtype=np.int64)
    remap[front] = np.arange(len(front))
    vtri_start = np.zeros(len(world) + 1, dtype=np.int32)
    vtri_idx = []
    for k, faces in enumerate(incident):
        vtri_idx.extend(int(remap[t]) for t in faces if tri.side[t] == side)
        vtri_start[k + 1] = len(vtri_idx)
    status = {mode: init[mode][texel_off].astype(np.int16) for mode in init}
    meta = {
        'part_name': name, 'part_no': part_no, 'urdf': os.path.basename(urdf_path), 'width': width, 'height': height,
        'axes': [int(a) for a in axes], 'non_principal_axis': int(non_principal), 'front_normal': front_normal,
        'base_position': base, 'max_points': max_points, 'density': float(density), 'grid_granularity': GRID_GRANULARITY,
        'source': 'paintrl_b200.loader.load_part',
        'loader_stats': {'silhouette_scans': silhouette.scans, 'silhouette_batches': silhouette.batches, 'sparse_grid_rows': silhouette.sparse_rows,
                         'gpu_stages': bool(gpu_stages),
                         'hull_corrected_normals': int(hull_corrected), 'smoothed_normals': int(smoothed[side]),
                         'triangles': int(len(fv)), 'front_triangles': int((tri.side == side).sum())},
    }
    if texture_size is not None:
        # a texel count scales with the texel density (see PartPack.retextured)
        tw, th, _ = read_texture(texture_path)
        meta['max_points'] = float(max_points) * (width / float(tw)) * (height / float(th))
    arrays = dict(
        ranges=np.array(ranges, dtype=np.float64), length_width_ratio=np.float64(length_width_ratio),
        planes_n=planes_n, planes_off=planes_off,
        front_ij=np.asarray(front_ij, dtype=np.int32), front_pos=np.asarray(front_pos, dtype=np.float64), texel_off=texel_off,
        status_init_rgb=status['RGB'], status_init_hsi=status['HSI'],
        init_texture_rgb=init['RGB'], init_texture_hsi=init['HSI'],
        vertices=masked[side], vtri_start=vtri_start, vtri_idx=np.array(vtri_idx, dtype=np.int32),
        tri_id=front.astype(np.int32),
        tri_a=tri.a[front], tri_v0=tri.v0[front], tri_v1=tri.v1[front], tri_b=tri.b[front], tri_c=tri.c[front],
        tri_uv=uv[front], tri_d00=tri.d00[front], tri_d01=tri.d01[front], tri_d11=tri.d11[front],
        tri_inv_denom=tri.inv_denom[front], tri_n=normals[front],
        grid_lo=lo, grid_hi=hi,
        start_fixed=starts['fixed'], start_anchor=starts['anchor'], start_edge=starts['edge'], start_all=starts['all'],
    )
    return PartPack(meta, arrays)
