"""Shim S1 -- a headless stand-in for the `pybullet` wheel (TEST INFRASTRUCTURE ONLY).

PaintRL's own Python sources (`/root/reference/PaintRLEnv/*.py`) are imported verbatim by
`oracle/ref_env.py`; the un-vendored, un-pinned pybullet wheel they call into is absent from this
image, so this module supplies exactly the calls the paint-step path makes:

  step path : rayTestBatch (bullet_paint_wrapper.py:873), multiplyTransforms (robot.py:104,267,274),
              changeTexture (bullet_paint_wrapper.py:537, a render-only side effect -> no-op)
  load path : loadURDF, getBasePositionAndOrientation, loadTexture, changeVisualShape,
              getQuaternionFromEuler, connect/..., the debug-draw calls (all no-ops)

The arithmetic below is the frozen contract the CUDA engine is parity-checked against
(real Bullet's iterative convex cast is not bit-reproducible):

* collision shape of a URDF mesh without a concave flag = convex hull of the OBJ vertices
  (+ base position, identity orientation), margin 0; planes are the de-duplicated Qhull facet
  equations  n.x <= off.
* ray test = exact slab test against those half-spaces, FP64, every product and sum rounded
  separately (no FMA), left to right:
      d   = to - from
      den = nx*d0 + ny*d1 + nz*d2          num = off - (nx*f0 + ny*f1 + nz*f2)
      t   = num / den                      t_in = max t | den<0      t_out = min t | den>0
      den == 0 and num < 0  -> miss        hit <=> t_in <= t_out and 0 <= t_in <= 1
      hit point = from + d * t_in          (per component: one multiply, one add)
* multiplyTransforms = Bullet's btMatrix3x3::setRotation + btTransform::operator() in scalar FP64:
      s = 2/(x*x+y*y+z*z+w*w); xs,ys,zs = x*s,y*s,z*s; wx,wy,wz = w*xs,w*ys,w*zs; xx,xy,xz = x*xs,x*ys,x*zs;
      yy,yz,zz = y*ys,y*zs,z*zs;  R = [[1-(yy+zz), xy-wz, xz+wy],[xy+wz, 1-(xx+zz), yz-wx],[xz-wy, yz+wx, 1-(xx+yy)]]
      out_i = ((R_i0*v0 + R_i1*v1) + R_i2*v2) + pos_i
"""
import os
import xml.etree.ElementTree as _Et

import numpy as _np

# ----------------------------------------------------------------------------- constants
SHARED_MEMORY = 3
GUI = 1
DIRECT = 2
URDF_ENABLE_SLEEPING = 2048
URDF_USE_SELF_COLLISION = 8
POSITION_CONTROL = 2
ER_BULLET_HARDWARE_OPENGL = 131072
COV_ENABLE_RENDERING = 7
COV_ENABLE_GUI = 1


class error(Exception):
    """pybullet.error"""


# ----------------------------------------------------------------------------- world state
class _Body:
    def __init__(self, uid, path, pos, orn, normals=None, offsets=None):
        self.uid = uid
        self.path = path
        self.pos = tuple(float(v) for v in pos)
        self.orn = tuple(float(v) for v in orn)
        self.normals = normals      # (P, 3) float64 or None (no collision shape)
        self.offsets = offsets      # (P,)


class _World:
    def __init__(self):
        self.bodies = []
        self.textures = 0
        self.debug_items = 0
        self.search_path = ''
        self.ray_count = 0


_world = _World()


def _reset_world():
    global _world
    _world = _World()


def connect(mode, *args, **kwargs):
    # One PaintGymEnv per process is the reference's own constraint (module-global client and
    # `_urdf_cache`, bullet_paint_wrapper.py:9,15); every connect() starts from an empty world.
    if mode == SHARED_MEMORY:
        return -1
    _reset_world()
    return 0


def disconnect(*args, **kwargs):
    return None


def resetSimulation(*args, **kwargs):
    _reset_world()


def _noop(*args, **kwargs):
    return None


setTimeStep = _noop
setPhysicsEngineParameter = _noop
setGravity = _noop
stepSimulation = _noop
resetDebugVisualizerCamera = _noop
configureDebugVisualizer = _noop
changeVisualShape = _noop
changeTexture = _noop
removeAllUserDebugItems = _noop
removeUserDebugItem = _noop
resetJointState = _noop
setJointMotorControlArray = _noop


def setAdditionalSearchPath(path):
    _world.search_path = path


def _debug_item(*args, **kwargs):
    _world.debug_items += 1
    return _world.debug_items


addUserDebugLine = _debug_item
addUserDebugText = _debug_item


def loadTexture(path):
    _world.textures += 1
    return _world.textures


def getQuaternionFromEuler(euler):
    roll, pitch, yaw = (float(v) for v in euler)
    cr, sr = _np.cos(roll * 0.5), _np.sin(roll * 0.5)
    cp, sp = _np.cos(pitch * 0.5), _np.sin(pitch * 0.5)
    cy, sy = _np.cos(yaw * 0.5), _np.sin(yaw * 0.5)
    return (float(sr * cp * cy - cr * sp * sy), float(cr * sp * cy + sr * cp * sy),
            float(cr * cp * sy - sr * sp * cy), float(cr * cp * cy + sr * sp * sy))


def computeViewMatrixFromYawPitchRoll(*args, **kwargs):
    return tuple([0.0] * 16)


def getCameraImage(width, height, **kwargs):
    px = _np.zeros((height, width, 4), dtype=_np.uint8)
    return width, height, px, None, None


# ----------------------------------------------------------------------------- collision hull
def hull_planes(points):
    """De-duplicated outward half-spaces  n.x <= off  of the convex hull of `points`.

    Qhull returns one equation per (triangulated) facet; coplanar facets repeat the same
    equation up to round-off, so rows are merged when they agree to 1e-9 and the first
    occurrence's unrounded values are kept.
    """
    from scipy.spatial import ConvexHull
    eq = ConvexHull(_np.asarray(points, dtype=_np.float64)).equations
    key = _np.round(eq, 9) + 0.0
    _, first = _np.unique(key, axis=0, return_index=True)
    eq = eq[_np.sort(first)]
    normals = _np.ascontiguousarray(eq[:, :3])
    offsets = _np.ascontiguousarray(-eq[:, 3])
    return normals, offsets


def _read_collision_mesh(urdf_path):
    root = _Et.parse(urdf_path).getroot()
    meshes = root.findall('./link/collision/geometry/mesh')
    if not meshes:
        return None
    obj_path = os.path.join(os.path.dirname(urdf_path), meshes[0].get('filename'))
    vertices = []
    with open(obj_path, 'r') as f:
        for line in f:
            content = line.split()
            if content and content[0] == 'v':
                vertices.append([float(v) for v in content[1:4]])
    return vertices


def loadURDF(path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False,
             flags=0, **kwargs):
    uid = len(_world.bodies)
    full = path
    if not os.path.isfile(full):
        full = os.path.join(_world.search_path, path)
    if not os.path.isfile(full):
        # plane.urdf / kuka from pybullet_data: present in the real wheel, shape-less here.  The
        # ground plane (z = 0) can never be the closest hit of a paint-path ray that also hits
        # the part (the part hull lies wholly above it), so leaving it out changes no result.
        _world.bodies.append(_Body(uid, path, basePosition, baseOrientation))
        return uid
    if tuple(float(v) for v in baseOrientation) != (0.0, 0.0, 0.0, 1.0):
        raise error('shim supports identity base orientation only')
    vertices = _read_collision_mesh(full)
    if vertices is None:
        _world.bodies.append(_Body(uid, full, basePosition, baseOrientation))
        return uid
    base = [float(v) for v in basePosition]
    # same arithmetic as multiplyTransforms(base, identity, v, identity): R = I exactly
    pts = [list(multiplyTransforms(base, (0, 0, 0, 1), v, (0, 0, 0, 1))[0]) for v in vertices]
    normals, offsets = hull_planes(pts)
    _world.bodies.append(_Body(uid, full, basePosition, baseOrientation, normals, offsets))
    return uid


def getBasePositionAndOrientation(uid):
    body = _world.bodies[uid]
    return body.pos, body.orn


def get_collision_planes(uid):
    """Shim-only accessor used by the golden/part-pack exporter."""
    body = _world.bodies[uid]
    return body.normals, body.offsets


# ----------------------------------------------------------------------------- transforms
def _rotation_rows(q):
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return ((1.0 - (yy + zz), xy - wz, xz + wy),
            (xy + wz, 1.0 - (xx + zz), yz - wx),
            (xz - wy, yz + wx, 1.0 - (xx + yy)))


def multiplyTransforms(positionA, orientationA, positionB, orientationB):
    rows = _rotation_rows(orientationA)
    v0, v1, v2 = float(positionB[0]), float(positionB[1]), float(positionB[2])
    pos = tuple(((r[0] * v0 + r[1] * v1) + r[2] * v2) + float(p) for r, p in zip(rows, positionA))
    ax, ay, az, aw = (float(v) for v in orientationA)
    bx, by, bz, bw = (float(v) for v in orientationB)
    orn = (aw * bx + ax * bw + ay * bz - az * by,
           aw * by + ay * bw + az * bx - ax * bz,
           aw * bz + az * bw + ax * by - ay * bx,
           aw * bw - ax * bx - ay * by - az * bz)
    return pos, orn


# ----------------------------------------------------------------------------- ray test
_MISS = (-1, -1, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0))


def _ray_vs_body(body, frm, to):
    n, off = body.normals, body.offsets
    f0, f1, f2 = float(frm[0]), float(frm[1]), float(frm[2])
    d0, d1, d2 = float(to[0]) - f0, float(to[1]) - f1, float(to[2]) - f2
    den = n[:, 0] * d0 + n[:, 1] * d1 + n[:, 2] * d2
    num = off - (n[:, 0] * f0 + n[:, 1] * f1 + n[:, 2] * f2)
    par = den == 0.0
    if par.any() and (num[par] < 0.0).any():
        return None
    with _np.errstate(divide='ignore', invalid='ignore'):
        t = num / den
    ent = den < 0.0
    ext = den > 0.0
    t_in, k_in = -_np.inf, -1
    if ent.any():
        idx = _np.flatnonzero(ent)
        k = idx[_np.argmax(t[idx])]
        t_in, k_in = float(t[k]), int(k)
    t_out = float(t[ext].min()) if ext.any() else _np.inf
    if not (t_in <= t_out and 0.0 <= t_in <= 1.0):
        return None
    hit = (f0 + d0 * t_in, f1 + d1 * t_in, f2 + d2 * t_in)
    return t_in, hit, tuple(float(v) for v in n[k_in])


def rayTestBatch(rayFromPositions, rayToPositions, *args, **kwargs):
    results = []
    for frm, to in zip(rayFromPositions, rayToPositions):
        _world.ray_count += 1
        best = None
        for body in _world.bodies:
            if body.normals is None:
                continue
            res = _ray_vs_body(body, frm, to)
            if res is not None and (best is None or res[0] < best[1][0]):
                best = (body.uid, res)
        if best is None:
            results.append(_MISS)
        else:
            uid, (t_in, hit, normal) = best
            results.append((uid, -1, t_in, hit, normal))
    return results


def rayTest(rayFromPosition, rayToPosition, *args, **kwargs):
    return rayTestBatch([rayFromPosition], [rayToPosition])


# ----------------------------------------------------------------------------- arm (with_robot)
def _no_arm(*args, **kwargs):
    raise error('the shim has no articulated bodies: construct PaintGymEnv(with_robot=False)')


getNumJoints = _no_arm
getJointStates = _no_arm
getJointInfo = _no_arm
getLinkState = _no_arm
calculateInverseKinematics = _no_arm
