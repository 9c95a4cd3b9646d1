"""Mint part packs and golden step traces from the VERBATIM reference under shims S1-S5.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # everything (a few minutes, one process per job)
    python -m oracle.make_golden packs      # only paintrl_b200/data/partpacks/*.npz
    python -m oracle.make_golden g1_door_zigzag ...

Outputs (committed; they are what travels to the GPU box, the reference cannot):
    paintrl_b200/data/partpacks/<part>_<W>x<H>.npz   constant per-part tables
    tests/golden/<name>.npz                           action traces + per-step reference outputs

The reference holds no golden vectors of its own ("parity unpinned" upstream); these
files are outputs of the reference's own code run here, which is what pins the C restatement
(oracle/paint_oracle.c) and, through it and directly, the CUDA engine.
"""
import json
import multiprocessing as mp
import os
import random
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACK_DIR = os.path.join(ROOT, 'paintrl_b200', 'data', 'partpacks')
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

PART_NAMES = {0: 'door_test', 1: 'square', 5: 'door_rr', 9: 'test'}
MAX_POINTS = {0: 9148, 1: 14350, 5: 17000, 9: 9148}      # robot_gym_env.py:106-117 (the other parts carry 0: unusable)


# ------------------------------------------------------------------------------ part packs
# the repository's own synthetic part (tests/data/make_synthetic_part.py), loaded by the reference like one of its own
SYNTHETIC_ROOT = os.path.join(ROOT, 'tests', 'data')
SYNTHETIC_PARTS = {100: ['bulge.urdf', 5200]}
SYNTHETIC_PACK_DIR = os.path.join(ROOT, 'tests', 'golden')


def export_partpack(part_no, pack_dir=None):
    from oracle.ref_env import ReferenceEnv
    kw = {}
    if part_no in SYNTHETIC_PARTS:
        kw = dict(urdf_root=SYNTHETIC_ROOT, extra_parts=SYNTHETIC_PARTS)
        PART_NAMES[part_no] = os.path.splitext(SYNTHETIC_PARTS[part_no][0])[0]
        MAX_POINTS[part_no] = SYNTHETIC_PARTS[part_no][1]
        pack_dir = pack_dir or SYNTHETIC_PACK_DIR
    pack_dir = pack_dir or PACK_DIR
    ref = ReferenceEnv({'Part_NO': part_no, 'START_POINT_MODE': 'anchor', 'COLOR_MODE': 'RGB'}, **kw)
    shim = sys.modules['pybullet']          # the S1 shim instance this env was built against
    part, bpw = ref.part, ref.bpw
    side = part.side
    ax0, ax1 = part.principal_axes
    normals, offsets = shim.get_collision_planes(ref.env._part_id)
    base_pos, _ = shim.getBasePositionAndOrientation(ref.env._part_id)

    profile = part.profile[side]
    front_ij = np.array(profile, dtype=np.int32)
    front_pos = np.array(part.pixel_kd_tree[side].tree_points, dtype=np.float64)
    assert front_pos.shape == (len(profile), 3)
    # the section observation iterates profile_dicts; same texels, same positions
    for k in (0, len(profile) // 2, len(profile) - 1):
        assert tuple(front_pos[k]) == tuple(part.profile_dicts[side][profile[k]])
    texel_off = np.array([part.get_texel(i, j) for (i, j) in profile], dtype=np.int64)
    assert len(np.unique(texel_off)) == len(texel_off), 'front texels alias in the byte plane'

    vertices = np.array(part.vertices_kd_tree[side].tree_points, dtype=np.float64)
    tri_index = {id(b): k for k, b in enumerate(part.bary_list)}
    front_tris = [k for k, b in enumerate(part.bary_list) if b.is_in_same_side(side)]
    remap = {k: j for j, k in enumerate(front_tris)}
    vtri_start = np.zeros(len(vertices) + 1, dtype=np.int32)
    vtri_idx = []
    for v in range(len(vertices)):
        for b in part.uv_map.get(v, []):
            if b.is_in_same_side(side):
                vtri_idx.append(remap[tri_index[id(b)]])
        vtri_start[v + 1] = len(vtri_idx)
    tris = [part.bary_list[k] for k in front_tris]

    def col(fn):
        return np.array([fn(b) for b in tris], dtype=np.float64)

    # start points for every mode, produced by the reference's own get_start_points
    anchors = [[list(p), list(n)] for p, n in part._start_points[side]]
    starts = {}
    for mode in ('fixed', 'anchor', 'edge', 'all'):
        part._start_points[side] = [[list(p), list(n)] for p, n in anchors]
        pts = part.get_start_points(mode)
        starts[mode] = np.array([[list(p), list(n)] for p, n in pts], dtype=np.float64)
    part._start_points[side] = anchors

    grid = part.grid_dict[side]
    grid_lo = np.array([grid[i][0] for i in range(part.GRID_GRANULARITY)], dtype=np.float64)
    grid_hi = np.array([grid[i][1] for i in range(part.GRID_GRANULARITY)], dtype=np.float64)

    # grid-observation cell lists of the reference for two granularities (checked by tests
    # against the product's own derivation from positions)
    grid_cells = {}
    for g in (4, 10):
        try:
            handler = bpw.GridObservation(part, g)
        except KeyError:        # door_lf: a texel left of its row's silhouette gives cell -1 (bullet_paint_wrapper.py:1085-1101)
            continue
        cell_of = {}
        for i in range(g):
            for j in range(g):
                for px in handler._grid_pixels[side][i][j]:
                    cell_of[px] = i * g + j
        grid_cells[g] = np.array([cell_of[tuple(px)] for px in profile], dtype=np.int32)

    init_rgb = np.array([int(v) for v in part.init_texture], dtype=np.uint8)
    status_rgb = ref.front_status()
    density = part.get_density()

    # HSI labelling differs only in the front colour (bullet_paint_wrapper.py:586)
    ref_h = ReferenceEnv({'Part_NO': part_no, 'START_POINT_MODE': 'anchor', 'COLOR_MODE': 'HSI'}, **kw)
    assert ref_h.part.profile[ref_h.part.side] == profile
    init_hsi = np.array([int(v) for v in ref_h.part.init_texture], dtype=np.uint8)
    status_hsi = ref_h.front_status()

    os.makedirs(pack_dir, exist_ok=True)
    name = '%s_%dx%d' % (PART_NAMES[part_no], part.texture_width, part.texture_height)
    meta = {
        'part_name': PART_NAMES[part_no], 'part_no': part_no, 'urdf': ref.env._part_name,
        'width': part.texture_width, 'height': part.texture_height,
        'axes': [int(ax0), int(ax1)], 'non_principal_axis': int(part.non_principal_axis),
        'front_normal': [int(v) for v in part.front_normal],
        'base_position': [float(v) for v in base_pos],
        'max_points': MAX_POINTS[part_no], 'density': float(density),
        'grid_granularity': int(part.GRID_GRANULARITY),
        'source': 'reference Part object under shims S1-S5 (oracle/make_golden.py)',
    }
    np.savez_compressed(
        os.path.join(pack_dir, name + '.npz'),
        meta=json.dumps(meta),
        ranges=np.array(part.ranges, dtype=np.float64),
        length_width_ratio=np.float64(part._length_width_ratio),
        planes_n=normals, planes_off=offsets,
        front_ij=front_ij, front_pos=front_pos, texel_off=texel_off,
        status_init_rgb=status_rgb.astype(np.int16), status_init_hsi=status_hsi.astype(np.int16),
        init_texture_rgb=init_rgb, init_texture_hsi=init_hsi,
        vertices=vertices, vtri_start=vtri_start, vtri_idx=np.array(vtri_idx, dtype=np.int32),
        tri_id=np.array(front_tris, dtype=np.int32),
        tri_a=col(lambda b: b._a), tri_v0=col(lambda b: b._v0), tri_v1=col(lambda b: b._v1),
        # raw corners and UVs of the front triangles, in bary_list order: the inputs of the texel
        # rasterisation (bullet_paint_wrapper.py:191-212), so packs at other texture sizes can be derived
        tri_b=col(lambda b: b._b), tri_c=col(lambda b: b._c),
        tri_uv=col(lambda b: [list(b._uva), list(b._uvb), list(b._uvc)]),
        tri_d00=col(lambda b: b._d00), tri_d01=col(lambda b: b._d01), tri_d11=col(lambda b: b._d11),
        tri_inv_denom=col(lambda b: b._inv_denom), tri_n=col(lambda b: b.get_normal()),
        grid_lo=grid_lo, grid_hi=grid_hi,
        start_fixed=starts['fixed'], start_anchor=starts['anchor'],
        start_edge=starts['edge'], start_all=starts['all'],
        **{'grid_cells_%d' % g: cells for g, cells in grid_cells.items()}
    )
    return name


# ------------------------------------------------------------------------------ traces
def _crc(status):
    return zlib.crc32(np.ascontiguousarray(status, dtype=np.int16).tobytes())


class _Recorder(object):
    def __init__(self, ref):
        self.ref = ref
        self.apply_log = []
        robot = ref.env.robot
        real_apply = robot.apply_action

        def logged_apply(action, part_id):
            out = real_apply(action, part_id)
            self.apply_log.append((float(out[0]), float(out[1])))
            return out

        robot.apply_action = logged_apply
        self.episodes = []

    def _snap(self):
        env, robot = self.ref.env, self.ref.env.robot
        pos, orn = self.ref.pose()
        return dict(pose=pos, quat=orn, total_reward=float(env._total_reward),
                    total_return=float(env._total_return), step_counter=int(env._step_counter),
                    term_counter=int(robot._terminate_counter), last_on_part=bool(robot._last_on_part),
                    terminate=bool(robot._terminate), angle_diff=float(robot.angle_diff),
                    last_angle=float(robot._last_turning_angle))

    def run_episode(self, start_index, policy, max_steps, set_pose=None):
        """policy(obs, t) -> action or None to stop.  Returns the episode record."""
        ref = self.ref
        obs = ref.reset(start_index)
        if set_pose is not None:
            ref.env.robot.reset(set_pose)           # spiral.py:28-38
            obs = ref.env._augmented_observation()
        rec = dict(start_index=-1 if start_index is None else start_index, actions=[],
                   obs=[np.asarray(obs, dtype=np.float64)], reward=[], penalty=[], actual=[],
                   done=[], rate=[], succeeded=[], crc=[], painted=[], snaps=[self._snap()])
        if set_pose is not None:
            rec['set_pose'] = np.array(set_pose, dtype=np.float64)
        for t in range(max_steps):
            action = policy(obs, t)
            if action is None:
                break
            obs, actual, done, info = ref.step(action)
            status = ref.front_status()
            rec['actions'].append(np.atleast_1d(np.asarray(action, dtype=np.float64)))
            rec['obs'].append(np.asarray(obs, dtype=np.float64))
            rec['reward'].append(float(info['reward']))
            rec['penalty'].append(float(info['penalty']))
            rec['actual'].append(float(actual))
            rec['done'].append(bool(done))
            rate, succ = self.apply_log[-1]
            rec['rate'].append(rate)
            rec['succeeded'].append(succ)
            rec['crc'].append(_crc(status))
            rec['painted'].append(int(np.count_nonzero(status != ref.part_init_value)))
            rec['snaps'].append(self._snap())
            if done:
                break
        rec['status_final'] = ref.front_status().astype(np.int16)
        self.episodes.append(rec)
        return rec


def _pack_episodes(episodes):
    out = {}
    n = len(episodes)
    lengths = np.array([len(e['actions']) for e in episodes], dtype=np.int32)
    tmax = int(lengths.max()) if n else 0
    adim = episodes[0]['actions'][0].shape[0] if tmax else 1
    odim = episodes[0]['obs'][0].shape[0]
    out['lengths'] = lengths
    out['start_index'] = np.array([e['start_index'] for e in episodes], dtype=np.int32)
    acts = np.zeros((n, tmax, adim), dtype=np.float64)
    obs = np.full((n, tmax + 1, odim), np.nan)
    scal = {k: np.zeros((n, tmax), dtype=np.float64) for k in
            ('reward', 'penalty', 'actual', 'rate', 'succeeded')}
    done = np.zeros((n, tmax), dtype=np.uint8)
    crc = np.zeros((n, tmax), dtype=np.int64)
    painted = np.zeros((n, tmax), dtype=np.int32)
    snap_keys = ('total_reward', 'total_return', 'angle_diff', 'last_angle')
    snap_int = ('step_counter', 'term_counter', 'last_on_part', 'terminate')
    snaps = {k: np.zeros((n, tmax + 1), dtype=np.float64) for k in snap_keys}
    snaps.update({k: np.zeros((n, tmax + 1), dtype=np.int32) for k in snap_int})
    pose = np.zeros((n, tmax + 1, 3))
    quat = np.zeros((n, tmax + 1, 4))
    for e, ep in enumerate(episodes):
        L = lengths[e]
        if L:
            acts[e, :L] = np.stack(ep['actions'])
        obs[e, :L + 1] = np.stack(ep['obs'])
        for k in scal:
            scal[k][e, :L] = ep[k]
        done[e, :L] = ep['done']
        crc[e, :L] = ep['crc']
        painted[e, :L] = ep['painted']
        for t, s in enumerate(ep['snaps']):
            pose[e, t] = s['pose']
            quat[e, t] = s['quat']
            for k in snap_keys + snap_int:
                snaps[k][e, t] = s[k]
    out.update(actions=acts, obs=obs, done=done, status_crc=crc, painted=painted, pose=pose,
               quat=quat, status_final=np.stack([e['status_final'] for e in episodes]))
    out.update(scal)
    out.update({'snap_' + k: v for k, v in snaps.items()})
    if 'set_pose' in episodes[0]:
        out['set_pose'] = np.stack([e['set_pose'] for e in episodes])
    return out


def _zigzag_policy_simple(state):
    """zigzag.py:77-104 (simple_rgb1_zigzag), keyed on the last observation entry."""
    def policy(obs, t):
        while True:
            if state['up']:
                if obs[-1] < 0.95:
                    return 1
                elif state['h'] < 2:
                    state['h'] += 1
                    return 0
                else:
                    state['h'] = 0
                    state['up'] = False
            else:
                if obs[-1] > 0.05:
                    return 3
                elif state['h'] < 2:
                    state['h'] += 1
                    return 0
                else:
                    state['h'] = 0
                    state['up'] = True
    return policy


def _zigzag_policy_discrete(state):
    """zigzag.py:22-62 (simple_rgb_zigzag), keyed on the discrete position code."""
    def policy(obs, t):
        while True:
            current_pos = 0 if obs[-1] == 0 else round(1 / obs[-1])
            if state['up']:
                if current_pos % 22 != 19:
                    return 1
                elif state['h'] < 2:
                    state['h'] += 1
                    return 0
                else:
                    state['h'] = 0
                    state['up'] = False
            else:
                if current_pos % 22 != 2:
                    return 3
                elif state['h'] < 2:
                    state['h'] += 1
                    return 0
                else:
                    state['h'] = 0
                    state['up'] = True
    return policy


def _zigzag_policy_hsi(state):
    """zigzag.py:150-191 (simple_hsi_zigzag): continuous 2-D actions."""
    def policy(obs, t):
        while True:
            if state['up']:
                if obs[1] < 0.95:
                    return [0, 1]
                elif state['h'] < 2:
                    state['h'] += 1
                    return [0.5, 0]
                else:
                    state['h'] = 0
                    state['up'] = False
            else:
                if obs[1] > 0.05:
                    return [0, -1]
                elif state['h'] < 2:
                    state['h'] += 1
                    return [0.5, 0]
                else:
                    state['h'] = 0
                    state['up'] = True
    return policy


def _random_discrete(rng, n):
    return lambda obs, t: int(rng.integers(0, n))


def _random_box(rng, dim):
    return lambda obs, t: [float(v) for v in rng.uniform(-1.0, 1.0, size=dim)]


JOBS = {}


def job(fn):
    JOBS[fn.__name__] = fn
    return fn


def _make(extra, **kw):
    from oracle.ref_env import ReferenceEnv
    ref = ReferenceEnv(extra, **kw)
    ref.part_init_value = 191 if ref.extra_config['COLOR_MODE'] == 'RGB' else 255
    return ref, _Recorder(ref)


def _save(name, ref, rec, kw, note):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cfg = dict(extra_config=ref.extra_config, note=note, n_start_points=len(ref.env._start_points),
               part=PART_NAMES[ref.extra_config['Part_NO']], **kw)
    extra = {}
    if kw.get('paint_method') == 'normal':
        # Robot._paint_plain as the reference built it (robot.py:244-249; the HSI table draws from `random`)
        extra['beam_plain'] = np.array(ref.env.robot._paint_plain, dtype=np.float64).reshape(-1, 3)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'), config=json.dumps(cfg),
                        **_pack_episodes(rec.episodes), **extra)
    return name


@job
def g1_door_zigzag():
    """BASELINE config C1: door, RGB, discrete-4, section-4, rollout start, 245-step zigzag."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=True)
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'anchor'}, **kw)
    rec.run_episode(None, _zigzag_policy_simple({'up': True, 'h': 0}), 245)
    return _save('g1_door_zigzag', ref, rec, kw, 'C1: zigzag.py:77-104 policy on obs[-1]')


@job
def g10_door_rr_random():
    """Part_NO 5 (door_rr, the third part with a usable max-points entry): RGB, anchor, random discrete-4."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=False)
    ref, rec = _make({'Part_NO': 5, 'START_POINT_MODE': 'anchor'}, **kw)
    rng = np.random.default_rng(510)
    for ep in range(3):
        rec.run_episode(ep % len(ref.env._start_points), _random_discrete(rng, 4), 60)
    return _save('g10_door_rr_random', ref, rec, kw, 'Part_NO 5, random discrete-4, 3 episodes of <= 60 steps')


@job
def g2_door_random():
    """C2 at small size: door, RGB, anchor starts, random discrete actions, late termination."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=False)
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'anchor'}, **kw)
    rng = np.random.default_rng(1234)
    for e in range(12):
        rec.run_episode(e % 4, _random_discrete(rng, 4), 245)
    return _save('g2_door_random', ref, rec, kw, 'C2 semantics, 12 episodes')


@job
def g3_sheet_hsi_hybrid():
    """C3 at small size: sheet, HSI, turning+overlap penalties, hybrid termination."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=False)
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'anchor', 'COLOR_MODE': 'HSI',
                      'TURNING_PENALTY': True, 'OVERLAP_PENALTY': True,
                      'TERMINATION_MODE': 'hybrid'}, **kw)
    rng = np.random.default_rng(1234)
    for e in range(16):
        rec.run_episode(e % 4, _random_discrete(rng, 4), 245)
    return _save('g3_sheet_hsi_hybrid', ref, rec, kw, 'C3 semantics, 16 episodes')


@job
def g3b_sheet_hsi_late():
    """sheet, HSI, penalties, late termination: long overlapping random walks (S5 saturation)."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=False)
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'anchor', 'COLOR_MODE': 'HSI',
                      'TURNING_PENALTY': True, 'OVERLAP_PENALTY': True,
                      'TERMINATION_MODE': 'late'}, **kw)
    rng = np.random.default_rng(77)
    # back-and-forth strokes over the same band force repeated coats -> thickness reaches <= 0
    seq = ([1] * 12 + [3] * 12) * 10
    rec.run_episode(0, lambda obs, t: seq[t] if t < len(seq) else None, 245)
    for e in range(3):
        rec.run_episode(1 + e, _random_discrete(rng, 4), 245)
    return _save('g3b_sheet_hsi_late', ref, rec, kw, 'HSI saturation + random, late')


@job
def g4_door_grid_continuous():
    """C4 at 240x240: door, RGB, continuous 2-D actions, grid-4 observation, all start points."""
    kw = dict(action_mode='continuous', action_shape=2, obs_mode='grid', obs_grad=4, rollout=False)
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'all'}, **kw)
    rng = np.random.default_rng(1234)
    n = len(ref.env._start_points)
    for e in range(10):
        rec.run_episode(int(rng.integers(0, n)), _random_box(rng, 2), 120)
    return _save('g4_door_grid_continuous', ref, rec, kw, 'C4 semantics at 240x240, 10 episodes')


@job
def g5_sheet_zigzag_discrete():
    """zigzag.py:22-62: sheet, RGB, fixed start, 'discrete' observation; ends by full coverage."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='discrete',
              obs_grad=4, rollout=True)
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'fixed'}, **kw)
    rec.run_episode(None, _zigzag_policy_discrete({'up': True, 'h': 0}), 245)
    return _save('g5_sheet_zigzag_discrete', ref, rec, kw, 'simple_rgb_zigzag')


@job
def g6_door_section8_early():
    """door, RGB, 1-D continuous actions, section-8 (atan2 sectors), early termination, edge
    starts, both penalties, discrete-20 in a second batch."""
    kw = dict(action_mode='continuous', action_shape=1, obs_mode='section', obs_grad=8, rollout=False)
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'edge', 'TURNING_PENALTY': True,
                      'OVERLAP_PENALTY': True, 'TERMINATION_MODE': 'early',
                      'Expected_Episode_Length': 400}, **kw)
    rng = np.random.default_rng(5)
    n = len(ref.env._start_points)
    for e in range(8):
        rec.run_episode(int(rng.integers(0, n)), _random_box(rng, 1), 80)
    return _save('g6_door_section8_early', ref, rec, kw, 'atan2 sectors, early termination')


@job
def g7_sheet_spiral_simple():
    """spiral.py:22-55: sheet, RGB, 'simple' observation, robot.reset(centre) then a spiral."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='simple',
              obs_grad=4, rollout=True)
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'all'}, **kw)
    sp = ref.env._start_points
    a1 = [p[0][1] for p in sp]
    a2 = [p[0][2] for p in sp]
    x = min(a1) + (max(a1) - min(a1)) / 2
    y = min(a2) + (max(a2) - min(a2)) / 2
    centre = [[sp[0][0][0], x, y], sp[0][1]]
    st = {'direction': 0, 'strait': 1, 'cur': 1}

    def policy(obs, t):
        st['cur'] -= 1
        a = st['direction'] % 4
        if st['cur'] == 0:
            st['strait'] += 1
            st['direction'] += 1
            st['cur'] = st['strait']
        return a
    rec.run_episode(None, policy, 245, set_pose=centre)
    return _save('g7_sheet_spiral_simple', ref, rec, kw, 'simple_rgb_spiral')


@job
def g8_sheet_hsi_zigzag_continuous():
    """zigzag.py:150-191: sheet, HSI, continuous 2-D actions, 'simple' observation."""
    kw = dict(action_mode='continuous', action_shape=2, obs_mode='simple', obs_grad=4, rollout=True)
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'fixed', 'COLOR_MODE': 'HSI'}, **kw)
    rec.run_episode(None, _zigzag_policy_hsi({'up': True, 'h': 0}), 245)
    return _save('g8_sheet_hsi_zigzag_continuous', ref, rec, kw, 'simple_hsi_zigzag')


@job
def g9_door_discrete20_grid10():
    """door, RGB, discrete-20 actions, grid-10 observation, fixed start, hybrid termination."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=20, obs_mode='grid',
              obs_grad=10, rollout=False)
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'fixed', 'TERMINATION_MODE': 'hybrid',
                      'Expected_Episode_Length': 2000, 'OVERLAP_PENALTY': True}, **kw)
    rng = np.random.default_rng(9)
    for e in range(4):
        rec.run_episode(0, _random_discrete(rng, 20), 150)
    return _save('g9_door_discrete20_grid10', ref, rec, kw, 'discrete-20, grid-10, hybrid')


@job
def g11_door_normal_rgb():
    """Robot.PAINT_METHOD = 'normal' (robot.py:172, 414-417): door, RGB, beam fan per shot, overlap penalty."""
    kw = dict(action_mode='discrete', action_shape=1, discrete_granularity=4, obs_mode='section',
              obs_grad=4, rollout=False, paint_method='normal')
    ref, rec = _make({'Part_NO': 0, 'START_POINT_MODE': 'anchor', 'OVERLAP_PENALTY': True}, **kw)
    rng = np.random.default_rng(11)
    rec.run_episode(0, _zigzag_policy_simple({'up': True, 'h': 0}), 40)
    for e in range(3):
        rec.run_episode(1 + e, _random_discrete(rng, 4), 40)
    return _save('g11_door_normal_rgb', ref, rec, kw, 'normal paint method, RGB, 4 episodes of <= 40 steps')


@job
def g12_sheet_normal_hsi():
    """Normal paint method with the HSI handler (beta-distributed beam rings, one subtraction per beam hit): sheet,
    both penalties, continuous 2-D actions."""
    kw = dict(action_mode='continuous', action_shape=2, obs_mode='section', obs_grad=4, rollout=False, paint_method='normal')
    ref, rec = _make({'Part_NO': 1, 'START_POINT_MODE': 'anchor', 'COLOR_MODE': 'HSI', 'TURNING_PENALTY': True,
                      'OVERLAP_PENALTY': True}, **kw)
    rng = np.random.default_rng(12)
    seq = ([[0, 1]] * 8 + [[0, -1]] * 8) * 2
    rec.run_episode(0, lambda obs, t: seq[t] if t < len(seq) else None, 40)
    for e in range(2):
        rec.run_episode(1 + e, _random_box(rng, 2), 30)
    return _save('g12_sheet_normal_hsi', ref, rec, kw, 'normal paint method, HSI, repeated coats + random walks')


def _run(name):
    random.seed(20261017)
    np.random.seed(20261017)
    if name.startswith('pack'):
        return export_partpack(int(name[4:]))
    return JOBS[name]()


UNPACKED_PARTS = {2: 'door_lf', 3: 'door_lr', 4: 'door_rf', 6: 'roof', 7: 'bonnet', 8: 'door_rr_big'}   # max points 0: no stored pack


def export_pack_digests(pack_dir=None):
    """tests/golden/pack_digests.json: per-table sha256 of the packs the REFERENCE makes of the parts that have no
    stored pack (`Part_Dict` max points 0); tests/test_loader.py holds the loader to them.  `pack_dir` reuses packs
    minted earlier (about 1.5-4 minutes of reference load time per part otherwise)."""
    import tempfile
    for path in (os.path.join(ROOT, 'tests'), ROOT):
        sys.path.insert(0, path)
    from pack_util import pack_digest
    from paintrl_b200.partpack import PartPack
    pack_dir = pack_dir or tempfile.mkdtemp(prefix='refpacks_')
    out = {}
    for no, name in sorted(UNPACKED_PARTS.items()):
        PART_NAMES[no], MAX_POINTS[no] = name, 0
        import glob
        found = glob.glob(os.path.join(pack_dir, name + '_[0-9]*x[0-9]*.npz'))
        if not found:
            if os.environ.get('PAINTRL_DIGESTS_EXISTING_ONLY'):
                continue
            found = [os.path.join(pack_dir, export_partpack(no, pack_dir=pack_dir) + '.npz')]
        out[name] = pack_digest(PartPack.load(found[0]))
    with open(os.path.join(GOLDEN_DIR, 'pack_digests.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return 'pack_digests.json'


def main(argv):
    if argv and argv[0] == 'digests':
        print('wrote', export_pack_digests(*argv[1:2]))
        return
    names = argv or (['pack0', 'pack1'] + sorted(JOBS))
    if names == ['packs']:
        names = ['pack0', 'pack1']
    with mp.get_context('spawn').Pool(min(len(names), os.cpu_count() or 1)) as pool:
        for done in pool.imap_unordered(_run, names):
            print('wrote', done, flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
