// paintrl_device.cuh -- device-side tables, state and FP64 geometry helpers of the batched paint step.
//
// Arithmetic contract (bit-exact against the oracle, see DESIGN.md "Arithmetic"):
//   * this translation unit is compiled with -fmad=false: every product and sum below is rounded
//     separately, in the written (= the reference's) order; FP64 division and sqrt are IEEE.
//   * np.dot / np.linalg.norm on 3-vectors (reference call sites: bullet_paint_wrapper.py:156-157,
//     robot.py:97, 269) are OpenBLAS ddot = fma(x2,y2, fma(x1,y1, x0*y0)); npdot3() is that chain,
//     written with explicit fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace paintrl {

constexpr double kPaintRadius = 0.051;        // bullet_paint_wrapper.py:42
constexpr double kStepSize = kPaintRadius;    // bullet_paint_wrapper.py:43
constexpr int kPaintPerAction = 5;            // robot.py:165
constexpr int kNotOnPartTerminateSteps = 1000;  // robot.py:167
constexpr double kHookDistance = 0.1;         // bullet_paint_wrapper.py:443
constexpr int kHsiTargetMax = 25;             // bullet_paint_wrapper.py:388
constexpr int kPainted = 255;                 // bullet_paint_wrapper.py:354, 496
constexpr double kPi = 3.141592653589793;     // math.pi / np.pi
constexpr int kMaxObs = 128;                  // largest observation vector (OBS_GRAD^2 or OBS_GRAD+2)
constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

enum : int { kFlagLastOnPart = 1, kFlagTerminate = 2, kFlagHasLast = 4 };

// Per-environment record: one 128-byte line, read and written once per step.
// Robot._pose/_orn (robot.py:235-242), Part._last_painted_pixels (bullet_paint_wrapper.py:483,
// represented by the centre of the last shot: the set is the ball query of that centre),
// Robot turning/termination state (robot.py:201-218), PaintGymEnv counters (robot_gym_env.py:219-221).
struct alignas(128) EnvState {
    double pose[3];
    double quat[4];
    double last_center[3];
    double last_angle;
    double angle_diff;
    double total_reward;
    double total_return;
    int32_t step_counter;
    int32_t term_counter;
    int32_t flags;
    int32_t episode;
};
static_assert(sizeof(EnvState) == 128, "EnvState must be one 128-byte line");

// Constant per-part tables in device memory (shared by all environments, L1/L2 resident).
struct DevPack {
    int n_texels, n_pad;          // n_pad: status-plane length per env, multiple of 128
    int axis0, axis1;
    int status_init;
    // collision hull: (nx, ny, nz, off) per plane
    int n_planes;
    const double4 *planes;
    // ray-test cells over (axis0, axis1): per cell the planes not trivially satisfied in its box
    int rc_nx, rc_ny;
    double rc_o0, rc_o1, rc_inv;
    const int *rc_start;          // [rc_nx*rc_ny + 1]
    const uint16_t *rc_idx;       // plane indices
    const double *rc_dlo, *rc_dhi;  // depth (non-principal axis) range of each cell's box
    // nearest-vertex grid over (axis0, axis1): front vertices sorted by cell
    int vg_nx, vg_ny;
    double vg_o0, vg_o1, vg_cs, vg_inv;
    const int *vg_start;          // [vg_nx*vg_ny + 1]
    const double *vx, *vy, *vz;   // sorted vertex coordinates
    const int *vid;               // sorted -> pack vertex index
    const int *vtri_start;        // CSR over pack vertex index
    const int *vtri_idx;
    const double *tri;            // [n_tris][16]: a(3) v0(3) v1(3) d00 d01 d11 inv_denom n(3)
    // texel bins over (axis0, axis1): texels sorted by cell (row-major, axis1 = row)
    int tb_nx, tb_ny;
    double tb_o0, tb_o1, tb_inv;
    const int *tb_start;          // [tb_nx*tb_ny + 1]
    const double *tx, *ty, *tz;   // [n_pad] sorted texel positions (world x, y, z)
    // section observation: rank of each sorted texel's axis0 / axis1 coordinate among the
    // sorted distinct values (0xFFFF / 0xFFFFFFFF marks padding)
    const void *rank0, *rank1;    // uint16_t or uint32_t [n_pad]
    const void *chunk_box;        // [n_pad/16] rank bounding box (r0min, r0max, r1min, r1max) per 16-texel chunk
    int rank_bytes;
    int n_uniq0, n_uniq1;
    const double *uniq0, *uniq1;
    // grid observation (bullet_paint_wrapper.py:1072-1112)
    const uint16_t *gcell;        // [n_pad] cell per sorted texel (0xFFFF padding)
    const int *gtotal;            // [obs_grad^2]
    // normalised pose (bullet_paint_wrapper.py:965-978)
    int grid_granularity;
    const double *grid_lo, *grid_hi;
    double range0_min, range0_max, range1_min, range1_max, lwr;
    // start points
    int n_starts;
    const double *start_pos, *start_normal;
    // pack order <-> sorted order (state export)
    const int *sorted_to_pack;    // [n_texels]
};

struct DevConfig {
    int action_mode, action_shape, discrete_granularity;
    const double *discrete_table;   // [n][3] u1, u2, angle
    int obs_mode, obs_grad, obs_dim;
    int color_mode, termination_mode;
    double switch_threshold;
    int expected_episode_length, episode_max_length;
    int turning_penalty, overlap_penalty;
    double max_possible_point;
    int auto_reset;
    unsigned long long seed;
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double npdot3(double x0, double x1, double x2, double y0, double y1, double y2) {
    return fma(x2, y2, fma(x1, y1, x0 * y0));
}

struct Vec3 { double x, y, z; };

__device__ __forceinline__ double comp(const Vec3 &v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ void add_comp(Vec3 &v, int a, double d) {
    if (a == 0) v.x += d; else if (a == 1) v.y += d; else v.z += d;
}

// oracle/shims/pybullet.py multiplyTransforms (Bullet btMatrix3x3::setRotation + btTransform())
__device__ __forceinline__ Vec3 transform_point(const Vec3 &pos, const double q[4], double v0, double v1, double v2) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    Vec3 o;
    o.x = (((1.0 - (yy + zz)) * v0 + (xy - wz) * v1) + (xz + wy) * v2) + pos.x;
    o.y = (((xy + wz) * v0 + (1.0 - (xx + zz)) * v1) + (yz - wx) * v2) + pos.y;
    o.z = (((xz - wy) * v0 + (yz + wx) * v1) + (1.0 - (xx + yy)) * v2) + pos.z;
    return o;
}

// robot.py:93-100 get_pose_orn + bullet_paint_wrapper.py:32-37 normalize
__device__ __forceinline__ void quat_from_normal(const Vec3 &n, double q[4]) {
    double x = 0.0 * n.z - 1.0 * n.y;
    double y = 1.0 * n.x - 0.0 * n.z;
    double z = 0.0 * n.y - 0.0 * n.x;
    double w = 1.0 + npdot3(0.0, 0.0, 1.0, n.x, n.y, n.z);
    double mag2 = ((x * x + y * y) + z * z) + w * w;
    if (fabs(mag2 - 1.0) > 0.00001) {
        double mag = sqrt(mag2);
        x /= mag; y /= mag; z /= mag; w /= mag;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

// robot.py:266-271 _get_tcp_orn_norm
__device__ __forceinline__ Vec3 tcp_orn_norm(const Vec3 &pose, const double q[4]) {
    Vec3 along = transform_point(pose, q, 0.0, 0.0, 1.0);
    double vx = along.x - pose.x, vy = along.y - pose.y, vz = along.z - pose.z;
    double norm = sqrt(npdot3(vx, vy, vz, vx, vy, vz));
    Vec3 o = {vx / norm, vy / norm, vz / norm};
    return o;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// One pass of the slab test over a list of planes (all of them, or one cell's active list),
// split across the warp; max/min are order-independent so the result equals the serial one.
struct SlabResult { double t_in, t_out; bool outside; };

template <bool INDEXED>
__device__ __forceinline__ SlabResult slab_pass(const DevPack &pk, const Vec3 &frm, double d0, double d1, double d2,
                                                int begin, int end, int lane) {
    double t_in = -INFINITY, t_out = INFINITY;
    bool outside = false;
    for (int i = begin + lane; i < end; i += 32) {
        int pi = INDEXED ? (int)__ldg(&pk.rc_idx[i]) : i;
        const double2 *p2 = reinterpret_cast<const double2 *>(pk.planes) + 2 * pi;
        double2 lo = __ldg(p2), hi2 = __ldg(p2 + 1);
        double den = (lo.x * d0 + lo.y * d1) + hi2.x * d2;
        double num = hi2.y - ((lo.x * frm.x + lo.y * frm.y) + hi2.x * frm.z);
        if (den == 0.0) {
            if (num < 0.0) outside = true;
        } else {
            double t = num / den;
            if (den < 0.0) t_in = fmax(t_in, t);
            else t_out = fmin(t_out, t);
        }
    }
    SlabResult r;
    r.t_in = warp_max(t_in);
    r.t_out = warp_min(t_out);
    r.outside = __any_sync(kFull, outside);
    return r;
}

// Exact slab test of the ray frm -> to against the hull half-spaces (shim S1).
//
// Fast path: the hull footprint is covered by a grid of cells over (axis0, axis1); each cell
// lists the planes that are NOT satisfied with a safety margin everywhere in the cell's box
// (footprint x the cell's depth range).  If the entry point h* found from one cell's list lies in
// that same box, every unlisted plane j satisfies n_j.h* < off_j - margin, i.e. t_j < t* if it is
// an entering plane and t_j > t* if it is an exiting one, so max/min over the list decide exactly
// what max/min over all planes decide, and t* is the global t_in bit for bit.  Otherwise the
// full plane list is scanned.  Either way the result is the serial slab test's.
__device__ __forceinline__ bool ray_test(const DevPack &pk, const Vec3 &frm, const Vec3 &to, int lane, Vec3 &hit,
                                         int &full_scans) {
    double d0 = to.x - frm.x, d1 = to.y - frm.y, d2 = to.z - frm.z;
    SlabResult r;
    bool accepted = false;
    if (pk.rc_nx > 0) {
        // the TCP hovers kHookDistance above the surface: first guess = the point that far along the ray
        Vec3 g = {frm.x + d0 * kHookDistance, frm.y + d1 * kHookDistance, frm.z + d2 * kHookDistance};
#pragma unroll 1
        for (int attempt = 0; attempt < 2 && !accepted; ++attempt) {
            int cx = (int)floor((comp(g, pk.axis0) - pk.rc_o0) * pk.rc_inv);
            int cy = (int)floor((comp(g, pk.axis1) - pk.rc_o1) * pk.rc_inv);
            if (cx < 0 || cy < 0 || cx >= pk.rc_nx || cy >= pk.rc_ny) break;
            int cell = cy * pk.rc_nx + cx;
            int begin = __ldg(&pk.rc_start[cell]), end = __ldg(&pk.rc_start[cell + 1]);
            if (end <= begin) break;
            r = slab_pass<true>(pk, frm, d0, d1, d2, begin, end, lane);
            if (!(r.t_in > -INFINITY) || !(r.t_in < INFINITY)) break;
            Vec3 h = {frm.x + d0 * r.t_in, frm.y + d1 * r.t_in, frm.z + d2 * r.t_in};
            int hx = (int)floor((comp(h, pk.axis0) - pk.rc_o0) * pk.rc_inv);
            int hy = (int)floor((comp(h, pk.axis1) - pk.rc_o1) * pk.rc_inv);
            double depth = comp(h, 3 - pk.axis0 - pk.axis1);
            if (hx == cx && hy == cy && depth >= __ldg(&pk.rc_dlo[cell]) && depth <= __ldg(&pk.rc_dhi[cell])) accepted = true;
            else g = h;
        }
    }
    if (!accepted) {
        r = slab_pass<false>(pk, frm, d0, d1, d2, 0, pk.n_planes, lane);
        full_scans += 1;
    }
    if (r.outside || !(r.t_in <= r.t_out && 0.0 <= r.t_in && r.t_in <= 1.0)) return false;
    hit.x = frm.x + d0 * r.t_in;
    hit.y = frm.y + d1 * r.t_in;
    hit.z = frm.z + d2 * r.t_in;
    return true;
}

// cKDTree.query(point, k=1) over the side-masked vertices (bullet_paint_wrapper.py:526): grid
// search over (axis0, axis1) with ring expansion; exact FP64 squared distances, lowest pack
// index on ties.  Returns the pack vertex index.
__device__ __forceinline__ int nearest_vertex(const DevPack &pk, const Vec3 &p, int lane) {
    double q0 = comp(p, pk.axis0), q1 = comp(p, pk.axis1);
    int cx = (int)floor((q0 - pk.vg_o0) * pk.vg_inv);
    int cy = (int)floor((q1 - pk.vg_o1) * pk.vg_inv);
    cx = min(max(cx, 0), pk.vg_nx - 1);
    cy = min(max(cy, 0), pk.vg_ny - 1);
    double best_d = INFINITY;
    int best_i = 0x7fffffff;
    for (int k = 1;; ++k) {
        int x0 = max(cx - k, 0), x1 = min(cx + k, pk.vg_nx - 1);
        int y0 = max(cy - k, 0), y1 = min(cy + k, pk.vg_ny - 1);
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int row = y0; row <= y1; ++row) {
            int begin = __ldg(&pk.vg_start[row * pk.vg_nx + x0]);
            int end = __ldg(&pk.vg_start[row * pk.vg_nx + x1 + 1]);
            for (int j = begin + lane; j < end; j += 32) {
                double dx = __ldg(&pk.vx[j]) - p.x, dy = __ldg(&pk.vy[j]) - p.y, dz = __ldg(&pk.vz[j]) - p.z;
                double d = dx * dx + dy * dy + dz * dz;
                int id = __ldg(&pk.vid[j]);
                if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(kFull, bd, o);
            int oi = __shfl_xor_sync(kFull, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        best_d = bd;
        best_i = bi;
        bool whole = (x0 == 0 && y0 == 0 && x1 == pk.vg_nx - 1 && y1 == pk.vg_ny - 1);
        if (whole) break;
        // every vertex outside the scanned block is at least m away in the (axis0, axis1) plane
        double m = INFINITY;
        if (x0 > 0) m = fmin(m, q0 - (pk.vg_o0 + x0 * pk.vg_cs));
        if (x1 < pk.vg_nx - 1) m = fmin(m, (pk.vg_o0 + (x1 + 1) * pk.vg_cs) - q0);
        if (y0 > 0) m = fmin(m, q1 - (pk.vg_o1 + y0 * pk.vg_cs));
        if (y1 < pk.vg_ny - 1) m = fmin(m, (pk.vg_o1 + (y1 + 1) * pk.vg_cs) - q1);
        m -= 1e-9;
        if (m > 0.0 && best_d < m * m) break;
    }
    return best_i;
}

// Part._get_hook_point + _get_closest_bary (bullet_paint_wrapper.py:525-534, 508-523, 154-185):
// incident front triangles of the nearest vertex, one per lane.
__device__ __forceinline__ bool hook_point(const DevPack &pk, const Vec3 &point, int lane, Vec3 &pose, Vec3 &orn) {
    int v = nearest_vertex(pk, point, lane);
    if (v == 0x7fffffff) return false;
    int begin = __ldg(&pk.vtri_start[v]), end = __ldg(&pk.vtri_start[v + 1]);
    int deg = end - begin;
    if (deg <= 0) return false;
    int pick = -1;
    double run_max = -INFINITY;   // max of min_uvw over the lanes scanned so far
    int run_arg = -1;             // last list position attaining it
    for (int base = 0; base < deg; base += 32) {
        int k = base + lane;
        bool inside = false;
        double m = -INFINITY;
        if (k < deg) {
            const double *t = pk.tri + 16 * (size_t)__ldg(&pk.vtri_idx[begin + k]);
            double v2x = point.x - __ldg(t + 0), v2y = point.y - __ldg(t + 1), v2z = point.z - __ldg(t + 2);
            double d20 = npdot3(v2x, v2y, v2z, __ldg(t + 3), __ldg(t + 4), __ldg(t + 5));
            double d21 = npdot3(v2x, v2y, v2z, __ldg(t + 6), __ldg(t + 7), __ldg(t + 8));
            double d00 = __ldg(t + 9), d01 = __ldg(t + 10), d11 = __ldg(t + 11), inv = __ldg(t + 12);
            double bv = (d11 * d20 - d01 * d21) * inv;
            double bw = (d00 * d21 - d01 * d20) * inv;
            double bu = 1.0 - bv - bw;
            if (inv == 0.0) { bu = -1.0; bv = -1.0; bw = -1.0; }
            inside = (0.0 <= bu && bu <= 1.0 && 0.0 <= bv && bv <= 1.0 && 0.0 <= bw && bw <= 1.0);
            m = fmin(fmin(bu, bv), bw);
        }
        unsigned in_mask = __ballot_sync(kFull, inside);
        if (in_mask) { pick = base + __ffs(in_mask) - 1; break; }
        double cm = warp_max(m);
        if (cm >= run_max) {   // `>=`: a later triangle wins ties (bullet_paint_wrapper.py:520)
            unsigned eq = __ballot_sync(kFull, k < deg && m == cm);
            run_max = cm;
            run_arg = base + 31 - __clz(eq);
        }
    }
    if (pick < 0) pick = (run_max >= -1.0) ? run_arg : 0;   // closest_uvw starts at -1 (:509)
    const double *t = pk.tri + 16 * (size_t)__ldg(&pk.vtri_idx[begin + pick]);
    double nx = __ldg(t + 13), ny = __ldg(t + 14), nz = __ldg(t + 15);
    pose.x = point.x + nx * kHookDistance;
    pose.y = point.y + ny * kHookDistance;
    pose.z = point.z + nz * kHookDistance;
    orn.x = -nx; orn.y = -ny; orn.z = -nz;
    return true;
}

// bullet_paint_wrapper.py:844-851
__device__ __forceinline__ int grid_index_2(const DevPack &pk, double v) {
    double rel = (v - pk.range1_min) / (pk.range1_max - pk.range1_min);
    double scaled = rel * pk.grid_granularity;
    if (!(scaled > -1.0)) return 0;
    if (scaled >= (double)pk.grid_granularity) return pk.grid_granularity - 1;
    return (int)scaled;
}

__device__ __forceinline__ double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

// bullet_paint_wrapper.py:965-978
__device__ __forceinline__ void normalized_pose(const DevPack &pk, const Vec3 &pose, double &a1, double &a2) {
    const double radius = kPaintRadius;
    double axis1_real = comp(pose, pk.axis0), axis2_real = comp(pose, pk.axis1);
    double axis2_in = (axis2_real - pk.range1_min + radius) / (pk.range1_max - pk.range1_min + 2 * radius);
    int gi = grid_index_2(pk, axis2_real);
    double lo = __ldg(&pk.grid_lo[gi]), hi = __ldg(&pk.grid_hi[gi]);
    double axis1_in;
    if (hi - lo == 0.0) axis1_in = 0.0;
    else axis1_in = (axis1_real - lo + radius) / (hi - lo + 2 * radius);
    a1 = clip01(axis1_in);
    a2 = clip01(axis2_in);
}

// robot_gym_env.py:92-98
__device__ __forceinline__ int handle_pos(double pos) {
    if (pos == 0.0) return 0;
    if (pos == 1.0) return 21;
    return (int)(pos * 20) + 1;
}

// numpy float64 floor_divide (npy_divmod), used by `angle // basis` (bullet_paint_wrapper.py:1030)
__device__ __forceinline__ double np_floor_divide(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) { mod += b; div -= 1.0; }
    }
    double fd;
    if (div != 0.0) {
        fd = floor(div);
        if (div - fd > 0.5) fd += 1.0;
    } else {
        fd = copysign(0.0, a / b);
    }
    return fd;
}

// counter-based start-index stream for auto-reset (not in the reference: it draws from the
// process-global `random`, robot_gym_env.py:381)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace paintrl
