/*
 * paint_oracle.c -- CPU restatement of PaintRL's paint-simulation step.  TEST INFRASTRUCTURE.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  It restates, line for line and
 * in the reference's own FP64 operation order, the single-environment path
 *     PaintGymEnv.step            PaintRLEnv/robot_gym_env.py:349-368
 *     Robot.apply_action          PaintRLEnv/robot.py:383-433
 *     Part.get_guided_point / fast_paint / get_normalized_pose / observations
 *                                 PaintRLEnv/bullet_paint_wrapper.py:865-880, 568-577, 965-978,
 *                                 1045-1061, 1126-1139
 * on top of the constant per-part tables of a PaintrlPartPack (include/paintrl.h) and the frozen
 * shim arithmetic of oracle/shims/pybullet.py (ray test, multiplyTransforms).  It is pinned by
 * tests/test_oracle_golden.py against traces minted from the reference's own Python sources
 * (oracle/make_golden.py -> tests/golden/g*.npz files).
 *
 * Everything is brute force on purpose (no bins, no ranks, no incremental counters): it shares
 * no acceleration structure with the CUDA engine.
 *
 * Arithmetic notes:
 *   - build with -ffp-contract=off: every product and sum is rounded separately, as CPython /
 *     NumPy element-wise arithmetic does.
 *   - np.dot / np.linalg.norm on 3-vectors go through OpenBLAS ddot, which on every FMA-capable
 *     x86 kernel evaluates  fma(x2,y2, fma(x1,y1, x0*y0))  (measured in the build container,
 *     200000/200000 random vectors); npdot3() below is that chain.
 *   - direction_normalize (robot.py:151-160) uses NumPy's cos/sin/arctan2, which are NumPy's own
 *     SIMD kernels, not libm; the Python wrapper (oracle/oracle.py) evaluates it with NumPy and
 *     passes the unit direction (u1,u2) in, so this file needs libm only for sqrt, atan, atan2,
 *     fmod, floor and pow -- the same libm CPython's math module and NumPy's scalar paths call.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/paintrl.h"

#define PAINT_RADIUS 0.051          /* bullet_paint_wrapper.py:42 */
#define STEP_SIZE PAINT_RADIUS      /* bullet_paint_wrapper.py:43 */
#define PAINT_PER_ACTION 5          /* robot.py:165 */
#define NOT_ON_PART_TERMINATE_STEPS 1000 /* robot.py:167 */
#define HOOK_DISTANCE_TO_PART 0.1   /* bullet_paint_wrapper.py:443 */
#define HSI_TARGET_MAX 25           /* bullet_paint_wrapper.py:388  int(255 / 10) */
#define PAINTED 255                 /* Part.color = (1,0,0) -> first channel 255 (:496, :354) */
#define MAX_OBS 512

typedef struct OracleEnv {
    PaintrlPartPack pack;
    PaintrlConfig cfg;
    int n;                     /* n_texels */
    int obs_dim;
    int32_t *grid_cell;        /* [n] grid-observation cell per texel, or NULL */
    int32_t *grid_total;       /* [G*G] */
    /* Part state */
    int16_t *status;           /* first-channel value per front texel (Part.texels) */
    uint8_t *last_affected;    /* Part._last_painted_pixels as a membership mask */
    uint8_t *affected;         /* scratch */
    uint8_t *possible;         /* scratch: robot.py:401 possible_pixels */
    double *dist;              /* scratch: HSI distances */
    int32_t *beam_texel;       /* scratch (normal paint): nearest texel of every beam that hit, in beam order */
    /* Robot state (robot.py:201-218, 235-242) */
    double pose[3], orn[4];
    int terminate, terminate_counter, last_on_part;
    double last_turning_angle, angle_diff;
    /* PaintGymEnv state (robot_gym_env.py:219-221) */
    int step_counter;
    double total_reward, total_return;
    /* probes for the tests */
    double last_rate, last_succeeded;
    int last_pixel_counter;
    int anomalies;
} OracleEnv;

/* ---------------------------------------------------------------------------------- helpers */
static double npdot3(const double *x, const double *y) {
    return fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0]));
}

/* oracle/shims/pybullet.py: _rotation_rows + multiplyTransforms (position part) */
static void transform_point(const double *pos, const double *q, const double *v, double *out) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    double r[3][3] = {{1.0 - (yy + zz), xy - wz, xz + wy},
                      {xy + wz, 1.0 - (xx + zz), yz - wx},
                      {xz - wy, yz + wx, 1.0 - (xx + yy)}};
    for (int i = 0; i < 3; ++i)
        out[i] = ((r[i][0] * v[0] + r[i][1] * v[1]) + r[i][2] * v[2]) + pos[i];
}

/* robot.py:93-100 get_pose_orn + bullet_paint_wrapper.py:32-37 normalize */
static void quat_from_normal(const double *n, double *q) {
    /* np.cross((0,0,1), n) = (0*n2 - 1*n1, 1*n0 - 0*n2, 0*n1 - 0*n0) */
    double x = 0.0 * n[2] - 1.0 * n[1];
    double y = 1.0 * n[0] - 0.0 * n[2];
    double z = 0.0 * n[1] - 0.0 * n[0];
    const double old_z[3] = {0.0, 0.0, 1.0};
    double w = 1.0 + npdot3(old_z, n);
    double mag2 = ((x * x + y * y) + z * z) + w * w;
    if (fabs(mag2 - 1.0) > 0.00001) {
        double mag = sqrt(mag2);
        x /= mag; y /= mag; z /= mag; w /= mag;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

/* robot.py:266-271 _get_tcp_orn_norm */
static void tcp_orn_norm(const OracleEnv *e, double *out) {
    const double unit_z[3] = {0.0, 0.0, 1.0};
    double along[3], v[3];
    transform_point(e->pose, e->orn, unit_z, along);
    for (int i = 0; i < 3; ++i) v[i] = along[i] - e->pose[i];
    double norm = sqrt(npdot3(v, v));
    for (int i = 0; i < 3; ++i) out[i] = v[i] / norm;
}

/* oracle/shims/pybullet.py:_ray_vs_body */
static int ray_test(const PaintrlPartPack *p, const double *frm, const double *to, double *hit) {
    double d0 = to[0] - frm[0], d1 = to[1] - frm[1], d2 = to[2] - frm[2];
    double t_in = -INFINITY, t_out = INFINITY;
    for (int i = 0; i < p->n_planes; ++i) {
        const double *n = p->plane_n + 3 * i;
        double den = (n[0] * d0 + n[1] * d1) + n[2] * d2;
        double num = p->plane_off[i] - ((n[0] * frm[0] + n[1] * frm[1]) + n[2] * frm[2]);
        if (den == 0.0) {
            if (num < 0.0) return 0;
            continue;
        }
        double t = num / den;
        if (den < 0.0) { if (t > t_in) t_in = t; }
        else           { if (t < t_out) t_out = t; }
    }
    if (!(t_in <= t_out && 0.0 <= t_in && t_in <= 1.0)) return 0;
    hit[0] = frm[0] + d0 * t_in;
    hit[1] = frm[1] + d1 * t_in;
    hit[2] = frm[2] + d2 * t_in;
    return 1;
}

/* bullet_paint_wrapper.py:154-163 _get_bary_coordinate */
static void bary_uvw(const PaintrlPartPack *p, int t, const double *point, double *uvw) {
    const double *a = p->tri_a + 3 * t;
    double v2[3] = {point[0] - a[0], point[1] - a[1], point[2] - a[2]};
    double d20 = npdot3(v2, p->tri_v0 + 3 * t);
    double d21 = npdot3(v2, p->tri_v1 + 3 * t);
    double inv = p->tri_inv_denom[t];
    double v = (p->tri_d11[t] * d20 - p->tri_d01[t] * d21) * inv;
    double w = (p->tri_d00[t] * d21 - p->tri_d01[t] * d20) * inv;
    double u = 1.0 - v - w;
    if (inv == 0) { uvw[0] = uvw[1] = uvw[2] = -1; return; }
    uvw[0] = u; uvw[1] = v; uvw[2] = w;
}

/* bullet_paint_wrapper.py:525-534 _get_hook_point (+ 508-523 _get_closest_bary) */
static int hook_point(const PaintrlPartPack *p, const double *point, double *pose, double *orn) {
    /* cKDTree.query(point, k=1): nearest side-masked vertex */
    int best = -1;
    double best_d = INFINITY;
    for (int i = 0; i < p->n_vertices; ++i) {
        const double *v = p->vertices + 3 * i;
        double dx = v[0] - point[0], dy = v[1] - point[1], dz = v[2] - point[2];
        double d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) { best_d = d; best = i; }
    }
    double closest_uvw = -1;
    int closest = -1;
    for (int k = p->vtri_start[best]; k < p->vtri_start[best + 1]; ++k) {
        int t = p->vtri_idx[k];
        double uvw[3];
        bary_uvw(p, t, point, uvw);
        if (0 <= uvw[0] && uvw[0] <= 1 && 0 <= uvw[1] && uvw[1] <= 1 && 0 <= uvw[2] && uvw[2] <= 1) {
            closest = t;
            break;
        }
        if (closest < 0) closest = t;
        double m = uvw[0] < uvw[1] ? uvw[0] : uvw[1];
        if (uvw[2] < m) m = uvw[2];
        if (m >= closest_uvw) { closest_uvw = m; closest = t; }
    }
    if (closest < 0) return 0;
    const double *n = p->tri_n + 3 * closest;
    for (int i = 0; i < 3; ++i) {
        pose[i] = point[i] + n[i] * HOOK_DISTANCE_TO_PART;   /* :56-58 */
        orn[i] = -n[i];
    }
    return 1;
}

/* bullet_paint_wrapper.py:865-880 get_guided_point; returns 1 on hit */
static int guided_point(const OracleEnv *e, const double *point_in, const double *normal,
                        double delta_axis1, double delta_axis2, double *pos, double *orn) {
    const PaintrlPartPack *p = &e->pack;
    double point[3] = {point_in[0], point_in[1], point_in[2]};
    double delta_2 = delta_axis2 * p->length_width_ratio;
    double delta_1 = delta_axis1;
    point[p->axis0] += delta_1;
    point[p->axis1] += delta_2;
    double end_point[3] = {point[0] + normal[0], point[1] + normal[1], point[2] + normal[2]};
    double surface[3];
    if (!ray_test(p, point, end_point, surface)) {
        memcpy(orn, normal, 3 * sizeof(double));
        return 0;
    }
    if (!hook_point(p, surface, pos, orn)) {
        memcpy(orn, normal, 3 * sizeof(double));
        return 0;
    }
    return 1;
}

/* robot.py:292-300 */
static void count_not_on_part(OracleEnv *e) {
    if (e->last_on_part) { e->last_on_part = 0; return; }
    e->terminate_counter += 1;
    e->last_on_part = 0;
    if (e->terminate_counter > NOT_ON_PART_TERMINATE_STEPS) e->terminate = 1;
}

/* bullet_paint_wrapper.py:844-851 */
static int grid_index_2(const PaintrlPartPack *p, double val_axis_2) {
    double rel = (val_axis_2 - p->range1_min) / (p->range1_max - p->range1_min);
    double scaled = rel * p->grid_granularity;
    int idx = (int)scaled;                    /* Python int(): truncation toward zero */
    if (scaled <= -2147483648.0 || scaled >= 2147483647.0) idx = scaled < 0 ? -1 : p->grid_granularity;
    if (idx < 0) return 0;
    if (idx > p->grid_granularity - 1) return p->grid_granularity - 1;
    return idx;
}

static double clip01(double v) { return v < 0 ? 0.0 : (v > 1 ? 1.0 : v); }

/* bullet_paint_wrapper.py:965-978 */
static void normalized_pose(const OracleEnv *e, const double *pose, double *out) {
    const PaintrlPartPack *p = &e->pack;
    double radius = PAINT_RADIUS;
    double axis1_real = pose[p->axis0], axis2_real = pose[p->axis1];
    double axis2_in_range = (axis2_real - p->range1_min + radius) /
                            (p->range1_max - p->range1_min + 2 * radius);
    int gi = grid_index_2(p, axis2_real);
    double lo = p->grid_lo[gi], hi = p->grid_hi[gi];
    double axis1_in_range;
    if (hi - lo == 0) axis1_in_range = 0;
    else axis1_in_range = (axis1_real - lo + radius) / (hi - lo + 2 * radius);
    out[0] = clip01(axis1_in_range);
    out[1] = clip01(axis2_in_range);
}

/* numpy npy_divmod -> floor_divide for float64 scalars (`angle // self._basis`, :1030) */
static double np_floor_divide(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0) {
        if ((b < 0) != (mod < 0)) { mod += b; div -= 1.0; }
    }
    double floordiv;
    if (div != 0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
    return floordiv;
}

/* bullet_paint_wrapper.py:1045-1061 SectionObservation.get_observation */
static void section_observation(OracleEnv *e, const double *pose, int section, double *out) {
    const PaintrlPartPack *p = &e->pack;
    int done[MAX_OBS], total[MAX_OBS];
    memset(done, 0, sizeof(int) * section);
    memset(total, 0, sizeof(int) * section);
    double basis = 2 * M_PI / section;
    for (int i = 0; i < e->n; ++i) {
        const double *c = p->texel_pos + 3 * i;
        double rx = c[p->axis0] - pose[p->axis0];
        double ry = c[p->axis1] - pose[p->axis1];
        if (rx == 0 && ry == 0) continue;
        int valid = e->status[i] == PAINTED ? 0 : 1;     /* :723-725, :354 */
        int index;
        if (section == 4) {                              /* :1033-1043 */
            if (rx > 0 && ry > 0) index = 0;
            else if (rx < 0 && 0 < ry) index = 1;
            else if (rx < 0 && ry < 0) index = 2;
            else index = 3;
        } else {                                         /* :1026-1031 */
            double angle = atan2(ry, rx);
            if (angle < 0) angle = 2 * M_PI + angle;
            index = (int)np_floor_divide(angle, basis);
            if (index >= section) { index = section - 1; e->anomalies++; }
        }
        done[index] += valid;
        total[index] += 1;
    }
    for (int s = 0; s < section; ++s)
        out[s] = total[s] == 0 ? 0.0 : (double)done[s] / (double)total[s];
}

/* bullet_paint_wrapper.py:1072-1112 cell lists; returns 0 on a layout the reference cannot index */
static int build_grid_cells(OracleEnv *e) {
    const PaintrlPartPack *p = &e->pack;
    int g = e->cfg.obs_grad, vgran = p->grid_granularity;
    int v_interval = (int)((double)vgran / (double)g);
    double axis_2_step = (p->range1_max - p->range1_min) / vgran;
    e->grid_cell = (int32_t *)malloc(sizeof(int32_t) * e->n);
    e->grid_total = (int32_t *)calloc((size_t)g * g, sizeof(int32_t));
    for (int i = 0; i < e->n; ++i) {
        const double *c = p->texel_pos + 3 * i;
        int y_grid = (int)((c[p->axis1] - p->range1_min) / axis_2_step);
        if (y_grid > vgran - 1) y_grid = vgran - 1;
        if (y_grid < 0) return 0;
        double range = p->grid_hi[y_grid] - p->grid_lo[y_grid];
        int x_grid;
        if (range == 0) x_grid = 0;
        else {
            double x_step = range / g;
            x_grid = (int)((c[p->axis0] - p->grid_lo[y_grid]) / x_step);
            if (x_grid > g - 1) x_grid = g - 1;
        }
        int v_target = y_grid / v_interval;
        if (x_grid < 0 || v_target >= g) return 0;
        e->grid_cell[i] = v_target * g + x_grid;
        e->grid_total[v_target * g + x_grid] += 1;
    }
    return 1;
}

/* bullet_paint_wrapper.py:1126-1139 GridObservation.get_observation */
static void grid_observation(OracleEnv *e, double *out) {
    int g = e->cfg.obs_grad;
    int done[MAX_OBS];
    memset(done, 0, sizeof(int) * g * g);
    for (int i = 0; i < e->n; ++i)
        if (e->status[i] == PAINTED) done[e->grid_cell[i]] += 1;
    for (int c = 0; c < g * g; ++c)
        out[c] = e->grid_total[c] == 0 ? 0.0 : 1.0 - (double)done[c] / (double)e->grid_total[c];
}

/* robot_gym_env.py:92-103 */
static int handle_pos(double pos) {
    if (pos == 0) return 0;
    if (pos == 1) return 21;
    return (int)(pos * 20) + 1;
}

/* robot_gym_env.py:306-319 _augmented_observation */
static void augmented_observation(OracleEnv *e, double *obs) {
    double np_[2];
    normalized_pose(e, e->pose, np_);
    int grad = e->cfg.obs_grad;
    switch (e->cfg.obs_mode) {
    case PAINTRL_OBS_SIMPLE:
        obs[0] = np_[0]; obs[1] = np_[1];
        break;
    case PAINTRL_OBS_GRID:
        grid_observation(e, obs);
        break;
    case PAINTRL_OBS_DISCRETE: {
        section_observation(e, e->pose, grad, obs);
        int position = (handle_pos(np_[0]) + 1) * 22 + handle_pos(np_[1]);
        obs[grad] = 1.0 / position;
        break;
    }
    default:
        section_observation(e, e->pose, grad, obs);
        obs[grad] = np_[0]; obs[grad + 1] = np_[1];
    }
}

/* Part.fast_paint + _paint (bullet_paint_wrapper.py:568-577) with the colour handlers
 * (:352-375 RGB, :392-434 HSI).  Returns succeed_counter; marks e->affected. */
static double fast_paint(OracleEnv *e, const double *center, int *n_valid_out) {
    const PaintrlPartPack *p = &e->pack;
    const double r2 = PAINT_RADIUS * PAINT_RADIUS;
    int n = e->n;
    double succeed = 0;
    /* cKDTree.query_ball_point(center, r):  dx*dx + dy*dy + dz*dz <= r*r   */
    double rmax = 0;
    int any = 0;
    for (int i = 0; i < n; ++i) {
        const double *c = p->texel_pos + 3 * i;
        double dx = c[0] - center[0], dy = c[1] - center[1], dz = c[2] - center[2];
        double d2 = dx * dx + dy * dy + dz * dz;
        int in = d2 <= r2;
        e->affected[i] = (uint8_t)in;
        if (in && e->cfg.color_mode == PAINTRL_COLOR_HSI) {
            /* minkowski_distance: sum(|y - x|**2, axis=-1) ** 0.5 */
            double ex = fabs(center[0] - c[0]), ey = fabs(center[1] - c[1]), ez = fabs(center[2] - c[2]);
            double dist = sqrt((ex * ex + ey * ey) + ez * ez);
            e->dist[i] = dist;
            if (!any || dist > rmax) rmax = dist;
            any = 1;
        }
    }
    if (e->cfg.color_mode == PAINTRL_COLOR_RGB) {
        for (int i = 0; i < n; ++i)
            if (e->affected[i] && e->status[i] != PAINTED) { e->status[i] = PAINTED; succeed += 1; }
    } else {
        for (int i = 0; i < n; ++i) {
            if (!e->affected[i]) continue;
            double ratio = e->dist[i] / rmax;
            /* int(TARGET_MAX * (1 - ratio ** 2) ** (BETA - 1)) + 1, BETA = 2 */
            int quantity = (int)(HSI_TARGET_MAX * pow(1 - pow(ratio, 2.0), 1.0)) + 1;
            if (e->status[i] <= 0) continue;             /* is_changed (:392-394) */
            e->status[i] = (int16_t)(e->status[i] - quantity);
            succeed += quantity / 255.0;
        }
    }
    /* valid_pixels = affected \ last_painted ; last_painted = affected  (:575-576) */
    int n_valid = 0;
    for (int i = 0; i < n; ++i) {
        if (e->affected[i] && !e->last_affected[i]) { e->possible[i] = 1; n_valid++; }
    }
    memcpy(e->last_affected, e->affected, (size_t)n);
    *n_valid_out = n_valid;
    return succeed;
}

/* Robot._paint (robot.py:280-285) + Part.paint / _paint (bullet_paint_wrapper.py:562-566, 572-577): the beam fan of
 * the TCP (robot.py:251-258: every point of Robot._paint_plain, given in the TCP frame, is a ray end point), the hits
 * on the part (shim S1 rayTestBatch), the texel nearest to each hit (cKDTree.query, k = 1: brute force here), and the
 * colour handlers on that list -- duplicates included: RGB repaints are no-ops (:358-365), HSI subtracts once per
 * occurrence while the texel is still above 0 (:411-434).  Uses e->pose / e->orn of the current sub-step. */
static double normal_paint(OracleEnv *e, const double *center, int *n_valid_out) {
    const PaintrlPartPack *p = &e->pack;
    const int n = e->n;
    int n_hits = 0;
    for (int b = 0; b < e->cfg.n_beams; ++b) {
        double dst[3], hit[3];
        transform_point(e->pose, e->orn, e->cfg.beam_plain + 3 * b, dst);   /* _get_tcp_point_in_world */
        if (!ray_test(p, e->pose, dst, hit)) continue;
        int best = -1;
        double best_d = INFINITY;
        for (int i = 0; i < n; ++i) {
            const double *c = p->texel_pos + 3 * i;
            double dx = c[0] - hit[0], dy = c[1] - hit[1], dz = c[2] - hit[2];
            double d = dx * dx + dy * dy + dz * dz;
            if (d < best_d) { best_d = d; best = i; }
        }
        if (p->texel_nn_rep) best = p->texel_nn_rep[best];    /* twins at the same position: the kd-tree's pick */
        e->beam_texel[n_hits++] = best;
    }
    *n_valid_out = 0;
    if (n_hits == 0) return 0;                 /* Part.paint: `if not points: return [], 0` -- the last set is kept */
    memset(e->affected, 0, (size_t)n);
    double succeed = 0;
    if (e->cfg.color_mode == PAINTRL_COLOR_RGB) {
        for (int k = 0; k < n_hits; ++k) {
            int i = e->beam_texel[k];
            if (e->status[i] != PAINTED) { e->status[i] = PAINTED; succeed += 1; }
            e->affected[i] = 1;
        }
    } else {
        double rmax = 0;
        for (int k = 0; k < n_hits; ++k) {
            const double *c = p->texel_pos + 3 * e->beam_texel[k];
            double ex = fabs(center[0] - c[0]), ey = fabs(center[1] - c[1]), ez = fabs(center[2] - c[2]);
            double dist = sqrt((ex * ex + ey * ey) + ez * ez);
            e->dist[k] = dist;
            if (k == 0 || dist > rmax) rmax = dist;
        }
        if (rmax == 0) e->anomalies += 1;      /* 0 / 0: the reference would raise on int(nan) */
        for (int k = 0; k < n_hits; ++k) {
            int i = e->beam_texel[k];
            double ratio = e->dist[k] / rmax;
            int quantity = (int)(HSI_TARGET_MAX * pow(1 - pow(ratio, 2.0), 1.0)) + 1;
            e->affected[i] = 1;
            if (e->status[i] <= 0) continue;
            e->status[i] = (int16_t)(e->status[i] - quantity);
            succeed += quantity / 255.0;
        }
    }
    int n_valid = 0;
    for (int i = 0; i < n; ++i)
        if (e->affected[i] && !e->last_affected[i]) { e->possible[i] = 1; n_valid++; }
    memcpy(e->last_affected, e->affected, (size_t)n);
    *n_valid_out = n_valid;
    return succeed;
}

/* ---------------------------------------------------------------------------------- API */
OracleEnv *oracle_create(const PaintrlPartPack *pack, const PaintrlConfig *cfg) {
    OracleEnv *e = (OracleEnv *)calloc(1, sizeof(OracleEnv));
    e->pack = *pack;
    e->cfg = *cfg;
    e->n = pack->n_texels;
    int grad = cfg->obs_grad;
    switch (cfg->obs_mode) {
    case PAINTRL_OBS_SECTION: e->obs_dim = grad + 2; break;
    case PAINTRL_OBS_GRID: e->obs_dim = grad * grad; break;
    case PAINTRL_OBS_SIMPLE: e->obs_dim = 2; break;
    default: e->obs_dim = grad + 1;
    }
    if (e->obs_dim > MAX_OBS) { free(e); return NULL; }
    e->status = (int16_t *)malloc(sizeof(int16_t) * e->n);
    e->last_affected = (uint8_t *)calloc(e->n, 1);
    e->affected = (uint8_t *)calloc(e->n, 1);
    e->possible = (uint8_t *)calloc(e->n, 1);
    e->dist = (double *)calloc((size_t)(e->n > cfg->n_beams ? e->n : cfg->n_beams) + 1, sizeof(double));
    e->beam_texel = (int32_t *)calloc((size_t)(cfg->n_beams > 0 ? cfg->n_beams : 1), sizeof(int32_t));
    if (cfg->paint_method == PAINTRL_PAINT_NORMAL && (cfg->n_beams <= 0 || !cfg->beam_plain)) { free(e); return NULL; }
    if (cfg->obs_mode == PAINTRL_OBS_GRID && !build_grid_cells(e)) { free(e); return NULL; }
    return e;
}

void oracle_destroy(OracleEnv *e) {
    if (!e) return;
    free(e->status); free(e->last_affected); free(e->affected); free(e->possible); free(e->dist); free(e->beam_texel);
    free(e->grid_cell); free(e->grid_total);
    free(e);
}

int oracle_obs_dim(const OracleEnv *e) { return e->obs_dim; }
const int32_t *oracle_grid_cells(const OracleEnv *e) { return e->grid_cell; }

/* Robot.reset(pose) (robot.py:366-372, 208-212) */
void oracle_set_pose(OracleEnv *e, const double *pos, const double *normal, double *obs) {
    quat_from_normal(normal, e->orn);
    memcpy(e->pose, pos, 3 * sizeof(double));
    e->terminate = 0;
    e->terminate_counter = 0;
    e->last_on_part = 1;
    e->last_turning_angle = 0;
    if (obs) augmented_observation(e, obs);
}

/* PaintGymEnv.reset (robot_gym_env.py:370-387) with the start index chosen by the caller */
void oracle_reset(OracleEnv *e, int start_index, double *obs) {
    for (int i = 0; i < e->n; ++i) e->status[i] = (int16_t)e->pack.status_init;   /* :706-708 */
    memset(e->last_affected, 0, (size_t)e->n);
    e->step_counter = 0;
    e->total_return = 0;
    e->total_reward = 0;
    oracle_set_pose(e, e->pack.start_pos + 3 * start_index, e->pack.start_normal + 3 * start_index, obs);
}

/* PaintGymEnv.step (robot_gym_env.py:349-368); (u1,u2) = direction_normalize(clipped action) */
void oracle_step(OracleEnv *e, double u1, double u2, double *obs, double *reward_out,
                 double *penalty_out, double *actual_out, uint8_t *done_out) {
    /* robot.py:396-398 */
    double delta_axis1 = u1 * STEP_SIZE;
    double delta_axis2 = u2 * STEP_SIZE;
    /* robot.py:352-358 _set_turning_angle */
    double new_angle;
    if (delta_axis1 != 0) new_angle = atan(fabs(delta_axis2 / delta_axis1));
    else new_angle = M_PI / 2;
    e->angle_diff = fabs(new_angle - e->last_turning_angle);
    e->last_turning_angle = new_angle;
    int current_on_part_counter = e->terminate_counter;

    /* robot.py:302-329 _get_actions */
    double cur_pose[3], cur_n[3];
    memcpy(cur_pose, e->pose, sizeof cur_pose);
    tcp_orn_norm(e, cur_n);
    double delta1 = delta_axis1 / PAINT_PER_ACTION;
    double delta2 = delta_axis2 / PAINT_PER_ACTION;
    double poses[PAINT_PER_ACTION][3], orns[PAINT_PER_ACTION][4];
    for (int i = 0; i < PAINT_PER_ACTION; ++i) {
        double pos[3], orn_norm[3], q[4];
        int hit = guided_point(e, cur_pose, cur_n, delta1, delta2, pos, orn_norm);
        quat_from_normal(orn_norm, q);
        if (!hit) {
            double off[3] = {delta2, delta1, 0.0};          /* robot.py:317 (sic) */
            transform_point(cur_pose, q, off, pos);
            count_not_on_part(e);
        } else {
            e->last_on_part = 1;
        }
        memcpy(poses[i], pos, sizeof pos);
        memcpy(orns[i], q, sizeof q);
        memcpy(cur_pose, pos, sizeof pos);
        memcpy(cur_n, orn_norm, sizeof orn_norm);
    }

    /* robot.py:401-426 */
    memset(e->possible, 0, (size_t)e->n);
    double succeeded_counter = 0;
    int extended = 0;
    for (int i = 0; i < PAINT_PER_ACTION; ++i) {
        memcpy(e->pose, poses[i], sizeof e->pose);          /* _refresh_robot_pose */
        memcpy(e->orn, orns[i], sizeof e->orn);
        const double shot[3] = {0.0, 0.0, 0.1};
        double center[3];
        transform_point(e->pose, e->orn, shot, center);     /* _get_shot_center :277-278 */
        int n_valid;
        succeeded_counter += e->cfg.paint_method == PAINTRL_PAINT_NORMAL ? normal_paint(e, center, &n_valid)   /* robot.py:414-417 */
                                                                        : fast_paint(e, center, &n_valid);
        extended += n_valid;
    }
    int pixel_counter = 0;
    for (int i = 0; i < e->n; ++i) pixel_counter += e->possible[i];
    double success_rate = extended ? succeeded_counter / pixel_counter : 0;
    if (e->terminate_counter - current_on_part_counter >= PAINT_PER_ACTION && pixel_counter == 0)
        e->terminate = 1;
    e->last_rate = success_rate;
    e->last_succeeded = succeeded_counter;
    e->last_pixel_counter = pixel_counter;

    /* robot_gym_env.py:321-340 */
    double reward = succeeded_counter / 100;
    e->total_reward += reward;
    double penalty = 0.2;
    if (e->cfg.overlap_penalty) penalty += 0.1 * (1 - success_rate);
    if (e->cfg.turning_penalty) penalty += 0.1 * (e->angle_diff / M_PI);
    double actual = reward - penalty;

    /* robot_gym_env.py:289-304 _termination */
    e->step_counter += 1;
    double max_pts = e->cfg.max_possible_point;
    int finished = max_pts > e->total_reward * 100 ? 0 : 1;
    double avg_reward = e->total_reward / e->step_counter;
    double expected_avg = max_pts / (e->cfg.expected_episode_length * 100);
    int done;
    int decided = 0;
    if (avg_reward < expected_avg && e->cfg.termination_mode != PAINTRL_TERM_LATE) {
        if (e->cfg.termination_mode == PAINTRL_TERM_EARLY) { done = 1; decided = 1; }
        else if (e->total_reward < e->cfg.switch_threshold * max_pts / 100) { done = 1; decided = 1; }
    }
    if (!decided)
        done = finished || e->terminate || e->step_counter > e->cfg.episode_max_length - 1;

    augmented_observation(e, obs);
    if (!done) e->total_return += actual;
    *reward_out = reward;
    *penalty_out = penalty;
    *actual_out = actual;
    *done_out = (uint8_t)done;
}

/* Batched drivers (independent environments; one OpenMP thread per chunk). */
void oracle_reset_batch(OracleEnv **envs, int n, const int32_t *start_index, double *obs, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int i = 0; i < n; ++i)
        oracle_reset(envs[i], start_index[i], obs ? obs + (size_t)i * envs[i]->obs_dim : NULL);
}

void oracle_step_batch(OracleEnv **envs, int n, const double *dirs, double *obs, double *reward,
                       double *penalty, double *actual, uint8_t *done, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int i = 0; i < n; ++i)
        oracle_step(envs[i], dirs[2 * i], dirs[2 * i + 1], obs + (size_t)i * envs[i]->obs_dim,
                    reward + i, penalty + i, actual + i, done + i);
}

/* ---- probes */
const int16_t *oracle_status(const OracleEnv *e) { return e->status; }
void oracle_get_pose(const OracleEnv *e, double *pose, double *quat) {
    memcpy(pose, e->pose, 3 * sizeof(double));
    memcpy(quat, e->orn, 4 * sizeof(double));
}
void oracle_get_scalars(const OracleEnv *e, double *out) {
    out[0] = e->total_reward; out[1] = e->total_return; out[2] = e->step_counter;
    out[3] = e->terminate_counter; out[4] = e->last_on_part; out[5] = e->terminate;
    out[6] = e->last_turning_angle; out[7] = e->angle_diff;
    out[8] = e->last_rate; out[9] = e->last_succeeded; out[10] = e->last_pixel_counter;
    out[11] = e->anomalies;
}
void oracle_set_state(OracleEnv *e, const int16_t *status, const double *pose, const double *quat,
                      const double *scalars) {
    if (status) memcpy(e->status, status, sizeof(int16_t) * e->n);
    if (pose) memcpy(e->pose, pose, 3 * sizeof(double));
    if (quat) memcpy(e->orn, quat, 4 * sizeof(double));
    if (scalars) {
        e->total_reward = scalars[0]; e->total_return = scalars[1]; e->step_counter = (int)scalars[2];
        e->terminate_counter = (int)scalars[3]; e->last_on_part = (int)scalars[4];
        e->terminate = (int)scalars[5]; e->last_turning_angle = scalars[6]; e->angle_diff = scalars[7];
    }
}
/* stand-alone pieces for unit tests */
int oracle_ray_test(const OracleEnv *e, const double *frm, const double *to, double *hit) {
    return ray_test(&e->pack, frm, to, hit);
}
int oracle_ball_query(const OracleEnv *e, const double *center, uint8_t *mask) {
    const double r2 = PAINT_RADIUS * PAINT_RADIUS;
    int k = 0;
    for (int i = 0; i < e->n; ++i) {
        const double *c = e->pack.texel_pos + 3 * i;
        double dx = c[0] - center[0], dy = c[1] - center[1], dz = c[2] - center[2];
        mask[i] = (uint8_t)(dx * dx + dy * dy + dz * dz <= r2);
        k += mask[i];
    }
    return k;
}
int oracle_nearest_vertex(const OracleEnv *e, const double *point) {
    int best = -1;
    double best_d = INFINITY;
    for (int i = 0; i < e->pack.n_vertices; ++i) {
        const double *v = e->pack.vertices + 3 * i;
        double dx = v[0] - point[0], dy = v[1] - point[1], dz = v[2] - point[2];
        double d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) { best_d = d; best = i; }
    }
    return best;
}

/* ---------------------------------------------------------------------------------- load time
 * Texel rasterisation of Part.preprocess (bullet_paint_wrapper.py:604-618) for one side: every
 * triangle of the side, in bary_list order, contributes its three corner pixels and every pixel of
 * the corners' bounding box whose normalised coordinate lies inside the UV triangle
 * (BarycentricInterpolator.get_uv_pixels, :191-212); `profile_dicts[side].update(pixel_dict)` lets a
 * later triangle overwrite an earlier one, and inside a triangle the interior value overwrites a
 * corner's.  Restated sequentially with exactly those overwrite semantics.
 *
 * 2-vector np.dot is OpenBLAS ddot's scalar tail fma(x1,y1, x0*y0) (the n=3 chain of npdot3 cut
 * after two terms); np.dot(scalar, vector) is an element-wise product, no FMA.
 *
 * owner[u * height + v]: -1 (not a texel of the side) or tri * 4 + kind, kind 0/1/2 = value of
 * corner a/b/c, 3 = barycentric interior value; pos[(u * height + v) * 3 ..] its 3-D position.
 * Returns the number of texels, or -1 if some pixel coordinate falls outside the texture.
 */
static double npdot2(const double *x, const double *y) { return fma(x[1], y[1], x[0] * y[0]); }

/* bullet_paint_wrapper.py:165-171 _get_pixel_coordinate: Python round() is round-half-even */
static void pixel_coordinate(double u, double v, int width, int height, int *i, int *j) {
    double ri = nearbyint(width * u), rj = nearbyint(height * v);
    *i = ri < width - 1 ? (int)ri : width - 1;
    *j = rj < height - 1 ? (int)rj : height - 1;
}

int oracle_rasterize(const double *tri_a, const double *tri_b, const double *tri_c, const double *tri_uv,
                     int n_tris, int width, int height, int32_t *owner, double *pos) {
    for (long k = 0; k < (long)width * height; ++k) owner[k] = -1;
    for (int t = 0; t < n_tris; ++t) {
        const double *a = tri_a + 3 * t, *b = tri_b + 3 * t, *c = tri_c + 3 * t;
        const double *uva = tri_uv + 6 * t, *uvb = uva + 2, *uvc = uva + 4;
        const double *corner[3] = {a, b, c};
        const double *uv[3] = {uva, uvb, uvc};
        int ci[3], cj[3];
        for (int k = 0; k < 3; ++k) {
            pixel_coordinate(uv[k][0], uv[k][1], width, height, &ci[k], &cj[k]);
            if (ci[k] < 0 || cj[k] < 0) return -1;
        }
        /* pixel_dict = {uva: a, uvb: b, uvc: c}  (:198) -- later keys win */
        for (int k = 0; k < 3; ++k) {
            long o = (long)ci[k] * height + cj[k];
            owner[o] = t * 4 + k;
            memcpy(pos + 3 * o, corner[k], 3 * sizeof(double));
        }
        /* uv_bary = BarycentricInterpolator(uva, uvb, uvc)  (:200, :123-134) */
        double v0[2] = {uvb[0] - uva[0], uvb[1] - uva[1]}, v1[2] = {uvc[0] - uva[0], uvc[1] - uva[1]};
        double d00 = npdot2(v0, v0), d01 = npdot2(v0, v1), d11 = npdot2(v1, v1);
        double denom = d00 * d11 - d01 * d01;
        double inv = denom != 0 ? 1.0 / denom : 0;
        int x_min = ci[0], x_max = ci[0], y_min = cj[0], y_max = cj[0];
        for (int k = 1; k < 3; ++k) {
            if (ci[k] < x_min) x_min = ci[k];
            if (ci[k] > x_max) x_max = ci[k];
            if (cj[k] < y_min) y_min = cj[k];
            if (cj[k] > y_max) y_max = cj[k];
        }
        for (int u = x_min; u <= x_max; ++u)
            for (int v = y_min; v <= y_max; ++v) {
                double pt[2] = {(double)u / width, (double)v / height};
                double v2[2] = {pt[0] - uva[0], pt[1] - uva[1]};
                double d20 = npdot2(v2, v0), d21 = npdot2(v2, v1);
                double bv = (d11 * d20 - d01 * d21) * inv;
                double bw = (d00 * d21 - d01 * d20) * inv;
                double bu = 1.0 - bv - bw;
                if (inv == 0) continue;                          /* (-1,-1,-1): never inside */
                if (!(0 <= bu && bu <= 1 && 0 <= bv && bv <= 1 && 0 <= bw && bw <= 1)) continue;
                long o = (long)u * height + v;
                owner[o] = t * 4 + 3;
                for (int k = 0; k < 3; ++k)                      /* :222-224 */
                    pos[3 * o + k] = (bu * a[k] + bv * b[k]) + bw * c[k];
            }
    }
    int n = 0;
    for (long k = 0; k < (long)width * height; ++k) n += owner[k] >= 0;
    return n;
}
