// paintrl_kernels.cuh -- the fused step kernel and its helpers (one warp per environment).
//
// Per environment and step:
//   move    5 sub-steps of {ray vs hull, nearest vertex, closest triangle}    robot.py:302-329
//   stamp   5 ball queries over the binned texel table, colour update, overlap bookkeeping
//                                                     bullet_paint_wrapper.py:568-577, 352-434
//   score   reward / penalty / termination                                   robot_gym_env.py:289-340
//   observe normalised pose + section/grid observation
//                                                     bullet_paint_wrapper.py:965-978, 1045-1139
// Per-environment arrays: the status plane (1 byte per front texel in RGB mode, int16 in HSI mode,
// texels sorted by spatial bin) and one 16-bit counter per bin = number of texels of the bin whose
// "painted" predicate (first channel == 255) differs from the fresh texture's.  The stamp keeps
// the counters current, so the 4-sector observation reads the counters of the bins that lie
// wholly inside one sector and classifies texel by texel only the pose's bin row and bin column.
#pragma once
#include "paintrl_device.cuh"

namespace paintrl {

template <int COLOR> struct StatusT;
template <> struct StatusT<0> { typedef uint8_t type; };
template <> struct StatusT<1> { typedef int16_t type; };

struct StepIO {
    const void *actions;
    double *obs, *reward, *penalty, *actual;
    uint8_t *done;
    int32_t *new_texels;
    double *next_obs;
    const int32_t *reset_start_idx;
    unsigned long long *stats;   // [0] env steps, [1] episodes ended, [2] footprint texels, [3] full-plane ray scans
};

// Per-environment dynamic arrays.
template <int COLOR>
struct EnvArrays {
    EnvState *states;
    typename StatusT<COLOR>::type *planes;   // [num_envs][n_pad]
    unsigned *bin_cnt;                       // [num_envs][n_bins_pad / 2]   two 16-bit counters per word
    unsigned *grid_cnt;                      // [num_envs][n_gcells_pad]     grid-observation cells (grid mode only)
};

// Per-warp shared scratch.
struct WarpScratch {
    double centers[kPaintPerAction][3];
    int seg_begin[32];
    int seg_cum[33];
    int hist[2 * kMaxObs];                   // K != 4 section histogram
};

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// Visits every index of a list of contiguous index ranges with all 32 lanes busy.
//   seg(s, begin, len)   range s of n_seg (called by lane s % 32)
//   body(j, active)      called warp-uniformly; j is valid when active
template <typename SegFn, typename BodyFn>
__device__ __forceinline__ void for_each_in_segments(int n_seg, int lane, WarpScratch &ws, SegFn seg, BodyFn body) {
    for (int s0 = 0; s0 < n_seg; s0 += 32) {
        int b = 0, l = 0;
        if (s0 + lane < n_seg) seg(s0 + lane, b, l);
        const int incl = warp_inclusive_scan(l, lane);
        __syncwarp();
        ws.seg_begin[lane] = b;
        ws.seg_cum[lane + 1] = incl;
        if (lane == 0) ws.seg_cum[0] = 0;
        __syncwarp();
        const int total = __shfl_sync(kFull, incl, 31);
        int k = 0;
        for (int it0 = 0; it0 < total; it0 += 32) {
            const int it = it0 + lane;
            const bool active = it < total;
            int j = 0;
            if (active) {
                while (it >= ws.seg_cum[k + 1]) ++k;
                j = ws.seg_begin[k] + (it - ws.seg_cum[k]);
            }
            body(j, active);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------ observation
// 4-sector observation (bullet_paint_wrapper.py:1033-1061): for every front texel, rx / ry = texel
// position - TCP position along the principal axes; skipped if both are 0; sector 0 if rx>0,ry>0,
// 1 if rx<0,ry>0, 2 if rx<0,ry<0, else 3; obs[s] = #(status != 255) / #texels of the sector.
//
// A texel's bin index along an axis is a monotone function of its coordinate, evaluated with the
// same two FP64 operations for texels (host) and pose (here), so every texel of a bin left of /
// right of / above / below the pose's bin has rx < 0 / rx > 0 / ry > 0 / ry < 0: such bins belong
// wholly to one sector and contribute their size (2-D prefix table) and their flip counter.
// Only the texels of the pose's own bin row and bin column are compared coordinate by coordinate.
template <int COLOR>
__device__ __forceinline__ void section4_counts(const DevPack &pk, const typename StatusT<COLOR>::type *status,
                                                const unsigned *bin_cnt, const Vec3 &pose, int lane, WarpScratch &ws,
                                                int tot[4], int open[4]) {
    const double p0 = comp(pose, pk.axis0), p1 = comp(pose, pk.axis1);
    const int nx = pk.tb_nx, ny = pk.tb_ny;
    double f0 = floor((p0 - pk.tb_o0) * pk.tb_inv), f1 = floor((p1 - pk.tb_o1) * pk.tb_inv);
    const int px = (f0 < 0.0) ? -1 : (f0 >= (double)nx ? nx : (int)f0);
    const int py = (f1 < 0.0) ? -1 : (f1 >= (double)ny ? ny : (int)f1);

    // ---- bins wholly inside a sector: sizes from the static prefix table
    const int W = nx + 1;
    auto P = [&](int iy, int ix) { return __ldg(&pk.tb_prefix[iy * W + ix]); };   // sum over rows < iy, cols < ix
    const int xl = max(px, 0), xr = min(px + 1, nx), yb = max(py, 0), ya = min(py + 1, ny);
    const int P_ny_nx = P(ny, nx), P_ny_xr = P(ny, xr), P_ny_xl = P(ny, xl);
    const int P_ya_nx = P(ya, nx), P_ya_xr = P(ya, xr), P_ya_xl = P(ya, xl);
    const int P_yb_nx = P(yb, nx), P_yb_xr = P(yb, xr), P_yb_xl = P(yb, xl);
    int pure_tot[4];
    pure_tot[0] = (P_ny_nx - P_ny_xr) - (P_ya_nx - P_ya_xr);   // right, above
    pure_tot[1] = P_ny_xl - P_ya_xl;                           // left, above
    pure_tot[2] = P_yb_xl;                                     // left, below
    pure_tot[3] = P_yb_nx - P_yb_xr;                           // right, below

    // ---- their flip counters: two 16-bit counters per word, lanes own word columns
    int flips[4] = {0, 0, 0, 0};
    const int wpr = nx >> 1;
    for (int w0 = 0; w0 < wpr; w0 += 32) {
        const int w = w0 + lane;
        unsigned mask_l = 0, mask_r = 0;
        if (w < wpr) {
            mask_l = (2 * w < px ? 0x0000ffffu : 0u) | (2 * w + 1 < px ? 0xffff0000u : 0u);
            mask_r = (2 * w > px ? 0x0000ffffu : 0u) | (2 * w + 1 > px ? 0xffff0000u : 0u);
        }
        unsigned al = 0, ar = 0, bl = 0, br = 0;        // packed 2 x 16-bit partial sums
        const unsigned *col = bin_cnt + w;
        const int wl = (w < wpr) ? w : 0;
        (void)wl;
        int iy = 0;
        const int below_end = min(max(py, 0), ny);
        if (w < wpr) {
#pragma unroll 4
            for (iy = 0; iy < below_end; ++iy) {
                unsigned c = __ldcg(col + (size_t)iy * wpr);
                bl += c & mask_l;
                br += c & mask_r;
            }
#pragma unroll 4
            for (iy = min(py + 1, ny); iy < ny; ++iy) {
                if (iy < 0) continue;
                unsigned c = __ldcg(col + (size_t)iy * wpr);
                al += c & mask_l;
                ar += c & mask_r;
            }
        }
        flips[0] += (int)(ar & 0xffffu) + (int)(ar >> 16);
        flips[1] += (int)(al & 0xffffu) + (int)(al >> 16);
        flips[2] += (int)(bl & 0xffffu) + (int)(bl >> 16);
        flips[3] += (int)(br & 0xffffu) + (int)(br >> 16);
    }

    // ---- the pose's bin row and bin column, texel by texel
    const double *c0 = pk.axis0 == 0 ? pk.tx : (pk.axis0 == 1 ? pk.ty : pk.tz);
    const double *c1 = pk.axis1 == 0 ? pk.tx : (pk.axis1 == 1 ? pk.ty : pk.tz);
    const bool row_ok = (py >= 0 && py < ny), col_ok = (px >= 0 && px < nx);
    const int n_seg = (row_ok ? 1 : 0) + (col_ok ? ny - (row_ok ? 1 : 0) : 0);
    unsigned long long ptot = 0, popen = 0;   // 4 x 16-bit fields per lane
    int carry_tot[4] = {0, 0, 0, 0}, carry_open[4] = {0, 0, 0, 0};
    int since_flush = 0;
    for_each_in_segments(
        n_seg, lane, ws,
        [&](int s, int &b, int &l) {
            if (row_ok && s == 0) {
                b = __ldg(&pk.tb_start[py * nx]);
                l = __ldg(&pk.tb_start[py * nx + nx]) - b;
            } else {
                int iy = s - (row_ok ? 1 : 0);
                if (row_ok && iy >= py) ++iy;
                b = __ldg(&pk.tb_start[iy * nx + px]);
                l = __ldg(&pk.tb_start[iy * nx + px + 1]) - b;
            }
        },
        [&](int j, bool active) {
            if (active) {
                const double x0 = __ldg(&c0[j]), x1 = __ldg(&c1[j]);
                const int s = (int)__ldcg(&status[j]);
                const bool gx = x0 > p0, lx = x0 < p0, gy = x1 > p1, ly = x1 < p1;
                const bool skip = !(gx || lx || gy || ly);
                const int q = (gx && gy) ? 0 : ((lx && gy) ? 1 : ((lx && ly) ? 2 : 3));
                const unsigned long long one = skip ? 0ull : (1ull << (16 * q));
                ptot += one;
                popen += (s != kPainted) ? one : 0ull;
            }
            if (++since_flush == 0xffff) {      // keep the 16-bit fields from overflowing
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    carry_tot[q] += (int)((ptot >> (16 * q)) & 0xffff);
                    carry_open[q] += (int)((popen >> (16 * q)) & 0xffff);
                }
                ptot = popen = 0;
                since_flush = 0;
            }
        });
    const bool init_painted = (pk.status_init == kPainted);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int st = __reduce_add_sync(kFull, carry_tot[q] + (int)((ptot >> (16 * q)) & 0xffff));
        const int so = __reduce_add_sync(kFull, carry_open[q] + (int)((popen >> (16 * q)) & 0xffff));
        const int fl = __reduce_add_sync(kFull, flips[q]);
        tot[q] = pure_tot[q] + st;
        open[q] = (init_painted ? fl : pure_tot[q] - fl) + so;
    }
}

// Section observation with K != 4 sectors (atan2 path, bullet_paint_wrapper.py:1026-1031): full scan.
template <int COLOR>
__device__ __forceinline__ void sectionk_counts(const DevPack &pk, const typename StatusT<COLOR>::type *status,
                                                const Vec3 &pose, int section, int lane, int *hist /*[2*kMaxObs] smem*/) {
    for (int i = lane; i < 2 * kMaxObs; i += 32) hist[i] = 0;
    __syncwarp();
    double p0 = comp(pose, pk.axis0), p1 = comp(pose, pk.axis1);
    const double *c0 = pk.axis0 == 0 ? pk.tx : (pk.axis0 == 1 ? pk.ty : pk.tz);
    const double *c1 = pk.axis1 == 0 ? pk.tx : (pk.axis1 == 1 ? pk.ty : pk.tz);
    double basis = 2 * kPi / section;
    for (int j = lane; j < pk.n_texels; j += 32) {
        double rx = __ldg(&c0[j]) - p0, ry = __ldg(&c1[j]) - p1;
        if (rx == 0.0 && ry == 0.0) continue;
        double angle = atan2(ry, rx);
        if (angle < 0.0) angle = 2 * kPi + angle;
        int idx = (int)np_floor_divide(angle, basis);
        if (idx >= section) idx = section - 1;
        atomicAdd(&hist[idx], 1);
        if ((int)__ldcg(&status[j]) != kPainted) atomicAdd(&hist[kMaxObs + idx], 1);
    }
    __syncwarp();
}

// robot_gym_env.py:306-319 _augmented_observation; every lane returns, lanes < obs_dim write.
template <int COLOR>
__device__ __forceinline__ void write_observation(const DevPack &pk, const DevConfig &cfg,
                                                  const typename StatusT<COLOR>::type *status, const unsigned *bin_cnt,
                                                  const unsigned *grid_cnt, const Vec3 &pose, int lane, WarpScratch &ws,
                                                  double *obs_a, double *obs_b) {
    double a1, a2;
    normalized_pose(pk, pose, a1, a2);
    const int grad = cfg.obs_grad;
    if (cfg.obs_mode == 2) {   // simple
        if (lane < 2) {
            double v = lane == 0 ? a1 : a2;
            if (obs_a) obs_a[lane] = v;
            if (obs_b) obs_b[lane] = v;
        }
        return;
    }
    if (cfg.obs_mode == 1) {   // grid (bullet_paint_wrapper.py:1126-1139): painted texels per cell from the flip counters
        const int cells = grad * grad;
        const bool init_painted = (pk.status_init == kPainted);
        for (int c = lane; c < cells; c += 32) {
            const int total = __ldg(&pk.gtotal[c]);
            const int fl = (int)__ldcg(&grid_cnt[c]);
            const int painted = init_painted ? total - fl : fl;
            double v = total == 0 ? 0.0 : 1.0 - (double)painted / (double)total;
            if (obs_a) obs_a[c] = v;
            if (obs_b) obs_b[c] = v;
        }
        return;
    }
    // section / discrete
    if (grad == 4) {
        int tot[4], open[4];
        section4_counts<COLOR>(pk, status, bin_cnt, pose, lane, ws, tot, open);
        if (lane < 4) {
            int t = lane == 0 ? tot[0] : (lane == 1 ? tot[1] : (lane == 2 ? tot[2] : tot[3]));
            int o = lane == 0 ? open[0] : (lane == 1 ? open[1] : (lane == 2 ? open[2] : open[3]));
            double v = t == 0 ? 0.0 : (double)o / (double)t;
            if (obs_a) obs_a[lane] = v;
            if (obs_b) obs_b[lane] = v;
        }
    } else {
        sectionk_counts<COLOR>(pk, status, pose, grad, lane, ws.hist);
        for (int s = lane; s < grad; s += 32) {
            int t = ws.hist[s], o = ws.hist[kMaxObs + s];
            double v = t == 0 ? 0.0 : (double)o / (double)t;
            if (obs_a) obs_a[s] = v;
            if (obs_b) obs_b[s] = v;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (cfg.obs_mode == 3) {   // discrete: robot_gym_env.py:101-103, 314-318
            int position = (handle_pos(a1) + 1) * 22 + handle_pos(a2);
            double v = 1.0 / position;
            if (obs_a) obs_a[grad] = v;
            if (obs_b) obs_b[grad] = v;
        } else {
            if (obs_a) { obs_a[grad] = a1; obs_a[grad + 1] = a2; }
            if (obs_b) { obs_b[grad] = a1; obs_b[grad + 1] = a2; }
        }
    }
}

// ------------------------------------------------------------------------------ reset pieces
// Part.reset_part (bullet_paint_wrapper.py:706-708): restore the init colour and clear the flip
// counters, 128-bit stores.
template <int COLOR>
__device__ __forceinline__ void fill_status(const DevPack &pk, typename StatusT<COLOR>::type *status, unsigned *bin_cnt,
                                            unsigned *grid_cnt, int lane) {
    typedef typename StatusT<COLOR>::type S;
    constexpr int kPer = 16 / sizeof(S);
    uint4 v;
    S *e = reinterpret_cast<S *>(&v);
#pragma unroll
    for (int k = 0; k < kPer; ++k) e[k] = (S)pk.status_init;
    for (int j0 = lane * kPer; j0 < pk.n_pad; j0 += 32 * kPer) *reinterpret_cast<uint4 *>(status + j0) = v;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int w = lane * 4; w < pk.n_bins_pad / 2; w += 128) *reinterpret_cast<uint4 *>(bin_cnt + w) = z;
    if (grid_cnt)
        for (int w = lane * 4; w < pk.n_gcells_pad; w += 128) *reinterpret_cast<uint4 *>(grid_cnt + w) = z;
}

// Robot.reset(pose) (robot.py:366-372, 208-212)
__device__ __forceinline__ void robot_reset(EnvState &st, const double *pos, const double *normal) {
    Vec3 n = {normal[0], normal[1], normal[2]};
    quat_from_normal(n, st.quat);
    st.pose[0] = pos[0]; st.pose[1] = pos[1]; st.pose[2] = pos[2];
    st.flags = (st.flags & kFlagHasLast) | kFlagLastOnPart;   // terminate cleared, last_on_part = True
    st.term_counter = 0;
    st.last_angle = 0.0;
}

// PaintGymEnv.reset (robot_gym_env.py:370-387) minus the observation
template <int COLOR>
__device__ __forceinline__ void env_reset(const DevPack &pk, EnvState &st, typename StatusT<COLOR>::type *status,
                                          unsigned *bin_cnt, unsigned *grid_cnt, int start_index, int lane) {
    fill_status<COLOR>(pk, status, bin_cnt, grid_cnt, lane);
    st.flags &= ~kFlagHasLast;                 // _last_painted_pixels = []
    st.step_counter = 0;
    st.total_return = 0.0;
    st.total_reward = 0.0;
    robot_reset(st, pk.start_pos + 3 * start_index, pk.start_normal + 3 * start_index);
    st.episode += 1;
}

__device__ __forceinline__ void load_state(const EnvState *g, EnvState &st) {
    const double2 *src = reinterpret_cast<const double2 *>(g);
    double2 *dst = reinterpret_cast<double2 *>(&st);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = src[i];
}
__device__ __forceinline__ void store_state(EnvState *g, const EnvState &st, int lane) {
    // lanes 0..7 each write one 16-byte piece of the record (the record is warp-uniform)
    const double2 *src = reinterpret_cast<const double2 *>(&st);
    double2 v = src[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) if (lane == i) v = src[i];
    if (lane < 8) reinterpret_cast<double2 *>(g)[lane] = v;
}

template <int COLOR>
__device__ __forceinline__ unsigned *bin_cnt_of(const DevPack &pk, const EnvArrays<COLOR> &ea, int env) {
    return ea.bin_cnt + (size_t)env * (pk.n_bins_pad >> 1);
}
template <int COLOR>
__device__ __forceinline__ unsigned *grid_cnt_of(const DevPack &pk, const EnvArrays<COLOR> &ea, int env) {
    return ea.grid_cnt ? ea.grid_cnt + (size_t)env * pk.n_gcells_pad : nullptr;
}

__device__ __forceinline__ int auto_start_index(const DevPack &pk, const DevConfig &cfg, int env, int episode) {
    return (int)(splitmix64(cfg.seed ^ splitmix64(((unsigned long long)env << 32) | (unsigned)episode)) %
                 (unsigned long long)pk.n_starts);
}

// ------------------------------------------------------------------------------ stamp
// Part.fast_paint x 5 (bullet_paint_wrapper.py:568-577) + colour handlers (:352-434).
//
// Candidates come from the texel bins overlapping the shots' bounding box, flattened over the bin
// rows.  The ball test `dx*dx + dy*dy + dz*dz <= r*r` (FP64, exactly the kd-tree's) is decided in
// FP32 on origin-relative coordinates whenever the FP32 value is further than kBallEps from r*r --
// the FP32 evaluation differs from the FP64 one by < 2e-8 for |d| <= 2r -- and in FP64 otherwise.
constexpr float kBallEps = 1e-7f;

template <int COLOR>
__device__ __forceinline__ void stamp(const DevPack &pk, const DevConfig &cfg, typename StatusT<COLOR>::type *status,
                                      unsigned *bin_cnt, unsigned *grid_cnt, const Vec3 &lastc, bool has_last, int lane,
                                      WarpScratch &ws, int &n_new_out, int &n_possible_out) {
    typedef typename StatusT<COLOR>::type S;
    constexpr int NS = kPaintPerAction;
    const double r2 = kPaintRadius * kPaintRadius;
    const float r2f = (float)r2;
    double cx[NS + 1], cy[NS + 1], cz[NS + 1];
#pragma unroll
    for (int s = 0; s < NS; ++s) { cx[s] = ws.centers[s][0]; cy[s] = ws.centers[s][1]; cz[s] = ws.centers[s][2]; }
    cx[NS] = lastc.x; cy[NS] = lastc.y; cz[NS] = lastc.z;
    float fx[NS + 1], fy[NS + 1], fz[NS + 1];
#pragma unroll
    for (int s = 0; s <= NS; ++s) {
        fx[s] = (float)(cx[s] - pk.org0); fy[s] = (float)(cy[s] - pk.org1); fz[s] = (float)(cz[s] - pk.org2);
    }
    double lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double c0 = pk.axis0 == 0 ? cx[s] : (pk.axis0 == 1 ? cy[s] : cz[s]);
        double c1 = pk.axis1 == 0 ? cx[s] : (pk.axis1 == 1 ? cy[s] : cz[s]);
        lo0 = fmin(lo0, c0); hi0 = fmax(hi0, c0);
        lo1 = fmin(lo1, c1); hi1 = fmax(hi1, c1);
    }
    const double margin = kPaintRadius + 1e-9;
    // clamp in FP64 first: an off-part TCP can be far away from the bin grid
    double b0 = floor((lo0 - margin - pk.tb_o0) * pk.tb_inv), b1 = floor((hi0 + margin - pk.tb_o0) * pk.tb_inv);
    double b2 = floor((lo1 - margin - pk.tb_o1) * pk.tb_inv), b3 = floor((hi1 + margin - pk.tb_o1) * pk.tb_inv);
    const int bx0 = (int)fmax(b0, 0.0), bx1 = (int)fmin(b1, (double)(pk.tb_nx - 1));
    const int by0 = (int)fmax(b2, 0.0), by1 = (int)fmin(b3, (double)(pk.tb_ny - 1));
    const bool empty = !(b1 >= 0.0 && b3 >= 0.0 && b0 <= (double)(pk.tb_nx - 1) && b2 <= (double)(pk.tb_ny - 1));
    const int n_rows = empty ? 0 : (by1 - by0 + 1);
    const int nx = pk.tb_nx;
    auto row_seg = [&](int s, int &b, int &l) {
        b = __ldg(&pk.tb_start[(by0 + s) * nx + bx0]);
        l = __ldg(&pk.tb_start[(by0 + s) * nx + bx1 + 1]) - b;
    };
    // which shots (bits 0..4) and the previous step's last shot (bit 5) contain texel j
    auto ball_mask = [&](int j, const float4 &t) -> unsigned {
        unsigned in = 0, amb = 0;
#pragma unroll
        for (int s = 0; s <= NS; ++s) {
            float dx = t.x - fx[s], dy = t.y - fy[s], dz = t.z - fz[s];
            float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            in |= (d2 <= r2f ? 1u : 0u) << s;
            amb |= (fabsf(d2 - r2f) <= kBallEps ? 1u : 0u) << s;
        }
        if (amb) {
            const double x = __ldg(&pk.tx[j]), y = __ldg(&pk.ty[j]), z = __ldg(&pk.tz[j]);
            in = 0;
#pragma unroll
            for (int s = 0; s <= NS; ++s) {
                double dx = x - cx[s], dy = y - cy[s], dz = z - cz[s];
                in |= ((dx * dx + dy * dy + dz * dz) <= r2 ? 1u : 0u) << s;
            }
        }
        if (!has_last) in &= (1u << NS) - 1u;
        return in;
    };

    double rmax[NS];
    if (COLOR == 1) {   // HSI: r = distances.max() per shot (bullet_paint_wrapper.py:423-424)
#pragma unroll
        for (int s = 0; s < NS; ++s) rmax[s] = -1.0;
        for_each_in_segments(n_rows, lane, ws, row_seg, [&](int j, bool active) {
            if (!active) return;
            const float4 t = __ldg(&pk.trel[j]);
            const unsigned in = ball_mask(j, t) & ((1u << NS) - 1u);
            if (in) {
                const double x = __ldg(&pk.tx[j]), y = __ldg(&pk.ty[j]), z = __ldg(&pk.tz[j]);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (in & (1u << s)) {
                        double dx = x - cx[s], dy = y - cy[s], dz = z - cz[s];
                        rmax[s] = fmax(rmax[s], dx * dx + dy * dy + dz * dz);
                    }
                }
            }
        });
#pragma unroll
        for (int s = 0; s < NS; ++s) rmax[s] = sqrt(warp_max(rmax[s]));   // sqrt is monotone
    }

    int n_new = 0, n_possible = 0;   // per-lane partial counts
    for_each_in_segments(n_rows, lane, ws, row_seg, [&](int j, bool active) {
        if (!active) return;
        const float4 t = __ldg(&pk.trel[j]);
        const unsigned in = ball_mask(j, t);
        const unsigned shots = in & ((1u << NS) - 1u);
        if (!shots) return;
        // affected \ last_affected, shot by shot (:575): shot s counts if the previous shot missed the texel
        const unsigned prev = ((shots << 1) | (in >> NS)) & ((1u << NS) - 1u);
        n_possible += (shots & ~prev) ? 1 : 0;
        bool flipped = false;
        if (COLOR == 0) {                              // :358-365
            if ((int)__ldcg(&status[j]) != kPainted) { status[j] = (S)kPainted; n_new += 1; flipped = true; }
        } else {                                       // :411-434
            const int sv0 = (int)__ldcg(&status[j]);
            int sv = sv0;
            const double x = __ldg(&pk.tx[j]), y = __ldg(&pk.ty[j]), z = __ldg(&pk.tz[j]);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if ((shots & (1u << s)) && sv > 0) {
                    double dx = x - cx[s], dy = y - cy[s], dz = z - cz[s];
                    double ratio = sqrt(dx * dx + dy * dy + dz * dz) / rmax[s];
                    int quantity = (int)(kHsiTargetMax * (1.0 - ratio * ratio)) + 1;   // :429
                    sv -= quantity;
                    n_new += quantity;
                }
            }
            if (sv != sv0) {
                status[j] = (S)sv;
                flipped = (sv0 == kPainted);           // values only decrease: 255 is left once
            }
        }
        if (flipped) {
            const unsigned w = __float_as_uint(t.w);
            const unsigned bin = w & 0xffffu;
            atomicAdd(bin_cnt + (bin >> 1), 1u << ((bin & 1u) * 16));
            if (grid_cnt) atomicAdd(grid_cnt + (w >> 16), 1u);
        }
    });
    n_new_out = __reduce_add_sync(kFull, n_new);
    n_possible_out = __reduce_add_sync(kFull, n_possible);
    __syncwarp();
}

// ------------------------------------------------------------------------------ kernels
template <int COLOR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
reset_kernel(DevPack pk, DevConfig cfg, EnvArrays<COLOR> ea, const int32_t *env_ids, int n, const int32_t *start_idx,
             const double *set_pos, const double *set_normal, double *obs_out, int mode /*0 reset, 1 set_pose*/) {
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = blockIdx.x * kWarpsPerBlock + warp;
    if (k >= n) return;
    const int env = env_ids ? env_ids[k] : k;
    typename StatusT<COLOR>::type *status = ea.planes + (size_t)env * pk.n_pad;
    unsigned *bin_cnt = bin_cnt_of<COLOR>(pk, ea, env), *grid_cnt = grid_cnt_of<COLOR>(pk, ea, env);
    EnvState st;
    load_state(&ea.states[env], st);
    if (mode == 0) {
        int idx = start_idx ? start_idx[k] : auto_start_index(pk, cfg, env, st.episode);
        idx = min(max(idx, 0), pk.n_starts - 1);
        env_reset<COLOR>(pk, st, status, bin_cnt, grid_cnt, idx, lane);
    } else {
        robot_reset(st, set_pos + 3 * k, set_normal + 3 * k);
    }
    __syncwarp();
    Vec3 pose = {st.pose[0], st.pose[1], st.pose[2]};
    write_observation<COLOR>(pk, cfg, status, bin_cnt, grid_cnt, pose, lane, scratch[warp],
                             obs_out ? obs_out + (size_t)k * cfg.obs_dim : nullptr, nullptr);
    store_state(&ea.states[env], st, lane);
}

template <int COLOR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
step_kernel(DevPack pk, DevConfig cfg, EnvArrays<COLOR> ea, int num_envs, StepIO io) {
    typedef typename StatusT<COLOR>::type S;
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env = blockIdx.x * kWarpsPerBlock + warp;
    if (env >= num_envs) return;
    WarpScratch &ws = scratch[warp];
    S *status = ea.planes + (size_t)env * pk.n_pad;
    unsigned *bin_cnt = bin_cnt_of<COLOR>(pk, ea, env), *grid_cnt = grid_cnt_of<COLOR>(pk, ea, env);
    EnvState st;
    load_state(&ea.states[env], st);

    // ---- action -> direction (robot_gym_env.py:342-347, robot.py:390-398, 352-358)
    double u1, u2, new_angle;
    if (cfg.action_mode == 0) {
        long long a = reinterpret_cast<const long long *>(io.actions)[env];
        int ai = (int)min(max(a, 0ll), (long long)cfg.discrete_granularity - 1);
        u1 = __ldg(&cfg.discrete_table[3 * ai]);
        u2 = __ldg(&cfg.discrete_table[3 * ai + 1]);
        new_angle = __ldg(&cfg.discrete_table[3 * ai + 2]);
    } else {
        const double *a = reinterpret_cast<const double *>(io.actions) + (size_t)env * cfg.action_shape;
        double a0 = a[0];
        if (!(-1.0 <= a0 && a0 <= 1.0)) a0 = a0 < -1.0 ? -1.0 : (a0 > 1.0 ? 1.0 : a0);
        if (cfg.action_shape == 1) {
            double phi = (a0 + 1.0) * kPi;
            u1 = cos(phi);
            u2 = sin(phi);
        } else {
            double a1v = a[1];
            if (!(-1.0 <= a1v && a1v <= 1.0)) a1v = a1v < -1.0 ? -1.0 : (a1v > 1.0 ? 1.0 : a1v);
            double phi = atan2(a1v, a0);
            double x = fabs(a0), y = fabs(a1v);
            if (x == 0.0 && y == 0.0) { u1 = x; u2 = y; }
            else { double m = fmax(x, y); u1 = m * cos(phi); u2 = m * sin(phi); }
        }
        double da1 = u1 * kStepSize, da2 = u2 * kStepSize;
        new_angle = (da1 != 0.0) ? atan(fabs(da2 / da1)) : kPi / 2;
    }
    const double delta_axis1 = u1 * kStepSize, delta_axis2 = u2 * kStepSize;
    st.angle_diff = fabs(new_angle - st.last_angle);
    st.last_angle = new_angle;
    const int counter_before = st.term_counter;

    // ---- Robot._get_actions: 5 guided sub-steps (robot.py:302-329, bullet_paint_wrapper.py:865-880)
    Vec3 cur_p = {st.pose[0], st.pose[1], st.pose[2]};
    Vec3 cur_n = tcp_orn_norm(cur_p, st.quat);
    const double delta1 = delta_axis1 / kPaintPerAction, delta2 = delta_axis2 / kPaintPerAction;
    const double delta2_scaled = delta2 * pk.lwr;
    int full_scans = 0;
    double quat[4] = {st.quat[0], st.quat[1], st.quat[2], st.quat[3]};
    bool miss_quat_valid = false;   // quat == quat_from_normal(cur_n) from an earlier miss of this step
#pragma unroll 1
    for (int s = 0; s < kPaintPerAction; ++s) {
        Vec3 p = cur_p;
        add_comp(p, pk.axis0, delta1);
        add_comp(p, pk.axis1, delta2_scaled);
        Vec3 end = {p.x + cur_n.x, p.y + cur_n.y, p.z + cur_n.z};
        Vec3 hit, pos, center;
        int cell;
        const double *rec = nullptr;
        if (ray_test(pk, p, end, lane, hit, cell, full_scans)) rec = hook_triangle(pk, hit, cell, lane);
        if (rec) {
            // pose = hit + 0.1 n, orn = -n (bullet_paint_wrapper.py:529-530); quaternion and shot-centre
            // offset of -n come from the record (computed by the host with these same operations)
            const double2 *t = reinterpret_cast<const double2 *>(rec + 12);   // [12] inv, [13..15] n, [16..19] q, [20..22] off
            const double2 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2), t3 = __ldg(t + 3), t4 = __ldg(t + 4),
                          t5 = __ldg(t + 5);
            const double nx = t0.y, ny = t1.x, nz = t1.y;
            pos.x = hit.x + nx * kHookDistance;
            pos.y = hit.y + ny * kHookDistance;
            pos.z = hit.z + nz * kHookDistance;
            quat[0] = t2.x; quat[1] = t2.y; quat[2] = t3.x; quat[3] = t3.y;
            center.x = t4.x + pos.x; center.y = t4.y + pos.y; center.z = t5.x + pos.z;   // robot.py:277-278
            cur_n.x = -nx; cur_n.y = -ny; cur_n.z = -nz;
            st.flags |= kFlagLastOnPart;
            miss_quat_valid = false;
        } else {
            if (!miss_quat_valid) { quat_from_normal(cur_n, quat); miss_quat_valid = true; }
            pos = transform_point(cur_p, quat, delta2, delta1, 0.0);   // robot.py:317 (sic)
            center = transform_point(pos, quat, 0.0, 0.0, 0.1);        // robot.py:277-278
            if (st.flags & kFlagLastOnPart) {                           // robot.py:292-300
                st.flags &= ~kFlagLastOnPart;
            } else {
                st.term_counter += 1;
                if (st.term_counter > kNotOnPartTerminateSteps) st.flags |= kFlagTerminate;
            }
        }
        if (lane == 0) { ws.centers[s][0] = center.x; ws.centers[s][1] = center.y; ws.centers[s][2] = center.z; }
        cur_p = pos;
    }
    st.pose[0] = cur_p.x; st.pose[1] = cur_p.y; st.pose[2] = cur_p.z;
    st.quat[0] = quat[0]; st.quat[1] = quat[1]; st.quat[2] = quat[2]; st.quat[3] = quat[3];
    __syncwarp();

    // ---- stamp the 5 shots (bullet_paint_wrapper.py:568-577)
    const bool has_last = (st.flags & kFlagHasLast) != 0;
    const Vec3 lastc = {st.last_center[0], st.last_center[1], st.last_center[2]};
    int n_new, n_possible;
    stamp<COLOR>(pk, cfg, status, bin_cnt, grid_cnt, lastc, has_last, lane, ws, n_new, n_possible);
    st.last_center[0] = ws.centers[kPaintPerAction - 1][0];
    st.last_center[1] = ws.centers[kPaintPerAction - 1][1];
    st.last_center[2] = ws.centers[kPaintPerAction - 1][2];
    st.flags |= kFlagHasLast;

    // ---- robot.py:425-433, robot_gym_env.py:321-340
    const double succeeded = (COLOR == 0) ? (double)n_new : (double)n_new / 255.0;
    const double rate = n_possible ? succeeded / (double)n_possible : 0.0;
    if (st.term_counter - counter_before >= kPaintPerAction && n_possible == 0) st.flags |= kFlagTerminate;
    const double reward = succeeded / 100;
    st.total_reward += reward;
    double penalty = 0.2;
    if (cfg.overlap_penalty) penalty += 0.1 * (1 - rate);
    if (cfg.turning_penalty) penalty += 0.1 * (st.angle_diff / kPi);
    const double actual = reward - penalty;

    // ---- robot_gym_env.py:289-304 _termination
    st.step_counter += 1;
    const double max_pts = cfg.max_possible_point;
    const bool finished = !(max_pts > st.total_reward * 100);
    const double avg_reward = st.total_reward / st.step_counter;
    bool done, decided = false;
    if (avg_reward < cfg.expected_avg_reward && cfg.termination_mode != 0) {
        if (cfg.termination_mode == 1) { done = true; decided = true; }
        else if (st.total_reward < cfg.hybrid_threshold) { done = true; decided = true; }
    }
    if (!decided) done = finished || (st.flags & kFlagTerminate) || st.step_counter > cfg.episode_max_length - 1;

    // ---- observation (computed even when done, robot_gym_env.py:358)
    const bool resetting = done && cfg.auto_reset;
    double *obs = io.obs + (size_t)env * cfg.obs_dim;
    double *next_obs = io.next_obs ? io.next_obs + (size_t)env * cfg.obs_dim : nullptr;
    write_observation<COLOR>(pk, cfg, status, bin_cnt, grid_cnt, cur_p, lane, ws, obs, resetting ? nullptr : next_obs);
    if (!done) st.total_return += actual;
    if (lane == 0) {
        io.reward[env] = reward;
        io.penalty[env] = penalty;
        io.actual[env] = actual;
        io.done[env] = done ? 1 : 0;
        if (io.new_texels) io.new_texels[env] = n_new;
        atomicAdd(&io.stats[0], 1ull);
        atomicAdd(&io.stats[2], (unsigned long long)n_possible);
        if (done) atomicAdd(&io.stats[1], 1ull);
        if (full_scans) atomicAdd(&io.stats[3], (unsigned long long)full_scans);
    }

    // ---- same-step auto-reset: `obs` keeps the terminal observation, `next_obs` gets reset()'s
    if (resetting) {
        int idx = io.reset_start_idx ? io.reset_start_idx[env] : auto_start_index(pk, cfg, env, st.episode);
        idx = min(max(idx, 0), pk.n_starts - 1);
        __syncwarp();
        env_reset<COLOR>(pk, st, status, bin_cnt, grid_cnt, idx, lane);
        __syncwarp();
        Vec3 pose = {st.pose[0], st.pose[1], st.pose[2]};
        write_observation<COLOR>(pk, cfg, status, bin_cnt, grid_cnt, pose, lane, ws, next_obs, nullptr);
    }
    store_state(&ea.states[env], st, lane);
}

// ------------------------------------------------------------------------------ state access
template <int COLOR>
__global__ void get_state_kernel(DevPack pk, EnvArrays<COLOR> ea, const int32_t *env_ids, int n, int16_t *status_out,
                                 double *pose_out, double *quat_out, double *scalars_out) {
    const int k = blockIdx.y;
    const int env = env_ids ? env_ids[k] : k;
    const typename StatusT<COLOR>::type *status = ea.planes + (size_t)env * pk.n_pad;
    if (status_out) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pk.n_texels; j += gridDim.x * blockDim.x)
            status_out[(size_t)k * pk.n_texels + pk.sorted_to_pack[j]] = (int16_t)status[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const EnvState &st = ea.states[env];
        if (pose_out) for (int i = 0; i < 3; ++i) pose_out[3 * k + i] = st.pose[i];
        if (quat_out) for (int i = 0; i < 4; ++i) quat_out[4 * k + i] = st.quat[i];
        if (scalars_out) {
            double *o = scalars_out + 8 * (size_t)k;
            o[0] = st.total_reward; o[1] = st.total_return; o[2] = st.step_counter; o[3] = st.term_counter;
            o[4] = (st.flags & kFlagLastOnPart) ? 1.0 : 0.0; o[5] = (st.flags & kFlagTerminate) ? 1.0 : 0.0;
            o[6] = st.last_angle; o[7] = st.angle_diff;
        }
    }
}

template <int COLOR>
__global__ void set_state_kernel(DevPack pk, EnvArrays<COLOR> ea, const int32_t *env_ids, int n, const int16_t *status_in,
                                 const double *pose_in, const double *quat_in, const double *scalars_in) {
    typedef typename StatusT<COLOR>::type S;
    const int k = blockIdx.y;
    const int env = env_ids ? env_ids[k] : k;
    S *status = ea.planes + (size_t)env * pk.n_pad;
    if (status_in) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pk.n_texels; j += gridDim.x * blockDim.x)
            status[j] = (S)status_in[(size_t)k * pk.n_texels + pk.sorted_to_pack[j]];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        EnvState &st = ea.states[env];
        if (pose_in) for (int i = 0; i < 3; ++i) st.pose[i] = pose_in[3 * k + i];
        if (quat_in) for (int i = 0; i < 4; ++i) st.quat[i] = quat_in[4 * k + i];
        if (scalars_in) {
            const double *o = scalars_in + 8 * (size_t)k;
            st.total_reward = o[0]; st.total_return = o[1]; st.step_counter = (int)o[2]; st.term_counter = (int)o[3];
            int f = st.flags & kFlagHasLast;
            if (o[4] != 0.0) f |= kFlagLastOnPart;
            if (o[5] != 0.0) f |= kFlagTerminate;
            st.flags = f;
            st.last_angle = o[6]; st.angle_diff = o[7];
        }
        // the overlap reference set cannot be expressed through this interface: clear it, as
        // reset_part does (bullet_paint_wrapper.py:708)
        if (status_in) st.flags &= ~kFlagHasLast;
    }
}

// Rebuilds the flip counters of the listed environments from their status planes (after set_state).
template <int COLOR>
__global__ void recount_kernel(DevPack pk, EnvArrays<COLOR> ea, const int32_t *env_ids, int n) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n) return;
    const int env = env_ids ? env_ids[k] : k;
    const typename StatusT<COLOR>::type *status = ea.planes + (size_t)env * pk.n_pad;
    unsigned *bin_cnt = bin_cnt_of<COLOR>(pk, ea, env), *grid_cnt = grid_cnt_of<COLOR>(pk, ea, env);
    for (int w = lane; w < pk.n_bins_pad / 2; w += 32) bin_cnt[w] = 0;
    if (grid_cnt) for (int w = lane; w < pk.n_gcells_pad; w += 32) grid_cnt[w] = 0;
    __syncwarp();
    const bool init_painted = (pk.status_init == kPainted);
    for (int j = lane; j < pk.n_texels; j += 32) {
        const bool painted = ((int)status[j] == kPainted);
        if (painted != init_painted) {
            const unsigned w = __float_as_uint(__ldg(&pk.trel[j]).w);
            const unsigned bin = w & 0xffffu;
            atomicAdd(bin_cnt + (bin >> 1), 1u << ((bin & 1u) * 16));
            if (grid_cnt) atomicAdd(grid_cnt + (w >> 16), 1u);
        }
    }
}

// get_job_status (bullet_paint_wrapper.py:727-732): painted front texels per env, one warp each
template <int COLOR>
__global__ void job_status_kernel(DevPack pk, const typename StatusT<COLOR>::type *planes, int num_envs, int32_t *out) {
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (env >= num_envs) return;
    const typename StatusT<COLOR>::type *status = planes + (size_t)env * pk.n_pad;
    int c = 0;
    for (int j = lane; j < pk.n_texels; j += 32) c += ((int)status[j] == kPainted) ? 1 : 0;
    c = __reduce_add_sync(kFull, c);
    if (lane == 0) out[env] = c;
}

}  // namespace paintrl
