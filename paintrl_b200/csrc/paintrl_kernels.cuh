// paintrl_kernels.cuh -- the fused step kernel and its helpers (one warp per environment).
//
// Per environment and step:
//   move    5 sub-steps of {ray vs hull, nearest vertex, closest triangle}    robot.py:302-329
//   stamp   5 ball queries over the binned texel table, colour update, overlap bookkeeping
//                                                     bullet_paint_wrapper.py:568-577, 352-434
//   score   reward / penalty / termination                                   robot_gym_env.py:289-340
//   observe normalised pose + section/grid histogram over the env's whole status plane
//                                                     bullet_paint_wrapper.py:965-978, 1045-1139
// The status plane (1 byte per front texel in RGB mode, int16 in HSI mode) is the only per-env
// array; it is read once in full per step (128-bit loads) and written only inside the footprint.
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

__device__ __forceinline__ uint4 ldcg16(const void *p) { return __ldcg(reinterpret_cast<const uint4 *>(p)); }

// ------------------------------------------------------------------------------ observation
// Section observation with 4 sectors (bullet_paint_wrapper.py:1033-1061).
//
// rx > 0 <=> rank >= hi, rx < 0 <=> rank < lo, where `rank` is the texel coordinate's index among
// the sorted distinct coordinates of the part and [lo, hi) is the rank interval equal to the
// pose's coordinate (found once per step by binary search): integer compares replace the FP64
// subtraction and are exactly equivalent.  Texels are laid out in chunks of 16 (one 128-bit load
// in RGB mode) with a precomputed rank bounding box per chunk; a chunk whose box lies strictly on
// one side of the pose on both axes belongs wholly to one sector and is counted with byte-SIMD
// popcounts, the others (the pose's row and column, ~15 %) are queued in shared memory and then
// classified texel by texel, one queued chunk per lane.
template <typename RankT>
__device__ __forceinline__ void rank_bounds(const double *uniq, int n, double v, int lane, RankT &lo, RankT &hi) {
    // 32-ary search: every round each lane probes one pivot; ~3 rounds for 10^4 values
    int a = 0, b = n;            // lower_bound: first index with uniq[i] >= v lies in [a, b]
    while (b - a > 0) {
        int span = b - a, step = (span + 31) >> 5;
        int i = a + lane * step;
        bool less = (i < b) && (__ldg(&uniq[i]) < v);
        unsigned m = __ballot_sync(kFull, less);
        int k = __popc(m);       // pivots 0..k-1 are < v (monotone)
        int na = (k == 0) ? a : a + (k - 1) * step + 1;
        int nb = (k == 32 || a + k * step >= b) ? b : a + k * step;
        a = na; b = nb;
        if (step == 1) break;
    }
    int l = a;
    int c = l, d = n;            // upper_bound from l
    while (d - c > 0) {
        int span = d - c, step = (span + 31) >> 5;
        int i = c + lane * step;
        bool le = (i < d) && (__ldg(&uniq[i]) <= v);
        unsigned m = __ballot_sync(kFull, le);
        int k = __popc(m);
        int nc = (k == 0) ? c : c + (k - 1) * step + 1;
        int nd = (k == 32 || c + k * step >= d) ? d : c + k * step;
        c = nc; d = nd;
        if (step == 1) break;
    }
    lo = (RankT)l;
    hi = (RankT)c;
}

// number of painted (== 255) entries among the 16 status values of one chunk
__device__ __forceinline__ int painted_in_word_u8(unsigned w) {
    unsigned x = ~w;                                   // zero byte <=> painted
    unsigned y = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    y = ~(y | x | 0x7F7F7F7Fu);                        // 0x80 in every zero byte, exact
    return __popc(y);
}
__device__ __forceinline__ int painted_in_word_i16(unsigned w) {
    unsigned x = w ^ 0x00FF00FFu;                      // zero halfword <=> value == 255
    unsigned y = (x & 0x7FFF7FFFu) + 0x7FFF7FFFu;
    y = ~(y | x | 0x7FFF7FFFu);                        // 0x8000 in every zero halfword, exact
    return __popc(y);
}
template <int COLOR>
__device__ __forceinline__ int painted_in_chunk(const typename StatusT<COLOR>::type *p) {
    if (COLOR == 0) {
        uint4 v = ldcg16(p);
        return painted_in_word_u8(v.x) + painted_in_word_u8(v.y) + painted_in_word_u8(v.z) + painted_in_word_u8(v.w);
    } else {
        uint4 v = ldcg16(p), w = ldcg16(p + 8);
        return painted_in_word_i16(v.x) + painted_in_word_i16(v.y) + painted_in_word_i16(v.z) + painted_in_word_i16(v.w) +
               painted_in_word_i16(w.x) + painted_in_word_i16(w.y) + painted_in_word_i16(w.z) + painted_in_word_i16(w.w);
    }
}

template <typename RankT> struct ChunkBox { RankT r0min, r0max, r1min, r1max; };

constexpr int kMixedQueue = 128;   // per-warp queue of mixed chunks (ints, aliases the histogram area)

// texel-by-texel classification of one chunk (lane-private); adds into 4x16-bit packed counters
template <int COLOR, typename RankT>
__device__ __forceinline__ void classify_chunk(const DevPack &pk, const typename StatusT<COLOR>::type *status, int chunk,
                                               RankT lo0, RankT hi0, RankT lo1, RankT hi1,
                                               unsigned long long &ptot, unsigned long long &popen) {
    typedef typename StatusT<COLOR>::type S;
    const RankT *r0 = reinterpret_cast<const RankT *>(pk.rank0) + (size_t)chunk * 16;
    const RankT *r1 = reinterpret_cast<const RankT *>(pk.rank1) + (size_t)chunk * 16;
    const RankT pad = (RankT)~(RankT)0;
    RankT a0[16], a1[16];
    S sv[16];
    constexpr int kRankLoads = 16 * sizeof(RankT) / 16;
#pragma unroll
    for (int q = 0; q < kRankLoads; ++q) {
        reinterpret_cast<uint4 *>(a0)[q] = __ldg(reinterpret_cast<const uint4 *>(r0) + q);
        reinterpret_cast<uint4 *>(a1)[q] = __ldg(reinterpret_cast<const uint4 *>(r1) + q);
    }
#pragma unroll
    for (int q = 0; q < (int)(16 * sizeof(S) / 16); ++q)
        reinterpret_cast<uint4 *>(sv)[q] = ldcg16(status + (size_t)chunk * 16 + q * (16 / sizeof(S)));
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        RankT x = a0[e], y = a1[e];
        bool px = x >= hi0, nx = x < lo0, py = y >= hi1, ny = y < lo1;
        bool skip = (x == pad) || !(px || nx || py || ny);
        int q = (px && py) ? 0 : ((nx && py) ? 1 : ((nx && ny) ? 2 : 3));
        unsigned long long one = skip ? 0ull : (1ull << (16 * q));
        ptot += one;
        popen += ((int)sv[e] != kPainted) ? one : 0ull;
    }
}

template <int COLOR, typename RankT>
__device__ __forceinline__ void section4_counts(const DevPack &pk, const typename StatusT<COLOR>::type *status,
                                                const Vec3 &pose, int lane, int *queue /*smem, kMixedQueue ints*/,
                                                int tot[4], int open[4]) {
    RankT lo0, hi0, lo1, hi1;
    rank_bounds<RankT>(pk.uniq0, pk.n_uniq0, comp(pose, pk.axis0), lane, lo0, hi0);
    rank_bounds<RankT>(pk.uniq1, pk.n_uniq1, comp(pose, pk.axis1), lane, lo1, hi1);
    const ChunkBox<RankT> *boxes = reinterpret_cast<const ChunkBox<RankT> *>(pk.chunk_box);
    const int n_chunks = pk.n_pad >> 4;
    // per-lane accumulators: pure chunks counted as (#chunks, open texels) per sector
    int pure_chunks[4] = {0, 0, 0, 0}, pure_open[4] = {0, 0, 0, 0};
    unsigned long long ptot = 0, popen = 0;   // texel-wise path, 4 x 16-bit fields
    int mix_tot[4] = {0, 0, 0, 0}, mix_open[4] = {0, 0, 0, 0};
    int queued = 0;                            // warp-uniform
    auto drain = [&]() {
        __syncwarp();
        for (int base = 0; base < queued; base += 32) {
            if (base + lane < queued)
                classify_chunk<COLOR, RankT>(pk, status, queue[base + lane], lo0, hi0, lo1, hi1, ptot, popen);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {          // fields hold at most 16 * (kMixedQueue/32) = 64 each
            mix_tot[q] += (int)((ptot >> (16 * q)) & 0xffff);
            mix_open[q] += (int)((popen >> (16 * q)) & 0xffff);
        }
        ptot = popen = 0;
        queued = 0;
        __syncwarp();
    };
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
        const int c = c0 + lane;
        bool mixed = false;
        if (c < n_chunks) {
            ChunkBox<RankT> bx = boxes[c];
            bool nx = bx.r0max < lo0, px = bx.r0min >= hi0, ny = bx.r1max < lo1, py = bx.r1min >= hi1;
            if ((nx || px) && (ny || py) && bx.r0min <= bx.r0max) {   // inverted box = chunk with padding
                int painted = painted_in_chunk<COLOR>(status + (size_t)c * 16);
                int q = py ? (px ? 0 : 1) : (px ? 3 : 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pure_chunks[k] += (q == k) ? 1 : 0;
                    pure_open[k] += (q == k) ? 16 - painted : 0;
                }
            } else {
                mixed = true;
            }
        }
        unsigned mm = __ballot_sync(kFull, mixed);
        if (mm) {
            if (mixed) queue[queued + __popc(mm & ((1u << lane) - 1))] = c;
            queued += __popc(mm);
            if (queued > kMixedQueue - 32) drain();
        }
    }
    if (queued) drain();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        tot[q] = __reduce_add_sync(kFull, pure_chunks[q] * 16 + mix_tot[q]);
        open[q] = __reduce_add_sync(kFull, pure_open[q] + mix_open[q]);
    }
}

// Section observation with K != 4 sectors (atan2 path, bullet_paint_wrapper.py:1026-1031) and the
// grid observation (bullet_paint_wrapper.py:1126-1139) histogram into per-warp shared memory.
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

template <int COLOR>
__device__ __forceinline__ void grid_counts(const DevPack &pk, const typename StatusT<COLOR>::type *status,
                                            int cells, int lane, int *hist /*smem*/) {
    for (int i = lane; i < cells; i += 32) hist[i] = 0;
    __syncwarp();
    for (int j = lane; j < pk.n_texels; j += 32) {
        if ((int)__ldcg(&status[j]) == kPainted) atomicAdd(&hist[__ldg(&pk.gcell[j])], 1);
    }
    __syncwarp();
}

// robot_gym_env.py:306-319 _augmented_observation; every lane returns, lanes < obs_dim write.
template <int COLOR, typename RankT>
__device__ __forceinline__ void write_observation(const DevPack &pk, const DevConfig &cfg,
                                                  const typename StatusT<COLOR>::type *status, const Vec3 &pose,
                                                  int lane, int *hist, double *obs_a, double *obs_b) {
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
    if (cfg.obs_mode == 1) {   // grid
        const int cells = grad * grad;
        grid_counts<COLOR>(pk, status, cells, lane, hist);
        for (int c = lane; c < cells; c += 32) {
            int total = __ldg(&pk.gtotal[c]);
            double v = total == 0 ? 0.0 : 1.0 - (double)hist[c] / (double)total;
            if (obs_a) obs_a[c] = v;
            if (obs_b) obs_b[c] = v;
        }
        __syncwarp();
        return;
    }
    // section / discrete
    if (grad == 4) {
        int tot[4], open[4];
        section4_counts<COLOR, RankT>(pk, status, pose, lane, hist, tot, open);
        if (lane < 4) {
            int t = lane == 0 ? tot[0] : (lane == 1 ? tot[1] : (lane == 2 ? tot[2] : tot[3]));
            int o = lane == 0 ? open[0] : (lane == 1 ? open[1] : (lane == 2 ? open[2] : open[3]));
            double v = t == 0 ? 0.0 : (double)o / (double)t;
            if (obs_a) obs_a[lane] = v;
            if (obs_b) obs_b[lane] = v;
        }
    } else {
        sectionk_counts<COLOR>(pk, status, pose, grad, lane, hist);
        for (int s = lane; s < grad; s += 32) {
            int t = hist[s], o = hist[kMaxObs + s];
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
// Part.reset_part (bullet_paint_wrapper.py:706-708): restore the init colour, 128-bit stores.
template <int COLOR>
__device__ __forceinline__ void fill_status(const DevPack &pk, typename StatusT<COLOR>::type *status, int lane) {
    typedef typename StatusT<COLOR>::type S;
    constexpr int kPer = 16 / sizeof(S);
    uint4 v;
    S *e = reinterpret_cast<S *>(&v);
#pragma unroll
    for (int k = 0; k < kPer; ++k) e[k] = (S)pk.status_init;
    for (int j0 = lane * kPer; j0 < pk.n_pad; j0 += 32 * kPer) *reinterpret_cast<uint4 *>(status + j0) = v;
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
                                          int start_index, int lane) {
    fill_status<COLOR>(pk, status, lane);
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

// ------------------------------------------------------------------------------ kernels
template <int COLOR, typename RankT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
reset_kernel(DevPack pk, DevConfig cfg, EnvState *states, typename StatusT<COLOR>::type *planes,
             const int32_t *env_ids, int n, const int32_t *start_idx, const double *set_pos,
             const double *set_normal, double *obs_out, int mode /*0 reset, 1 set_pose*/) {
    __shared__ int hist_all[kWarpsPerBlock][2 * kMaxObs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = blockIdx.x * kWarpsPerBlock + warp;
    if (k >= n) return;
    const int env = env_ids ? env_ids[k] : k;
    typename StatusT<COLOR>::type *status = planes + (size_t)env * pk.n_pad;
    EnvState st;
    load_state(&states[env], st);
    if (mode == 0) {
        int idx;
        if (start_idx) idx = start_idx[k];
        else idx = (int)(splitmix64(cfg.seed ^ splitmix64(((unsigned long long)env << 32) | (unsigned)st.episode)) %
                         (unsigned long long)pk.n_starts);
        idx = min(max(idx, 0), pk.n_starts - 1);
        env_reset<COLOR>(pk, st, status, idx, lane);
    } else {
        robot_reset(st, set_pos + 3 * k, set_normal + 3 * k);
    }
    __syncwarp();
    Vec3 pose = {st.pose[0], st.pose[1], st.pose[2]};
    write_observation<COLOR, RankT>(pk, cfg, status, pose, lane, hist_all[warp],
                                    obs_out ? obs_out + (size_t)k * cfg.obs_dim : nullptr, nullptr);
    store_state(&states[env], st, lane);
}

template <int COLOR, typename RankT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
step_kernel(DevPack pk, DevConfig cfg, EnvState *states, typename StatusT<COLOR>::type *planes, int num_envs,
            StepIO io) {
    typedef typename StatusT<COLOR>::type S;
    __shared__ int hist_all[kWarpsPerBlock][2 * kMaxObs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env = blockIdx.x * kWarpsPerBlock + warp;
    if (env >= num_envs) return;
    S *status = planes + (size_t)env * pk.n_pad;
    EnvState st;
    load_state(&states[env], st);

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
    Vec3 centers[kPaintPerAction];
    int full_scans = 0;
    double quat[4] = {st.quat[0], st.quat[1], st.quat[2], st.quat[3]};
#pragma unroll 1
    for (int s = 0; s < kPaintPerAction; ++s) {
        Vec3 p = cur_p;
        add_comp(p, pk.axis0, delta1);
        add_comp(p, pk.axis1, delta2_scaled);
        Vec3 end = {p.x + cur_n.x, p.y + cur_n.y, p.z + cur_n.z};
        Vec3 hit, pos, orn = cur_n;
        bool ok = ray_test(pk, p, end, lane, hit, full_scans);
        if (ok) ok = hook_point(pk, hit, lane, pos, orn);
        if (!ok) orn = cur_n;
        quat_from_normal(orn, quat);
        if (!ok) {
            pos = transform_point(cur_p, quat, delta2, delta1, 0.0);   // robot.py:317 (sic)
            if (st.flags & kFlagLastOnPart) {                           // robot.py:292-300
                st.flags &= ~kFlagLastOnPart;
            } else {
                st.term_counter += 1;
                if (st.term_counter > kNotOnPartTerminateSteps) st.flags |= kFlagTerminate;
            }
        } else {
            st.flags |= kFlagLastOnPart;
        }
        centers[s] = transform_point(pos, quat, 0.0, 0.0, 0.1);        // robot.py:277-278
        cur_p = pos;
        cur_n = orn;
    }
    st.pose[0] = cur_p.x; st.pose[1] = cur_p.y; st.pose[2] = cur_p.z;
    st.quat[0] = quat[0]; st.quat[1] = quat[1]; st.quat[2] = quat[2]; st.quat[3] = quat[3];

    // ---- stamp the 5 shots, texel-major over the binned candidates (bullet_paint_wrapper.py:568-577)
    const double r2 = kPaintRadius * kPaintRadius;
    const bool has_last = (st.flags & kFlagHasLast) != 0;
    const Vec3 lastc = {st.last_center[0], st.last_center[1], st.last_center[2]};
    double lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
#pragma unroll
    for (int s = 0; s < kPaintPerAction; ++s) {
        double c0 = comp(centers[s], pk.axis0), c1 = comp(centers[s], pk.axis1);
        lo0 = fmin(lo0, c0); hi0 = fmax(hi0, c0);
        lo1 = fmin(lo1, c1); hi1 = fmax(hi1, c1);
    }
    const double margin = kPaintRadius + 1e-9;
    int bx0 = (int)floor((lo0 - margin - pk.tb_o0) * pk.tb_inv), bx1 = (int)floor((hi0 + margin - pk.tb_o0) * pk.tb_inv);
    int by0 = (int)floor((lo1 - margin - pk.tb_o1) * pk.tb_inv), by1 = (int)floor((hi1 + margin - pk.tb_o1) * pk.tb_inv);
    bx0 = max(bx0, 0); by0 = max(by0, 0);
    bx1 = min(bx1, pk.tb_nx - 1); by1 = min(by1, pk.tb_ny - 1);

    double rmax[kPaintPerAction];
    if (COLOR == 1) {   // HSI: r = distances.max() per shot (bullet_paint_wrapper.py:423-424)
#pragma unroll
        for (int s = 0; s < kPaintPerAction; ++s) rmax[s] = -1.0;
        for (int row = by0; row <= by1; ++row) {
            if (bx0 > bx1) break;
            int begin = __ldg(&pk.tb_start[row * pk.tb_nx + bx0]), end = __ldg(&pk.tb_start[row * pk.tb_nx + bx1 + 1]);
            for (int j = begin + lane; j < end; j += 32) {
                double x = __ldg(&pk.tx[j]), y = __ldg(&pk.ty[j]), z = __ldg(&pk.tz[j]);
#pragma unroll
                for (int s = 0; s < kPaintPerAction; ++s) {
                    double dx = x - centers[s].x, dy = y - centers[s].y, dz = z - centers[s].z;
                    double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 <= r2) rmax[s] = fmax(rmax[s], d2);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < kPaintPerAction; ++s) rmax[s] = sqrt(warp_max(rmax[s]));   // sqrt is monotone
    }

    int n_new = 0, n_possible = 0;   // per-lane partial counts
    for (int row = by0; row <= by1; ++row) {
        if (bx0 > bx1) break;
        int begin = __ldg(&pk.tb_start[row * pk.tb_nx + bx0]), end = __ldg(&pk.tb_start[row * pk.tb_nx + bx1 + 1]);
        for (int j = begin + lane; j < end; j += 32) {
            double x = __ldg(&pk.tx[j]), y = __ldg(&pk.ty[j]), z = __ldg(&pk.tz[j]);
            bool prev_in = false;
            if (has_last) {
                double dx = x - lastc.x, dy = y - lastc.y, dz = z - lastc.z;
                prev_in = (dx * dx + dy * dy + dz * dz) <= r2;
            }
            bool any = false, valid = false;
            int sv = 0;
            bool loaded = false, dirty = false;
#pragma unroll
            for (int s = 0; s < kPaintPerAction; ++s) {
                double dx = x - centers[s].x, dy = y - centers[s].y, dz = z - centers[s].z;
                double d2 = dx * dx + dy * dy + dz * dz;
                bool in = d2 <= r2;
                if (in) {
                    any = true;
                    if (!prev_in) valid = true;          // affected \ last_affected (:575)
                    if (COLOR == 1) {
                        if (!loaded) { sv = (int)__ldcg(&status[j]); loaded = true; }
                        double ratio = sqrt(d2) / rmax[s];
                        int quantity = (int)(kHsiTargetMax * (1.0 - ratio * ratio)) + 1;   // :429
                        if (sv > 0) { sv -= quantity; n_new += quantity; dirty = true; }  // :411-418
                    }
                }
                prev_in = in;
            }
            if (COLOR == 0 && any) {                      // :358-365
                if ((int)__ldcg(&status[j]) != kPainted) { status[j] = (S)kPainted; n_new += 1; }
            }
            if (COLOR == 1 && dirty) status[j] = (S)sv;
            n_possible += valid ? 1 : 0;
        }
    }
    n_new = __reduce_add_sync(kFull, n_new);
    n_possible = __reduce_add_sync(kFull, n_possible);
    st.last_center[0] = centers[kPaintPerAction - 1].x;
    st.last_center[1] = centers[kPaintPerAction - 1].y;
    st.last_center[2] = centers[kPaintPerAction - 1].z;
    st.flags |= kFlagHasLast;
    __syncwarp();

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
    const double expected_avg = max_pts / (double)(cfg.expected_episode_length * 100);
    bool done, decided = false;
    if (avg_reward < expected_avg && cfg.termination_mode != 0) {
        if (cfg.termination_mode == 1) { done = true; decided = true; }
        else if (st.total_reward < cfg.switch_threshold * max_pts / 100) { done = true; decided = true; }
    }
    if (!decided) done = finished || (st.flags & kFlagTerminate) || st.step_counter > cfg.episode_max_length - 1;

    // ---- observation (computed even when done, robot_gym_env.py:358)
    const bool resetting = done && cfg.auto_reset;
    double *obs = io.obs + (size_t)env * cfg.obs_dim;
    double *next_obs = io.next_obs ? io.next_obs + (size_t)env * cfg.obs_dim : nullptr;
    write_observation<COLOR, RankT>(pk, cfg, status, cur_p, lane, hist_all[warp], obs, resetting ? nullptr : next_obs);
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
        int idx;
        if (io.reset_start_idx) idx = io.reset_start_idx[env];
        else idx = (int)(splitmix64(cfg.seed ^ splitmix64(((unsigned long long)env << 32) | (unsigned)st.episode)) %
                         (unsigned long long)pk.n_starts);
        idx = min(max(idx, 0), pk.n_starts - 1);
        __syncwarp();
        env_reset<COLOR>(pk, st, status, idx, lane);
        __syncwarp();
        Vec3 pose = {st.pose[0], st.pose[1], st.pose[2]};
        write_observation<COLOR, RankT>(pk, cfg, status, pose, lane, hist_all[warp], next_obs, nullptr);
    }
    store_state(&states[env], st, lane);
}

// ------------------------------------------------------------------------------ state access
template <int COLOR>
__global__ void get_state_kernel(DevPack pk, const EnvState *states, const typename StatusT<COLOR>::type *planes,
                                 const int32_t *env_ids, int n, int16_t *status_out, double *pose_out,
                                 double *quat_out, double *scalars_out) {
    const int k = blockIdx.y;
    const int env = env_ids ? env_ids[k] : k;
    const typename StatusT<COLOR>::type *status = planes + (size_t)env * pk.n_pad;
    if (status_out) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pk.n_texels; j += gridDim.x * blockDim.x)
            status_out[(size_t)k * pk.n_texels + pk.sorted_to_pack[j]] = (int16_t)status[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const EnvState &st = states[env];
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
__global__ void set_state_kernel(DevPack pk, EnvState *states, typename StatusT<COLOR>::type *planes,
                                 const int32_t *env_ids, int n, const int16_t *status_in, const double *pose_in,
                                 const double *quat_in, const double *scalars_in) {
    typedef typename StatusT<COLOR>::type S;
    const int k = blockIdx.y;
    const int env = env_ids ? env_ids[k] : k;
    S *status = planes + (size_t)env * pk.n_pad;
    if (status_in) {
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pk.n_texels; j += gridDim.x * blockDim.x)
            status[j] = (S)status_in[(size_t)k * pk.n_texels + pk.sorted_to_pack[j]];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        EnvState &st = states[env];
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
