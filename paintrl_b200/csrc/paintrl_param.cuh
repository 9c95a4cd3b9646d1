// paintrl_param.cuh -- the reference's grid-world ParamTestEnv (PaintRLEnv/param_test_env.py:96-246), batched.
//
// The reference uses this N x N world to tune RL hyper-parameters before the paint task (param_test_*.py):
// the agent walks a grid, collects 1 from every interior cell it leaves or enters for the first time, pays
// 0.2 per step, and the episode ends at a wall, when nothing is left, at the length limit, or (optionally) on
// a repeated visit.  One thread per environment; all tables are laid out [cell][env] so the threads of a warp
// touch consecutive bytes.  Integer state; the observations' divisions are single IEEE FP64 operations on the
// reference's operands, so every output is bit-exact against the reference (tests/test_gpu_param_env.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace paintrl {

enum : int { kParamObsSection = 0, kParamObsSimple = 1, kParamObsDirect = 2, kParamObsGrid = 3 };

struct ParamWorld {
    int num_envs, size, episode_max_length, repeat_termination, obs_mode, obs_dim, auto_reset;
    int init_reward_counter;
    uint8_t *world;        // [size * size][num_envs]  world[(i, j)]  (param_test_env.py:122-130)
    uint16_t *visit;       // [size * size][num_envs]  visit_table[(i, j)] (saturates at 65535: beyond any EPISODE_MAX_LENGTH in use)
    int *pos_i, *pos_j;    // [num_envs]
    int *reward_counter, *step_counter;
    uint8_t *flags;        // [num_envs] bit 0 violated_wall, bit 1 repeat_visit
    unsigned long long *stats;   // [2] env-steps, episodes ended
};

__device__ __forceinline__ void param_reset_env(const ParamWorld &w, int e) {      // param_test_env.py:150-160
    const int s = w.size;
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
            const bool edge = i == 0 || i == s - 1 || j == 0 || j == s - 1;
            w.world[(size_t)(i * s + j) * w.num_envs + e] = edge ? 0 : 1;
            w.visit[(size_t)(i * s + j) * w.num_envs + e] = (i == 1 && j == 1) ? 1 : 0;
        }
    w.pos_i[e] = 1; w.pos_j[e] = 1;
    w.reward_counter[e] = w.init_reward_counter;
    w.step_counter[e] = 0;
    w.flags[e] = 0;
}

__device__ __forceinline__ void param_observation(const ParamWorld &w, int e, double *obs) {   // :199-204
    const int s = w.size, x = w.pos_i[e], y = w.pos_j[e];
    int k = 0;
    if (w.obs_mode == kParamObsSection) {                                          // :66-93
        int cnt[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
        // the reference's own conditions: a position on the far edge (x or y == size - 1) pulls that edge's
        // cells (value 0) into the first / second half's cell count
        for (int i = 1; i < s; ++i) {
            const int qi = i <= x ? 0 : (i < s - 1 ? 2 : -1);
            if (qi < 0) continue;
            for (int j = 1; j < s; ++j) {
                const int qj = j <= y ? 0 : (j < s - 1 ? 1 : -1);
                if (qj < 0) continue;
                cnt[qi + qj] += w.world[(size_t)(i * s + j) * w.num_envs + e];
                mx[qi + qj] += 1;
            }
        }
        for (int q = 0; q < 4; ++q) obs[k++] = mx[q] == 0 ? 0.0 : (double)cnt[q] / (double)mx[q];
    } else if (w.obs_mode == kParamObsDirect) {                                    // :24-30
        for (int c = 0; c < s * s; ++c) obs[k++] = (double)w.world[(size_t)c * w.num_envs + e];
    } else if (w.obs_mode == kParamObsGrid) {                                      // :50-63 (size 22: indices 0..9)
        const double max_counter = (double)(w.init_reward_counter / 100);
        for (int c = 0; c < 100; ++c) obs[c] = 0.0;
        for (int i = 1; i < s - 1; ++i)
            for (int j = 1; j < s - 1; ++j) {
                const int gx = (int)((double)i / 2 + 0.5) - 1, gy = (int)((double)j / 2 + 0.5) - 1;
                obs[gx * 10 + gy] += (double)w.world[(size_t)(i * s + j) * w.num_envs + e] / max_counter;
            }
        k = 100;
    }
    obs[k] = (double)x / (double)s;
    obs[k + 1] = (double)y / (double)s;
}

__device__ __forceinline__ int param_immediate(const ParamWorld &w, int e, int i, int j, int &reward_counter) {   // :206-211
    uint8_t *cell = &w.world[(size_t)(i * w.size + j) * w.num_envs + e];
    if (*cell > 0) { *cell -= 1; reward_counter -= 1; return 1; }
    return 0;
}

__global__ void param_reset_kernel(ParamWorld w, const int32_t *env_ids, int n, double *obs_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int e = env_ids ? env_ids[k] : k;
    if ((unsigned)e >= (unsigned)w.num_envs) return;
    param_reset_env(w, e);
    if (obs_out) param_observation(w, e, obs_out + (size_t)k * w.obs_dim);
}

// ParamTestEnv.step (param_test_env.py:218-240) for every environment.  An action outside 0..3 (the reference raises
// IndexError, :173-174) raises *bad_action and leaves that world untouched, but its output row is still written
// -- current observation, zero reward / penalty, done = 1 -- so that a caller who skips the check never trains
// on the previous step's stale transition.  The host mirror checks the actions (or the flag) and raises.
__global__ void param_step_kernel(ParamWorld w, const long long *actions, double *obs, double *reward_out, double *penalty_out,
                                  double *actual_out, uint8_t *done_out, double *next_obs, int *bad_action) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.num_envs) return;
    const long long a = actions[e];
    if (a < 0 || a > 3) {
        atomicExch(bad_action, 1);
        param_observation(w, e, obs + (size_t)e * w.obs_dim);
        if (next_obs) param_observation(w, e, next_obs + (size_t)e * w.obs_dim);
        reward_out[e] = 0.0; penalty_out[e] = 0.0; actual_out[e] = 0.0; done_out[e] = 1;
        return;
    }
    const int s = w.size;
    int i = w.pos_i[e], j = w.pos_j[e], rc = w.reward_counter[e];
    int flags = w.flags[e];
    const int immediate = param_immediate(w, e, i, j, rc);                         // :163
    const int step_counter = w.step_counter[e] + 1;
    if (a == 0) i += 1; else if (a == 1) j += 1; else if (a == 2) i -= 1; else j -= 1;
    if (i < 0 || i >= s || j < 0 || j >= s) {                                      // :175-179
        i = min(max(i, 0), s - 1);
        j = min(max(j, 0), s - 1);
        flags |= 1;
    } else {
        uint16_t *v = &w.visit[(size_t)(i * s + j) * w.num_envs + e];
        if (*v < 65535) *v += 1;
        if (*v > 1) flags |= 2;
    }
    int reward = (flags & 1) ? 0 : param_immediate(w, e, i, j, rc);                // :213-216
    reward += immediate;
    const double penalty = 0.2;
    const bool done = (flags & 1) || rc <= 0 || step_counter >= w.episode_max_length - 1 ||
                      ((flags & 2) && w.repeat_termination);                       // :192-197
    w.pos_i[e] = i; w.pos_j[e] = j; w.reward_counter[e] = rc; w.step_counter[e] = step_counter; w.flags[e] = (uint8_t)flags;
    param_observation(w, e, obs + (size_t)e * w.obs_dim);
    reward_out[e] = (double)reward;
    penalty_out[e] = penalty;
    actual_out[e] = (double)reward - penalty;
    done_out[e] = done ? 1 : 0;
    atomicAdd(&w.stats[0], 1ull);
    if (done) atomicAdd(&w.stats[1], 1ull);
    if (next_obs) {
        if (done && w.auto_reset) param_reset_env(w, e);
        param_observation(w, e, next_obs + (size_t)e * w.obs_dim);
    }
}

// world / visit tables of the listed environments as int32 [n][size * size] (the reference's Visualizer input)
__global__ void param_tables_kernel(ParamWorld w, const int32_t *env_ids, int n, int32_t *world_out, int32_t *visit_out) {
    const int k = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || c >= w.size * w.size) return;
    const int e = env_ids ? env_ids[k] : k;
    if ((unsigned)e >= (unsigned)w.num_envs) return;
    if (world_out) world_out[(size_t)k * w.size * w.size + c] = w.world[(size_t)c * w.num_envs + e];
    if (visit_out) visit_out[(size_t)k * w.size * w.size + c] = w.visit[(size_t)c * w.num_envs + e];
}

}  // namespace paintrl
