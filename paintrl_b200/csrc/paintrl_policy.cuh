// paintrl_policy.cuh -- the rollout policy of the reference's PPO script as ONE kernel per step.
//
// paint_ppo.py:179-183 trains a fully connected policy obs -> 256 -> 128 -> {logits | mean, value} (tanh, value branch
// sharing the hidden layers); its rollout workers evaluate it once per environment step and sample an action.  Here
// that is one launch for the whole batch, observations in, actions straight into the rollout fragment out:
//
//   CTA = 128 environments (one per TMEM lane), 512 threads: four threads share a row -- each computes a quarter of the
//         row's layer-1 features and reduces a quarter of its layer-2 columns (warps w, w + 4, w + 8, w + 12 own TMEM
//         lanes 32 (w % 4) ...), so the SM runs sixteen warps instead of four
//   layer 1  (obs_dim x 256, obs_dim <= 32): FP32 FFMA per thread, weights broadcast from shared memory, tanh, the
//            activations written as BF16 into shared memory in the tensor core's canonical K-major layout
//   layer 2  (256 x 128): sixteen tcgen05.mma (M 128, N 128, K 16, BF16 x BF16 -> FP32) issued by one thread, A = the
//            activations, B = the weights (pre-packed by the host in the same canonical layout, brought in with one
//            cp.async.bulk), accumulator in TMEM (128 lanes x 128 columns)
//   epilogue tcgen05.ld of each thread's row, + bias, tanh, the 128 x (n_out + 1) head in FP32 FFMA
//   sampling discrete: Gumbel-max over the logits; continuous: tanh(mean) + unit Gaussian noise; log-probability and
//            value; counter-based random numbers (per-environment counters in device memory, so a captured CUDA graph
//            draws fresh noise on every replay)
//
// Shared-memory operand layout (no swizzle, K-major; cute::UMMA "INTERLEAVE"): 8 x 8 BF16 core matrices of 128
// contiguous bytes (row r at +16 r); core (mi, kj) of an [rows x 256] operand sits at kj * (rows / 8) * 128 + mi * 128,
// i.e. stride-byte-offset (next 8 rows) = 128 B, leading-byte-offset (next 8 columns of K) = rows * 16 B.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace paintrl {

constexpr int kPolH1 = 256, kPolH2 = 128;      // hidden sizes (paint_ppo.py:180 fcnet_hiddens)
constexpr int kPolRows = 128;                  // environments per CTA = MMA M
constexpr int kPolMaxObs = 32, kPolMaxOut = 16;
constexpr int kPolSplit = 4;                   // threads per environment row
constexpr int kPolThreads = kPolRows * kPolSplit;

struct PolicyParams {
    int obs_dim, n_out, discrete;
    const float *w1, *b1;                      // [obs_dim][256], [256]
    const __nv_bfloat16 *w2_packed;            // [128 x 256] in the canonical layout above (64 KB)
    const float *b2;                           // [128]
    const float *w3, *b3;                      // [128][n_out + 1] (last column: value), [n_out + 1]
    unsigned long long seed;
    unsigned *counters;                        // [capacity] per-environment draw counters
};

struct PolicyIO {
    const double *obs;        // [B][obs_dim]
    int batch;
    long long *act_discrete;  // [B]               (discrete)
    double *act_continuous;   // [B][n_out]        (continuous)
    float *logp, *value;      // [B]
    float *logits;            // [B][n_out + 1] or nullptr (tests)
    int sample;               // 0: value / logits only (bootstrap), 1: sample actions
};

__device__ __forceinline__ unsigned long long pol_mix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// uniform in (0, 1): 24 random bits, never 0 or 1
__device__ __forceinline__ float pol_uniform(unsigned long long seed, unsigned env, unsigned ctr, unsigned k) {
    const unsigned long long h = pol_mix(seed ^ pol_mix(((unsigned long long)env << 32) | ctr) ^ ((unsigned long long)k * 0xD6E8FEB86659FD93ull));
    return ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ unsigned long long pol_smem_desc(unsigned smem_byte_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all >> 4),
    // version [46,48) = 1 on sm_100, layout type [61,64) = 0 (no swizzle)
    return (unsigned long long)((smem_byte_addr & 0x3FFFFu) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ float pol_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// dynamic shared memory: A (64 KB) | B (64 KB) | w1 | b1 | b2 | w3 | b3 | barriers
// MO / MX: compile-time caps of the observation and output loops (register arrays, fully unrolled): 8 / 8 for the
// reference's spaces (6-18 observations are 8 or 32; 4 logits + value), 32 / 16 in general.  The shared-memory layout
// does not depend on them.
template <int MO, int MX>
__global__ void __launch_bounds__(kPolThreads, 1) policy_act_kernel(PolicyParams pp, PolicyIO io) {
    extern __shared__ __align__(1024) unsigned char pol_smem[];
    __nv_bfloat16 *sA = reinterpret_cast<__nv_bfloat16 *>(pol_smem);
    unsigned char *sB = pol_smem + kPolRows * kPolH1 * 2;
    float *sw1 = reinterpret_cast<float *>(sB + kPolH2 * kPolH1 * 2);
    float *sb1 = sw1 + kPolMaxObs * kPolH1;
    float *sb2 = sb1 + kPolH1;
    float *sw3 = sb2 + kPolH2;
    float *sb3 = sw3 + kPolH2 * kPolMaxOut;
    unsigned long long *bar_w = reinterpret_cast<unsigned long long *>(sb3 + kPolMaxOut);   // W2 landed
    unsigned long long *bar_mma = bar_w + 1;                                                 // accumulator ready
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bar_mma + 1);
    float *spart = reinterpret_cast<float *>(tmem_slot + 2);          // [kPolSplit - 1][kPolRows][kPolMaxOut] partial heads

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kPolRows - 1), part = tid / kPolRows;      // warp % 4 == (row / 32): the TMEM lanes this warp may read
    const int env = blockIdx.x * kPolRows + row;
    const int nout1 = pp.n_out + 1;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar_w)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar_mma)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // B operand: 64 KB, already in the canonical layout
        const unsigned bytes = kPolH2 * kPolH1 * 2;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(sB)),
                     "l"(pp.w2_packed), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar_w))
                     : "memory");
    }
    if (warp == 0) {   // TMEM: 128 columns for the 128 x 128 FP32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((unsigned)__cvta_generic_to_shared(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // small FP32 tables
    for (int i = tid; i < pp.obs_dim * kPolH1; i += kPolThreads) sw1[i] = __ldg(&pp.w1[i]);
    for (int i = tid; i < kPolH1; i += kPolThreads) sb1[i] = __ldg(&pp.b1[i]);
    for (int i = tid; i < kPolH2; i += kPolThreads) sb2[i] = __ldg(&pp.b2[i]);
    for (int i = tid; i < kPolH2 * nout1; i += kPolThreads) sw3[i] = __ldg(&pp.w3[i]);
    if (tid < nout1) sb3[tid] = __ldg(&pp.b3[tid]);
    float x[MO];
#pragma unroll
    for (int i = 0; i < MO; ++i) x[i] = (i < pp.obs_dim && env < io.batch) ? (float)io.obs[(size_t)env * pp.obs_dim + i] : 0.f;
    __syncthreads();

    // ---- layer 1: h1 = tanh(x W1 + b1) -> BF16, canonical layout; this thread's quarter of the row's features
    {
        unsigned char *rowbase = reinterpret_cast<unsigned char *>(sA) + (row >> 3) * 128 + (row & 7) * 16;
        for (int kj = part * (kPolH1 / 8 / kPolSplit); kj < (part + 1) * (kPolH1 / 8 / kPolSplit); ++kj) {
            float h[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) h[c] = sb1[kj * 8 + c];
#pragma unroll
            for (int i = 0; i < MO; ++i) {
                if (i < pp.obs_dim) {
                    const float xi = x[i];
                    const float *wr = sw1 + i * kPolH1 + kj * 8;
#pragma unroll
                    for (int c = 0; c < 8; ++c) h[c] = fmaf(xi, wr[c], h[c]);
                }
            }
            __nv_bfloat162 p0 = __floats2bfloat162_rn(pol_tanh(h[0]), pol_tanh(h[1]));
            __nv_bfloat162 p1 = __floats2bfloat162_rn(pol_tanh(h[2]), pol_tanh(h[3]));
            __nv_bfloat162 p2 = __floats2bfloat162_rn(pol_tanh(h[4]), pol_tanh(h[5]));
            __nv_bfloat162 p3 = __floats2bfloat162_rn(pol_tanh(h[6]), pol_tanh(h[7]));
            uint4 v;
            v.x = *reinterpret_cast<unsigned *>(&p0); v.y = *reinterpret_cast<unsigned *>(&p1);
            v.z = *reinterpret_cast<unsigned *>(&p2); v.w = *reinterpret_cast<unsigned *>(&p3);
            *reinterpret_cast<uint4 *>(rowbase + kj * (kPolRows / 8) * 128) = v;
        }
    }
    // the activations were written through the generic proxy; the tensor core reads through the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    // ---- layer 2 on the tensor core: D[128 x 128] = A[128 x 256] B[128 x 256]^T, one thread issues
    if (tid == 0) {
        {   // W2 has landed?
            const unsigned bw = (unsigned)__cvta_generic_to_shared(bar_w);
            asm volatile("{\n.reg .pred p;\nWAITW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONEW_%=;\nbra WAITW_%=;\nDONEW_%=:\n}" ::"r"(bw) : "memory");
        }
        // instruction descriptor (cute::UMMA::InstrDescriptor): D FP32, A / B BF16, both K-major, N = 128, M = 128
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kPolH2 >> 3) << 17) | ((unsigned)(kPolRows >> 4) << 24);
        const unsigned a0 = (unsigned)__cvta_generic_to_shared(sA), b0 = (unsigned)__cvta_generic_to_shared(sB);
#pragma unroll 1
        for (int ks = 0; ks < kPolH1 / 16; ++ks) {
            // K step of 16 = two core matrices along K; LBO = distance between them, SBO = distance between 8-row groups
            const unsigned long long da = pol_smem_desc(a0 + ks * 2 * (kPolRows / 8) * 128, (kPolRows / 8) * 128, 128);
            const unsigned long long db = pol_smem_desc(b0 + ks * 2 * (kPolH2 / 8) * 128, (kPolH2 / 8) * 128, 128);
            const unsigned accumulate = ks > 0 ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(da), "l"(db),
                         "r"(idesc), "r"(accumulate)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar_mma)) : "memory");
    }
    {   // everyone: accumulator ready
        const unsigned bm = (unsigned)__cvta_generic_to_shared(bar_mma);
        asm volatile("{\n.reg .pred p;\nWAITM_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONEM_%=;\nbra WAITM_%=;\nDONEM_%=:\n}" ::"r"(bm) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: this thread's quarter of its row of D (TMEM lane = row), + b2, tanh, head
    float out[MX];
#pragma unroll
    for (int o = 0; o < MX; ++o) out[o] = (o < nout1 && part == 0) ? sb3[o] : 0.f;
    const unsigned lane_base = tmem + ((unsigned)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c0 = part * (kPolH2 / kPolSplit); c0 < (part + 1) * (kPolH2 / kPolSplit); c0 += 16) {
        unsigned r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(lane_base + (unsigned)c0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float h2 = pol_tanh(__uint_as_float(r[c]) + sb2[c0 + c]);
            const float *wr = sw3 + (c0 + c) * nout1;
#pragma unroll
            for (int o = 0; o < MX; ++o)
                if (o < nout1) out[o] = fmaf(h2, wr[o], out[o]);
        }
    }
    // partial heads of the row's other three threads -> the row's first thread
    if (part > 0) {
#pragma unroll
        for (int o = 0; o < MX; ++o) spart[((part - 1) * kPolRows + row) * kPolMaxOut + o] = out[o];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    if (part > 0 || env >= io.batch) return;
#pragma unroll
    for (int q = 0; q < kPolSplit - 1; ++q)
#pragma unroll
        for (int o = 0; o < MX; ++o) out[o] += spart[(q * kPolRows + row) * kPolMaxOut + o];

    // ---- outputs
    float v = 0.f;                              // out[n_out], selected without dynamic register indexing
#pragma unroll
    for (int o = 0; o < MX; ++o) if (o == pp.n_out) v = out[o];
    io.value[env] = v;
    if (io.logits) {
#pragma unroll
        for (int o = 0; o < MX; ++o) if (o < nout1) io.logits[(size_t)env * nout1 + o] = out[o];
    }
    if (!io.sample) return;
    const unsigned ctr = pp.counters[env];
    pp.counters[env] = ctr + 1u;
    if (pp.discrete) {
        // log-softmax + Gumbel-max
        float mx = -INFINITY;
#pragma unroll
        for (int o = 0; o < MX; ++o) if (o < pp.n_out) mx = fmaxf(mx, out[o]);
        float se = 0.f;
#pragma unroll
        for (int o = 0; o < MX; ++o) if (o < pp.n_out) se += __expf(out[o] - mx);
        const float lse = mx + __logf(se);
        float best = -INFINITY, best_l = 0.f;
        int arg = 0;
#pragma unroll
        for (int o = 0; o < MX; ++o) {
            if (o < pp.n_out) {
                const float u = pol_uniform(pp.seed, (unsigned)env, ctr, (unsigned)o);
                const float g = out[o] - __logf(-__logf(u));
                if (g > best) { best = g; arg = o; best_l = out[o]; }
            }
        }
        io.act_discrete[env] = arg;
        io.logp[env] = best_l - lse;
    } else {
        float lp = 0.f;
#pragma unroll
        for (int o = 0; o < MX; ++o) {
            if (o < pp.n_out) {
                const float u1 = pol_uniform(pp.seed, (unsigned)env, ctr, (unsigned)(2 * o));
                const float u2 = pol_uniform(pp.seed, (unsigned)env, ctr, (unsigned)(2 * o + 1));
                const float z = sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853071795865f * u2);      // Box-Muller
                io.act_continuous[(size_t)env * pp.n_out + o] = (double)(pol_tanh(out[o]) + z);
                lp += -0.5f * z * z - 0.9189385332046727f;
            }
        }
        io.logp[env] = lp;
    }
}

constexpr size_t kPolSmemBytes = (size_t)kPolRows * kPolH1 * 2 + (size_t)kPolH2 * kPolH1 * 2 +
                                 sizeof(float) * (kPolMaxObs * kPolH1 + kPolH1 + kPolH2 + kPolH2 * kPolMaxOut + kPolMaxOut) + 64 +
                                 sizeof(float) * (kPolSplit - 1) * kPolRows * kPolMaxOut;

}  // namespace paintrl
