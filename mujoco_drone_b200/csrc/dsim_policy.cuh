// dsim_policy.cuh — the action-sampling step on the caller side of the env path:
// distributions.py:6-38 `MyBetaDist` (inputs clamped to +-50, alpha / beta = log(exp(x) + 1) + 1, Beta(alpha, beta);
// sample -> action in (0,1); deterministic_sample -> mean; logp with x clamped to [0.01, 0.99], summed over actions).
// One thread per env: reads the policy head's [n][2A] logits, writes the [n][A] action rows the step kernel bulk-loads
// (16-byte aligned for A = 4) and the per-env log-probability PPO needs.
//
// Gamma(a >= 1) by Marsaglia-Tsang squeeze/rejection on a counter-based Philox4x32-10 stream keyed by
// (seed, GLOBAL env id) with counter (attempt block, step, stream 2, variate index); the first attempt of an action's two
// variates shares one Philox block: runs on 1 or 8 GPUs, and replays
// of a CUDA graph with a new `step`, draw identical numbers.  alpha, beta >= 1 always (softplus + 1), so no a < 1 boost.
#pragma once
#include "dsim_device.cuh"

namespace dsim {

constexpr int kBetaMaxBlocks = 8;     // 16 attempts; acceptance >= 95 % per attempt for a >= 1

template <typename T> DSIM_DEV T softplus1(T x) {     // torch.log(torch.exp(clamp(x, -50, 50)) + 1.0) + 1.0
    x = clamp_(x, T(-50), T(50));
    if constexpr (std::is_same<T, float>::value) return __logf(__expf(x) + 1.0f) + 1.0f;     // MUFU ex2 / lg2: ~1e-6 relative, 6 instructions instead of libm's ~30
    else return log(exp(x) + 1.0) + 1.0;
}
template <typename T> DSIM_DEV T log1p_(T x) { if constexpr (std::is_same<T, float>::value) return log1pf(x); else return log1p(x); }
// log Gamma(x) for x >= 1 (alpha, beta = softplus + 1 >= 1, their sum >= 2).  FP32: shift the argument up by 4 and use the
// Stirling series there (z >= 5: truncation error < 1e-8), two MUFU logs instead of libm's ~50-instruction lgammaf;
// FP64 keeps libm.
template <typename T> DSIM_DEV T lgamma_ge1(T x) {
    if constexpr (std::is_same<T, float>::value) {
        const float z = x + 4.0f, iz = rcp_(z), iz2 = iz * iz;
        const float series = iz * (0.0833333333f + iz2 * (-0.00277777778f + iz2 * 0.000793650794f));
        return (z - 0.5f) * __logf(z) - z + 0.918938533f + series - __logf(x * (x + 1.0f) * (x + 2.0f) * (x + 3.0f));
    } else return lgamma(x);
}
template <typename T> DSIM_DEV T flog_(T x) { if constexpr (std::is_same<T, float>::value) return __logf(x); else return log(x); }

// Marsaglia-Tsang acceptance test for Gamma(a, 1), a >= 1, given one normal z and one uniform u
template <typename T> DSIM_DEV bool mt_accept(T d, T c, T z, T u, T &out) {
    const T t = T(1) + c * z;
    if (!(t > T(0))) return false;
    const T v = t * t * t;
    if (!(flog_(u) < T(0.5) * z * z + d - d * v + d * flog_(v))) return false;
    out = d * v;
    return true;
}
// One env's action row: x[2A] head outputs -> out[A] actions in (0, 1) and their summed log-probability.  Called by the
// stand-alone sampling kernel below and from the logits epilogue of the fused policy kernel (dsim_policy_mlp.cu); every
// lane of the calling warp must enter (the retry loop is warp-wide).
template <typename T, int A>
DSIM_DEV void beta_row(const T (&x)[2 * A], uint32_t seed, uint32_t env, uint32_t step, int deterministic, T (&out)[A], T &lp) {
    // alpha_k = al[k], beta_k = al[A + k]  (torch.chunk(inputs, 2, dim=-1): alpha first, beta second); variate index 2k / 2k+1
    T al[2 * A], g[2 * A];
    #pragma unroll
    for (int k = 0; k < 2 * A; k++) { al[k] = softplus1(x[k]); g[k] = T(0); }
    if (!deterministic) {
        // first attempt of all 2A variates: one Philox block per action (Box-Muller pair -> two normals, two more words -> two uniforms)
        unsigned pend = 0;
        #pragma unroll
        for (int k = 0; k < A; k++) {
            const T da = al[k] - T(1.0 / 3.0), db = al[A + k] - T(1.0 / 3.0);
            const U4 r = philox4x32(0u, step, 2u, 2u * k, seed, env);
            T z0, z1;
            box_muller(r.x, r.y, z0, z1);
            if (!mt_accept(da, rsqrt_(T(9) * da), z0, u01<T>(r.z), g[2 * k])) pend |= 1u << (2 * k);
            if (!mt_accept(db, rsqrt_(T(9) * db), z1, u01<T>(r.w), g[2 * k + 1])) pend |= 1u << (2 * k + 1);
        }
        // retries (>= 95 % of first attempts are accepted): ONE warp-wide loop in which every lane works on its lowest
        // pending variate, instead of a divergent retry branch per variate.  Stream of variate vi: counter blocks 1, 2, ...
        uint32_t blk = 1;
        while (__any_sync(__activemask(), pend != 0)) {
            if (pend) {
                const int vi = __ffs(pend) - 1;
                T a_sel = al[0];
                #pragma unroll
                for (int j = 1; j < 2 * A; j++) a_sel = ((vi >> 1) + (vi & 1) * A == j) ? al[j] : a_sel;
                const T d = a_sel - T(1.0 / 3.0), c = rsqrt_(T(9) * d);
                const U4 r = philox4x32(blk, step, 2u, (uint32_t)vi, seed, env);
                T z0, z1, got = d;
                box_muller(r.x, r.y, z0, z1);
                const bool acc = mt_accept(d, c, z0, u01<T>(r.z), got) || mt_accept(d, c, z1, u01<T>(r.w), got) || (blk + 1 >= (uint32_t)kBetaMaxBlocks);
                if (acc) {
                    #pragma unroll
                    for (int j = 0; j < 2 * A; j++) g[j] = (j == vi) ? got : g[j];
                    pend &= pend - 1;
                    blk = 1;
                } else blk++;
            }
        }
    }
    lp = T(0);
    #pragma unroll
    for (int k = 0; k < A; k++) {
        const T a = al[k], b = al[A + k];
        const T s = deterministic ? a / (a + b) : g[2 * k] / (g[2 * k] + g[2 * k + 1]);      // Beta mean (distributions.py:24-26) | Ga / (Ga + Gb)
        out[k] = s;
        const T xc = clamp_(s, T(1e-2), T(1 - 1e-2));              // logp clamps (distributions.py:19-22)
        lp += lgamma_ge1(a + b) - lgamma_ge1(a) - lgamma_ge1(b) + (a - T(1)) * flog_(xc) + (b - T(1)) * flog_(T(1) - xc);
    }
}

template <typename T, int A>
__global__ void __launch_bounds__(128) beta_policy_kernel(int n, const T *logits, uint32_t seed, uint32_t env_base, uint32_t step,
                                                          const uint32_t *step_dev, int deterministic, T *actions, T *logp) {
    // programmatic dependent launch: the next kernel of the stream (the env step) may be scheduled while this grid drains;
    // nothing an earlier kernel wrote (the logits) is read before the dependency wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (step_dev) step += *step_dev;                     // device-resident step counter: CUDA-graph replays draw fresh numbers
    T x[2 * A], out[A], lp;
    #pragma unroll
    for (int k = 0; k < 2 * A; k++) x[k] = logits[(size_t)i * (2 * A) + k];
    beta_row<T, A>(x, seed, env_base + (uint32_t)i, step, deterministic, out, lp);
    #pragma unroll
    for (int k = 0; k < A; k++) actions[(size_t)i * A + k] = out[k];
    if (logp) logp[i] = lp;
}

}  // namespace dsim
