// dsim_policy.cuh — the action-sampling step on the caller side of the env path:
// distributions.py:6-38 `MyBetaDist` (inputs clamped to +-50, alpha / beta = log(exp(x) + 1) + 1, Beta(alpha, beta);
// sample -> action in (0,1); deterministic_sample -> mean; logp with x clamped to [0.01, 0.99], summed over actions).
// One thread per env: reads the policy head's [n][2A] logits, writes the [n][A] action rows the step kernel bulk-loads
// (16-byte aligned for A = 4) and the per-env log-probability PPO needs.
//
// Gamma(a >= 1) by Marsaglia-Tsang squeeze/rejection on a counter-based Philox4x32-10 stream keyed by
// (seed, GLOBAL env id) with counter (attempt block, step, stream 2, variate index): runs on 1 or 8 GPUs, and replays
// of a CUDA graph with a new `step`, draw identical numbers.  alpha, beta >= 1 always (softplus + 1), so no a < 1 boost.
#pragma once
#include "dsim_device.cuh"

namespace dsim {

constexpr int kBetaMaxBlocks = 8;     // 16 attempts; acceptance >= 95 % per attempt for a >= 1

template <typename T> DSIM_DEV T softplus1(T x) {     // torch.log(torch.exp(clamp(x, -50, 50)) + 1.0) + 1.0
    x = clamp_(x, T(-50), T(50));
    if constexpr (std::is_same<T, float>::value) return logf(expf(x) + 1.0f) + 1.0f; else return log(exp(x) + 1.0) + 1.0;
}
template <typename T> DSIM_DEV T lgamma_(T x) { if constexpr (std::is_same<T, float>::value) return lgammaf(x); else return lgamma(x); }
template <typename T> DSIM_DEV T log1p_(T x) { if constexpr (std::is_same<T, float>::value) return log1pf(x); else return log1p(x); }

// one Gamma(a, 1) variate, a >= 1; `vi` = variate index inside the env's step (0 .. 2A-1)
template <typename T> DSIM_DEV T gamma_mt(T a, uint32_t seed, uint32_t env, uint32_t step, uint32_t vi) {
    const T d = a - T(1.0 / 3.0), c = rsqrt_(T(9) * d);
    for (uint32_t blk = 0; blk < (uint32_t)kBetaMaxBlocks; blk++) {
        const U4 x = philox4x32(blk, step, 2u, vi, seed, env);
        T z[2];
        {   // exact (libm) Box-Muller here: the accept test compares against log(u)
            const T r = sqrt_(T(-2) * log_(u01<T>(x.x)));
            T sn, cs;
            sincos_(T(2 * kPi) * u01<T>(x.y), &sn, &cs);
            z[0] = r * cs; z[1] = r * sn;
        }
        const uint32_t uw[2] = {x.z, x.w};
        #pragma unroll
        for (int j = 0; j < 2; j++) {
            const T t = T(1) + c * z[j];
            if (t > T(0)) {
                const T v = t * t * t, u = u01<T>(uw[j]);
                if (log_(u) < T(0.5) * z[j] * z[j] + d - d * v + d * log_(v)) return d * v;
            }
        }
    }
    return d;                                          // (probability < 1e-20) never silent garbage: the mode-ish value
}

template <typename T, int A>
__global__ void __launch_bounds__(128) beta_policy_kernel(int n, const T *logits, uint32_t seed, uint32_t env_base, uint32_t step,
                                                          const uint32_t *step_dev, int deterministic, T *actions, T *logp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (step_dev) step += *step_dev;                     // device-resident step counter: CUDA-graph replays draw fresh numbers
    T x[2 * A];
    #pragma unroll
    for (int k = 0; k < 2 * A; k++) x[k] = logits[(size_t)i * (2 * A) + k];
    T lp = T(0), out[A];
    #pragma unroll
    for (int k = 0; k < A; k++) {
        const T a = softplus1(x[k]), b = softplus1(x[A + k]);      // torch.chunk(inputs, 2, dim=-1): alpha first, beta second
        T s;
        if (deterministic) s = a / (a + b);                         // Beta mean (distributions.py:24-26)
        else {
            const T ga = gamma_mt(a, seed, env_base + (uint32_t)i, step, (uint32_t)(2 * k));
            const T gb = gamma_mt(b, seed, env_base + (uint32_t)i, step, (uint32_t)(2 * k + 1));
            s = ga / (ga + gb);
        }
        out[k] = s;
        const T xc = clamp_(s, T(1e-2), T(1 - 1e-2));              // logp clamps (distributions.py:19-22)
        lp += lgamma_(a + b) - lgamma_(a) - lgamma_(b) + (a - T(1)) * log_(xc) + (b - T(1)) * log1p_(-xc);
    }
    #pragma unroll
    for (int k = 0; k < A; k++) actions[(size_t)i * A + k] = out[k];
    if (logp) logp[i] = lp;
}

}  // namespace dsim
