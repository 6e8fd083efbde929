// Micro-benchmark for the next step-kernel lever: the SAME templated physics (dsim::substep) instantiated on a
// two-wide numeric type F2 whose +, -, *, fma are the packed f32x2 instructions of sm_100a, i.e. two envs per lane, against
// the scalar float instantiation (one env per lane).  Compute only (state in registers), no memory traffic.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o substep_f2_bench substep_f2_bench.cu && ./substep_f2_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mujoco_drone_b200/csrc/dsim_device.cuh"

namespace dsim {
struct F2 {
    unsigned long long v;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ F2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); }
    __device__ __forceinline__ F2(double c) : F2((float)c, (float)c) {}
    __device__ __forceinline__ F2(float c) : F2(c, c) {}
    __device__ __forceinline__ F2(int c) : F2((float)c, (float)c) {}
    __device__ __forceinline__ float lo() const { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
    __device__ __forceinline__ float hi() const { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
};
#define F2OP __device__ __forceinline__
F2OP F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
F2OP F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
F2OP F2 operator-(F2 a) { F2 r; r.v = a.v ^ 0x8000000080000000ull; return r; }
F2OP F2 operator-(F2 a, F2 b) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(b.v), "l"(F2(-1.0f).v), "l"(a.v)); return r; }
F2OP F2 &operator+=(F2 &a, F2 b) { a = a + b; return a; }
F2OP F2 &operator-=(F2 &a, F2 b) { a = a - b; return a; }
F2OP bool operator>=(F2 a, F2 b) { return a.lo() >= b.lo() && a.hi() >= b.hi(); }      // (benchmark only: both halves take one branch)
F2OP bool operator<=(F2 a, F2 b) { return a.lo() <= b.lo() && a.hi() <= b.hi(); }
F2OP float rcpf_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
F2OP float sqrtf_(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
F2OP F2 operator/(double one, F2 x) { return F2((float)one * rcpf_(x.lo()), (float)one * rcpf_(x.hi())); }
F2OP F2 operator/(F2 a, F2 b) { return a * F2(rcpf_(b.lo()), rcpf_(b.hi())); }
F2OP F2 sqrt(F2 x) { return F2(sqrtf_(x.lo()), sqrtf_(x.hi())); }
F2OP F2 fabs(F2 x) { F2 r; r.v = x.v & 0x7fffffff7fffffffull; return r; }
F2OP F2 fmax(F2 a, F2 b) { return F2(fmaxf(a.lo(), b.lo()), fmaxf(a.hi(), b.hi())); }
F2OP F2 fmin(F2 a, F2 b) { return F2(fminf(a.lo(), b.lo()), fminf(a.hi(), b.hi())); }
F2OP void sincos(F2 a, F2 *s, F2 *c) {
    float s0, c0, s1, c1;
    sincos_hinge(a.lo(), &s0, &c0); sincos_hinge(a.hi(), &s1, &c1);
    *s = F2(s0, s1); *c = F2(c0, c1);
}
}  // namespace dsim
using namespace dsim;

template <typename T> __device__ T mkT(float a, float b);
template <> __device__ float mkT<float>(float a, float) { return a; }
template <> __device__ F2 mkT<F2>(float a, float b) { return F2(a, b); }
template <typename T> __device__ float sumT(T x);
template <> __device__ float sumT<float>(float x) { return x; }
template <> __device__ float sumT<F2>(F2 x) { return x.lo() + x.hi(); }

template <typename T>
__global__ void __launch_bounds__(128) bench(int iters, float *out) {
    const float u = (threadIdx.x + blockIdx.x * blockDim.x) * 1e-6f, w = u + 3e-4f;
    EnvState<T> s;
    s.pos = mk(mkT<T>(u, w), mkT<T>(-u, -w), mkT<T>(0.1f + u, 0.1f + w));
    s.qw = mkT<T>(0.9f, 0.92f); s.qx = mkT<T>(0.1f + u, 0.12f); s.qy = mkT<T>(-0.2f, -0.18f + w); s.qz = mkT<T>(0.3f, 0.28f);
    s.hx = mkT<T>(0.2f + u, 0.25f); s.hy = mkT<T>(-0.3f, -0.2f + w);
    s.vel = mk(mkT<T>(0.5f, 0.4f), mkT<T>(-0.2f + u, -0.1f), mkT<T>(0.1f, 0.2f + w));
    s.om = mk(mkT<T>(0.3f, 0.2f + w), mkT<T>(-0.4f + u, -0.3f), mkT<T>(0.2f, 0.1f));
    s.hvx = mkT<T>(0.1f, 0.2f); s.hvy = mkT<T>(-0.1f, -0.2f);
    for (int k = 0; k < 4; k++) s.act[k] = mkT<T>(0.4f + 0.1f * k, 0.5f + 0.05f * k);
    s.acc = mk(T(0.f), T(0.f), T(0.f));
    EnvConsts<T> c;
    c.mB = mkT<T>(1.0f, 1.1f); c.cz = mkT<T>(0.002f, 0.0021f); c.IBx = mkT<T>(0.006f, 0.0065f); c.IBy = mkT<T>(0.006f, 0.0065f); c.IBz = mkT<T>(0.011f, 0.012f);
    c.mD = mkT<T>(0.54f, 0.5f); c.zD = mkT<T>(-0.93f, -0.9f); c.IDx = mkT<T>(0.05f, 0.045f); c.IDz = mkT<T>(0.0004f, 0.00035f);
    c.Fs = mkT<T>(0.84f, 0.8f); c.F = mkT<T>(7.0f, 6.8f); c.kq = mkT<T>(0.07f, 0.068f); c.inv_tau = mkT<T>(100.f, 95.f);
    T ctrl[4];
    for (int k = 0; k < 4; k++) ctrl[k] = mkT<T>(0.5f + 0.05f * k, 0.45f + 0.05f * k);
    const T h = T(0.005f);
    #pragma unroll 1
    for (int it = 0; it < iters; it++) substep<T, true, true>(s, c, ctrl, h);
    out[threadIdx.x + blockIdx.x * blockDim.x] = sumT<T>(s.pos.x) + sumT<T>(s.qw) + sumT<T>(s.hx) + sumT<T>(s.acc.z) + sumT<T>(s.hvy);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; mode++) {
        const bool packed = mode & 1;
        const int ctas_per_sm = mode < 2 ? (packed ? 2 : 4) : (packed ? 4 : 8);
        const int grid = sms * ctas_per_sm;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (packed) bench<F2><<<grid, 128>>>(iters, out); else bench<float><<<grid, 128>>>(iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double envs = (double)grid * 128 * (packed ? 2 : 1);
        printf("%-6s %d CTAs/SM: %8.3f ms  %.3e env-substeps/s  (%s)\n", packed ? "f32x2" : "float", ctas_per_sm, ms, envs * iters / (ms * 1e-3),
               cudaGetErrorString(cudaGetLastError()));
    }
    float hsum[4];
    cudaMemcpy(hsum, out, sizeof hsum, cudaMemcpyDeviceToHost);
    printf("checksum %g %g\n", hsum[0], hsum[1]);
    return 0;
}
