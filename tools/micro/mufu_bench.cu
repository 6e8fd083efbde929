// micro-benchmark: SM-wide throughput of the special-function unit for the tanh flavours (B200, sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE> __global__ void k(float *out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    unsigned pa = threadIdx.x * 0x00010001u, pb = pa + 0x01000100u, pc = pa + 7, pd = pa + 11;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d)); }
        if (MODE == 1) { asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(pa)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(pb)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(pc)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(pd)); }
        if (MODE == 2) { asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(pa)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(pb)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(pc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(pd)); }
        if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
        if (MODE == 4) { asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(pa)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(pb)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(pc)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(pd)); }
        if (MODE == 5) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(b)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(c)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(d)); }
        if (MODE == 6) { asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(pa)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(pb)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(pc)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(pd)); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(pa ^ pb ^ pc ^ pd);
}
template <int MODE> void run(const char *name, float *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    k<MODE><<<blocks, threads>>>(out, 16);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * threads / 32 * iters * 4;
    double per_sm_clk = warp_instr / 148 / (ms * 1e-3 * 1.965e9);       // warp-instructions per SM per clock (at 1965 MHz)
    printf("%-22s %8.3f ms  %.3f warp-instr/clk/SM = %.1f lanes/clk/SM\n", name, ms, per_sm_clk, per_sm_clk * 32);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<0>("tanh.approx.f32", out); run<1>("tanh.approx.bf16x2", out); run<2>("tanh.approx.f16x2", out); run<3>("ex2.approx.f32", out);
    run<4>("fma.f16x2", out); run<5>("fma.f32", out); run<6>("fma.bf16x2", out);
    return 0;
}
