// dsim_policy_mlp.cu — RMA_full policy inference (models/PPO/RMA/RMA_model.py:48-71, 79-116, train_adaptation=False)
// as ONE fused tcgen05 kernel for sm_100a: obs + previous action in, Beta-head logits + value out, nothing in between
// touches HBM.  SURVEY.md §8f-1 (the caller side of the env-step path, BASELINE config 5).
//
//   z   = E2 tanh(E1 e + e1) + e2                       6 -> 32 -> 8      (CUDA cores, per row)
//   h1  = tanh(W1 [s, a_prev, z] + b1)                  28 -> 256         tcgen05, A from shared memory
//   h2  = tanh(W2 h1 + b2)                              256 -> 128        tcgen05, A from TMEM
//   (BatchNorm1d in eval mode is an affine map: folded into W3 / V1 on the host)
//   l1  = tanh(W3' h2 + b3'),  logits = W4 l1 + b4      128 -> 128 -> 8
//   v1  = tanh(V1' h2 + c1'),  v2 = tanh(V2 v1 + c2),  value = V3 v2 + c3
//
// One CTA per SM, persistent, 256 threads = two warpgroups.  All bf16 weights (180 KB, pre-packed on the host into the
// UMMA K-major no-swizzle canonical layout) stay in shared memory for the life of the CTA.  Each warpgroup owns a tile of
// 128 envs (thread = env row = TMEM lane) and half of the SM's tensor memory: a 128-column FP32 accumulator D and a
// 128-column operand region A that holds the bf16 activations of the previous layer (tcgen05.mma with A in TMEM, the
// FlashAttention-4 pattern), so activations go accumulator -> registers (bias, tanh, bf16 pack) -> TMEM and never to
// shared memory or HBM.  While one warpgroup runs an epilogue on the CUDA cores the other one's MMAs occupy the tensor
// core.  Accumulation is FP32; operands are bf16 (the same numerics as the torch autocast path it replaces).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "../../include/dronesim_b200.h"
#include "dsim_policy.cuh"

namespace {

constexpr int S_DIM = 16, P_DIM = 6, A_DIM = 4, E_DIM = 8, OBS_DIM = S_DIM + P_DIM, ENC_H = 32;
constexpr int K1 = 32;                                   // 28 inputs padded to a multiple of 16
// packed bf16 weight blob (element offsets); each matrix [N][K] is stored as [K/8][N][8] (core matrices of 8 rows x 16 B)
constexpr int W1_OFF = 0, W1_N = 256;                    // K = 32
constexpr int W2_OFF = W1_OFF + 256 * 32, W2_N = 128;    // K = 256
constexpr int W3_OFF = W2_OFF + 128 * 256, W3_N = 256;   // K = 128; rows 0-127 logits hidden (W3'), rows 128-255 value hidden (V1')
constexpr int W4_OFF = W3_OFF + 256 * 128, W4_N = 16;    // K = 128; rows 0-7 real
constexpr int V2_OFF = W4_OFF + 16 * 128, V2_N = 128;    // K = 128
constexpr int W_ELEMS = V2_OFF + 128 * 128;              // 92160 bf16 = 180 KB
// fp32 constants blob (float offsets)
constexpr int C_B1 = 0, C_B2 = 256, C_B3 = 384, C_B4 = 640, C_C2 = 656, C_V3 = 784, C_C3 = 912, C_E1 = 913, C_E1B = C_E1 + ENC_H * P_DIM,
              C_E2 = C_E1B + ENC_H, C_E2B = C_E2 + E_DIM * ENC_H, C_ELEMS = 1408;
static_assert(C_E2B + E_DIM <= C_ELEMS, "constants blob too small");
constexpr int X0_BYTES = 128 * K1 * 2;                   // per-warpgroup first-layer operand tile (canonical layout), 8 KB
constexpr int SMEM_BYTES = W_ELEMS * 2 + C_ELEMS * 4 + 2 * X0_BYTES + 64 + 128;   // + b4 padded to 32 columns

#define DEV __device__ __forceinline__
DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp): K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in
// 16-byte units: 8 rows x 16 B core matrices, SBO = stride between 8-row groups, LBO = stride between the two K halves
DEV uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);                                 // version = 1 (Blackwell), base_offset = 0, layout_type = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// next K-step of the same operand: only the 14-bit start-address field moves (no carry out of it for our < 227 KB layout)
DEV uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

DEV void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
DEV void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
DEV void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEV void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEV void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Ping-pong schedule (named barriers 3 + w): the special-function unit (tanh) is the scarce pipe of this kernel and both
// warpgroups run the same program, so left alone they fall into lock-step, fight over the SFU in their epilogues and
// leave the tensor core idle meanwhile.  A warpgroup may only run an epilogue while it holds the turn; it passes the turn
// to the other warpgroup when done, which forces the two to alternate: one in its epilogue, the other issuing MMAs.
DEV void turn_wait(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 3) : "memory"); }
DEV void turn_pass(int wg) { asm volatile("bar.arrive %0, 256;" ::"r"((wg ^ 1) + 3) : "memory"); }
DEV void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
DEV void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: a descriptor bug must end in an error flag, never in a hung GPU
DEV bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); it++) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
DEV float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
DEV uint32_t pack_bf16(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// TMEM -> registers: 32 consecutive FP32 columns of this thread's lane
DEV void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}
// software-pipelined form: issue the load of the next 32 columns, compute on the current ones, wait only then.  The wait
// "touches" the destination registers so that the compiler cannot schedule their consumers above it.
DEV void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
DEV void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                      "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]),
                      "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}
DEV void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
}
// registers -> TMEM: 16 consecutive 32-bit columns (32 packed bf16) of this thread's lane
DEV void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                   "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
DEV void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                 "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                   "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
                   "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// The NEXT layer's bias goes into the accumulator columns as soon as this layer's values have left them (every MMA
// accumulates): one tcgen05.st per 32 columns replaces 32 FADDs in the next epilogue, whose single warp per scheduler is
// bound by its own issue rate.
DEV void bias_to_tmem32(uint32_t taddr, const float *bias) {
    uint32_t r[32];
    #pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + j);
        r[j] = __float_as_uint(b.x); r[j + 1] = __float_as_uint(b.y); r[j + 2] = __float_as_uint(b.z); r[j + 3] = __float_as_uint(b.w);
    }
    tmem_st32(taddr, r);
}
DEV void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct MlpParams {
    const uint16_t *w;        // packed bf16 weights [W_ELEMS]
    const float *c;           // fp32 constants [C_ELEMS]
    const float *obs;         // [n][22]
    const float *prev_action; // [n][4]
    const unsigned char *reset_mask;   // [n] or nullptr: rows with mask != 0 see a zero previous action (episode start)
    float *logits;            // [n][8] or nullptr (fused sampling only)
    float *value;             // [n]
    float *actions;           // [n][4] or nullptr: Beta-head sample drawn in the logits epilogue (dsim_policy.cuh::beta_row)
    float *logp;              // [n] or nullptr
    const uint32_t *step_dev; // optional device step counter added to `step`
    uint32_t seed, env_base, step;
    int deterministic;
    int *error;               // set to 1 on a barrier timeout
    long long *dbg;           // debug: clock64 stamps [2 warpgroups][32] of CTA 0's first tile, or nullptr
    int n, ntiles;
};

// packed activation: two accumulator values + bias -> bf16x2 -> tanh.approx.bf16x2.  One MUFU op yields two activations
// already in the operand format (the special-function unit, 16 ops/clk/SM, is the epilogue's scarce pipe: ~900 tanh per row).
DEV uint32_t tanh_pack(float lo, float hi) {
    uint32_t r = pack_bf16(lo, hi);
    asm("tanh.approx.bf16x2 %0, %0;" : "+r"(r));
    return r;
}
// The same activation on the FMA pipe, for a share of the columns: the SFU delivers 16 tanh/clk/SM and is the epilogue's
// bound, while the FP32 pipe idles.  tanh(x) ~ x P(x^2) on |x| <= 3.5 (clamped beyond: 1 - tanh(3.5) = 1.8e-3), degree-13
// odd polynomial fitted for minimax RELATIVE error (3.1e-3, the size of the bf16 rounding that follows), evaluated for two
// columns at once with packed fma.rn.f32x2: 14 issue slots per column pair instead of 16 SFU cycles.
#ifndef MLP_POLY
#define MLP_POLY 2           // of every 8 column pairs, this many go to the FMA pipe (sweep 0/2/3/4/5: 187/180/181/190/198 us)
#endif
DEV uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DEV void unpk2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEV uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DEV uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEV void tanh_poly2(float &lo, float &hi) {
    constexpr float R = 3.5f;
    lo = fminf(fmaxf(lo, -R), R); hi = fminf(fmaxf(hi, -R), R);
    const uint64_t x = pk2(lo, hi), s = mul2(x, x);
    #define K2(c) pk2(c, c)
    uint64_t q = fma2(s, K2(2.013471598e-06f), K2(-8.968956237e-05f));
    q = fma2(q, s, K2(1.616502277e-03f)); q = fma2(q, s, K2(-1.532690791e-02f)); q = fma2(q, s, K2(8.490159052e-02f));
    q = fma2(q, s, K2(-3.053695960e-01f)); q = fma2(q, s, K2(9.969368461e-01f));
    #undef K2
    unpk2(mul2(q, x), lo, hi);
}
DEV uint32_t tanh_poly_pack(float lo, float hi) { tanh_poly2(lo, hi); return pack_bf16(lo, hi); }
// epilogue: D[0..127] (bias already inside) -> tanh -> bf16 -> A region at column `dst_col` (64 columns); the first
// `next_cols` accumulator columns are re-armed with the next layer's bias.  The first accumulator load is issued BEFORE the
// warpgroup waits for its turn (the turn rations the SFU, not tensor memory), and the load of columns c+32.. is in flight
// while columns c.. are computed.  One copy (noinline): four unrolled 32-column stages, five call sites.
__device__ __noinline__ void epilogue_to_tmem(uint32_t tD, uint32_t tAdst, const float *next_bias, int next_cols, int wg) {
    uint32_t ra[32], rb[32];
    tmem_ld32_issue(tD, ra);
    turn_wait(wg);
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t (&cur)[32] = (k & 1) ? rb : ra;
        uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
        tmem_ld32_wait(cur);
        if (k < 3) { tmem_ld32_issue(tD + 32 * (k + 1), nxt); __syncwarp(); }   // (the barrier keeps ptxas from sinking the load below the math)
        const int c = 32 * k;
        if (c < next_cols) bias_to_tmem32(tD + c, next_bias + c);
        uint32_t pk[16];
        #pragma unroll
        for (int j = 0; j < 16; j++) {
            const float lo = __uint_as_float(cur[2 * j]), hi = __uint_as_float(cur[2 * j + 1]);
#if defined(MLP_DBG_NOPACK)
            pk[j] = __float_as_uint(lo) ^ __float_as_uint(hi);
#elif defined(MLP_DBG_NOTANH)
            pk[j] = pack_bf16(lo, hi);
#else
            pk[j] = (j & 7) < MLP_POLY ? tanh_poly_pack(lo, hi) : tanh_pack(lo, hi);
#endif
        }
        tmem_st16(tAdst + c / 2, pk);
    }
    tmem_st_wait();
}

__global__ void __launch_bounds__(256, 1) rma_full_forward_kernel(const MlpParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint16_t *s_w = reinterpret_cast<uint16_t *>(smem);
    float *s_c = reinterpret_cast<float *>(smem + W_ELEMS * 2);
    unsigned char *s_x0 = smem + W_ELEMS * 2 + C_ELEMS * 4;                  // [2 warpgroups][8 KB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_x0 + 2 * X0_BYTES);      // [2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);
    float *s_b4 = reinterpret_cast<float *>(s_tmem + 12);                      // b4 (16 logits columns, 8 real) padded to 32
    const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, wq = (tid >> 5) & 3;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- one-time setup: weights + constants -> shared memory, TMEM allocation, barriers
    {
        const uint4 *gw = reinterpret_cast<const uint4 *>(p.w);
        uint4 *sw = reinterpret_cast<uint4 *>(s_w);
        for (int i = tid; i < W_ELEMS * 2 / 16; i += 256) sw[i] = __ldg(gw + i);
        for (int i = tid; i < C_ELEMS; i += 256) s_c[i] = __ldg(p.c + i);
        if (tid < 32) s_b4[tid] = tid < 16 ? __ldg(p.c + C_B4 + tid) : 0.0f;
    }
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {                                                          // warp 0 allocates all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // generic smem writes -> visible to the tensor core's async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t lane_off = (uint32_t)(32 * wq) << 16;
    const uint32_t tD = tmem_base + wg * 256 + lane_off;                     // accumulator, 128 columns
    const uint32_t tA = tD + 128;                                            // operand region, 128 columns = 256 bf16 per lane
    const uint32_t tD0 = tmem_base + wg * 256, tA0 = tD0 + 128;              // lane-0 addresses for the MMA operands
    const uint32_t w_addr = smem_u32(s_w), x0_addr = smem_u32(s_x0 + wg * X0_BYTES);
    uint64_t *bar = &s_bar[wg];
    uint32_t parity = 0;
    bool ok = true;
    #pragma unroll 1
    for (int c = 0; c < 128; c += 32) bias_to_tmem32(tD + c, s_c + C_B1 + c);   // first layer's bias for the first tile
    tmem_st_wait();

    // input rows are software-pipelined: nxt_* hold the obs / previous action of the tile this warpgroup will process next
    float nxt_o[OBS_DIM], nxt_a[A_DIM];
    auto load_row = [&](int tile_) {
        const int r_ = tile_ * 128 + wt;
        const bool live_ = tile_ < p.ntiles && r_ < p.n;
        #pragma unroll
        for (int k = 0; k < OBS_DIM; k += 2) {
            const float2 v = live_ ? __ldg(reinterpret_cast<const float2 *>(p.obs + (size_t)r_ * OBS_DIM + k)) : make_float2(0.f, 0.f);
            nxt_o[k] = v.x; nxt_o[k + 1] = v.y;
        }
        const bool fresh = live_ && p.reset_mask && p.reset_mask[r_];
        const float4 v = (live_ && !fresh) ? __ldg(reinterpret_cast<const float4 *>(p.prev_action) + r_) : make_float4(0.f, 0.f, 0.f, 0.f);
        nxt_a[0] = v.x; nxt_a[1] = v.y; nxt_a[2] = v.z; nxt_a[3] = v.w;
    };
    // programmatic dependent launch: everything above (weights into shared memory, tensor-memory allocation) reads nothing an
    // earlier kernel of the stream writes and may overlap its tail; the observation rows come after the dependency wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    load_row(2 * blockIdx.x + wg);
    const uint32_t step_now = p.step + (p.step_dev ? __ldg(p.step_dev) : 0u);
    constexpr int kTurnsPerTile = 6;                                       // epilogues that take a turn (the 16-column logits read does not)
    if (wg == 1) turn_pass(1);                                             // warpgroup 0 holds the first turn
    int dbg_k = 0;
    auto stamp = [&]() { if (p.dbg && blockIdx.x == 0 && wt == 0 && dbg_k < 32) p.dbg[wg * 32 + dbg_k++] = clock64(); };
    for (int tile = 2 * blockIdx.x + wg; tile < p.ntiles; tile += 2 * gridDim.x) {
        stamp();
        const int row = tile * 128 + wt;
        const bool live = row < p.n;
        // ---- layer-0 operand: [s(16), a_prev(4), z(8), 0(4)] as bf16 into the canonical smem tile
        {
            float o[OBS_DIM], a[A_DIM];
            #pragma unroll
            for (int k = 0; k < OBS_DIM; k++) o[k] = nxt_o[k];
            #pragma unroll
            for (int k = 0; k < A_DIM; k++) a[k] = nxt_a[k];
            load_row(tile + 2 * (int)gridDim.x);                           // prefetch: in flight during this tile's seven MMA / epilogue rounds
            float hdn[ENC_H];                                                // parameter encoder on the CUDA cores (448 MACs)
            #pragma unroll
            for (int j = 0; j < ENC_H; j++) {
                float acc = s_c[C_E1B + j];
                #pragma unroll
                for (int k = 0; k < P_DIM; k++) acc = fmaf(s_c[C_E1 + j * P_DIM + k], o[S_DIM + k], acc);
                hdn[j] = tanh_fast(acc);
            }
            float x[K1];
            #pragma unroll
            for (int k = 0; k < S_DIM; k++) x[k] = o[k];
            #pragma unroll
            for (int k = 0; k < A_DIM; k++) x[S_DIM + k] = a[k];
            #pragma unroll
            for (int e = 0; e < E_DIM; e++) {
                float acc = s_c[C_E2B + e];
                #pragma unroll
                for (int j = 0; j < ENC_H; j++) acc = fmaf(s_c[C_E2 + e * ENC_H + j], hdn[j], acc);
                x[S_DIM + A_DIM + e] = acc;
            }
            #pragma unroll
            for (int k = S_DIM + A_DIM + E_DIM; k < K1; k++) x[k] = 0.f;
            #pragma unroll
            for (int kc = 0; kc < K1 / 8; kc++) {                            // element (row, k) at (k/8) * 2048 + row * 16 + (k%8) * 2
                uint4 q;
                q.x = pack_bf16(x[8 * kc + 0], x[8 * kc + 1]); q.y = pack_bf16(x[8 * kc + 2], x[8 * kc + 3]);
                q.z = pack_bf16(x[8 * kc + 4], x[8 * kc + 5]); q.w = pack_bf16(x[8 * kc + 6], x[8 * kc + 7]);
                *reinterpret_cast<uint4 *>(s_x0 + wg * X0_BYTES + kc * 2048 + wt * 16) = q;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fence_before_sync();
        wg_sync(wg);

        // ---- h1 = tanh(W1 x0 + b1): two N = 128 halves, A from shared memory
        #pragma unroll 1
        for (int half = 0; half < 2; half++) {
            if (wt == 0) {
                fence_after_sync();
                const uint64_t da = umma_desc(x0_addr, 2048, 128), db = umma_desc(w_addr + (W1_OFF + half * 128 * 8) * 2, W1_N * 16, 128);
                #pragma unroll
                for (int s = 0; s < K1 / 16; s++)
                    mma_ss(tD0, desc_advance(da, s * 2 * 2048), desc_advance(db, s * 2 * W1_N * 16), umma_idesc(128), true);
                mma_commit(bar);
            }
            stamp();
            ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
            stamp();
            fence_after_sync();
            epilogue_to_tmem(tD, tA + half * 64, half == 0 ? s_c + C_B1 + 128 : s_c + C_B2, 128, wg);
            turn_pass(wg);
            fence_before_sync();
            wg_sync(wg);
            stamp();
        }
        // ---- h2 = tanh(W2 h1 + b2): K = 256, A from TMEM (operand region columns 0..127)
        if (wt == 0) {
            fence_after_sync();
            const uint64_t db = umma_desc(w_addr + W2_OFF * 2, W2_N * 16, 128);
            #pragma unroll 4
            for (int s = 0; s < 256 / 16; s++) mma_ts(tD0, tA0 + s * 8, desc_advance(db, s * 2 * W2_N * 16), umma_idesc(128), true);
            mma_commit(bar);
        }
        stamp();
        ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
        stamp();
        fence_after_sync();
        epilogue_to_tmem(tD, tA, s_c + C_B3, 128, wg);                       // h2 -> columns 0..63
        turn_pass(wg);
        fence_before_sync();
        wg_sync(wg);
        stamp();
        // ---- l1 = tanh(W3' h2 + b3') -> columns 64..127
        if (wt == 0) {
            fence_after_sync();
            const uint64_t db = umma_desc(w_addr + W3_OFF * 2, W3_N * 16, 128);
            #pragma unroll 4
            for (int s = 0; s < 128 / 16; s++) mma_ts(tD0, tA0 + s * 8, desc_advance(db, s * 2 * W3_N * 16), umma_idesc(128), true);
            mma_commit(bar);
        }
        ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
        fence_after_sync();
        epilogue_to_tmem(tD, tA + 64, s_b4, 32, wg);
        turn_pass(wg);
        fence_before_sync();
        wg_sync(wg);
        // ---- logits = W4 l1 + b4 (N = 16, 8 real)
        if (wt == 0) {
            fence_after_sync();
            const uint64_t db = umma_desc(w_addr + W4_OFF * 2, W4_N * 16, 128);
            #pragma unroll 4
            for (int s = 0; s < 128 / 16; s++) mma_ts(tD0, tA0 + 64 + s * 8, desc_advance(db, s * 2 * W4_N * 16), umma_idesc(16), true);
            mma_commit(bar);
        }
        ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
        fence_after_sync();
        {
            float v[16];
            tmem_ld16(tD, v);
            float x[8];
            #pragma unroll
            for (int k = 0; k < 8; k++) x[k] = live ? v[k] : 0.0f;          // (b4 was in the accumulator)
            #pragma unroll 1
            for (int c = 0; c < 128; c += 32) bias_to_tmem32(tD + c, s_c + C_B3 + 128 + c);   // value-branch hidden layer comes next
            tmem_st_wait();
            if (live && p.logits) {
                float4 *out = reinterpret_cast<float4 *>(p.logits + (size_t)row * 8);
                out[0] = make_float4(x[0], x[1], x[2], x[3]);
                out[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
            if (p.actions) {     // warp-uniform; every lane enters (rows past n sample from zero logits and store nothing)
                float act[4], lp;
                dsim::beta_row<float, 4>(x, p.seed, p.env_base + (uint32_t)row, step_now, p.deterministic, act, lp);
                if (live) {
                    *reinterpret_cast<float4 *>(p.actions + (size_t)row * 4) = make_float4(act[0], act[1], act[2], act[3]);
                    if (p.logp) p.logp[row] = lp;
                }
            }
        }
        fence_before_sync();
        wg_sync(wg);
        // ---- v1 = tanh(V1' h2 + c1') -> columns 64..127 (l1 is dead)
        if (wt == 0) {
            fence_after_sync();
            const uint64_t db = umma_desc(w_addr + (W3_OFF + 128 * 8) * 2, W3_N * 16, 128);
            #pragma unroll 4
            for (int s = 0; s < 128 / 16; s++) mma_ts(tD0, tA0 + s * 8, desc_advance(db, s * 2 * W3_N * 16), umma_idesc(128), true);
            mma_commit(bar);
        }
        ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
        fence_after_sync();
        epilogue_to_tmem(tD, tA + 64, s_c + C_C2, 128, wg);
        turn_pass(wg);
        fence_before_sync();
        wg_sync(wg);
        // ---- v2 = tanh(V2 v1 + c2); value = V3 v2 + c3
        if (wt == 0) {
            fence_after_sync();
            const uint64_t db = umma_desc(w_addr + V2_OFF * 2, V2_N * 16, 128);
            #pragma unroll 4
            for (int s = 0; s < 128 / 16; s++) mma_ts(tD0, tA0 + 64 + s * 8, desc_advance(db, s * 2 * V2_N * 16), umma_idesc(128), true);
            mma_commit(bar);
        }
        ok = mbar_wait_bounded(bar, parity) && ok; parity ^= 1;
        fence_after_sync();
        turn_wait(wg);
        {
            float val = s_c[C_C3];
            #pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
                float v[32];
                tmem_ld32(tD + c, v);
                bias_to_tmem32(tD + c, s_c + C_B1 + c);                      // the next tile's first layer
                #pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float lo = v[j], hi = v[j + 1];
                    if (((j >> 1) & 7) < MLP_POLY) tanh_poly2(lo, hi);
                    else {
                        const uint32_t t2 = tanh_pack(lo, hi);
                        lo = __uint_as_float(t2 << 16); hi = __uint_as_float(t2 & 0xFFFF0000u);
                    }
                    val = fmaf(s_c[C_V3 + c + j], lo, val);
                    val = fmaf(s_c[C_V3 + c + j + 1], hi, val);
                }
            }
            if (live) p.value[row] = val;
            tmem_st_wait();
        }
        turn_pass(wg);
        fence_before_sync();
        wg_sync(wg);                                                         // D and the operand region are free for the next tile
    }
    {   // the two warpgroups may own different numbers of tiles: keep passing the turn for the other one's remaining epilogues
        const int first = 2 * blockIdx.x, stride = 2 * gridDim.x;
        const int mine = (first + wg < p.ntiles) ? (p.ntiles - 1 - (first + wg)) / stride + 1 : 0;
        const int other = (first + (wg ^ 1) < p.ntiles) ? (p.ntiles - 1 - (first + (wg ^ 1))) / stride + 1 : 0;
        // turns alternate 0,1,0,1,...: after my last epilogue the other warpgroup still needs (its epilogues - mine) grants, minus
        // the one warpgroup 1 handed out before the loop
        int extra = (other - mine) * kTurnsPerTile;
        for (int k = 0; k < extra; k++) { turn_wait(wg); turn_pass(wg); }
        if (wg == 0) turn_wait(0);                                         // balance the grant warpgroup 1 handed out before the loop
    }
    if (!ok) atomicExch(p.error, 1);
    fence_before_sync();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}


// =====================================================================================================================
// v2 schedule (DSIM_MLP_V2=1): every epilogue is shared by ALL eight compute warps (two per scheduler: warp w and w + 4 read
// the same tensor-memory lane quarter and split the 128 columns in halves), and a ninth warp issues every MMA.
// Why: within one warp the accumulator read (tcgen05.ld occupies the warp, 190 cycles per 32 columns) and the SFU work
// (256 cycles per 32 columns) ADD; with two warps per scheduler one reads while the other computes.  The epilogue of tile A
// runs while the tensor core works on tile B's layer and vice versa (the MMA warp is released by an mbarrier that all 256
// compute threads arrive on), so no turn-taking is needed.
constexpr int V2_THREADS = 288;
constexpr int SMEM_BYTES_V2 = SMEM_BYTES + 64 + 2 * 2 * 128 * 4;            // + 4 mbarriers + value partial sums [2 tiles][2 halves][128]

DEV void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// first-layer operand row [s(16), a_prev(4), z(8), 0(4)] -> bf16, canonical smem tile (the parameter encoder runs on the CUDA cores)
DEV void build_x0_row(const float *s_c, const float (&o)[OBS_DIM], const float (&a)[A_DIM], unsigned char *x0_tile, int r) {
    float hdn[ENC_H];
    #pragma unroll
    for (int j = 0; j < ENC_H; j++) {
        float acc = s_c[C_E1B + j];
        #pragma unroll
        for (int k = 0; k < P_DIM; k++) acc = fmaf(s_c[C_E1 + j * P_DIM + k], o[S_DIM + k], acc);
        hdn[j] = tanh_fast(acc);
    }
    float x[K1];
    #pragma unroll
    for (int k = 0; k < S_DIM; k++) x[k] = o[k];
    #pragma unroll
    for (int k = 0; k < A_DIM; k++) x[S_DIM + k] = a[k];
    #pragma unroll
    for (int e = 0; e < E_DIM; e++) {
        float acc = s_c[C_E2B + e];
        #pragma unroll
        for (int j = 0; j < ENC_H; j++) acc = fmaf(s_c[C_E2 + e * ENC_H + j], hdn[j], acc);
        x[S_DIM + A_DIM + e] = acc;
    }
    #pragma unroll
    for (int k = S_DIM + A_DIM + E_DIM; k < K1; k++) x[k] = 0.f;
    #pragma unroll
    for (int kc = 0; kc < K1 / 8; kc++) {                                    // element (row, k) at (k/8) * 2048 + row * 16 + (k%8) * 2
        uint4 q;
        q.x = pack_bf16(x[8 * kc + 0], x[8 * kc + 1]); q.y = pack_bf16(x[8 * kc + 2], x[8 * kc + 3]);
        q.z = pack_bf16(x[8 * kc + 4], x[8 * kc + 5]); q.w = pack_bf16(x[8 * kc + 6], x[8 * kc + 7]);
        *reinterpret_cast<uint4 *>(x0_tile + kc * 2048 + r * 16) = q;
    }
}

// this thread's 64 accumulator columns [d, d + 64) -> tanh -> bf16 -> 32 operand columns at dst; nb0 / nb1: the next layer's
// bias for the two 32-column chunks (nullptr: leave the columns alone)
__device__ __noinline__ void epi_half(uint32_t d, uint32_t dst, const float *nb0, const float *nb1) {
    uint32_t ra[32], rb[32];
    tmem_ld32_issue(d, ra);
    tmem_ld32_wait(ra);
    tmem_ld32_issue(d + 32, rb);
    __syncwarp();
    if (nb0) bias_to_tmem32(d, nb0);
    uint32_t pk[16];
    #pragma unroll
    for (int j = 0; j < 16; j++) {
        const float lo = __uint_as_float(ra[2 * j]), hi = __uint_as_float(ra[2 * j + 1]);
        pk[j] = (j & 7) < MLP_POLY ? tanh_poly_pack(lo, hi) : tanh_pack(lo, hi);
    }
    tmem_st16(dst, pk);
    tmem_ld32_wait(rb);
    if (nb1) bias_to_tmem32(d + 32, nb1);
    #pragma unroll
    for (int j = 0; j < 16; j++) {
        const float lo = __uint_as_float(rb[2 * j]), hi = __uint_as_float(rb[2 * j + 1]);
        pk[j] = (j & 7) < MLP_POLY ? tanh_poly_pack(lo, hi) : tanh_pack(lo, hi);
    }
    tmem_st16(dst + 16, pk);
    tmem_st_wait();
}

__global__ void __launch_bounds__(V2_THREADS, 1) rma_full_forward_kernel_v2(const MlpParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint16_t *s_w = reinterpret_cast<uint16_t *>(smem);
    float *s_c = reinterpret_cast<float *>(smem + W_ELEMS * 2);
    unsigned char *s_x0 = smem + W_ELEMS * 2 + C_ELEMS * 4;                  // [2 tiles][8 KB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_x0 + 2 * X0_BYTES);      // [0..1] MMA done (tile), [2..3] epilogue done (tile)
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 4);
    float *s_b4 = reinterpret_cast<float *>(s_tmem + 8);                      // 32 floats
    float *s_val = s_b4 + 32;                                                 // [2][2][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma = warp == 8;
    const int half = (warp >> 2) & 1, wq = warp & 3, r128 = tid & 127, my_tile = (tid >> 7) & 1;

    {
        const uint4 *gw = reinterpret_cast<const uint4 *>(p.w);
        uint4 *sw = reinterpret_cast<uint4 *>(s_w);
        for (int i = tid; i < W_ELEMS * 2 / 16; i += V2_THREADS) sw[i] = __ldg(gw + i);
        for (int i = tid; i < C_ELEMS; i += V2_THREADS) s_c[i] = __ldg(p.c + i);
        if (tid < 32) s_b4[tid] = tid < 16 ? __ldg(p.c + C_B4 + tid) : 0.0f;
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_init(&s_bar[2], 256); mbar_init(&s_bar[3], 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t w_addr = smem_u32(s_w);
    const int npairs = (p.ntiles + 1) / 2;
    bool ok = true;

    if (is_mma) {
        if (lane == 0) {
            uint32_t dpar[2] = {0u, 0u};
            #pragma unroll 1
            for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
                #pragma unroll 1
                for (int st = 0; st < 7; st++) {
                    #pragma unroll 1
                    for (int T = 0; T < 2; T++) {
                        ok = mbar_wait_bounded(&s_bar[2 + T], dpar[T]) && ok; dpar[T] ^= 1u;
                        fence_after_sync();
                        const uint32_t d0 = tmem_base + T * 256, a0 = d0 + 128;
                        if (st < 2) {
                            const uint64_t da = umma_desc(smem_u32(s_x0 + T * X0_BYTES), 2048, 128);
                            const uint64_t db = umma_desc(w_addr + (W1_OFF + st * 128 * 8) * 2, W1_N * 16, 128);
                            #pragma unroll
                            for (int k = 0; k < K1 / 16; k++) mma_ss(d0, desc_advance(da, k * 2 * 2048), desc_advance(db, k * 2 * W1_N * 16), umma_idesc(128), true);
                        } else if (st == 2) {
                            const uint64_t db = umma_desc(w_addr + W2_OFF * 2, W2_N * 16, 128);
                            #pragma unroll 4
                            for (int k = 0; k < 256 / 16; k++) mma_ts(d0, a0 + k * 8, desc_advance(db, k * 2 * W2_N * 16), umma_idesc(128), true);
                        } else if (st == 3 || st == 5) {
                            const uint64_t db = umma_desc(w_addr + (W3_OFF + (st == 5 ? 128 * 8 : 0)) * 2, W3_N * 16, 128);
                            #pragma unroll 4
                            for (int k = 0; k < 128 / 16; k++) mma_ts(d0, a0 + k * 8, desc_advance(db, k * 2 * W3_N * 16), umma_idesc(128), true);
                        } else if (st == 4) {
                            const uint64_t db = umma_desc(w_addr + W4_OFF * 2, W4_N * 16, 128);
                            #pragma unroll 4
                            for (int k = 0; k < 128 / 16; k++) mma_ts(d0, a0 + 64 + k * 8, desc_advance(db, k * 2 * W4_N * 16), umma_idesc(16), true);
                        } else {
                            const uint64_t db = umma_desc(w_addr + V2_OFF * 2, V2_N * 16, 128);
                            #pragma unroll 4
                            for (int k = 0; k < 128 / 16; k++) mma_ts(d0, a0 + 64 + k * 8, desc_advance(db, k * 2 * V2_N * 16), umma_idesc(128), true);
                        }
                        mma_commit(&s_bar[T]);
                    }
                }
            }
        }
    } else {
        const uint32_t lane_off = (uint32_t)(32 * wq) << 16;
        const uint32_t step_now = p.step + (p.step_dev ? __ldg(p.step_dev) : 0u);
        float nxt_o[OBS_DIM], nxt_a[A_DIM];
        auto load_row = [&](int pair_) {
            const int tile_ = 2 * pair_ + my_tile, r_ = tile_ * 128 + r128;
            const bool live_ = pair_ < npairs && tile_ < p.ntiles && r_ < p.n;
            #pragma unroll
            for (int k = 0; k < OBS_DIM; k += 2) {
                const float2 v = live_ ? __ldg(reinterpret_cast<const float2 *>(p.obs + (size_t)r_ * OBS_DIM + k)) : make_float2(0.f, 0.f);
                nxt_o[k] = v.x; nxt_o[k + 1] = v.y;
            }
            const bool fresh = live_ && p.reset_mask && p.reset_mask[r_];
            const float4 v = (live_ && !fresh) ? __ldg(reinterpret_cast<const float4 *>(p.prev_action) + r_) : make_float4(0.f, 0.f, 0.f, 0.f);
            nxt_a[0] = v.x; nxt_a[1] = v.y; nxt_a[2] = v.z; nxt_a[3] = v.w;
        };
        auto finish = [&](int T) { fence_before_sync(); mbar_arrive(&s_bar[2 + T]); };
        // prologue: first layer's bias into both accumulators (my 64 columns), first pair's operand rows
        #pragma unroll 1
        for (int T = 0; T < 2; T++) {
            const uint32_t d = tmem_base + T * 256 + lane_off + 64 * half;
            bias_to_tmem32(d, s_c + C_B1 + 64 * half); bias_to_tmem32(d + 32, s_c + C_B1 + 64 * half + 32);
        }
        tmem_st_wait();
        load_row(blockIdx.x);
        build_x0_row(s_c, nxt_o, nxt_a, s_x0 + my_tile * X0_BYTES, r128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        finish(0); finish(1);
        uint32_t bpar[2] = {0u, 0u};
        #pragma unroll 1
        for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
            load_row(pair + (int)gridDim.x);                                 // next pair's rows: in flight during the first two stages
            #pragma unroll 1
            for (int st = 0; st < 7; st++) {
                #pragma unroll 1
                for (int T = 0; T < 2; T++) {
                    ok = mbar_wait_bounded(&s_bar[T], bpar[T]) && ok; bpar[T] ^= 1u;
                    fence_after_sync();
                    const uint32_t tD = tmem_base + T * 256 + lane_off, tA = tD + 128;
                    const uint32_t d = tD + 64 * half;
                    const int row = (2 * pair + T) * 128 + r128;
                    const bool live = row < p.n;
                    if (st == 4) {                                           // logits: 16 FP32 columns, rows handled by the half-0 warps
                        if (half == 0) {
                            float v[16];
                            tmem_ld16(tD, v);
                            float x[8];
                            #pragma unroll
                            for (int k = 0; k < 8; k++) x[k] = live ? v[k] : 0.0f;
                            if (live && p.logits) {
                                float4 *out = reinterpret_cast<float4 *>(p.logits + (size_t)row * 8);
                                out[0] = make_float4(x[0], x[1], x[2], x[3]);
                                out[1] = make_float4(x[4], x[5], x[6], x[7]);
                            }
                            if (p.actions) {
                                float act[4], lp;
                                dsim::beta_row<float, 4>(x, p.seed, p.env_base + (uint32_t)row, step_now, p.deterministic, act, lp);
                                if (live) {
                                    *reinterpret_cast<float4 *>(p.actions + (size_t)row * 4) = make_float4(act[0], act[1], act[2], act[3]);
                                    if (p.logp) p.logp[row] = lp;
                                }
                            }
                        }
                        bias_to_tmem32(d, s_c + C_B3 + 128 + 64 * half); bias_to_tmem32(d + 32, s_c + C_B3 + 128 + 64 * half + 32);
                        tmem_st_wait();
                    } else if (st == 6) {                                    // value head: partial dot product over my 64 columns
                        float val = 0.f;
                        #pragma unroll 1
                        for (int c = 0; c < 64; c += 32) {
                            float v[32];
                            tmem_ld32(d + c, v);
                            bias_to_tmem32(d + c, s_c + C_B1 + 64 * half + c);   // the next pair's first layer
                            #pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float lo = v[j], hi = v[j + 1];
                                if (((j >> 1) & 7) < MLP_POLY) tanh_poly2(lo, hi);
                                else {
                                    const uint32_t t2 = tanh_pack(lo, hi);
                                    lo = __uint_as_float(t2 << 16); hi = __uint_as_float(t2 & 0xFFFF0000u);
                                }
                                val = fmaf(s_c[C_V3 + 64 * half + c + j], lo, val);
                                val = fmaf(s_c[C_V3 + 64 * half + c + j + 1], hi, val);
                            }
                        }
                        tmem_st_wait();
                        s_val[(T * 2 + half) * 128 + r128] = val;
                        asm volatile("bar.sync 5, 256;" ::: "memory");
                        if (half == 0 && live) p.value[row] = s_c[C_C3] + s_val[(T * 2) * 128 + r128] + s_val[(T * 2 + 1) * 128 + r128];
                    } else {
                        const float *nb = st == 0 ? s_c + C_B1 + 128 : st == 1 ? s_c + C_B2 : st == 2 ? s_c + C_B3 : st == 5 ? s_c + C_C2 : nullptr;
                        const float *nb0 = nb ? nb + 64 * half : (half == 0 ? s_b4 : nullptr);     // st == 3: FP32 b4 into columns 0..31
                        const float *nb1 = nb ? nb + 64 * half + 32 : nullptr;
                        const uint32_t dst = tA + ((st == 1 || st == 3 || st == 5) ? 64 : 0) + 32 * half;
                        epi_half(d, dst, nb0, nb1);
                    }
                    finish(T);
                    if (st == 1 && T == 1) {                                 // both first-layer MMAs of this pair are done: the operand tiles are free
                        build_x0_row(s_c, nxt_o, nxt_a, s_x0 + my_tile * X0_BYTES, r128);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                }
            }
        }
    }
    if (!ok) atomicExch(p.error, 1);
    fence_before_sync();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

struct DsimPolicy {
    int device, sms, v2;
    uint16_t *w;
    float *c;
    int *error;
    long long *dbg;
    char err[256];
};

extern "C" int dsim_policy_blob_sizes(int64_t *weight_elems, int64_t *const_elems) {
    if (weight_elems) *weight_elems = W_ELEMS;
    if (const_elems) *const_elems = C_ELEMS;
    return DSIM_OK;
}

extern "C" int dsim_policy_create(int device, const uint16_t *weights_host, const float *consts_host, DsimPolicy **out) {
    if (!out || !weights_host || !consts_host) return DSIM_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DSIM_ECUDA;   // no CPU fallback
    DsimPolicy *h = new (std::nothrow) DsimPolicy();
    if (!h) return DSIM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->w, (size_t)W_ELEMS * 2);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->c, (size_t)C_ELEMS * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->error, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->error, 0, sizeof(int));
    if (e == cudaSuccess && getenv("DSIM_MLP_DEBUG")) { e = cudaMalloc((void **)&h->dbg, 64 * sizeof(long long)); if (e == cudaSuccess) e = cudaMemset(h->dbg, 0, 64 * sizeof(long long)); }
    if (e == cudaSuccess) e = cudaMemcpy(h->w, weights_host, (size_t)W_ELEMS * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->c, consts_host, (size_t)C_ELEMS * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rma_full_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rma_full_forward_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_V2);
    h->v2 = getenv("DSIM_MLP_V2") && atoi(getenv("DSIM_MLP_V2")) != 0;
    if (e != cudaSuccess) {
        if (h->w) cudaFree(h->w);
        if (h->c) cudaFree(h->c);
        if (h->error) cudaFree(h->error);
        delete h;
        return DSIM_ECUDA;
    }
    *out = h;
    return DSIM_OK;
}

extern "C" void dsim_policy_destroy(DsimPolicy *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->w); cudaFree(h->c); cudaFree(h->error);
    delete h;
}

static int launch_policy(DsimPolicy *h, MlpParams &p, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                         void *stream) {
    if (((uintptr_t)obs_dev & 7) || ((uintptr_t)prev_action_dev & 15) || ((uintptr_t)p.logits & 15) || ((uintptr_t)p.actions & 15)) return DSIM_EINVAL;
    if (cudaSetDevice(h->device) != cudaSuccess) return DSIM_ECUDA;
    p.w = h->w; p.c = h->c; p.obs = obs_dev; p.prev_action = prev_action_dev; p.reset_mask = reset_mask_dev; p.error = h->error;
    p.n = n; p.ntiles = (n + 127) / 128; p.dbg = h->dbg;
    const int pairs = (p.ntiles + 1) / 2;
    const int grid = pairs < h->sms ? pairs : h->sms;
    if (h->v2) {
        rma_full_forward_kernel_v2<<<grid, V2_THREADS, SMEM_BYTES_V2, (cudaStream_t)stream>>>(p);
        return cudaGetLastError() == cudaSuccess ? DSIM_OK : DSIM_ECUDA;
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(grid); lc.blockDim = dim3(256); lc.dynamicSmemBytes = SMEM_BYTES; lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // pairs with griddepcontrol.* in the kernel
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    void *args[] = {&p};
    return cudaLaunchKernelExC(&lc, (const void *)rma_full_forward_kernel, args) == cudaSuccess ? DSIM_OK : DSIM_ECUDA;
}

extern "C" int dsim_policy_forward(DsimPolicy *h, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                                   float *logits_dev, float *value_dev, void *stream) {
    if (!h || !obs_dev || !prev_action_dev || !logits_dev || !value_dev || n <= 0) return DSIM_EINVAL;
    MlpParams p{};
    p.logits = logits_dev; p.value = value_dev;
    return launch_policy(h, p, obs_dev, prev_action_dev, reset_mask_dev, n, stream);
}

// forward + Beta-head sampling in one launch: the logits never leave the SM unless `logits_dev` is given.  `actions_dev` may
// alias `prev_action_dev` (each row is read by the thread that later writes it).  Same Philox streams as dsim_beta_policy.
extern "C" int dsim_policy_forward_sample(DsimPolicy *h, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                                          uint32_t seed, uint32_t env_id_offset, uint32_t step, const uint32_t *step_dev, int deterministic,
                                          float *logits_dev, float *value_dev, float *actions_dev, float *logp_dev, void *stream) {
    if (!h || !obs_dev || !prev_action_dev || !value_dev || !actions_dev || n <= 0) return DSIM_EINVAL;
    MlpParams p{};
    p.logits = logits_dev; p.value = value_dev; p.actions = actions_dev; p.logp = logp_dev;
    p.seed = seed; p.env_base = env_id_offset; p.step = step; p.step_dev = step_dev; p.deterministic = deterministic;
    return launch_policy(h, p, obs_dev, prev_action_dev, reset_mask_dev, n, stream);
}

extern "C" int dsim_policy_debug(DsimPolicy *h, long long *out64) {
    if (!h || !h->dbg || !out64) return DSIM_EINVAL;
    cudaDeviceSynchronize();
    return cudaMemcpy(out64, h->dbg, 64 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? DSIM_OK : DSIM_ECUDA;
}

// 1 if any launch since creation hit a tensor-core barrier timeout (results of that launch are invalid); syncs the device
extern "C" int dsim_policy_error(DsimPolicy *h) {
    if (!h) return DSIM_EINVAL;
    int v = 0;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return DSIM_ECUDA;
    if (cudaMemcpy(&v, h->error, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return DSIM_ECUDA;
    return v;
}
