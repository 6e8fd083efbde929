// dsim_policy_fp32.cu — RMA_full inference (models/PPO/RMA/RMA_model.py:48-71, 79-116, train_adaptation=False) at the
// REFERENCE's precision: FP32 operands, FP32 accumulation, libm-accurate tanh.  The reference's policy is an FP32 torch
// module; the tcgen05 kernel (dsim_policy_mlp.cu) trades precision for speed (bf16 operands, ~1e-2 on the logits).  An
// FP32-faithful tensor-core variant does not fit the fused design: bf16 hi + lo operand splitting (3 MMAs per product)
// needs both halves of every weight matrix in shared memory, 2 x 180 KB > 227 KB, and kind::tf32 carries 10 mantissa bits
// (~1e-3 on the logits, and its FP32-sized weights do not fit either).  This kernel is the faithful mode: the whole network
// fused on the FP32 pipe, activations in shared memory, weights streamed through L1 from their 363 KB L2-resident blob.
//
//   tile = 128 rows per CTA pass, 512 threads = 16 warps: warp w owns rows 8 w .. +7 and ALL output columns (every A operand
//   is a warp-wide broadcast read of the transposed activation tile At[k][row]; lane l owns columns l, l+32, ..., so every W
//   operand is a coalesced 128-byte line of the transposed weights Wt[k][n], shared by the sixteen warps through L1):
//   32 packed FFMA2 (64 FMAs) per 2 LDS.128 + 8 LDG at N = 256, 16 per 2 + 4 at N = 128 (the 64-row tile with the columns
//   split over two warp groups and scalar FFMAs: 3.06 ms per 524 288 rows; 128-row tile 2.77 ms; FFMA2 2.31 ms).  Layers ping-pong between a 256-deep and
//   a 128-deep activation tile (203 KB); BatchNorm (eval) is folded into its two consumers on the host.  tanh(x) = 1 - 2 / (exp(2x) + 1) on the MUFU exp2 / rcp (abs. error
//   ~2e-7: the logits stay within 1e-6 of the torch FP32 module, tests/test_policy_reference.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "../../include/dronesim_b200.h"

namespace {

constexpr int S_DIM = 16, P_DIM = 6, A_DIM = 4, E_DIM = 8, OBS_DIM = S_DIM + P_DIM, ENC_H = 32, K1 = S_DIM + A_DIM + E_DIM;   // 28
#ifndef DSIM_P32_TM
#define DSIM_P32_TM 128
#endif
constexpr int TM = DSIM_P32_TM, LDT = TM + 4;      // rows per tile; row pitch of the transposed tiles (132 = 4 mod 32: conflict-free 16-byte stores)
// fp32 blob (float offsets): transposed weights Wt[k][n] then bias, per layer
constexpr int F_W1 = 0, F_B1 = F_W1 + K1 * 256;                 // 28 -> 256
constexpr int F_W2 = F_B1 + 256, F_B2 = F_W2 + 256 * 128;       // 256 -> 128
constexpr int F_W3 = F_B2 + 128, F_B3 = F_W3 + 128 * 256;       // 128 -> 256: columns 0-127 logits hidden (BN folded), 128-255 value hidden
constexpr int F_V2 = F_B3 + 256, F_C2 = F_V2 + 128 * 128;       // 128 -> 128 (value branch)
constexpr int F_W4 = F_C2 + 128, F_B4 = F_W4 + 128 * 8;         // 128 -> 8, stored [k][8]
constexpr int F_V3 = F_B4 + 8, F_C3 = F_V3 + 128;               // 128 -> 1
constexpr int F_E1 = F_C3 + 1, F_E1B = F_E1 + ENC_H * P_DIM;    // encoder 6 -> 32 ([j][k])
constexpr int F_E2 = F_E1B + ENC_H, F_E2B = F_E2 + E_DIM * ENC_H;   // 32 -> 8 ([e][j])
constexpr int F_ELEMS = F_E2B + E_DIM;
constexpr int SMEM_BYTES = (256 + 128) * LDT * 4;               // activation tiles A [256][132] and B [128][132] fp32 = 202 752 B

struct P32 {
    const float *w, *obs, *prev_action;
    const unsigned char *reset_mask;
    float *logits, *value;
    int n, ntiles;
};

constexpr int NT = TM * 4;                 // threads per CTA: one warp per 8 rows
constexpr int kCtasPerSm = TM == 128 ? 1 : 2;
__device__ __forceinline__ float tanh_acc(float x) {
    // 1 - 2 / (e^{2x} + 1): exact limits for |x| large (e -> inf: 1; e -> 0: -1), no cancellation near 0 beyond 1 ulp of 1
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
// Blackwell's packed FP32 FMA: two IEEE fma.rn per lane and instruction (SASS FFMA2).  A three-register FFMA issues every other
// cycle per scheduler (register read ports), so a scalar-FFMA kernel tops out at half the FP32 peak (measured: FMA pipe 50 %
// active, 34 TFLOP/s, with eligible warps held at dispatch); FFMA2 moves 64-bit operands through the same ports.  Results are
// bit-identical to fmaf.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64 &d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
// Ot[n][row] = act(bias[n] + sum_k Wt[k][n] * At[k][row]) for the 128 rows of the tile; N a multiple of 32
template <int K, int N, bool TANH>
__device__ __forceinline__ void dense(const float *__restrict__ Wt, const float *__restrict__ bias, const float *At, float *Ot) {
    constexpr int CN = N / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    u64 acc[4][CN];                                                           // [row pair][column]: rows (2 rp, 2 rp + 1) of this warp's eight
    #pragma unroll
    for (int j = 0; j < CN; j++) {
        const float b = __ldg(bias + tx + 32 * j);
        #pragma unroll
        for (int r = 0; r < 4; r++) acc[r][j] = pack2(b, b);
    }
    // software pipeline over k in blocks of KB: the W and A operands of block b + 1 are fetched while block b is multiplied
    // (an unpipelined loop sat 28 % of its time on the first FFMA of every block waiting for the L1 / L2 round trip)
    constexpr int KB = CN >= 8 ? 1 : 2;                                       // 64 accumulators leave room for one operand block in flight
    static_assert(K % KB == 0, "K must be a multiple of the pipeline block");
    const float *wq = Wt + tx;
    const float *ap = At + ty * 8;
    float w[2][KB][CN];
    float4 a[2][KB][2];
    #pragma unroll
    for (int u = 0; u < KB; u++) {
        #pragma unroll
        for (int j = 0; j < CN; j++) w[0][u][j] = __ldg(wq + u * N + 32 * j);
        a[0][u][0] = *reinterpret_cast<const float4 *>(ap + u * LDT); a[0][u][1] = *reinterpret_cast<const float4 *>(ap + u * LDT + 4);
    }
    #pragma unroll 2
    for (int kb = 0; kb < K / KB; kb++) {
        const int cur = kb & 1, nxt = cur ^ 1;
        if (kb + 1 < K / KB) {
            const float *wn = wq + (size_t)(kb + 1) * KB * N;
            const float *an = ap + (kb + 1) * KB * LDT;
            #pragma unroll
            for (int u = 0; u < KB; u++) {
                #pragma unroll
                for (int j = 0; j < CN; j++) w[nxt][u][j] = __ldg(wn + u * N + 32 * j);
                a[nxt][u][0] = *reinterpret_cast<const float4 *>(an + u * LDT); a[nxt][u][1] = *reinterpret_cast<const float4 *>(an + u * LDT + 4);
            }
        }
        #pragma unroll
        for (int u = 0; u < KB; u++) {
            // row pairs come packed out of the 16-byte activation loads; the weight of a column is duplicated into both halves
            const u64 av[4] = {pack2(a[cur][u][0].x, a[cur][u][0].y), pack2(a[cur][u][0].z, a[cur][u][0].w),
                               pack2(a[cur][u][1].x, a[cur][u][1].y), pack2(a[cur][u][1].z, a[cur][u][1].w)};
            #pragma unroll
            for (int j = 0; j < CN; j++) {
                const u64 ww = pack2(w[cur][u][j], w[cur][u][j]);
                #pragma unroll
                for (int r = 0; r < 4; r++) ffma2(acc[r][j], av[r], ww);
            }
        }
    }
    #pragma unroll
    for (int j = 0; j < CN; j++) {
        float v[8];
        #pragma unroll
        for (int r = 0; r < 4; r++) unpack2(acc[r][j], v[2 * r], v[2 * r + 1]);
        if (TANH) {
            #pragma unroll
            for (int r = 0; r < 8; r++) v[r] = tanh_acc(v[r]);
        }
        float *o = Ot + (tx + 32 * j) * LDT + ty * 8;
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__global__ void __launch_bounds__(NT, kCtasPerSm) rma_full_forward_fp32_kernel(const P32 p) {
    extern __shared__ __align__(16) float sm[];
    float *A = sm, *B = sm + 256 * LDT;                              // A: up to 256 activation rows deep, B: up to 128
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        // ---- x0 = [s(16), a_prev(4), z(8)] with z = E2 tanh(E1 e + e1) + e2 (RMA_model.py:94-104), one thread per row -> B[k][row]
        if (tid < TM) {
            const int row = tile * TM + tid;
            const bool live = row < p.n;
            float o[OBS_DIM], a[A_DIM];
            #pragma unroll
            for (int k = 0; k < OBS_DIM; k++) o[k] = live ? __ldg(p.obs + (size_t)row * OBS_DIM + k) : 0.f;
            const bool fresh = live && p.reset_mask && p.reset_mask[row];
            #pragma unroll
            for (int k = 0; k < A_DIM; k++) a[k] = (live && !fresh) ? __ldg(p.prev_action + (size_t)row * A_DIM + k) : 0.f;
            float hdn[ENC_H];
            #pragma unroll
            for (int j = 0; j < ENC_H; j++) {
                float acc = __ldg(p.w + F_E1B + j);
                #pragma unroll
                for (int k = 0; k < P_DIM; k++) acc = fmaf(__ldg(p.w + F_E1 + j * P_DIM + k), o[S_DIM + k], acc);
                hdn[j] = tanh_acc(acc);
            }
            #pragma unroll
            for (int k = 0; k < S_DIM; k++) B[k * LDT + tid] = o[k];
            #pragma unroll
            for (int k = 0; k < A_DIM; k++) B[(S_DIM + k) * LDT + tid] = a[k];
            #pragma unroll
            for (int e = 0; e < E_DIM; e++) {
                float acc = __ldg(p.w + F_E2B + e);
                #pragma unroll
                for (int j = 0; j < ENC_H; j++) acc = fmaf(__ldg(p.w + F_E2 + e * ENC_H + j), hdn[j], acc);
                B[(S_DIM + A_DIM + e) * LDT + tid] = acc;
            }
        }
        __syncthreads();
        dense<K1, 256, true>(p.w + F_W1, p.w + F_B1, B, A);          // h1 -> A
        __syncthreads();
        dense<256, 128, true>(p.w + F_W2, p.w + F_B2, A, B);         // h2 -> B (BatchNorm folded into the consumers)
        __syncthreads();
        dense<128, 256, true>(p.w + F_W3, p.w + F_B3, B, A);         // [l1 | v1] -> A
        __syncthreads();
        dense<128, 128, true>(p.w + F_V2, p.w + F_C2, A + 128 * LDT, B);   // v2 = tanh(V2 v1 + c2) -> B
        // logits = W4 l1 + b4: 128 rows x 8 outputs, thread -> (row, output pair)
        {
            const int row = tid & (TM - 1), o2 = (tid / TM) * 2;
            float x0 = __ldg(p.w + F_B4 + o2), x1 = __ldg(p.w + F_B4 + o2 + 1);
            #pragma unroll 8
            for (int k = 0; k < 128; k++) {
                const float l = A[k * LDT + row];
                x0 = fmaf(__ldg(p.w + F_W4 + k * 8 + o2), l, x0);
                x1 = fmaf(__ldg(p.w + F_W4 + k * 8 + o2 + 1), l, x1);
            }
            const int r = tile * TM + row;
            if (r < p.n) *reinterpret_cast<float2 *>(p.logits + (size_t)r * 8 + o2) = make_float2(x0, x1);
        }
        __syncthreads();
        if (tid < TM) {                                              // value = V3 v2 + c3
            float v = __ldg(p.w + F_C3);
            #pragma unroll 8
            for (int k = 0; k < 128; k++) v = fmaf(__ldg(p.w + F_V3 + k), B[k * LDT + tid], v);
            const int r = tile * TM + tid;
            if (r < p.n) p.value[r] = v;
        }
        __syncthreads();
    }
}

}  // namespace

struct DsimPolicy32 {
    int device, sms;
    float *w;
};

extern "C" int64_t dsim_policy32_blob_elems(void) { return F_ELEMS; }

extern "C" int dsim_policy32_create(int device, const float *weights_host, DsimPolicy32 **out) {
    if (!weights_host || !out) return DSIM_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DSIM_ECUDA;     // no CPU fallback
    if (device < 0 || device >= ndev) return DSIM_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return DSIM_ECUDA;
    DsimPolicy32 *h = new (std::nothrow) DsimPolicy32();
    if (!h) return DSIM_ENOMEM;
    h->device = device;
    if (cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaMalloc((void **)&h->w, F_ELEMS * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(h->w, weights_host, F_ELEMS * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(rma_full_forward_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
        if (h->w) cudaFree(h->w);
        delete h;
        return DSIM_ECUDA;
    }
    *out = h;
    return DSIM_OK;
}

extern "C" void dsim_policy32_destroy(DsimPolicy32 *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->w);
    delete h;
}

extern "C" int dsim_policy32_forward(DsimPolicy32 *h, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                                     float *logits_dev, float *value_dev, void *stream) {
    if (!h || !obs_dev || !prev_action_dev || !logits_dev || !value_dev || n <= 0) return DSIM_EINVAL;
    if (cudaSetDevice(h->device) != cudaSuccess) return DSIM_ECUDA;
    P32 p;
    p.w = h->w; p.obs = obs_dev; p.prev_action = prev_action_dev; p.reset_mask = reset_mask_dev; p.logits = logits_dev; p.value = value_dev;
    p.n = n; p.ntiles = (n + TM - 1) / TM;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(p.ntiles < kCtasPerSm * h->sms ? p.ntiles : kCtasPerSm * h->sms); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = SMEM_BYTES; lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    void *args[] = {&p};
    if (cudaLaunchKernelExC(&lc, (const void *)rma_full_forward_fp32_kernel, args) != cudaSuccess) return DSIM_ECUDA;
    return cudaGetLastError() == cudaSuccess ? DSIM_OK : DSIM_ECUDA;
}
