// Strict per-frame online inference as ONE cooperative launch per frame (BASELINE configs[1]; rnn.py:51-71 with
// B <= 8 streams and T == 1).  The four GEMV launches of online_kernels.cuh are bound by launch ramps and by the
// bytes each of them keeps in flight; here every CTA owns a fixed slice of the weight ROWS of all three big
// matrices, keeps its 16-bit weights in flight as plain 16-byte register loads (the next phase's slice is requested
// BEFORE the grid barrier it depends on, so the L2 latency hides behind the barrier) and does the arithmetic on
// the legacy tensor-core path (mma.sync m16n8k16, fp32 accumulate): 16 weight rows x 8 streams per instruction, so
// up to 8 streams cost the same as one and no warp-shuffle reductions are needed.
//
//   phase A : y[n]  = W1[n, :] . [rgb | flow] + b1[n]      (rows n of this CTA, K split over the 16 warps; each warp
//             gh[u] = W_hh'[u, :] . h                        stages only ITS K slice of x / h: no CTA-wide sync first)
//             + per-CTA LayerNorm partials (mean, M2 of its y rows) published with the barrier arrival
//   -- grid barrier (y and the partials complete) --
//   phase B : mean / rstd from the 148 partials (Chan's combination: as exact as two passes), e = relu(LN(y)),
//             gi[u] = W_ih'[u, :] . e, gates, h' (in place); the classifier is split over K like everything else:
//             every CTA publishes Wc[:, its units] . relu(h') for all classes
//   tail    : the last CTA to arrive sums the per-CTA logit partials in a fixed order (deterministic), softmax,
//             first-max argmax, and re-arms the counters for the next frame.
//
// The K permutation inside a 64-wide chunk is free (weights and activations are permuted alike): lane (g, t) loads
// 32 contiguous bytes of weight row g (k = 16 t .. 16 t + 15 of the chunk) and feeds MMA j with its halves
// 4 j .. 4 j + 3, the B fragment being the same 8 bytes of the activation row of stream g.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "gemm_tc.cuh"

namespace prego {

constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxRows = 8;
constexpr int kFusedPartStride = 160;  // >= grid size (one logit partial per CTA, padded)
constexpr uint32_t kFusedSpinMax = 1u << 22;  // bounded spins: a lost CTA raises the error flag instead of hanging the GPU

struct OnlineFusedArgs {
    const float* rgb;
    const float* flow;
    const uint4* wstream;  // the three 16-bit weight matrices re-ordered into per-CTA, per-warp, per-load 512-byte
                           // blocks (online_pack_stream): every warp-level 16-byte load reads 4 full cache lines
    const float *b1, *ln_g, *ln_b, *bih, *bhh;  // bih / bhh in packed row order
    const float *wct, *bc;                       // fp32 classifier, transposed [H, K]
    float* y;       // [8, E] scratch
    float2* stats;  // [8, grid] scratch: (mean, M2) of each CTA's y rows
    float* gpart;   // [8, K, kFusedPartStride] scratch: per-CTA logit partials
    float* h;       // [rows, H] carried state, updated in place
    float* probs;
    float* logits;
    int32_t* labels;
    unsigned* sync;  // [2] zero before the first launch; the kernel re-arms them
    int* err_flag;
    long long* trace;  // optional [grid][16] SM-clock stamps of the phase boundaries (diagnostics; NULL = off)
    unsigned long long* host_seq;  // optional host-mapped doorbell words [rows]: {frame number << 32 | label} -- the label and
                                   // its completion flag travel in ONE 8-byte store, so no system-scope fence is needed
    unsigned seq;
    int fence_outputs;             // probs / logits also live in host memory: fence them before ringing the doorbell
    int rows, Dr, Df, E, H, K;
    int64_t T, t0;  // output row of stream r is r * T + t0
    float eps;
};

template <int FMT>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if constexpr (FMT == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 32 bytes of one weight row = two consecutive 16-byte loads of this lane from the packed stream
struct W32 {
    uint4 lo, hi;
};
__device__ __forceinline__ W32 ldw32(const uint4* ws, int i) {  // ws: this lane's slot of load 0; loads are 32 slots apart
    W32 w;
    w.lo = __ldg(ws + i * 32);
    w.hi = __ldg(ws + (i + 1) * 32);
    return w;
}

// Load order of one warp (index i of its 512-byte blocks), C1 = D / 1024:
//   [0, 6)            W_hh' rows of (unit u0 + g, gate i / 2), k = warp * 64 + 16 t + 8 (i % 2) ..
//   [6, 6 + 4 C1)     W1: chunk c = (i - 6) / 4, row n0 + g (+ 8 for (i - 6) % 4 >= 2), k = (warp * C1 + c) * 64 + 16 t + 8 (i % 2) ..
//   [.., + 12)        W_ih': chunk c = j / 6, gate (j % 6) / 2, k = (warp * 2 + c) * 64 + 16 t + 8 (j % 2) ..
// Rows a CTA does not own are zero blocks, so the kernel needs no predicates.
__host__ __device__ constexpr int online_stream_loads(int C1) { return 6 + 4 * C1 + 12; }

template <int FMT>
__global__ void online_pack_stream(const typename Op16<FMT>::T* __restrict__ w1, const typename Op16<FMT>::T* __restrict__ wih,
                                   const typename Op16<FMT>::T* __restrict__ whh, uint4* __restrict__ out, int G, int D, int E, int H) {
    const int C1 = D / 1024, NI = online_stream_loads(C1);
    const int upc = (H + G - 1) / G, rpc = (E + G - 1) / G;
    const int64_t total = static_cast<int64_t>(G) * kFusedWarps * NI * 32;
    for (int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; s < total; s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int lane = static_cast<int>(s % 32), i = static_cast<int>((s / 32) % NI);
        const int warp = static_cast<int>((s / (32 * NI)) % kFusedWarps), cta = static_cast<int>(s / (32 * NI * kFusedWarps));
        const int g = lane >> 2, t = lane & 3, half = i & 1;
        const int u0 = cta * upc, n0 = cta * rpc;
        const int nu = min(upc, max(H - u0, 0)), nr = min(rpc, max(E - n0, 0));
        const typename Op16<FMT>::T* src = nullptr;
        if (i < 6) {
            if (g < nu) src = whh + static_cast<int64_t>(((u0 + g) >> 6) * 192 + (i / 2) * 64 + ((u0 + g) & 63)) * H + warp * 64 + t * 16 + half * 8;
        } else if (i < 6 + 4 * C1) {
            const int j = i - 6, c = j / 4, row = g + ((j % 4) >= 2 ? 8 : 0);
            if (row < nr) src = w1 + static_cast<int64_t>(n0 + row) * D + (warp * C1 + c) * 64 + t * 16 + half * 8;
        } else {
            const int j = i - 6 - 4 * C1, c = j / 6, gt = (j % 6) / 2;
            if (g < nu) src = wih + static_cast<int64_t>(((u0 + g) >> 6) * 192 + gt * 64 + ((u0 + g) & 63)) * E + (warp * 2 + c) * 64 + t * 16 + half * 8;
        }
        out[s] = src != nullptr ? *reinterpret_cast<const uint4*>(src) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// one 64-wide K chunk: rows (ra | rb) x the activation fragment xa (32 bytes of this lane's stream row)
template <int FMT>
__device__ __forceinline__ void mma_chunk(float (&c)[4], const W32& ra, const W32& rb, const W32& x) {
    mma16816<FMT>(c, ra.lo.x, rb.lo.x, ra.lo.y, rb.lo.y, x.lo.x, x.lo.y);
    mma16816<FMT>(c, ra.lo.z, rb.lo.z, ra.lo.w, rb.lo.w, x.lo.z, x.lo.w);
    mma16816<FMT>(c, ra.hi.x, rb.hi.x, ra.hi.y, rb.hi.y, x.hi.x, x.hi.y);
    mma16816<FMT>(c, ra.hi.z, rb.hi.z, ra.hi.w, rb.hi.w, x.hi.z, x.hi.w);
}

__device__ __forceinline__ W32 lds_act(const uint8_t* row_base, int k, bool valid) {
    W32 x;
    if (valid) {
        const uint4* p = reinterpret_cast<const uint4*>(row_base + k * 2);
        x.lo = p[0];
        x.hi = p[1];
    } else {
        x.lo = make_uint4(0u, 0u, 0u, 0u);
        x.hi = x.lo;
    }
    return x;
}

__device__ __forceinline__ void fused_arrive(unsigned* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
}
__device__ __forceinline__ void fused_wait(unsigned* ctr, unsigned target, int* err_flag) {
    uint32_t spins = 0;
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) return;
        if (++spins > kFusedSpinMax) {
            atomicExch(err_flag, 3);
            return;
        }
    }
}

// slot i = SM clock, slot 8 + i = %globaltimer (ns; comparable across SMs) for i = 0 (entry) and 7 / 8 (exit paths)
#define FUSED_STAMP(i)                                                                              \
    do {                                                                                            \
        if (a.trace != nullptr && threadIdx.x == 0) {                                               \
            a.trace[blockIdx.x * 16 + (i)] = clock64();                                             \
            if ((i) == 0 || (i) >= 7) {                                                             \
                unsigned long long gt_;                                                             \
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                             \
                a.trace[blockIdx.x * 16 + ((i) == 0 ? 9 : 10)] = static_cast<long long>(gt_);       \
                if ((i) == 0) a.trace[blockIdx.x * 16 + 8] = 0;                                     \
            }                                                                                       \
        }                                                                                           \
    } while (0)

// Packed row of hidden unit u, gate gt (0 r, 1 z, 2 n) in W_ih' / W_hh' / bih / bhh.
__device__ __forceinline__ int packed_row(int u, int gt) { return (u >> 6) * 192 + gt * 64 + (u & 63); }

// C1 = D / 1024 (K chunks of W1 per warp), CH = H / 1024.  E is 2048 (two chunks of W_ih' per warp).
template <int FMT, int C1, int CH>
__global__ void __launch_bounds__(kFusedThreads, 1) online_fused_kernel(const OnlineFusedArgs a) {
    using Op = Op16<FMT>;
    constexpr int D = C1 * 1024, H = CH * 1024, E = 2048, C2 = 2;
    constexpr int XS = D * 2 + 16, HS = H * 2 + 16, ES = E * 2 + 16;  // padded activation row strides (bytes)
    extern __shared__ __align__(16) uint8_t fsm[];
    const int R = a.rows;
    uint8_t* xs = fsm;                                   // [R][XS]  (phase A)  /  es [R][ES] (phase B)
    uint8_t* hs = xs + R * (XS > ES ? XS : ES);          // [R][HS]
    float* part = reinterpret_cast<float*>(hs + R * HS);  // [3][16 warps][128]
    float* ghs = part + 3 * kFusedWarps * 128;           // [3 gates][8 units][8 streams]
    float* ysm = ghs + 192;                              // [8 streams][16 rows]   (phase A) / hrl [8 streams][8 units]
    float* lnp = ysm + 128;                              // [8 streams][2]: mean, rstd
    __shared__ int is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int G = gridDim.x, cta = blockIdx.x;
    const int upc = (H + G - 1) / G, rpc = (E + G - 1) / G;  // hidden units / W1 rows per CTA (<= 8 / <= 16: checked by the host)
    const int u0 = cta * upc, n0 = cta * rpc;
    const int nu = min(upc, max(H - u0, 0)), nr = min(rpc, max(E - n0, 0));
    FUSED_STAMP(0);

    // ---- this warp's K slice of the activations is requested FIRST (the L2 -> SM path is served in order and the
    //      weight requests below queue ~110 KB per SM): x = [rgb | flow] of frame t0 (k in [warp * C1 * 64, +C1 * 64)),
    //      h = carried state (slice of CH * 64); first batch of four x loads + the h loads
    constexpr int XV = C1 * 16, HV = CH * 16;  // float4 per stream in the slice
    const int kx0 = warp * C1 * 64, kh0 = warp * CH * 64;
    auto x_src = [&](int i) -> const float4* {
        const int r = i / XV, c4 = kx0 + (i % XV) * 4;
        return reinterpret_cast<const float4*>(c4 < a.Dr ? a.rgb + (static_cast<int64_t>(r) * a.T + a.t0) * a.Dr + c4
                                                          : a.flow + (static_cast<int64_t>(r) * a.T + a.t0) * a.Df + (c4 - a.Dr));
    };
    float4 hv[(kFusedMaxRows * HV + 31) / 32], xv[4];
#pragma unroll
    for (int q = 0; q < (kFusedMaxRows * HV + 31) / 32; ++q) {
        const int i = q * 32 + lane;
        if (i < R * HV) hv[q] = __ldcg(reinterpret_cast<const float4*>(a.h + static_cast<int64_t>(i / HV) * H + kh0 + (i % HV) * 4));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = q * 32 + lane;
        if (i < R * XV) xv[q] = __ldg(x_src(i));
    }

    // ---- weight requests of phase A (nothing they depend on): W_hh' rows of (u, r|z|n), first half of the W1 chunks
    const uint4* ws = a.wstream + (static_cast<int64_t>(cta) * kFusedWarps + warp) * online_stream_loads(C1) * 32 + lane;
    W32 wh[CH][3];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int gt = 0; gt < 3; ++gt)
            wh[c][gt] = ldw32(ws, 2 * gt);
    constexpr int C1A = C1 / 2;
    W32 w1a[C1A], w1b[C1A];
#pragma unroll
    for (int c = 0; c < C1A; ++c) {
        w1a[c] = ldw32(ws, 6 + 4 * c);
        w1b[c] = ldw32(ws, 6 + 4 * c + 2);
    }

    // ---- activations -> 16-bit smem rows (warp-private slices: __syncwarp only)
    {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = q * 32 + lane;
            if (i < R * XV)
                *reinterpret_cast<uint2*>(xs + (i / XV) * XS + (kx0 + (i % XV) * 4) * 2) = make_uint2(Op::pack2(xv[q].x, xv[q].y), Op::pack2(xv[q].z, xv[q].w));
        }
        for (int i0 = 128; i0 < R * XV; i0 += 128) {  // more than two streams: further batches of four loads per lane
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * 32 + lane;
                if (i < R * XV) v[q] = __ldg(x_src(i));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * 32 + lane;
                if (i < R * XV)
                    *reinterpret_cast<uint2*>(xs + (i / XV) * XS + (kx0 + (i % XV) * 4) * 2) = make_uint2(Op::pack2(v[q].x, v[q].y), Op::pack2(v[q].z, v[q].w));
            }
        }
#pragma unroll
        for (int q = 0; q < (kFusedMaxRows * HV + 31) / 32; ++q) {
            const int i = q * 32 + lane;
            if (i < R * HV)
                *reinterpret_cast<uint2*>(hs + (i / HV) * HS + (kh0 + (i % HV) * 4) * 2) = make_uint2(Op::pack2(hv[q].x, hv[q].y), Op::pack2(hv[q].z, hv[q].w));
        }
        __syncwarp();
    }
    FUSED_STAMP(1);

    // ---- phase A math
    const bool sval = g < R;  // this lane's B-fragment column (stream g) exists
    float crz[4] = {0.f, 0.f, 0.f, 0.f}, cn[4] = {0.f, 0.f, 0.f, 0.f}, cy[4] = {0.f, 0.f, 0.f, 0.f};
    const W32 zero{make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const W32 x = lds_act(hs + g * HS, (warp * CH + c) * 64 + t * 16, sval);
        mma_chunk<FMT>(crz, wh[c][0], wh[c][1], x);
        mma_chunk<FMT>(cn, wh[c][2], zero, x);
    }
    // second half of the W1 chunks goes out while the first half is consumed (requesting it at t = 0 as well was slower:
    // the x / h slices then queue behind 50 % more weight bytes on the L2 -> SM path)
    W32 w1c[C1 - C1A], w1d[C1 - C1A];
#pragma unroll
    for (int c = C1A; c < C1; ++c) {
        w1c[c - C1A] = ldw32(ws, 6 + 4 * c);
        w1d[c - C1A] = ldw32(ws, 6 + 4 * c + 2);
    }
#pragma unroll
    for (int c = 0; c < C1A; ++c) {
        const W32 x = lds_act(xs + g * XS, (warp * C1 + c) * 64 + t * 16, sval);
        mma_chunk<FMT>(cy, w1a[c], w1b[c], x);
    }
    // ---- weight requests of phase B (W_ih' rows of this CTA's units) fly across the barrier
    W32 wi[C2][3];
#pragma unroll
    for (int c = 0; c < C2; ++c)
#pragma unroll
        for (int gt = 0; gt < 3; ++gt)
            wi[c][gt] = ldw32(ws, 6 + 4 * C1 + 6 * c + 2 * gt);
#pragma unroll
    for (int c = C1A; c < C1; ++c) {
        const W32 x = lds_act(xs + g * XS, (warp * C1 + c) * 64 + t * 16, sval);
        mma_chunk<FMT>(cy, w1c[c - C1A], w1d[c - C1A], x);
    }
    // cross-warp K reduction (fixed order -> deterministic)
    *reinterpret_cast<float4*>(part + (0 * kFusedWarps + warp) * 128 + lane * 4) = make_float4(cy[0], cy[1], cy[2], cy[3]);
    *reinterpret_cast<float4*>(part + (1 * kFusedWarps + warp) * 128 + lane * 4) = make_float4(crz[0], crz[1], crz[2], crz[3]);
    *reinterpret_cast<float4*>(part + (2 * kFusedWarps + warp) * 128 + lane * 4) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    __syncthreads();
    if (tid < 384) {
        const int p = tid >> 7, i = tid & 127;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) s += part[(p * kFusedWarps + w) * 128 + i];
        const int el = i >> 2, reg = i & 3;                       // accumulator register `reg` of lane `el`
        const int row = (el >> 2) + 8 * (reg >> 1), col = (el & 3) * 2 + (reg & 1);  // weight row, stream
        if (p == 0) {
            float yv = 0.f;
            if (row < nr && col < R) {
                yv = s + __ldg(a.b1 + n0 + row);
                a.y[col * E + n0 + row] = yv;
            }
            ysm[col * 16 + row] = yv;
        } else {
            const int gt = p == 1 ? (row >> 3) : 2;  // pair 1 = (r | z), pair 2 = (n | -)
            if (p == 1 || row < 8) ghs[(gt * 8 + (row & 7)) * 8 + col] = s;
        }
    }
    __syncthreads();
    if (tid < R) {  // LayerNorm partials of this CTA's rows: (mean, sum of squared deviations)
        float m = 0.f, q = 0.f;
        if (nr > 0) {
            for (int r = 0; r < nr; ++r) m += ysm[tid * 16 + r];
            m /= static_cast<float>(nr);
            for (int r = 0; r < nr; ++r) {
                const float d = ysm[tid * 16 + r] - m;
                q += d * d;
            }
        }
        a.stats[tid * G + cta] = make_float2(m, q);
    }
    __syncthreads();
    FUSED_STAMP(2);
    if (tid == 0) {
        fused_arrive(a.sync + 0);
        fused_wait(a.sync + 0, static_cast<unsigned>(G), a.err_flag);
    }
    __syncthreads();
    FUSED_STAMP(3);

    // gate biases and the fp32 master state of this thread's (unit, stream) item: requested now, used after phase B
    float gb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, hp = 0.f;
    if (tid < 64 && (tid >> 3) < nu && (tid & 7) < R) {
        const int uu = u0 + (tid >> 3), pr = packed_row(uu, 0);
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) {
            gb[gt] = __ldg(a.bih + pr + gt * 64);
            gb[3 + gt] = __ldg(a.bhh + pr + gt * 64);
        }
        hp = __ldcg(a.h + static_cast<int64_t>(tid & 7) * H + uu);  // only this CTA writes it (after this point)
    }

    // ---- phase B: e = relu(LN(y)) for every stream (each CTA recomputes it), into the 16-bit rows es (alias of xs)
    uint8_t* es = xs;
    if (warp < R) {  // warp r combines the per-CTA partials of stream r (Chan et al.: exact merge of (n, mean, M2))
        float2 st[(kFusedPartStride + 31) / 32];
        float wsum = 0.f;
#pragma unroll
        for (int q = 0; q < (kFusedPartStride + 31) / 32; ++q) {
            const int c = q * 32 + lane;
            st[q] = c < G ? __ldcg(a.stats + warp * G + c) : make_float2(0.f, 0.f);
            const int nc = min(rpc, max(E - c * rpc, 0));
            wsum += c < G ? st[q].x * static_cast<float>(nc) : 0.f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        const float mu = wsum / static_cast<float>(E);
        float m2 = 0.f;
#pragma unroll
        for (int q = 0; q < (kFusedPartStride + 31) / 32; ++q) {
            const int c = q * 32 + lane;
            const int nc = min(rpc, max(E - c * rpc, 0));
            const float d = st[q].x - mu;
            m2 += c < G ? st[q].y + static_cast<float>(nc) * d * d : 0.f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        if (lane == 0) {
            lnp[warp * 2] = mu;
            lnp[warp * 2 + 1] = 1.0f / sqrtf(m2 / static_cast<float>(E) + a.eps);
        }
    }
    {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.ln_g) + tid), bt = __ldg(reinterpret_cast<const float4*>(a.ln_b) + tid);
        float4 v[kFusedMaxRows];
#pragma unroll
        for (int r = 0; r < kFusedMaxRows; ++r)
            if (r < R) v[r] = __ldcg(reinterpret_cast<const float4*>(a.y + r * E) + tid);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kFusedMaxRows; ++r)
            if (r < R) {
                const float mu = lnp[r * 2], rstd = lnp[r * 2 + 1];
                *reinterpret_cast<uint2*>(es + r * ES + tid * 8) =
                    make_uint2(Op::pack2(fmaxf((v[r].x - mu) * rstd * gm.x + bt.x, 0.f), fmaxf((v[r].y - mu) * rstd * gm.y + bt.y, 0.f)),
                               Op::pack2(fmaxf((v[r].z - mu) * rstd * gm.z + bt.z, 0.f), fmaxf((v[r].w - mu) * rstd * gm.w + bt.w, 0.f)));
            }
    }
    __syncthreads();
    FUSED_STAMP(4);
    float drz[4] = {0.f, 0.f, 0.f, 0.f}, dn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C2; ++c) {
        const W32 x = lds_act(es + g * ES, (warp * C2 + c) * 64 + t * 16, sval);
        mma_chunk<FMT>(drz, wi[c][0], wi[c][1], x);
        mma_chunk<FMT>(dn, wi[c][2], zero, x);
    }
    // classifier columns of this CTA's units (fp32 Wc^T rows u0 .. u0 + nu - 1), first (class, stream) item of this thread
    const int KR = a.K * R;
    float wcv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wcv[j] = (tid < KR && j < nu) ? __ldg(a.wct + static_cast<int64_t>(u0 + j) * a.K + tid % a.K) : 0.f;
    *reinterpret_cast<float4*>(part + (0 * kFusedWarps + warp) * 128 + lane * 4) = make_float4(drz[0], drz[1], drz[2], drz[3]);
    *reinterpret_cast<float4*>(part + (1 * kFusedWarps + warp) * 128 + lane * 4) = make_float4(dn[0], dn[1], dn[2], dn[3]);
    __syncthreads();
    float* gis = part + 2 * kFusedWarps * 128;  // [3][8][8], region of pair 2 (unused in phase B)
    if (tid < 256) {
        const int p = tid >> 7, i = tid & 127;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) s += part[(p * kFusedWarps + w) * 128 + i];
        const int el = i >> 2, reg = i & 3;
        const int row = (el >> 2) + 8 * (reg >> 1), col = (el & 3) * 2 + (reg & 1);
        const int gt = p == 0 ? (row >> 3) : 2;
        if (p == 0 || row < 8) gis[(gt * 8 + (row & 7)) * 8 + col] = s;
    }
    __syncthreads();
    float* hrl = ysm;  // [8 streams][8 units] relu(h') of this CTA's units
    if (tid < 64) {
        const int j = tid >> 3, n = tid & 7;  // unit j of this CTA, stream n
        float hr = 0.f;
        if (j < nu && n < R) {
            const int uu = u0 + j;
            // ATen's evaluation order: r, z from (gi + b_ih) + (gh + b_hh); n = tanh(gi_n + r * (gh_n + b_hn))
            const float rr = sigmoid_f((gis[(0 * 8 + j) * 8 + n] + gb[0]) + (ghs[(0 * 8 + j) * 8 + n] + gb[3]));
            const float zz = sigmoid_f((gis[(1 * 8 + j) * 8 + n] + gb[1]) + (ghs[(1 * 8 + j) * 8 + n] + gb[4]));
            const float nn = tanhf((gis[(2 * 8 + j) * 8 + n] + gb[2]) + rr * (ghs[(2 * 8 + j) * 8 + n] + gb[5]));
            const float hn = (hp - nn) * zz + nn;
            a.h[static_cast<int64_t>(n) * H + uu] = hn;
            hr = fmaxf(hn, 0.f);
        }
        hrl[n * 8 + j] = hr;
    }
    __syncthreads();
    FUSED_STAMP(5);
    // ---- classifier, split over the hidden units like the other layers: this CTA's share of every logit
    for (int i = tid; i < KR; i += kFusedThreads) {
        const int k = i % a.K, n = i / a.K;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float w = i == tid ? wcv[j] : (j < nu ? __ldg(a.wct + static_cast<int64_t>(u0 + j) * a.K + k) : 0.f);
            s = fmaf(w, hrl[n * 8 + j], s);
        }
        a.gpart[(static_cast<int64_t>(n) * a.K + k) * kFusedPartStride + cta] = s;
    }
    const float bc0 = (tid >> 4) < KR ? __ldg(a.bc + (tid >> 4) % a.K) : 0.f;
    __syncthreads();
    FUSED_STAMP(6);
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(a.sync + 1, 1u);
        is_last = prev + 1u == static_cast<unsigned>(G);
    }
    __syncthreads();
    FUSED_STAMP(7);
    if (!is_last) return;
    // ---- last CTA: logits = sum of the per-CTA partials in CTA order (16 lanes per logit, fixed shuffle tree),
    //      softmax + first-max argmax per stream (one warp each); counters re-armed for the next frame
    if (tid == 0) {
        a.sync[0] = 0u;
        a.sync[1] = 0u;
    }
    float* lgs = reinterpret_cast<float*>(xs);  // [R][K] logits; the activation rows (>= 4 KB per stream) are free now
    {
        const int sub = tid & 15;
        for (int o = tid >> 4; o < KR; o += kFusedThreads / 16) {
            const float* gp = a.gpart + static_cast<int64_t>(o) * kFusedPartStride;
            float v[kFusedPartStride / 16];
#pragma unroll
            for (int q = 0; q < kFusedPartStride / 16; ++q) v[q] = (q * 16 + sub) < G ? __ldcg(gp + q * 16 + sub) : 0.f;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < kFusedPartStride / 16; ++q) s += v[q];
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (sub == 0) lgs[o] = s + (o == (tid >> 4) ? bc0 : __ldg(a.bc + o % a.K));
        }
    }
    __syncthreads();
    if (warp < R) {
        const int r = warp, K = a.K;
        const float* lg = lgs + r * K;
        const int64_t go = static_cast<int64_t>(r) * a.T + a.t0;
        float mx = -INFINITY;
        for (int j = lane; j < K; j += 32) mx = fmaxf(mx, lg[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < K; j += 32) sum += expf(lg[j] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float best = -1.f;
        int arg = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
            const float l = lg[j];
            const float p = expf(l - mx) / sum;
            if (a.probs != nullptr) a.probs[go * K + j] = p;
            if (a.logits != nullptr) a.logits[go * K + j] = l;
            if (p > best) {
                best = p;
                arg = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
        }
        if (a.host_seq != nullptr) {  // completion doorbell for a host that polls instead of synchronizing the stream
            if (a.fence_outputs) __threadfence_system();  // this warp's probs / logits stores first
            if (lane == 0)
                *reinterpret_cast<volatile unsigned long long*>(a.host_seq + r) =
                    (static_cast<unsigned long long>(a.seq) << 32) | static_cast<unsigned>(arg);
        }
        if (lane == 0 && a.labels != nullptr) a.labels[go] = arg;  // (posted store: visible at the latest when the stream drains)
    }
    FUSED_STAMP(8);
}

inline size_t online_fused_smem(int rows, int D, int H) {
    const int XS = D * 2 + 16, ES = 2048 * 2 + 16, HS = H * 2 + 16;
    return static_cast<size_t>(rows) * ((XS > ES ? XS : ES) + HS) + (3 * kFusedWarps * 128 + 192 + 128 + 16) * sizeof(float);
}

// fp32 scratch of one online context: y [8, E] | stats [8, stride] float2 | gpart [8, K, stride] | 4 counters
inline size_t online_fused_scratch_floats(int E, int K) {
    return static_cast<size_t>(kFusedMaxRows) * (E + 2 * kFusedPartStride + static_cast<size_t>(K) * kFusedPartStride) + 4;
}

}  // namespace prego
