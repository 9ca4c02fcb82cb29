// Strict per-frame online inference as ONE cooperative launch per frame (BASELINE configs[1]; rnn.py:51-71 with
// B <= 8 streams and T == 1).  The four GEMV launches of online_kernels.cuh are bound by launch ramps and by the
// bytes each of them keeps in flight; here every CTA owns a fixed slice of the weight ROWS of all three big
// matrices, keeps its 16-bit weights in flight as plain 16-byte register loads (the next phase's slice is requested
// BEFORE the grid barrier it depends on, so the L2 latency hides behind the barrier) and does the arithmetic on
// the legacy tensor-core path (mma.sync m16n8k16, fp32 accumulate): 16 weight rows x 8 streams per instruction, so
// up to 8 streams cost the same as one and no warp-shuffle reductions are needed.
//
//   phase A : y[n]  = W1[n, :] . [rgb | flow] + b1[n]      (rows n of this CTA, K split over the 16 warps)
//             gh[u] = W_hh'[u, :] . h                       (units u of this CTA, gates r | z | n; stays in smem)
//   -- grid barrier 1 (y complete) --
//   phase B : e = relu(LN(y)) (every CTA, 8 KB per stream), gi[u] = W_ih'[u, :] . e, gates, h' (in place), relu(h')
//   -- barrier 2 (arrive: all CTAs; wait: only the CTAs that own a class row) --
//   head    : logit[k] = Wc[k, :] . relu(h') + bc[k]; the last CTA to finish does softmax + first-max argmax and
//             re-arms the three counters for the next frame.
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
constexpr uint32_t kFusedSpinMax = 1u << 22;  // bounded spins: a lost CTA raises the error flag instead of hanging the GPU

struct OnlineFusedArgs {
    const float* rgb;
    const float* flow;
    const void* w1;   // [E, D] 16-bit
    const void* wih;  // [3H, E] 16-bit, gate-interleaved rows (packed row = (u/64)*192 + gate*64 + u%64)
    const void* whh;  // [3H, H] 16-bit, same row order
    const float *b1, *ln_g, *ln_b, *bih, *bhh;  // bih / bhh in packed row order
    const float *wc, *bc;                        // fp32 classifier
    float* y;      // [8, E] scratch
    float* hrelu;  // [8, H] scratch
    float* lg;     // [8, K] scratch
    float* h;      // [rows, H] carried state, updated in place
    float* probs;
    float* logits;
    int32_t* labels;
    unsigned* sync;  // [3] zero before the first launch; the kernel re-arms them
    int* err_flag;
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

// 32 bytes of one weight row (two independent 16-byte loads; zeros when the row is not this lane's to fetch)
struct W32 {
    uint4 lo, hi;
};
__device__ __forceinline__ W32 ldw32(const void* base, int64_t elem_off, bool valid) {
    W32 w;
    if (valid) {
        const uint4* p = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + elem_off);
        w.lo = __ldg(p);
        w.hi = __ldg(p + 1);
    } else {
        w.lo = make_uint4(0u, 0u, 0u, 0u);
        w.hi = w.lo;
    }
    return w;
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
    __threadfence();
    atomicAdd(ctr, 1u);
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
    float* red = ghs + 192;                              // [32]
    __shared__ int is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int G = gridDim.x, cta = blockIdx.x;
    const int upc = (H + G - 1) / G, rpc = (E + G - 1) / G;  // hidden units / W1 rows per CTA (<= 8 / <= 16: checked by the host)
    const int u0 = cta * upc, n0 = cta * rpc;
    const int nu = min(upc, max(H - u0, 0)), nr = min(rpc, max(E - n0, 0));
    const bool uval = g < nu;
    const int u = u0 + g;

    // ---- weight requests of phase A (nothing they depend on): W_hh' rows of (u, r|z|n), first half of the W1 chunks
    W32 wh[CH][3];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int gt = 0; gt < 3; ++gt)
            wh[c][gt] = ldw32(a.whh, static_cast<int64_t>(packed_row(uval ? u : 0, gt)) * H + (warp * CH + c) * 64 + t * 16, uval);
    const bool v1a = g < nr, v1b = g + 8 < nr;
    const int64_t r1a = static_cast<int64_t>(n0 + (v1a ? g : 0)) * D, r1b = static_cast<int64_t>(n0 + (v1b ? g + 8 : 0)) * D;
    constexpr int C1A = C1 / 2;
    W32 w1a[C1A], w1b[C1A];
#pragma unroll
    for (int c = 0; c < C1A; ++c) {
        const int k = (warp * C1 + c) * 64 + t * 16;
        w1a[c] = ldw32(a.w1, r1a + k, v1a);
        w1b[c] = ldw32(a.w1, r1b + k, v1b);
    }

    // ---- activations -> 16-bit smem rows: x = [rgb | flow] of frame t0, h = carried state
    for (int i = tid; i < R * (D / 4); i += kFusedThreads) {
        const int r = i / (D / 4), c4 = (i % (D / 4)) * 4;
        const float* src = c4 < a.Dr ? a.rgb + (static_cast<int64_t>(r) * a.T + a.t0) * a.Dr + c4
                                     : a.flow + (static_cast<int64_t>(r) * a.T + a.t0) * a.Df + (c4 - a.Dr);
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        *reinterpret_cast<uint2*>(xs + r * XS + c4 * 2) = make_uint2(Op::pack2(v.x, v.y), Op::pack2(v.z, v.w));
    }
    for (int i = tid; i < R * (H / 4); i += kFusedThreads) {
        const int r = i / (H / 4), c4 = (i % (H / 4)) * 4;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.h + static_cast<int64_t>(r) * H + c4));
        *reinterpret_cast<uint2*>(hs + r * HS + c4 * 2) = make_uint2(Op::pack2(v.x, v.y), Op::pack2(v.z, v.w));
    }
    __syncthreads();

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
    // second half of the W1 chunks goes out while the first half is consumed
    W32 w1c[C1 - C1A], w1d[C1 - C1A];
#pragma unroll
    for (int c = C1A; c < C1; ++c) {
        const int k = (warp * C1 + c) * 64 + t * 16;
        w1c[c - C1A] = ldw32(a.w1, r1a + k, v1a);
        w1d[c - C1A] = ldw32(a.w1, r1b + k, v1b);
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
            wi[c][gt] = ldw32(a.wih, static_cast<int64_t>(packed_row(uval ? u : 0, gt)) * E + (warp * C2 + c) * 64 + t * 16, uval);
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
            if (row < nr && col < R) a.y[col * E + n0 + row] = s + __ldg(a.b1 + n0 + row);
        } else {
            const int gt = p == 1 ? (row >> 3) : 2;  // pair 1 = (r | z), pair 2 = (n | -)
            if (p == 1 || row < 8) ghs[(gt * 8 + (row & 7)) * 8 + col] = s;
        }
    }
    __syncthreads();
    if (tid == 0) {
        fused_arrive(a.sync + 0);
        fused_wait(a.sync + 0, static_cast<unsigned>(G), a.err_flag);
    }
    __syncthreads();

    // ---- phase B: e = relu(LN(y)) for every stream (each CTA recomputes it), into the 16-bit rows es (alias of xs)
    uint8_t* es = xs;
    for (int r = 0; r < R; ++r) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.y + r * E) + tid);
        float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) tot += red[w];
        const float mu = tot / static_cast<float>(E);
        const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
        float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red[16 + warp] = q;
        __syncthreads();
        float var = 0.f;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) var += red[16 + w];
        const float rstd = 1.0f / sqrtf(var / static_cast<float>(E) + a.eps);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.ln_g) + tid), bt = __ldg(reinterpret_cast<const float4*>(a.ln_b) + tid);
        *reinterpret_cast<uint2*>(es + r * ES + tid * 8) =
            make_uint2(Op::pack2(fmaxf(dx * rstd * gm.x + bt.x, 0.f), fmaxf(dy * rstd * gm.y + bt.y, 0.f)),
                       Op::pack2(fmaxf(dz * rstd * gm.z + bt.z, 0.f), fmaxf(dw * rstd * gm.w + bt.w, 0.f)));
    }
    __syncthreads();
    float drz[4] = {0.f, 0.f, 0.f, 0.f}, dn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C2; ++c) {
        const W32 x = lds_act(es + g * ES, (warp * C2 + c) * 64 + t * 16, sval);
        mma_chunk<FMT>(drz, wi[c][0], wi[c][1], x);
        mma_chunk<FMT>(dn, wi[c][2], zero, x);
    }
    // classifier row of this CTA (fp32, H floats) is requested before the second barrier
    const int kc = cta;  // class rows kc, kc + G, ...
    float2 wc0 = make_float2(0.f, 0.f);
    if (kc < a.K && tid * 2 < H) wc0 = __ldg(reinterpret_cast<const float2*>(a.wc + static_cast<int64_t>(kc) * H) + tid);
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
    if (tid < 64) {
        const int j = tid >> 3, n = tid & 7;  // unit j of this CTA, stream n
        if (j < nu && n < R) {
            const int uu = u0 + j, pr = packed_row(uu, 0);
            // ATen's evaluation order: r, z from (gi + b_ih) + (gh + b_hh); n = tanh(gi_n + r * (gh_n + b_hn))
            const float rr = sigmoid_f((gis[(0 * 8 + j) * 8 + n] + __ldg(a.bih + pr)) + (ghs[(0 * 8 + j) * 8 + n] + __ldg(a.bhh + pr)));
            const float zz = sigmoid_f((gis[(1 * 8 + j) * 8 + n] + __ldg(a.bih + pr + 64)) + (ghs[(1 * 8 + j) * 8 + n] + __ldg(a.bhh + pr + 64)));
            const float nn = tanhf((gis[(2 * 8 + j) * 8 + n] + __ldg(a.bih + pr + 128)) + rr * (ghs[(2 * 8 + j) * 8 + n] + __ldg(a.bhh + pr + 128)));
            const float hp = __ldcg(a.h + static_cast<int64_t>(n) * H + uu);  // fp32 master state (only this CTA writes it)
            const float hn = (hp - nn) * zz + nn;
            a.h[static_cast<int64_t>(n) * H + uu] = hn;
            a.hrelu[n * H + uu] = fmaxf(hn, 0.f);
        }
    }
    __syncthreads();
    if (tid == 0) fused_arrive(a.sync + 1);
    const int n_head = a.K < G ? a.K : G;
    if (cta >= n_head) return;
    if (tid == 0) fused_wait(a.sync + 1, static_cast<unsigned>(G), a.err_flag);
    __syncthreads();

    // ---- head: logit[k] for the class rows of this CTA
    for (int k = kc; k < a.K; k += G) {
        for (int r = 0; r < R; ++r) {
            float s = 0.f;
            for (int i = tid; i * 2 < H; i += kFusedThreads) {
                const float2 w = (k == kc && i == tid) ? wc0 : __ldg(reinterpret_cast<const float2*>(a.wc + static_cast<int64_t>(k) * H) + i);
                const float2 hv = __ldcg(reinterpret_cast<const float2*>(a.hrelu + r * H) + i);
                s = fmaf(w.x, hv.x, s);
                s = fmaf(w.y, hv.y, s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < kFusedWarps; ++w) tot += red[w];
                a.lg[r * a.K + k] = tot + __ldg(a.bc + k);
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(a.sync + 2, 1u);
        is_last = prev + 1u == static_cast<unsigned>(n_head);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last CTA: softmax + first-max argmax per stream (one warp each), counters re-armed for the next frame
    if (tid == 0) {
        a.sync[0] = 0u;
        a.sync[1] = 0u;
        a.sync[2] = 0u;
    }
    if (warp < R) {
        const int r = warp, K = a.K;
        const float* lg = a.lg + r * K;
        const int64_t go = static_cast<int64_t>(r) * a.T + a.t0;
        float mx = -INFINITY;
        for (int j = lane; j < K; j += 32) mx = fmaxf(mx, __ldcg(lg + j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < K; j += 32) sum += expf(__ldcg(lg + j) - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float best = -1.f;
        int arg = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
            const float l = __ldcg(lg + j);
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
        if (lane == 0 && a.labels != nullptr) a.labels[go] = arg;
    }
}

inline size_t online_fused_smem(int rows, int D, int H) {
    const int XS = D * 2 + 16, ES = 2048 * 2 + 16, HS = H * 2 + 16;
    return static_cast<size_t>(rows) * ((XS > ES ? XS : ES) + HS) + (3 * kFusedWarps * 128 + 192 + 32) * sizeof(float);
}

}  // namespace prego
