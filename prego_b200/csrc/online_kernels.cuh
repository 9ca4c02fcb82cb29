// Strict per-frame online inference (BASELINE configs[1]: one frame per call, carried GRU state) for a few
// streams (R = B rows <= 8, T == 1).  With a handful of rows every layer is a GEMV that must stream its
// weights once (W1 16.8 MB + W_ih 12.6 MB + W_hh 6.3 MB in 16 bits, L2-resident between frames), so the
// tensor-core tiles (128 rows) are the wrong tool: these kernels spread the weight ROWS over every warp of
// the chip (one warp = one output feature, 16-byte weight loads, fp32 accumulation) and fuse everything
// else into them:
//
//   online_proj1 : y = [rgb | flow] W1^T + b1           (feature concat + 16-bit rounding fused in the prologue)
//   online_proj2 : e = relu(LN(y)) in the prologue, gi = e W_ih'^T + b_ih'
//   online_gru   : gh = h W_hh'^T + b_hh', gates, h' = (h - n) z + n          (one warp per hidden unit)
//   online_head  : logits = relu(h') Wc^T + bc, softmax, first-max argmax
//
// Four launches per frame instead of the ten of the batched pipeline; no tile padding, no tensor maps.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "gemm_tc.cuh"

namespace prego {

constexpr int kOnlineMaxRows = 8;
constexpr int kOnlineThreads = 256;

// dot products of one 16-bit weight row (K elements, K % 256 == 0) with R fp16/bf16 activation rows held in
// shared memory (xs[r * K + k]); lane-strided 16-byte loads; returns the warp-reduced sums in acc[R].
template <int FMT, int R>
__device__ __forceinline__ void warp_row_dot(const typename Op16<FMT>::T* __restrict__ wrow, const typename Op16<FMT>::T* xs,
                                             int K, int lane, float* acc) {
    using Op = Op16<FMT>;
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const uint4* w4 = reinterpret_cast<const uint4*>(wrow);
    const int nvec = K / 8;
    // four 16-byte weight loads in flight per lane (the rows stream from L2: latency-bound without the batching)
    for (int i0 = lane; i0 < nvec; i0 += 128) {
        uint4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            w[u] = i < nvec ? __ldg(w4 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i >= nvec) break;
            const uint32_t wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
            float wf[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = Op::unpack2(wu[j]);
                wf[2 * j] = f.x;
                wf[2 * j + 1] = f.y;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint4 x = *reinterpret_cast<const uint4*>(xs + static_cast<int64_t>(r) * K + i * 8);
                const uint32_t xu[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = Op::unpack2(xu[j]);
                    acc[r] = fmaf(wf[2 * j], f.x, acc[r]);
                    acc[r] = fmaf(wf[2 * j + 1], f.y, acc[r]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
}

// y[r, n] = sum_k x[r, k] W1[n, k] + b1[n];  x row r = [rgb[r*T + t0, :] | flow[r*T + t0, :]] rounded to the operand format.
template <int FMT, int R>
__global__ void __launch_bounds__(kOnlineThreads)
online_proj1(const float* __restrict__ rgb, const float* __restrict__ flow, const typename Op16<FMT>::T* __restrict__ w1,
             const float* __restrict__ b1, float* __restrict__ y, int rows, int Dr, int Df, int E, int64_t T, int64_t t0) {
    using Op = Op16<FMT>;
    extern __shared__ __align__(16) uint8_t osm[];
    typename Op::T* xs = reinterpret_cast<typename Op::T*>(osm);
    const int D = Dr + Df;
    for (int i = threadIdx.x; i < R * D / 2; i += kOnlineThreads) {
        const int r = (i * 2) / D, c = (i * 2) % D;
        uint32_t v = 0u;  // rows beyond the real stream count are zero
        if (r < rows) {
            const float* src = c < Dr ? rgb + (static_cast<int64_t>(r) * T + t0) * Dr + c : flow + (static_cast<int64_t>(r) * T + t0) * Df + (c - Dr);
            v = Op::pack2(src[0], src[1]);
        }
        reinterpret_cast<uint32_t*>(xs)[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = gridDim.x * (kOnlineThreads / 32);
    for (int n = blockIdx.x * (kOnlineThreads / 32) + warp; n < E; n += warps) {
        float acc[R];
        warp_row_dot<FMT, R>(w1 + static_cast<int64_t>(n) * D, xs, D, lane, acc);
        if (lane < rows) {
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane == r) v = acc[r];
            y[static_cast<int64_t>(lane) * E + n] = v + __ldg(b1 + n);
        }
    }
}

// e = relu(LayerNorm(y)) (each CTA recomputes it for the R rows: 8 KB per row), gi[r, p] = e W_ih'[p, :] + b_ih'[p].
template <int FMT, int R>
__global__ void __launch_bounds__(kOnlineThreads)
online_proj2(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
             const typename Op16<FMT>::T* __restrict__ wih, const float* __restrict__ bih, float* __restrict__ gi, int rows,
             int E, int N3, float eps) {
    using Op = Op16<FMT>;
    extern __shared__ __align__(16) uint8_t osm[];
    typename Op::T* es = reinterpret_cast<typename Op::T*>(osm);
    __shared__ float red[2][kOnlineThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = 0; r < R; ++r) {
        if (r >= rows) {  // block-uniform
            for (int i = threadIdx.x; i < E; i += kOnlineThreads) es[static_cast<int64_t>(r) * E + i] = Op::from_float(0.f);
            continue;
        }
        const float* yr = y + static_cast<int64_t>(r) * E;
        float s = 0.f;
        for (int i = threadIdx.x; i < E; i += kOnlineThreads) s += yr[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[0][warp] = s;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kOnlineThreads / 32; ++w) tot += red[0][w];
        const float mu = tot / static_cast<float>(E);
        float q = 0.f;
        for (int i = threadIdx.x; i < E; i += kOnlineThreads) {
            const float d = yr[i] - mu;
            q += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red[1][warp] = q;
        __syncthreads();
        float var = 0.f;
#pragma unroll
        for (int w = 0; w < kOnlineThreads / 32; ++w) var += red[1][w];
        const float rstd = 1.0f / sqrtf(var / static_cast<float>(E) + eps);
        for (int i = threadIdx.x; i < E; i += kOnlineThreads)
            es[static_cast<int64_t>(r) * E + i] = Op::from_float(fmaxf((yr[i] - mu) * rstd * __ldg(gamma + i) + __ldg(beta + i), 0.f));
        __syncthreads();
    }
    const int warps = gridDim.x * (kOnlineThreads / 32);
    for (int n = blockIdx.x * (kOnlineThreads / 32) + warp; n < N3; n += warps) {
        float acc[R];
        warp_row_dot<FMT, R>(wih + static_cast<int64_t>(n) * E, es, E, lane, acc);
        if (lane < rows) {
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane == r) v = acc[r];
            gi[static_cast<int64_t>(lane) * N3 + n] = v + __ldg(bih + n);
        }
    }
}

// One GRU step, one warp per hidden unit (packed rows r | z | n of W_hh'); h_in / h_out must not alias.
template <int FMT, int R>
__global__ void __launch_bounds__(kOnlineThreads)
online_gru(const float* __restrict__ gi, const typename Op16<FMT>::T* __restrict__ whh, const float* __restrict__ bhh,
           const float* __restrict__ h_in, float* __restrict__ h_out, float* __restrict__ hrelu, int rows, int H) {
    using Op = Op16<FMT>;
    extern __shared__ __align__(16) uint8_t osm[];
    typename Op::T* hs = reinterpret_cast<typename Op::T*>(osm);
    for (int i = threadIdx.x; i < R * H; i += kOnlineThreads) hs[i] = Op::from_float(i < rows * H ? h_in[i] : 0.f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = gridDim.x * (kOnlineThreads / 32);
    for (int u = blockIdx.x * (kOnlineThreads / 32) + warp; u < H; u += warps) {
        const int p = (u / 64) * 192 + (u % 64);
        float ar[R], az[R], an[R];
        warp_row_dot<FMT, R>(whh + static_cast<int64_t>(p) * H, hs, H, lane, ar);
        warp_row_dot<FMT, R>(whh + static_cast<int64_t>(p + 64) * H, hs, H, lane, az);
        warp_row_dot<FMT, R>(whh + static_cast<int64_t>(p + 128) * H, hs, H, lane, an);
        if (lane < rows) {
            float gr = 0.f, gz = 0.f, gn = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (lane == r) {
                    gr = ar[r];
                    gz = az[r];
                    gn = an[r];
                }
            const float* g = gi + static_cast<int64_t>(lane) * 3 * H + p;
            const float rr = sigmoid_f(g[0] + (gr + __ldg(bhh + p)));
            const float zz = sigmoid_f(g[64] + (gz + __ldg(bhh + p + 64)));
            const float nn = tanhf(g[128] + rr * (gn + __ldg(bhh + p + 128)));
            const float hp = h_in[static_cast<int64_t>(lane) * H + u];  // fp32 master state for the update
            const float hn = (hp - nn) * zz + nn;
            h_out[static_cast<int64_t>(lane) * H + u] = hn;
            hrelu[static_cast<int64_t>(lane) * H + u] = fmaxf(hn, 0.f);
        }
    }
}

// logits = relu(h) Wc^T + bc (fp32 weights: 0.35 MB), softmax, first-max argmax.  One CTA per row.
__global__ void __launch_bounds__(kOnlineThreads)
online_head(const float* __restrict__ hrelu, const float* __restrict__ wc, const float* __restrict__ bc,
            float* __restrict__ probs, float* __restrict__ logits, int32_t* __restrict__ labels, int H, int K, int64_t T,
            int64_t t0) {
    __shared__ float lg[1024];
    const int r = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* h = hrelu + static_cast<int64_t>(r) * H;
    for (int k = warp; k < K; k += kOnlineThreads / 32) {
        const float* w = wc + static_cast<int64_t>(k) * H;
        float s = 0.f;
        for (int i = lane; i < H; i += 32) s = fmaf(h[i], __ldg(w + i), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) lg[k] = s + __ldg(bc + k);
    }
    __syncthreads();
    if (warp == 0) {
        const int64_t g = static_cast<int64_t>(r) * T + t0;
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
            const float p = expf(lg[j] - mx) / sum;
            if (probs != nullptr) probs[g * K + j] = p;
            if (logits != nullptr) logits[g * K + j] = lg[j];
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
        if (lane == 0 && labels != nullptr) labels[g] = arg;
    }
}

}  // namespace prego
