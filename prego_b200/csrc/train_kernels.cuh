// Training step of MiniROAD (reference: trainer/train.py:5-29, criterions/loss.py:15-34, rnn.py:51-71 in
// train mode): forward with saved activations and full BPTT backward, exact fp32 on CUDA cores.
//
// Everything is time-major inside the window (row m = t*B + b) so a time step is one contiguous [B, .] slab.
// GRU weights are used in the library's packed gate-interleaved row order (packed row p = nt*192 + g*64 + j
// <-> original row g*H + nt*64 + j); gradients are produced in packed order and un-permuted on the way out.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "simt_kernels.cuh"

namespace prego {

// ---------------------------------------------------------------------------------------------------------
// Generic fp32 GEMM, 128x128x16 tile, 256 threads, 8x8 register tile:
//     C[M, N] (+)= op(A) * op(B) (+ bias[N])
//   TA = false: A is [M, K] row-major (k contiguous);  TA = true: A is stored [K, M] (m contiguous)
//   TB = false: B is [N, K] row-major (k contiguous);  TB = true: B is stored [K, N] (n contiguous)
// With a_time_major_B > 0 (TA = false only) row m of A is gathered from a [B, T, lda] tensor: row (m % B) * T + m / B;
// with b_time_major_B > 0 (TB = true only) row k of B likewise (contraction over time-major rows against a
// caller tensor stored [B, T, ldb]).
// All extents are guarded (any M, N, K); vector loads are used only where alignment allows.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_gen_f32(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm, int64_t ldb,
              const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int accumulate,
              int a_time_major_B, int a_T, int b_time_major_B, int b_T, int k_chunk, int64_t c_split_stride) {
    constexpr int BM = 128, BN = 128, BK = 16;
    // split-K (k_chunk > 0, a multiple of BK): slice blockIdx.z contracts over [z * k_chunk, (z + 1) * k_chunk) and writes its
    // partial product to C + z * c_split_stride; splitk_reduce_f32 sums the slices in a fixed order (deterministic)
    const int k_begin = k_chunk > 0 ? static_cast<int>(blockIdx.z) * k_chunk : 0;
    if (k_chunk > 0) {
        K = min(K, k_begin + k_chunk);
        C += static_cast<int64_t>(blockIdx.z) * c_split_stride;
    }
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < K; k0 += BK) {
        // ---- A tile -> As[k][m]
        if constexpr (!TA) {
            const int lr = tid / 4, lk = (tid % 4) * 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + lr + 64 * h;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (m < M) {
                    const int64_t row = a_time_major_B > 0 ? static_cast<int64_t>(m % a_time_major_B) * a_T + m / a_time_major_B : m;
                    const float* p = A + row * lda + k0 + lk;
                    if (k0 + lk + 3 < K && (lda & 3) == 0) {
                        const float4 q = *reinterpret_cast<const float4*>(p);
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (k0 + lk + e < K) v[e] = p[e];
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) As[lk + e][lr + 64 * h] = v[e];
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = tid + 256 * h;
                const int k = idx / 32, m4 = (idx % 32) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (k0 + k < K) {
                    const float* p = A + static_cast<int64_t>(k0 + k) * lda + m0 + m4;
                    if (m0 + m4 + 3 < M && (lda & 3) == 0) {
                        const float4 q = *reinterpret_cast<const float4*>(p);
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (m0 + m4 + e < M) v[e] = p[e];
                    }
                }
                *reinterpret_cast<float4*>(&As[k][m4]) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        // ---- B tile -> Bs[k][n]
        if constexpr (!TB) {
            const int lr = tid / 4, lk = (tid % 4) * 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + lr + 64 * h;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (n < N) {
                    const float* p = Bm + static_cast<int64_t>(n) * ldb + k0 + lk;
                    if (k0 + lk + 3 < K && (ldb & 3) == 0) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (k0 + lk + e < K) v[e] = p[e];
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) Bs[lk + e][lr + 64 * h] = v[e];
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = tid + 256 * h;
                const int k = idx / 32, n4 = (idx % 32) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (k0 + k < K) {
                    const int kk = k0 + k;
                    const int64_t brow = b_time_major_B > 0 ? static_cast<int64_t>(kk % b_time_major_B) * b_T + kk / b_time_major_B : kk;
                    const float* p = Bm + brow * ldb + n0 + n4;
                    if (n0 + n4 + 3 < N && (ldb & 3) == 0) {
                        const float4 q = *reinterpret_cast<const float4*>(p);
                        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n0 + n4 + e < N) v[e] = p[e];
                    }
                }
                *reinterpret_cast<float4*>(&Bs[k][n4]) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
            float* dst = C + static_cast<int64_t>(m) * ldc + n;
            *dst = accumulate ? *dst + v : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Counter-based RNG for the dropout mask (one 32-bit hash per element; the reference uses torch's Philox
// stream, which cannot be reproduced bit-for-bit -- gradient parity is checked with dropout = 0).
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return static_cast<uint32_t>(x);
}

// layer1 forward in train mode: yn = (y - mu) * rstd (saved), e = dropout(relu(yn * gamma + beta)) (saved,
// includes the 1/(1-p) scale), rstd saved.  One warp per row of width E (multiple of 128, <= 4096).
template <int E>
__global__ void __launch_bounds__(256)
ln_relu_dropout_fwd(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ yn, float* __restrict__ e, float* __restrict__ rstd_out, int64_t M,
                    float eps, float p_drop, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    constexpr int nvec = E / 128;
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    const uint32_t thresh = p_drop > 0.f ? static_cast<uint32_t>(p_drop * 4294967296.0) : 0u;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; row < M; row += warps) {
        const float4* src = reinterpret_cast<const float4*>(y + row * E);
        float4 v[nvec];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < nvec; ++i)
            if (i < nvec) {
                v[i] = src[i * 32 + lane];
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum / static_cast<float>(E);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < nvec; ++i)
            if (i < nvec) {
                const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
                sq += (a * a + b * b) + (c * c + d * d);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.0f / sqrtf(sq / static_cast<float>(E) + eps);
        if (lane == 0) rstd_out[row] = rstd;
        float4* dyn = reinterpret_cast<float4*>(yn + row * E);
        float4* de = reinterpret_cast<float4*>(e + row * E);
#pragma unroll
        for (int i = 0; i < nvec; ++i)
            if (i < nvec) {
                const int c4 = i * 32 + lane;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
                const float n[4] = {(v[i].x - mu) * rstd, (v[i].y - mu) * rstd, (v[i].z - mu) * rstd, (v[i].w - mu) * rstd};
                const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float a = fmaxf(n[q] * gg[q] + bb[q], 0.f);
                    if (p_drop > 0.f) {
                        const uint32_t r = hash32(seed ^ (static_cast<uint64_t>(row) * E + c4 * 4 + q) * 0x9E3779B97F4A7C15ull);
                        a = r < thresh ? 0.f : a * keep_scale;
                    }
                    o[q] = a;
                }
                dyn[c4] = make_float4(n[0], n[1], n[2], n[3]);
                de[c4] = make_float4(o[0], o[1], o[2], o[3]);
            }
    }
}

// layer1 backward: given de (grad wrt the dropout output), saved yn / e / rstd:
//   dyhat = de * [e > 0] * keep_scale      (written back over de; dgamma = colsum(dyhat * yn), dbeta = colsum(dyhat))
//   dy = rstd * (g - mean(g) - yn * mean(g * yn)),  g = dyhat * gamma
// e == 0 exactly where relu clipped or dropout dropped (a kept positive activation is > 0), so the combined
// mask is (e > 0) and the surviving scale is keep_scale.
template <int E>
__global__ void __launch_bounds__(256)
ln_relu_dropout_bwd(float* __restrict__ de, const float* __restrict__ e, const float* __restrict__ yn,
                    const float* __restrict__ rstd_in, const float* __restrict__ gamma, float* __restrict__ dy,
                    int64_t M, float p_drop) {
    constexpr int NV = E / 128;
    const int lane = threadIdx.x & 31;
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; row < M; row += warps) {
        float4* pde = reinterpret_cast<float4*>(de + row * E);
        const float4* pe = reinterpret_cast<const float4*>(e + row * E);
        const float4* pyn = reinterpret_cast<const float4*>(yn + row * E);
        float4 g[NV], n[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = i * 32 + lane;
            const float4 d = pde[c4], ev = pe[c4], gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            n[i] = pyn[c4];
            float4 dh;
            dh.x = ev.x > 0.f ? d.x * keep_scale : 0.f;
            dh.y = ev.y > 0.f ? d.y * keep_scale : 0.f;
            dh.z = ev.z > 0.f ? d.z * keep_scale : 0.f;
            dh.w = ev.w > 0.f ? d.w * keep_scale : 0.f;
            pde[c4] = dh;
            g[i] = make_float4(dh.x * gm.x, dh.y * gm.y, dh.z * gm.z, dh.w * gm.w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * n[i].x + g[i].y * n[i].y) + (g[i].z * n[i].z + g[i].w * n[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / static_cast<float>(E), m2 = s2 / static_cast<float>(E), rs = rstd_in[row];
        float4* pdy = reinterpret_cast<float4*>(dy + row * E);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 o;
            o.x = rs * (g[i].x - m1 - n[i].x * m2);
            o.y = rs * (g[i].y - m1 - n[i].y * m2);
            o.z = rs * (g[i].z - m1 - n[i].z * m2);
            o.w = rs * (g[i].w - m1 - n[i].w * m2);
            pdy[i * 32 + lane] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// GRU forward step, train mode: gates from gi[t] (packed cols, b_ih folded) and gh = h_{t-1} W_hh'^T + b_hh'
// (packed cols); saves r, z, n, ghn for the backward pass and writes h_t into hall[t+1] and relu(h_t).
__global__ void __launch_bounds__(256)
gru_gates_train_fwd(const float* __restrict__ gi_t, const float* __restrict__ gh, const float* __restrict__ h_prev,
                    float* __restrict__ h_new, float* __restrict__ hrelu_t, float* __restrict__ r_t, float* __restrict__ z_t,
                    float* __restrict__ n_t, float* __restrict__ ghn_t, int B, int H) {
    const int64_t total = static_cast<int64_t>(B) * H;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / H), u = static_cast<int>(i % H);
        const int p = (u / 64) * 192 + (u % 64);
        const float* gi_p = gi_t + static_cast<int64_t>(b) * (3 * H) + p;
        const float* gh_p = gh + static_cast<int64_t>(b) * (3 * H) + p;
        const float r = sigmoid_f(gi_p[0] + gh_p[0]);
        const float z = sigmoid_f(gi_p[64] + gh_p[64]);
        const float ghn = gh_p[128];
        const float n = tanhf(gi_p[128] + r * ghn);
        const float hn = (h_prev[i] - n) * z + n;
        h_new[i] = hn;
        hrelu_t[i] = fmaxf(hn, 0.f);
        r_t[i] = r; z_t[i] = z; n_t[i] = n; ghn_t[i] = ghn;
    }
}

// GRU backward step: dh = dh_next (carried) + dhrelu_t * [h_t > 0];  produces dgi_t / dgh_t (packed cols) and the
// direct part of dh_{t-1} (dh * z); the W_hh part (dgh_t W_hh') is added by the following GEMM.
__global__ void __launch_bounds__(256)
gru_gates_train_bwd(const float* __restrict__ dh_carry, const float* __restrict__ dhrelu_t, const float* __restrict__ h_t,
                    const float* __restrict__ h_prev, const float* __restrict__ r_t, const float* __restrict__ z_t,
                    const float* __restrict__ n_t, const float* __restrict__ ghn_t, float* __restrict__ dgi_t,
                    float* __restrict__ dgh_t, float* __restrict__ dh_prev, int B, int H) {
    const int64_t total = static_cast<int64_t>(B) * H;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / H), u = static_cast<int>(i % H);
        const int p = (u / 64) * 192 + (u % 64);
        const float dh = dh_carry[i] + (h_t[i] > 0.f ? dhrelu_t[i] : 0.f);
        const float r = r_t[i], z = z_t[i], n = n_t[i], ghn = ghn_t[i];
        const float dn = dh * (1.0f - z);
        const float dz = dh * (h_prev[i] - n);
        const float dan = dn * (1.0f - n * n);
        const float dar = dan * ghn * r * (1.0f - r);
        const float daz = dz * z * (1.0f - z);
        float* gi_p = dgi_t + static_cast<int64_t>(b) * (3 * H) + p;
        float* gh_p = dgh_t + static_cast<int64_t>(b) * (3 * H) + p;
        gi_p[0] = dar; gi_p[64] = daz; gi_p[128] = dan;
        gh_p[0] = dar; gh_p[64] = daz; gh_p[128] = dan * r;
        dh_prev[i] = dh * z;
    }
}

// C[m, n] = (bias[n]) + sum_z part[z][m][n]   (second half of a split-K product; slices summed in z order)
__global__ void __launch_bounds__(256)
splitk_reduce_f32(const float* __restrict__ part, int splits, int64_t split_stride, const float* __restrict__ bias,
                  float* __restrict__ C, int64_t ldc, int M, int N) {
    const int64_t total = static_cast<int64_t>(M) * N;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
        float s = bias != nullptr ? __ldg(bias + n) : 0.f;
        for (int z = 0; z < splits; ++z) s += part[z * split_stride + i];
        C[static_cast<int64_t>(m) * ldc + n] = s;
    }
}

// out[y][c] = sum over the rows of slab y of in[r, c] (* mul[r, c])   (bias / LayerNorm-affine gradients).  One block per
// 32 columns x row slab, 8 row-lanes; gridDim.y slabs of rows_per_slab rows.  With gridDim.y == 1 `out` is the result;
// otherwise colsum_final_f32 adds the slabs in order.
__global__ void __launch_bounds__(256)
colsum_f32(const float* __restrict__ in, const float* __restrict__ mul, int64_t ld, float* __restrict__ out,
           int64_t rows, int cols, int64_t rows_per_slab) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    const int64_t r0 = blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float s0 = 0.f, s1 = 0.f;
    if (c < cols) {
        int64_t r = r0 + rl;
        if (mul != nullptr) {
            for (; r + 8 < r1; r += 16) {
                s0 = fmaf(in[r * ld + c], mul[r * ld + c], s0);
                s1 = fmaf(in[(r + 8) * ld + c], mul[(r + 8) * ld + c], s1);
            }
            if (r < r1) s0 = fmaf(in[r * ld + c], mul[r * ld + c], s0);
        } else {
            for (; r + 8 < r1; r += 16) {
                s0 += in[r * ld + c];
                s1 += in[(r + 8) * ld + c];
            }
            if (r < r1) s0 += in[r * ld + c];
        }
    }
    part[rl][threadIdx.x & 31] = s0 + s1;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
        out[static_cast<int64_t>(blockIdx.y) * cols + c] = t;
    }
}

__global__ void colsum_final_f32(const float* __restrict__ part, int slabs, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int y = 0; y < slabs; ++y) s += part[static_cast<int64_t>(y) * cols + c];
    out[c] = s;
}

// Packed -> original row order for GRU weight / bias gradients:  dst[orig_row, :] = src[packed_row, :].
__global__ void unpack_rows_f32(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int H) {
    const int64_t total = static_cast<int64_t>(rows) * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        dst[static_cast<int64_t>(packed_to_orig_row(p, H)) * cols + c] = src[i];
    }
}

// Operand split of the 3xTF32 products (PREGO_PREC_TF32X3): every fp32 value a = hi + lo with hi = a rounded to TF32 (10-bit
// mantissa) and lo = a - hi (exact in fp32; the tensor core reads its leading 11 bits).  The contraction is cut into chunks of
// kc original columns; chunk c of an activation-side row is stored as [lo | hi | hi] (3 kc values), of a weight-side row as
// [hi | lo | hi], so ONE kind::tf32 product over the 3 kc columns of a chunk sums lo hi + hi lo + hi hi -- the two small
// terms FIRST, while the accumulator is still small.  Measured (scripts/diag_x3_accuracy.py, K = 2048): the tensor core's
// accumulator drops low bits of what is added to a large running sum -- one product over 3K with the small terms last: 2.1e-5
// relative (Frobenius); small terms first: 4.9e-6; chunks of 1024 added in fp32 by the epilogue: ~1e-6 = cuBLAS fp32 (8e-7).
// src [rows, K] with row stride lds -> dst [rows, 3K]; K % kc == 0, kc % 4 == 0, lds % 4 == 0.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__global__ void __launch_bounds__(256)
split3_tf32(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t rows, int K, int kc, int weight_side) {
    const int k4 = K / 4;
    const int64_t total = rows * k4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / k4;
        const int c = static_cast<int>(i % k4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        float* d = dst + r * 3 * K + static_cast<int64_t>(c / kc) * 3 * kc + (c % kc);
        *reinterpret_cast<float4*>(d) = weight_side ? hi : lo;
        *reinterpret_cast<float4*>(d + kc) = weight_side ? lo : hi;
        *reinterpret_cast<float4*>(d + 2 * kc) = hi;
    }
}

// dst[c, r] = src[r, c] through 32 x 32 shared-memory tiles (both sides coalesced).  The TN / NN operand forms of the
// backward pass become the K-major NT form of the tensor-core GEMM through these (a few hundred MB per step, HBM-bound).
__global__ void __launch_bounds__(256)
transpose_tiled_f32(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 8 * j, c = c0 + tx;
        tile[ty + 8 * j][tx] = (r < rows && c < cols) ? src[static_cast<int64_t>(r) * lds + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + 8 * j, r = r0 + tx;
        if (c < cols && r < rows) dst[static_cast<int64_t>(c) * ldd + r] = tile[tx][ty + 8 * j];
    }
}

// Fused AdamW over up to 16 tensors in one launch (main.py:62-67; torch.optim.AdamW's single-tensor update order).
struct AdamWTensors {
    float* p[16];
    const float* g[16];
    float* m[16];
    float* v[16];
    int64_t start[17];  // prefix sums of numel (float4-granular work split uses the flat index)
    int n;
};
__global__ void __launch_bounds__(256)
adamw_fused(AdamWTensors t, float lr, float beta1, float beta2, float eps, float wd, float grad_scale, float bc1, float bc2_sqrt) {
    const int64_t total = t.start[t.n];
    const float step_size = lr / bc1, decay = 1.0f - lr * wd;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 16; ++j) k += (j < t.n && i >= t.start[j]) ? 1 : 0;
        const int64_t o = i - t.start[k];
        const float g = t.g[k][o] * grad_scale;
        float p = t.p[k][o] * decay;
        float m = t.m[k][o];
        m = m + (1.0f - beta1) * (g - m);                     // lerp_(grad, 1 - beta1)
        const float v = beta2 * t.v[k][o] + (1.0f - beta2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p = p - step_size * (m / denom);
        t.p[k][o] = p;
        t.m[k][o] = m;
        t.v[k][o] = v;
    }
}

// time-major [T, B, K] -> caller's [B, T, K] (logits out) and back (dlogits in).
__global__ void transpose_tb_f32(const float* __restrict__ src, float* __restrict__ dst, int B, int T, int K, int to_bt) {
    const int64_t total = static_cast<int64_t>(B) * T * K;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % K);
        const int64_t m = i / K;  // row index in the SOURCE ordering
        if (to_bt) {  // src row = t*B + b
            const int t = static_cast<int>(m / B), b = static_cast<int>(m % B);
            dst[(static_cast<int64_t>(b) * T + t) * K + k] = src[i];
        } else {      // src row = b*T + t
            const int b = static_cast<int>(m / T), t = static_cast<int>(m % T);
            dst[(static_cast<int64_t>(t) * B + b) * K + k] = src[i];
        }
    }
}

}  // namespace prego
