// CUDA-core kernels of the MiniROAD path: feature staging, LayerNorm+ReLU, the exact-fp32
// GEMM used by the PREGO_PREC_FP32 mode, the element-wise GRU gate update of that mode,
// and softmax/argmax.  All are HBM- or FFMA-bound helpers around the tcgen05 GEMMs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "gemm_tc.cuh"  // Op16, sigmoid_f

namespace prego {

// Row m of a time chunk -> row of the caller's [B, T, D] tensor.
__device__ __forceinline__ int64_t chunk_row_to_global(int64_t m, int Tc, int T, int t0) {
    return (m / Tc) * static_cast<int64_t>(T) + t0 + (m % Tc);
}

// ---------------------------------------------------------------------------------------
// Feature staging (replaces torch.cat of rnn.py:53 and the fp32->bf16 operand rounding):
//   xb[m, 0:Dr] = op16(rgb[b, t0+t, :]),  xb[m, Dr:Dr+Df] = op16(flow[b, t0+t, :]),  m = t*B + b
// Rows are written TIME-MAJOR inside the chunk so that every later stage (both projections, the
// per-step recurrence tiles) touches contiguous rows.  One thread converts 8 consecutive elements
// (2 x 16 B loads -> one 16 B store).
template <int FMT>
__global__ void __launch_bounds__(256)
stage_features_16(const float* __restrict__ rgb, const float* __restrict__ flow, typename Op16<FMT>::T* __restrict__ xb,
                  int64_t Mc, int Dr, int Df, int B, int T, int t0) {
    using Op = Op16<FMT>;
    const int D = Dr + Df;
    const int vec_per_row = D / 8;
    const int64_t total = Mc * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / vec_per_row;
        const int c = static_cast<int>(i % vec_per_row) * 8;
        const int64_t g = (m % B) * static_cast<int64_t>(T) + t0 + (m / B);
        const float* src = (c < Dr) ? (rgb + g * Dr + c) : (flow + g * Df + (c - Dr));
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 1);
        uint4 o;
        o.x = Op::pack2(a.x, a.y);
        o.y = Op::pack2(a.z, a.w);
        o.z = Op::pack2(b.x, b.y);
        o.w = Op::pack2(b.z, b.w);
        *reinterpret_cast<uint4*>(xb + m * D + c) = o;
    }
}

// Same rearrangement for features the caller already keeps in the 16-bit operand format (PREGO_FEAT_16): pure
// 16-byte copies.  Used when the projection GEMM cannot read the caller's tensors in place (B % 128 != 0).
__global__ void __launch_bounds__(256)
stage_features_16from16(const uint16_t* __restrict__ rgb, const uint16_t* __restrict__ flow, uint16_t* __restrict__ xb,
                        int64_t Mc, int Dr, int Df, int B, int T, int t0) {
    const int D = Dr + Df;
    const int vec_per_row = D / 8;
    const int64_t total = Mc * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / vec_per_row;
        const int c = static_cast<int>(i % vec_per_row) * 8;
        const int64_t g = (m % B) * static_cast<int64_t>(T) + t0 + (m / B);
        const uint16_t* src = (c < Dr) ? (rgb + g * Dr + c) : (flow + g * Df + (c - Dr));
        *reinterpret_cast<uint4*>(xb + m * D + c) = __ldcs(reinterpret_cast<const uint4*>(src));
    }
}

// ---------------------------------------------------------------------------------------
// LayerNorm (biased variance, eps inside the sqrt) + ReLU over rows of width E
// (rnn.py:41-42).  One warp per row, the row lives in registers; two-pass variance.
// In-place safe (each warp reads its whole row before writing).
// y is stored as fp16 (pre-activation, |y| = O(1..10)); e is written in the operand format FMT.
template <int E, int FMT>
__global__ void __launch_bounds__(256)
layernorm_relu_16(const __half* __restrict__ y, typename Op16<FMT>::T* __restrict__ e, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int64_t M, float eps) {
    using Op = Op16<FMT>;
    constexpr int kVec = E / (32 * 8);  // 16-byte vectors per lane
    const int lane = threadIdx.x & 31;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; row < M; row += warps) {
        const uint4* src = reinterpret_cast<const uint4*>(y + row * E);
        float v[kVec * 8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const uint4 u = src[i * 32 + lane];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
                v[i * 8 + 2 * j] = p.x;
                v[i * 8 + 2 * j + 1] = p.y;
                sum += p.x + p.y;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum * (1.0f / E);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kVec * 8; ++i) {
            const float d = v[i] - mu;
            sq += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.0f / sqrtf(sq * (1.0f / E) + eps);
        uint4* dst = reinterpret_cast<uint4*>(e + row * E);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const int col = (i * 32 + lane) * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col) + 1);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf((v[i * 8 + j] - mu) * rstd * gg[j] + bb[j], 0.0f);
            uint4 u;
            u.x = Op::pack2(o[0], o[1]);
            u.y = Op::pack2(o[2], o[3]);
            u.z = Op::pack2(o[4], o[5]);
            u.w = Op::pack2(o[6], o[7]);
            dst[i * 32 + lane] = u;
        }
    }
}

// fp32 variant for the exact mode; generic width (E multiple of 128, E <= 4096).
__global__ void __launch_bounds__(256)
layernorm_relu_f32(const float* __restrict__ y, float* __restrict__ e, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int64_t M, int E, float eps) {
    constexpr int kMaxVec = 32;  // float4 per lane -> E <= 4096
    const int lane = threadIdx.x & 31;
    const int nvec = E / 128;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; row < M; row += warps) {
        const float4* src = reinterpret_cast<const float4*>(y + row * E);
        float4 v[kMaxVec];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxVec; ++i) {
            if (i < nvec) {
                v[i] = src[i * 32 + lane];
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum / static_cast<float>(E);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxVec; ++i) {
            if (i < nvec) {
                const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
                sq += (a * a + b * b) + (c * c + d * d);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.0f / sqrtf(sq / static_cast<float>(E) + eps);
        float4* dst = reinterpret_cast<float4*>(e + row * E);
#pragma unroll
        for (int i = 0; i < kMaxVec; ++i) {
            if (i < nvec) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
                float4 o;
                o.x = fmaxf((v[i].x - mu) * rstd * g.x + b.x, 0.f);
                o.y = fmaxf((v[i].y - mu) * rstd * g.y + b.y, 0.f);
                o.z = fmaxf((v[i].z - mu) * rstd * g.z + b.z, 0.f);
                o.w = fmaxf((v[i].w - mu) * rstd * g.w + b.w, 0.f);
                dst[i * 32 + lane] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Exact fp32 GEMM on CUDA cores:  C[M, N] = [A0 | A1][M, K] * W[N, K]^T + bias.
// 128x128 tile, BK = 16, 256 threads, 8x8 register tile.  A is optionally split in two
// column segments (rgb | flow) and its rows optionally remapped from chunk order to the
// caller's [B, T, D] order, so no concatenated copy is ever made.
struct SgemmA {
    const float* a0;
    const float* a1;  // may be nullptr
    int k_split;      // columns [0, k_split) come from a0, the rest from a1
    int64_t lda0, lda1;
    int remap;        // 1: row m -> chunk_row_to_global(m, Tc, T, t0)
    int Tc, T, t0;
    int64_t ldw = 0;  // row stride of W (0 = K)
    int relu = 0;     // C = max(acc + bias, 0)
};

__global__ void __launch_bounds__(256)
sgemm_nt_f32(SgemmA A, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C, int M,
             int N, int K, int64_t ldc) {
    constexpr int BM = 128, BN = 128, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int tx = tid % 16;  // column group
    const int ty = tid / 16;  // row group
    const int64_t ldw = A.ldw > 0 ? A.ldw : K;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // each thread loads 2 float4 of A and 2 float4 of W per k-block: rows r = tid/4 (+64), k = (tid%4)*4
    const int lr = tid / 4;
    const int lk = (tid % 4) * 4;
    int64_t arow[2];
    bool aval[2], wval[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = m0 + lr + 64 * h;
        aval[h] = m < M;
        const int64_t mm = aval[h] ? m : 0;
        arow[h] = A.remap ? chunk_row_to_global(mm, A.Tc, A.T, A.t0) : mm;
        wval[h] = (n0 + lr + 64 * h) < N;
    }

    for (int k0 = 0; k0 < K; k0 += BK) {
        const int k = k0 + lk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (aval[h]) {
                const float* p = (k < A.k_split) ? (A.a0 + arow[h] * A.lda0 + k) : (A.a1 + arow[h] * A.lda1 + (k - A.k_split));
                a = *reinterpret_cast<const float4*>(p);
            }
            As[lk + 0][lr + 64 * h] = a.x;
            As[lk + 1][lr + 64 * h] = a.y;
            As[lk + 2][lr + 64 * h] = a.z;
            As[lk + 3][lr + 64 * h] = a.w;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (wval[h]) w = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(n0 + lr + 64 * h) * ldw + k));
            Ws[lk + 0][lr + 64 * h] = w.x;
            Ws[lk + 1][lr + 64 * h] = w.y;
            Ws[lk + 2][lr + 64 * h] = w.z;
            Ws[lk + 3][lr + 64 * h] = w.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], w[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
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
            if (n < N) {
                const float v = acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
                C[static_cast<int64_t>(m) * ldc + n] = A.relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Element-wise GRU gate update of the exact-fp32 mode (one thread per (stream, unit)).
// gh = h W_hh'^T + b_hh' was produced by sgemm_nt_f32 in the packed gate-interleaved order.
__global__ void __launch_bounds__(256)
gru_gates_f32(const float* __restrict__ gi, const float* __restrict__ gh, float* __restrict__ h32,
              float* __restrict__ hrelu, int B, int H, int Tc, int t) {
    const int64_t total = static_cast<int64_t>(B) * H;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / H);
        const int u = static_cast<int>(i % H);
        const int p = (u / 64) * 192 + (u % 64);
        const float* gi_p = gi + (static_cast<int64_t>(b) * Tc + t) * (3 * H) + p;
        const float* gh_p = gh + static_cast<int64_t>(b) * (3 * H) + p;
        const float r = sigmoid_f(gi_p[0] + gh_p[0]);
        const float z = sigmoid_f(gi_p[64] + gh_p[64]);
        const float n = tanhf(gi_p[128] + r * gh_p[128]);
        const float hn = (h32[i] - n) * z + n;
        h32[i] = hn;
        hrelu[(static_cast<int64_t>(b) * Tc + t) * H + u] = fmaxf(hn, 0.f);
    }
}

// ---------------------------------------------------------------------------------------
// softmax + first-max argmax over K classes, one warp per frame (exact-fp32 mode head).
__global__ void __launch_bounds__(256)
softmax_argmax_f32(const float* __restrict__ logits_chunk, float* __restrict__ probs, float* __restrict__ logits_out,
                   int32_t* __restrict__ labels, int64_t Mc, int K, int Tc, int T, int t0, int A = 1, int64_t row0 = 0) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t m = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; m < Mc; m += warps) {
        const float* src = logits_chunk + m * K;
        // anticipation head: chunk row = (row0 + m) / A, output rows [B, T, A]
        const int64_t g = A == 1 ? chunk_row_to_global(m, Tc, T, t0)
                                 : chunk_row_to_global((row0 + m) / A, Tc, T, t0) * A + (row0 + m) % A;
        float mx = -INFINITY;
        for (int j = lane; j < K; j += 32) mx = fmaxf(mx, src[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < K; j += 32) sum += expf(src[j] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float best = -1.f;
        int arg = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
            const float l = src[j];
            const float p = expf(l - mx) / sum;
            if (probs != nullptr) probs[g * K + j] = p;
            if (logits_out != nullptr) logits_out[g * K + j] = l;
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

// ---------------------------------------------------------------------------------------
// Weight packing helpers (run once per load).
// Gate-interleaved row order: packed row p = nt*192 + g*64 + j  <->  original row g*H + nt*64 + j.
__device__ __forceinline__ int packed_to_orig_row(int p, int H) {
    const int nt = p / 192, rem = p % 192;
    return (rem / 64) * H + nt * 64 + (rem % 64);
}

__global__ void pack_rows_f32(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int H,
                              int permute) {
    const int64_t total = static_cast<int64_t>(rows) * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        const int r = permute ? packed_to_orig_row(p, H) : p;
        dst[i] = src[static_cast<int64_t>(r) * cols + c];
    }
}

// fp32 GRU state between the caller's [B, H] rows and the batched recurrence's tiled order
// [ceil(B/128)][H/64][16][128][4] (see gru_step.cuh).  to_tiled = 1: plain -> tiled (padding rows zeroed).
__global__ void h32_retile(float* __restrict__ plain, float* __restrict__ tiled, int B, int H, int to_tiled) {
    const int64_t rows_pad = (static_cast<int64_t>(B) + 127) / 128 * 128;
    const int64_t total = rows_pad * (H / 4);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // i enumerates float4 slots of the TILED buffer
        const int r = static_cast<int>(i % 128);
        int64_t q = i / 128;
        const int c16 = static_cast<int>(q % 16); q /= 16;
        const int nt = static_cast<int>(q % (H / 64));
        const int64_t mb = q / (H / 64);
        const int64_t row = mb * 128 + r;
        float4* tp = reinterpret_cast<float4*>(tiled) + i;
        float4* pp = reinterpret_cast<float4*>(plain + row * H + nt * 64 + c16 * 4);
        if (to_tiled) *tp = row < B ? *pp : make_float4(0.f, 0.f, 0.f, 0.f);
        else if (row < B) *pp = *tp;
    }
}

// dst[c, r] = src[r, c]  (classifier weight [K, H] -> [H, K] for the split-K head of the per-frame kernel)
__global__ void transpose_f32(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    const int64_t total = static_cast<int64_t>(rows) * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        dst[static_cast<int64_t>(c) * rows + r] = src[i];
    }
}

// bgi[p] = bih[p] + (gate(p) is r or z ? bhh[p] : 0), packed order: columns [r64 | z64 | n64] per 192.
__global__ void presum_gate_bias(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ bgi, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bgi[i] = bih[i] + ((i % 192) < 128 ? bhh[i] : 0.f);
}

// dst rows beyond src_rows are zero (class padding of the head weight).
template <int FMT>
__global__ void pack_rows_16(const float* __restrict__ src, typename Op16<FMT>::T* __restrict__ dst, int rows,
                             int src_rows, int cols, int H, int permute) {
    const int64_t total = static_cast<int64_t>(rows) * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        const int r = permute ? packed_to_orig_row(p, H) : p;
        dst[i] = Op16<FMT>::from_float(r < src_rows ? src[static_cast<int64_t>(r) * cols + c] : 0.f);
    }
}

// Plain fp32 -> 16-bit operand conversion (weights; slot 0 of the state history = carried h).
template <int FMT>
__global__ void f32_to_16(const float* __restrict__ src, typename Op16<FMT>::T* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[i] = Op16<FMT>::from_float(src[i]);
}

// ---------------------------------------------------------------------------------------
// Split-fp16 operands ("fp16x3", the fp32-class tensor-core mode): an fp32 value v is carried as hi = fp16(v) and
// lo = fp16(v - hi) (22 significant bits together), and a product x . w is evaluated as
//     x_hi w_hi + x_lo w_hi + x_hi w_lo                 (the lo . lo term, 2^-22 relative, is dropped)
// by ONE tcgen05 GEMM over a three times longer K: activation rows [lo | hi | hi], weight rows [hi | lo | hi] -- the two
// small terms are contracted FIRST, while the accumulator is still small: the tensor core's accumulator drops low bits of
// what is added to a large running sum (scripts/diag_x3_accuracy.py: 4x smaller error than with the small terms last).
// fp16 x fp16 products are exact in the fp32 accumulator.  Weights are pre-scaled by a power of two so that their lo
// parts stay in fp16's normal range; the epilogue multiplies the accumulator by the inverse (exact).
__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
    const float c = fminf(fmaxf(v, -65504.f), 65504.f);
    hi = __float2half_rn(c);
    lo = __float2half_rn(c - __half2float(hi));
}

// dst[m, 0:K] = lo, dst[m, K:2K] = hi, dst[m, 2K:3K] = hi of src[m, :] (row-major fp32 [M, K]); 4 elements per thread.
__global__ void __launch_bounds__(256)
split3_rows_f32(const float* __restrict__ src, __half* __restrict__ dst, int64_t M, int K) {
    const int vec = K / 4;
    const int64_t total = M * vec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / vec;
        const int c = static_cast<int>(i % vec) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + m * K + c);
        __half h[4], l[4];
        split_hi_lo(v.x, h[0], l[0]); split_hi_lo(v.y, h[1], l[1]); split_hi_lo(v.z, h[2], l[2]); split_hi_lo(v.w, h[3], l[3]);
        __half* d = dst + m * 3 * static_cast<int64_t>(K) + c;
        *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(l);
        *reinterpret_cast<uint2*>(d + K) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(d + 2 * K) = *reinterpret_cast<const uint2*>(h);
    }
}

// Feature staging of the split mode: chunk row m (STREAM-major, m = b * Tc + t, as in the exact-fp32 path) of [rgb | flow]
// -> [lo | hi | hi] of width 3 D.
__global__ void __launch_bounds__(256)
stage_features_split3(const float* __restrict__ rgb, const float* __restrict__ flow, __half* __restrict__ dst, int64_t Mc, int Dr,
                      int Df, int Tc, int T, int t0) {
    const int D = Dr + Df, vec = D / 4;
    const int64_t total = Mc * vec;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / vec;
        const int c = static_cast<int>(i % vec) * 4;
        const int64_t g = chunk_row_to_global(m, Tc, T, t0);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // flow == nullptr: the caller declared the flow stream all-zero
        if (c < Dr) v = __ldcs(reinterpret_cast<const float4*>(rgb + g * Dr + c));
        else if (flow != nullptr) v = __ldcs(reinterpret_cast<const float4*>(flow + g * Df + (c - Dr)));
        __half h[4], l[4];
        split_hi_lo(v.x, h[0], l[0]); split_hi_lo(v.y, h[1], l[1]); split_hi_lo(v.z, h[2], l[2]); split_hi_lo(v.w, h[3], l[3]);
        __half* d = dst + m * 3 * static_cast<int64_t>(D) + c;
        *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(l);
        *reinterpret_cast<uint2*>(d + D) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(d + 2 * D) = *reinterpret_cast<const uint2*>(h);
    }
}

// Weights of the split mode: dst[p, 0:K] = hi, [K:2K] = lo, [2K:3K] = hi of scale * src[row(p), :]; rows optionally in the
// packed gate-interleaved order; columns [0, col_limit) of src only (zero-flow elision is not offered in this mode: col_limit = cols).
__global__ void pack_rows_split3(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols, int H, int permute, float scale) {
    const int64_t total = static_cast<int64_t>(rows) * cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        const int r = permute ? packed_to_orig_row(p, H) : p;
        __half hi, lo;
        split_hi_lo(src[static_cast<int64_t>(r) * cols + c] * scale, hi, lo);
        __half* d = dst + static_cast<int64_t>(p) * 3 * cols + c;
        d[0] = hi;
        d[cols] = lo;
        d[2 * cols] = hi;
    }
}

// max |src[i]| as the bit pattern of a non-negative float (atomicMax on the unsigned view is order-preserving)
__global__ void absmax_f32(const float* __restrict__ src, int64_t n, unsigned* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) m = fmaxf(m, fabsf(src[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

}  // namespace prego
