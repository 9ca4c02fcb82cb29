// Per-frame average precision on the device (reference: perframe_average_precision, utils/metrics.py:25-62 with
// metrics == 'AP' -> sklearn.metrics.average_precision_score per class; called from trainer/eval.py:67-76,124-141).
//
// For class k the reference sorts the N frame scores descending, takes the precision / recall pair at every DISTINCT
// score threshold and sums (R_n - R_{n-1}) * P_n in float64.  Here: one CTA per class.
//   1. ap_build_keys: key[k][n] = (float bits of score[n][k]) << 1 | positive.  Probabilities live in [0, 1], so their
//      bit patterns are < 2^30, order like the values, and the positive flag rides in bit 0 (inside a group of equal
//      scores the order is irrelevant: only group ends contribute).  [N, K] -> [K, N] through a shared-memory tile.
//   2. ap_sort_scan_kernel: four stable 8-bit LSD radix passes over the class's keys (ping-pong in the workspace),
//      then one descending scan: block prefix sums give tp at every position, a running maximum over group ends gives
//      the tp of the previous threshold.
// HBM-bound integer work: 4 passes x 8 B per key (+ 4 B to build, 4 B to scan).
#pragma once

#include <cstdint>

namespace prego {

constexpr int kApThreads = 1024;
constexpr int kApWarps = kApThreads / 32;
constexpr int kApItems = 4;                          // keys per thread per tile (striped: warp w, item i, lane l)
constexpr int kApTile = kApThreads * kApItems;       // 4096 keys
constexpr uint32_t kApOneBits = 0x3F800000u;         // 1.0f

__global__ void __launch_bounds__(256)
ap_build_keys(const float* __restrict__ scores, const float* __restrict__ targets, const int32_t* __restrict__ target_labels,
              int64_t N, int K, uint32_t* __restrict__ keys, int* __restrict__ err_flag) {
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int k0 = blockIdx.y * 32;
    for (int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32; n0 < N; n0 += static_cast<int64_t>(gridDim.x) * 32) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t n = n0 + r;
            const int k = k0 + tx;
            uint32_t key = 0;
            if (n < N && k < K) {
                const uint32_t bits = __float_as_uint(scores[n * K + k]);
                if (bits > kApOneBits) atomicExch(err_flag, 1);  // negative, > 1, NaN or inf: not a probability
                const bool pos = targets != nullptr ? (targets[n * K + k] != 0.f) : (target_labels[n] == k);
                key = (bits << 1) | (pos ? 1u : 0u);
            }
            tile[r][tx] = key;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int k = k0 + r;
            const int64_t n = n0 + tx;
            if (n < N && k < K) keys[static_cast<int64_t>(k) * N + n] = tile[tx][r];
        }
        __syncthreads();
    }
}

// Block-wide inclusive sum / exclusive max over one value per thread (1024 threads); `total` = reduction of the block.
__device__ __forceinline__ uint32_t ap_block_incl_sum(uint32_t v, uint32_t* wsum, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        wsum[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = warp > 0 ? wsum[warp - 1] : 0u;
    *total = wsum[kApWarps - 1];
    __syncthreads();  // wsum may be reused by the caller
    return v + before;
}

__device__ __forceinline__ uint32_t ap_block_excl_max(uint32_t v, uint32_t* wmax, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, u);
    }
    uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0u;
    if (lane == 31) wmax[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = wmax[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s = max(s, u);
        }
        wmax[lane] = s;
    }
    __syncthreads();
    const uint32_t before = warp > 0 ? wmax[warp - 1] : 0u;
    *total = wmax[kApWarps - 1];
    __syncthreads();
    return max(exc, before);
}

// One CTA per class.  keys_a / keys_b: [K, N] ping-pong; after the four passes the ascending order is back in keys_a.
__global__ void __launch_bounds__(kApThreads)
ap_sort_scan_kernel(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ keys_b, int64_t N, double* __restrict__ ap,
                    int64_t* __restrict__ num_pos) {
    __shared__ uint32_t base[256];             // running output cursor of every digit
    __shared__ uint32_t whist[kApWarps][256];  // per-warp digit counts of the tile -> per-warp output offsets
    __shared__ uint32_t wred[kApWarps];
    __shared__ double dred[kApWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* in = keys_a + static_cast<int64_t>(blockIdx.x) * N;
    uint32_t* out = keys_b + static_cast<int64_t>(blockIdx.x) * N;

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        // 1. digit histogram of the whole class (warp-aggregated shared atomics: the top byte is heavily skewed)
        if (tid < 256) base[tid] = 0;
        __syncthreads();
        for (int64_t i0 = 0; i0 < N; i0 += kApThreads) {
            const int64_t i = i0 + tid;
            const uint32_t d = i < N ? ((in[i] >> shift) & 0xFFu) : 0x100u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (d < 256u && (peers & lt_mask) == 0u) atomicAdd(&base[d], __popc(peers));
        }
        __syncthreads();
        // 2. exclusive prefix over the 256 digits (warp 0, 8 digits per lane)
        if (warp == 0) {
            uint32_t c[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = base[lane * 8 + j];
                s += c[j];
            }
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            uint32_t run = inc - s;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                base[lane * 8 + j] = run;
                run += c[j];
            }
        }
        __syncthreads();
        // 3. stable scatter, tile by tile; inside a tile key order = (warp, item, lane)
        for (int64_t ts = 0; ts < N; ts += kApTile) {
#pragma unroll
            for (int j = 0; j < 256 * kApWarps / kApThreads; ++j) (&whist[0][0])[j * kApThreads + tid] = 0u;
            __syncthreads();
            uint32_t key[kApItems], rank[kApItems];
#pragma unroll
            for (int it = 0; it < kApItems; ++it) {
                const int64_t i = ts + warp * (32 * kApItems) + it * 32 + lane;
                const bool valid = i < N;
                key[it] = valid ? in[i] : 0u;
                const uint32_t d = valid ? ((key[it] >> shift) & 0xFFu) : 0x100u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t prior = valid ? whist[warp][d] : 0u;
                __syncwarp();
                const uint32_t r = __popc(peers & lt_mask);
                if (valid && r == 0u) whist[warp][d] = prior + __popc(peers);
                __syncwarp();
                rank[it] = prior + r;
            }
            __syncthreads();
            if (tid < 256) {  // digit tid: counts of the 32 warps -> offsets, cursor advanced past the tile
                uint32_t run = base[tid];
#pragma unroll 8
                for (int w = 0; w < kApWarps; ++w) {
                    const uint32_t c = whist[w][tid];
                    whist[w][tid] = run;
                    run += c;
                }
                base[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kApItems; ++it) {
                const int64_t i = ts + warp * (32 * kApItems) + it * 32 + lane;
                if (i < N) out[whist[warp][(key[it] >> shift) & 0xFFu] + rank[it]] = key[it];
            }
            __syncthreads();
        }
        uint32_t* t = in;
        in = out;
        out = t;
    }
    // `in` = keys_a again, ascending.  Positives of the class:
    uint32_t cnt = 0;
    for (int64_t i = tid; i < N; i += kApThreads) cnt += in[i] & 1u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) wred[warp] = cnt;
    __syncthreads();
    uint32_t P = 0;
    for (int w = 0; w < kApWarps; ++w) P += wred[w];
    __syncthreads();
    if (P == 0u) {  // the reference skips classes without positives (metrics.py:54)
        if (tid == 0) {
            ap[blockIdx.x] = __longlong_as_double(0x7ff8000000000000LL);
            num_pos[blockIdx.x] = 0;
        }
        return;
    }
    // descending scan: position p = 0 is the highest score
    const double Pd = static_cast<double>(P);
    double acc = 0.0;
    uint32_t tp_carry = 0u, end_carry = 0u;  // tp up to the previous tile; tp at the last threshold so far
    for (int64_t ts = 0; ts < N; ts += kApThreads) {
        const int64_t p = ts + tid;
        const bool valid = p < N;
        const uint32_t key = valid ? in[N - 1 - p] : 0u;
        const bool is_end = valid && (p == N - 1 || (in[N - 2 - p] >> 1) != (key >> 1));
        uint32_t tile_tp, tile_end;
        const uint32_t tp = tp_carry + ap_block_incl_sum(valid ? (key & 1u) : 0u, wred, &tile_tp);
        const uint32_t prev_end = max(end_carry, ap_block_excl_max(is_end ? tp : 0u, wred, &tile_end));
        if (is_end && tp > prev_end) {
            // sklearn: precision = tps / (tps + fps), recall = tps / tps[-1], ap = sum diff(recall) * precision
            const double prec = static_cast<double>(tp) / static_cast<double>(p + 1);
            acc += (static_cast<double>(tp) / Pd - static_cast<double>(prev_end) / Pd) * prec;
        }
        tp_carry += tile_tp;
        end_carry = max(end_carry, tile_end);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dred[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kApWarps; ++w) s += dred[w];
        ap[blockIdx.x] = s;
        num_pos[blockIdx.x] = static_cast<int64_t>(P);
    }
}

}  // namespace prego
