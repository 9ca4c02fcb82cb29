// Frame -> step collapse on the GPU (reference: utils/aggregate.py).
//
//   window_mode_kernel : aggregate.py:55,65-72  -- every non-overlapping `window`-frame block
//                        of a stream is replaced by its mode, lowest label on ties
//                        (np.argmax(np.bincount(block))).  One warp per (stream, window).
//   rle_kernel         : aggregate.py:7-43      -- run values + change indices (+ final length)
//                        of a ragged batch of integer sequences.  One CTA per sequence,
//                        ballot/popc block scan with a running carry.
//
// Pure integer work: results are bit-exact with the reference.  HBM-bound: 4 B read per frame.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace prego {

constexpr int kAggThreads = 256;
constexpr int kAggMaxLabels = 1024;

// labels: concatenated int32 labels of all streams; offsets[B+1] frame offsets;
// win_offsets[B+1] window offsets (win_offsets[b+1]-win_offsets[b] = ceil(T_b/window)).
// modes[total_windows] out.  err_flag set to 1 if a label is outside [0, num_labels).
__global__ void __launch_bounds__(kAggThreads)
window_mode_kernel(const int32_t* __restrict__ labels, const int64_t* __restrict__ offsets,
                   const int64_t* __restrict__ win_offsets, int B, int window, int num_labels,
                   int32_t* __restrict__ modes, int* err_flag) {
    extern __shared__ int hist_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    int* hist = hist_all + warp * num_labels;
    const int64_t total_windows = win_offsets[B];
    for (int64_t w = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; w < total_windows;
         w += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
        // stream owning window w: largest b with win_offsets[b] <= w
        int lo = 0, hi = B - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (win_offsets[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int b = lo;
        const int64_t start = offsets[b] + (w - win_offsets[b]) * window;
        const int64_t end = min(start + static_cast<int64_t>(window), offsets[b + 1]);
        for (int k = lane; k < num_labels; k += 32) hist[k] = 0;
        __syncwarp();
        for (int64_t i = start + lane; i < end; i += 32) {
            const int32_t v = labels[i];
            if (v < 0 || v >= num_labels) *err_flag = 1;
            else atomicAdd(&hist[v], 1);
        }
        __syncwarp();
        int best_c = -1, best_l = 0x7fffffff;
        for (int k = lane; k < num_labels; k += 32) {
            const int c = hist[k];
            if (c > best_c) {  // ascending k: strict > keeps the lowest label
                best_c = c;
                best_l = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            const int ol = __shfl_xor_sync(0xffffffffu, best_l, o);
            if (oc > best_c || (oc == best_c && ol < best_l)) {
                best_c = oc;
                best_l = ol;
            }
        }
        if (lane == 0) modes[w] = best_l;
        __syncwarp();
    }
}

// Ragged run-length collapse.  Sequence b = in[seg_offsets[b] .. seg_offsets[b+1]).
// Run r of sequence b starts at element i_r: out_vals[seg_offsets[b]+r] = in[i_r] and, for r >= 1,
// out_changes[seg_offsets[b]+r-1] = i_r * scale; the last change entry is final_len[b]
// (len(arr) of aggregate.py:42).  counts[b] = number of runs, or -1 for an empty sequence
// (the reference raises IndexError there, aggregate.py:18).
__global__ void __launch_bounds__(kAggThreads)
rle_kernel(const int32_t* __restrict__ in, const int64_t* __restrict__ seg_offsets,
           const int64_t* __restrict__ final_len, int64_t scale, int32_t* __restrict__ out_vals,
           int64_t* __restrict__ out_changes, int32_t* __restrict__ counts) {
    __shared__ int warp_tot[kAggThreads / 32];
    __shared__ int64_t carry_s;
    const int b = blockIdx.x;
    const int64_t s0 = seg_offsets[b], s1 = seg_offsets[b + 1];
    const int64_t n = s1 - s0;
    if (n <= 0) {
        if (threadIdx.x == 0) counts[b] = -1;
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += kAggThreads) {
        const int64_t i = base + threadIdx.x;
        int32_t v = 0;
        bool head = false;
        if (i < n) {
            v = in[s0 + i];
            head = (i == 0) || (in[s0 + i - 1] != v);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        const int prefix = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kAggThreads / 32; ++w) {
            const int c = warp_tot[w];
            if (w < warp) before += c;
            total += c;
        }
        const int64_t carry = carry_s;
        if (head) {
            const int64_t r = carry + before + prefix;  // run index within the sequence
            out_vals[s0 + r] = v;
            if (r > 0) out_changes[s0 + r - 1] = i * scale;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int64_t runs = carry_s;
        out_changes[s0 + runs - 1] = final_len[b];
        counts[b] = static_cast<int32_t>(runs);
    }
}

}  // namespace prego
