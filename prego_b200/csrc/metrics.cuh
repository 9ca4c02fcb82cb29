// Per-frame average precision on the device (reference: perframe_average_precision, utils/metrics.py:25-62 with
// metrics == 'AP' -> sklearn.metrics.average_precision_score per class; called from trainer/eval.py:67-76,124-141).
//
// For class k the reference sorts the N frame scores descending, takes the precision / recall pair at every DISTINCT
// score threshold and sums (R_n - R_{n-1}) * P_n in float64.  Here, for all classes at once:
//   1. ap_build_keys: key[k][n] = (float bits of score[n][k]) << 1 | positive.  Probabilities live in [0, 1], so their
//      bit patterns are < 2^30, order like the values, and the positive flag rides in bit 0 (inside a group of equal
//      scores the order is irrelevant: only group ends contribute).  [N, K] -> [K, N] through a shared-memory tile.
//   2. four stable 8-bit LSD radix passes, each class cut into S slices so that K * S CTAs fill the machine:
//      ap_hist_kernel (digit counts per slice) -> ap_offsets_kernel (prefix over (digit, slice) per class) ->
//      ap_scatter_kernel (4096-key tiles: stable ranks from ballot-built peer masks + one shared atomic per digit
//      group, the tile is ordered by digit in shared memory and leaves in runs, so the global stores coalesce).
//   3. the precision-recall integral over the descending order, again per slice: ap_scan_local (positives and the tp
//      at the last threshold of each slice) -> ap_scan_final (carries from the slices above, float64 partial sums)
//      -> ap_reduce (fixed-order sum: deterministic).
// HBM-bound integer work: 4 B read + 4 B write to build, 4 x (4 + 4 + 4) B to sort, 8 B to scan = 64 B per (frame, class).
#pragma once

#include <cstdint>

namespace prego {

constexpr int kApThreads = 256;
constexpr int kApWarps = kApThreads / 32;
constexpr int kApItems = 16;                         // keys per thread per tile (striped: warp w, item i, lane l)
constexpr int kApTile = kApThreads * kApItems;       // 4096 keys
constexpr int kApScanItems = 8;                      // consecutive positions per thread in the PR scan
constexpr uint32_t kApOneBits = 0x3F800000u;         // 1.0f
constexpr uint32_t kApNoEnd = 0xFFFFFFFFu;

// Slices per class: enough CTAs to fill 148 SMs several times over, never less than one tile per slice.
inline int ap_num_slices(int64_t N, int K) {
    const int64_t tiles = (N + kApTile - 1) / kApTile;
    int64_t s = (148 * 16 + K - 1) / K;
    if (s > tiles) s = tiles;
    if (s > 1024) s = 1024;
    return s < 1 ? 1 : static_cast<int>(s);
}
inline int64_t ap_slice_len(int64_t N, int S) {
    const int64_t l = (N + S - 1) / S;
    return (l + kApTile - 1) / kApTile * kApTile;
}

__global__ void __launch_bounds__(256)
ap_build_keys(const float* __restrict__ scores, const float* __restrict__ targets, const int32_t* __restrict__ target_labels,
              int64_t N, int K, uint32_t* __restrict__ keys, int* __restrict__ err_flag) {
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int64_t n0 = static_cast<int64_t>(blockIdx.x) * 32; n0 < N; n0 += static_cast<int64_t>(gridDim.x) * 32) {
        // all class tiles of the same 32 frames back to back: the row segments that straddle two tiles are L1/L2 hits
        for (int k0 = 0; k0 < K; k0 += 32) {
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
}

// Lanes of the warp holding the same 8-bit digit, from eight ballots (constant cost; MATCH.ANY is far slower when the
// warp holds many distinct values -- measured: the histogram pass was 90 % SM-bound on it).
__device__ __forceinline__ uint32_t ap_peers8(uint32_t d, bool valid) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int32_t neg = static_cast<int32_t>(d << (31 - b)) >> 31;  // all ones when bit b of d is set
        const uint32_t bal = __ballot_sync(0xffffffffu, neg < 0);
        peers &= ~(bal ^ static_cast<uint32_t>(neg));                    // bit set: &= bal; clear: &= ~bal  (one LOP3)
    }
    return peers;
}

// ---- radix pass.  Grid = K * S; block (k, s) owns keys [s * L, min((s + 1) * L, N)) of class k.
// hist layout [K][256][S]: the S counts of one digit are contiguous for the prefix kernel.
__global__ void __launch_bounds__(kApThreads)
ap_hist_kernel(const uint32_t* __restrict__ keys, int64_t N, int S, int64_t L, int shift, uint32_t* __restrict__ hist) {
    // eight copies of the histogram, copy c shifted by c banks: lanes with the same digit (the top digits of
    // probabilities are heavily skewed) spread over 8 addresses in 8 banks instead of serialising on one
    __shared__ uint32_t h[8 * 257];
    const int k = blockIdx.x / S, s = blockIdx.x % S;
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * 257; i += kApThreads) h[i] = 0;
    __syncthreads();
    uint32_t* mine = h + (tid & 7) * 257;
    const uint32_t* in = keys + static_cast<int64_t>(k) * N;
    const int64_t lo = s * L, hi = lo + L < N ? lo + L : N;
    for (int64_t i0 = lo; i0 < hi; i0 += kApThreads * 8) {
        uint32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // eight independent loads in flight
            const int64_t i = i0 + j * kApThreads + tid;
            d[j] = i < hi ? ((in[i] >> shift) & 0xFFu) : 0x100u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (d[j] < 256u) atomicAdd(&mine[d[j]], 1u);
    }
    __syncthreads();
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) c += h[j * 257 + tid];
    hist[(static_cast<int64_t>(k) * 256 + tid) * S + s] = c;
}

// One CTA per class, thread d = digit d: hist[k][d][s] -> exclusive prefix over (d, s) in that order (in place).
__global__ void __launch_bounds__(256)
ap_offsets_kernel(uint32_t* __restrict__ hist, int S) {
    __shared__ uint32_t wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* row = hist + (static_cast<int64_t>(blockIdx.x) * 256 + tid) * S;
    uint32_t total = 0;
    for (int s = 0; s < S; ++s) total += row[s];
    uint32_t inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t run = inc - total;
    for (int w = 0; w < warp; ++w) run += wsum[w];
    for (int s = 0; s < S; ++s) {
        const uint32_t c = row[s];
        row[s] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(kApThreads, 4)
ap_scatter_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out, int64_t N, int S, int64_t L, int shift,
                  const uint32_t* __restrict__ offsets) {
    __shared__ uint32_t base[256];             // running output cursor of every digit for this slice
    __shared__ uint32_t delta[256];            // output index - tile-sorted index, per digit, for the current tile
    __shared__ uint32_t whist[kApWarps][256];  // per-warp digit counts of the tile -> per-warp tile-sorted offsets
    __shared__ uint32_t sorted[kApTile];       // the tile ordered by digit (stable)
    __shared__ uint32_t wsum[kApWarps];
    const int k = blockIdx.x / S, s = blockIdx.x % S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t* in = keys_in + static_cast<int64_t>(k) * N;
    uint32_t* out = keys_out + static_cast<int64_t>(k) * N;
    base[tid] = offsets[(static_cast<int64_t>(k) * 256 + tid) * S + s];
    const int64_t lo = s * L, hi = lo + L < N ? lo + L : N;
    for (int64_t ts = lo; ts < hi; ts += kApTile) {
#pragma unroll
        for (int j = 0; j < kApWarps; ++j) whist[j][tid] = 0u;
        __syncthreads();
        uint32_t key[kApItems];
        uint16_t rank[kApItems];
        const int64_t w0 = ts + warp * (32 * kApItems) + lane;
#pragma unroll
        for (int it = 0; it < kApItems; ++it) {  // all loads of the tile first (independent, in flight together)
            const int64_t i = w0 + it * 32;
            key[it] = i < hi ? in[i] : 0xFFFFFFFFu;  // real keys are < 2^31
        }
#pragma unroll
        for (int it = 0; it < kApItems; ++it) {  // stable rank inside the warp's 512 keys: order (item, lane)
            const bool valid = key[it] != 0xFFFFFFFFu;
            const uint32_t d = (key[it] >> shift) & 0xFFu;
            const uint32_t peers = ap_peers8(d, valid);
            const uint32_t r = __popc(peers & lt_mask);
            // the group's first lane bumps the warp's digit counter and hands the old value to its peers.  One shared
            // atomic per group instead of a load / store pair: the 16 items do not wait for each other's round trip
            uint32_t prior = 0u;
            if (valid && r == 0u) prior = atomicAdd(&whist[warp][d], static_cast<uint32_t>(__popc(peers)));
            __syncwarp();  // item order = counter order
            prior = __shfl_sync(0xffffffffu, prior, (__ffs(peers) - 1) & 31);
            rank[it] = static_cast<uint16_t>(prior + r);
        }
        __syncthreads();
        {   // digit tid: tile count -> start of the digit inside the digit-sorted tile (block exclusive scan), per-warp
            // offsets inside the tile, and the shift from tile-sorted index to output index
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < kApWarps; ++w) cnt += whist[w][tid];
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t start = inc - cnt;
#pragma unroll
            for (int w = 0; w < kApWarps; ++w)
                if (w < warp) start += wsum[w];
            uint32_t run = start;
#pragma unroll
            for (int w = 0; w < kApWarps; ++w) {
                const uint32_t c = whist[w][tid];
                whist[w][tid] = run;
                run += c;
            }
            delta[tid] = base[tid] - start;
            base[tid] += cnt;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kApItems; ++it)
            if (key[it] != 0xFFFFFFFFu) sorted[whist[warp][(key[it] >> shift) & 0xFFu] + rank[it]] = key[it];
        __syncthreads();
        // coalesced write-out: consecutive threads hold consecutive keys of the same digit run, which are consecutive in
        // the output (a plain scatter costs one 32-byte sector per key on the random middle digits)
        const int n_tile = static_cast<int>(hi - ts < kApTile ? hi - ts : kApTile);
#pragma unroll
        for (int it = 0; it < kApItems; ++it) {
            const int j = it * kApThreads + tid;
            if (j < n_tile) {
                const uint32_t kk = sorted[j];
                out[delta[(kk >> shift) & 0xFFu] + j] = kk;
            }
        }
        __syncthreads();
    }
}

// ---- precision-recall integral over the DESCENDING order: position p = 0 is the highest score, element N - 1 - p.
// Slice s owns positions [s * L, min((s + 1) * L, N)); thread t of a tile owns kApScanItems consecutive positions.
struct ApThreadScan {
    uint32_t tp_local;   // positives among the thread's positions
    uint32_t last_end;   // tp (thread-local, inclusive) at the thread's last threshold, kApNoEnd if none
};

// Loads the thread's positions and classifies them.  y[j] = positive flag, end[j] = position is the last of its group.
__device__ __forceinline__ void ap_load_positions(const uint32_t* __restrict__ in, int64_t N, int64_t p0, int64_t hi,
                                                  uint32_t (&y)[kApScanItems], bool (&end)[kApScanItems], bool (&valid)[kApScanItems]) {
    uint32_t key[kApScanItems + 1];
#pragma unroll
    for (int j = 0; j <= kApScanItems; ++j) {
        const int64_t p = p0 + j;
        key[j] = p < N ? in[N - 1 - p] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int j = 0; j < kApScanItems; ++j) {
        valid[j] = p0 + j < hi;
        y[j] = valid[j] ? (key[j] & 1u) : 0u;
        end[j] = valid[j] && ((key[j + 1] >> 1) != (key[j] >> 1));  // the sentinel differs from every real key
    }
}

// Block-wide exclusive sum and exclusive "last end" over one (count, last_end) pair per thread.
// Returns the sum of counts of all earlier threads; *prev_end = tp (block-local) at the last threshold owned by an earlier
// thread, kApNoEnd if none; totals of the whole block in *tot_cnt / *tot_end.
__device__ __forceinline__ uint32_t ap_block_scan(uint32_t cnt, uint32_t last_end_local, uint32_t* s_cnt, uint32_t* s_end,
                                                  uint32_t* prev_end, uint32_t* tot_cnt, uint32_t* tot_end) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_cnt[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kApWarps; ++w) {
        const uint32_t c = s_cnt[w];
        if (w < warp) before += c;
        total += c;
    }
    const uint32_t excl = before + inc - cnt;
    // tp at this thread's last threshold in block coordinates (+1 so that 0 means "none"; monotone -> max = latest)
    uint32_t e = last_end_local == kApNoEnd ? 0u : excl + last_end_local + 1u;
    uint32_t einc = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, einc, o);
        if (lane >= o) einc = max(einc, u);
    }
    uint32_t eexc = __shfl_up_sync(0xffffffffu, einc, 1);
    if (lane == 0) eexc = 0u;
    if (lane == 31) s_end[warp] = einc;
    __syncthreads();
    uint32_t ebefore = 0, etotal = 0;
#pragma unroll
    for (int w = 0; w < kApWarps; ++w) {
        const uint32_t c = s_end[w];
        if (w < warp) ebefore = max(ebefore, c);
        etotal = max(etotal, c);
    }
    __syncthreads();  // s_cnt / s_end are reused by the next tile
    const uint32_t pe = max(eexc, ebefore);
    *prev_end = pe == 0u ? kApNoEnd : pe - 1u;
    *tot_cnt = total;
    *tot_end = etotal == 0u ? kApNoEnd : etotal - 1u;
    return excl;
}

// part[k][s] = {positives of the slice, slice-local tp at its last threshold (kApNoEnd if the slice has none)}
__global__ void __launch_bounds__(kApThreads)
ap_scan_local(const uint32_t* __restrict__ keys, int64_t N, int S, int64_t L, uint2* __restrict__ part) {
    __shared__ uint32_t s_cnt[kApWarps], s_end[kApWarps];
    const int k = blockIdx.x / S, s = blockIdx.x % S;
    const uint32_t* in = keys + static_cast<int64_t>(k) * N;
    const int64_t lo = s * L, hi = lo + L < N ? lo + L : N;
    uint32_t cnt_carry = 0, end_carry = kApNoEnd;
    for (int64_t ts = lo; ts < hi; ts += kApThreads * kApScanItems) {
        uint32_t y[kApScanItems];
        bool end[kApScanItems], valid[kApScanItems];
        ap_load_positions(in, N, ts + static_cast<int64_t>(threadIdx.x) * kApScanItems, hi, y, end, valid);
        uint32_t c = 0, le = kApNoEnd;
#pragma unroll
        for (int j = 0; j < kApScanItems; ++j) {
            c += y[j];
            if (end[j]) le = c;
        }
        uint32_t pe, tc, te;
        ap_block_scan(c, le, s_cnt, s_end, &pe, &tc, &te);
        if (te != kApNoEnd) end_carry = cnt_carry + te;
        cnt_carry += tc;
    }
    if (threadIdx.x == 0) part[blockIdx.x] = make_uint2(cnt_carry, end_carry);
}

__global__ void __launch_bounds__(kApThreads)
ap_scan_final(const uint32_t* __restrict__ keys, int64_t N, int S, int64_t L, const uint2* __restrict__ part,
              double* __restrict__ acc_part) {
    __shared__ uint32_t s_cnt[kApWarps], s_end[kApWarps];
    __shared__ double s_acc[kApWarps];
    const int k = blockIdx.x / S, s = blockIdx.x % S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* in = keys + static_cast<int64_t>(k) * N;
    const int64_t lo = s * L, hi = lo + L < N ? lo + L : N;
    // carries from the slices above (higher scores) and the class total
    uint32_t tp_carry = 0, end_carry = 0, P = 0;  // end_carry: tp at the last threshold before this slice (0 = none yet)
    for (int q = 0; q < S; ++q) {
        const uint2 pq = part[k * S + q];
        if (q < s) {
            if (pq.y != kApNoEnd) end_carry = P + pq.y;
            tp_carry += pq.x;
        }
        P += pq.x;
    }
    double acc = 0.0;
    if (P > 0u) {
        const double Pd = static_cast<double>(P);
        for (int64_t ts = lo; ts < hi; ts += kApThreads * kApScanItems) {
            const int64_t p0 = ts + static_cast<int64_t>(threadIdx.x) * kApScanItems;
            uint32_t y[kApScanItems];
            bool end[kApScanItems], valid[kApScanItems];
            ap_load_positions(in, N, p0, hi, y, end, valid);
            uint32_t c = 0, le = kApNoEnd;
#pragma unroll
            for (int j = 0; j < kApScanItems; ++j) {
                c += y[j];
                if (end[j]) le = c;
            }
            uint32_t pe, tc, te;
            const uint32_t excl = ap_block_scan(c, le, s_cnt, s_end, &pe, &tc, &te);
            uint32_t prev_end = pe == kApNoEnd ? end_carry : tp_carry + pe;  // tp at the previous threshold
            uint32_t tp = tp_carry + excl;
#pragma unroll
            for (int j = 0; j < kApScanItems; ++j) {
                tp += y[j];
                if (end[j]) {
                    if (tp > prev_end) {
                        // sklearn: precision = tps / (tps + fps), recall = tps / tps[-1], ap = sum diff(recall) * precision
                        const double prec = static_cast<double>(tp) / static_cast<double>(p0 + j + 1);
                        acc += (static_cast<double>(tp) / Pd - static_cast<double>(prev_end) / Pd) * prec;
                    }
                    prev_end = tp;
                }
            }
            if (te != kApNoEnd) end_carry = tp_carry + te;
            tp_carry += tc;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_acc[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kApWarps; ++w) t += s_acc[w];
        acc_part[blockIdx.x] = t;
    }
}

// ap[k] = sum of the slice partials in slice order; NaN when the class has no positives (the reference skips such
// classes, metrics.py:54).
__global__ void ap_reduce(const uint2* __restrict__ part, const double* __restrict__ acc_part, int K, int S,
                          double* __restrict__ ap, int64_t* __restrict__ num_pos) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int64_t P = 0;
    double t = 0.0;
    for (int s = 0; s < S; ++s) {
        P += part[k * S + s].x;
        t += acc_part[k * S + s];
    }
    num_pos[k] = P;
    ap[k] = P > 0 ? t : __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace prego
