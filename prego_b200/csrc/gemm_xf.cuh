// CTA-pair tcgen05 GEMM whose A operand is TRANSFORMED in shared memory between the TMA load and the MMA:
//
//   D[M, N] = f(A)[M, K] * W[N, K]^T + bias,      f = LayerNorm (row statistics given) * gamma + beta, ReLU
//
// This is layer1's LayerNorm -> ReLU (rnn.py:41-42) folded into the GRU input-gate GEMM (rnn.py:38,61): the fp16
// pre-activations y written by the projection GEMM are normalised on their way INTO the tensor core, so the separate
// LayerNorm pass (read y, write e: 2 x 4 KiB per frame of HBM traffic, 6 % of the step) and the buffer e disappear.
// The row statistics (sum, sum of squares per 256-column tile) come out of the projection GEMM's own epilogue
// (EpiStoreStats below) and are reduced to (rstd, -mean * rstd) per row by ln_finalize_kernel.
//
// Same main loop as gemm_tc2_kernel (gemm_tc.cuh) with one more pipeline hop:
//   warp 0      TMA producer: A tile (raw y) -> the CTA's own `a_full` barrier; W half tile -> the leader's `full` barrier
//   warps 6..17 transform, two warps owning each of the 6 ring stages (64 rows each): wait a_full, read the swizzled A tile
//               (LDS.128), normalise / scale / shift / ReLU, write it back IN PLACE in the operand format (STS.128, same
//               swizzled position), fence.proxy.async, arrive on the leader's `a_ready` (4 arrivals: two warps per CTA)
//   warp 1      MMA issuer (leader CTA): wait full (W landed) + a_ready (A transformed in both CTAs), 4 MMAs, commit
//   warps 2..5  epilogue (unchanged)
// A k-block costs each of its transform warps 16 x (LDS.128 + 8 elements + STS.128) per lane; gamma / beta live in shared memory,
// the row statistics come through L1.
#pragma once
#include "gemm_tc.cuh"

namespace prego {

constexpr int kXfStages = 6;                               // ring depth of the transform GEMM: two transform warps own each stage
constexpr int kXfWarps = 2 * kXfStages;
constexpr int kXfThreads = kGemmThreads + kXfWarps * 32;  // 576

// Row statistics of LayerNorm, per row: x' = x * a + b with a = rstd, b = -mean * rstd.
struct LnXf {
    const float2* rowstat;  // [M] (a, b)
    const float* gamma;     // [K]
    const float* beta;      // [K]
    int dbg = 0;            // diagnostics: 1 = no shared-memory loads / stores in the transform, 2 = loads only, 4 = no proxy fence
};

// out = fp16(acc + bias) exactly as EpiStore<256, 0>, plus per-row partial sums of the ROUNDED values over this tile's
// columns: stats[(n0 / TILE_N) * M + row] = (sum, sum of squares).  One thread owns one row of the tile, so the partials
// need no reduction; eight tiles per row are combined by ln_finalize_kernel.
template <int TILE_N>
struct EpiStoreStats {
    __half* out;
    const float* bias;
    int64_t ldc;
    float2* stats;  // [N / TILE_N][M]
    int M;

    __device__ __forceinline__ void operator()(uint32_t taddr, int row, int n0, bool valid) const {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < TILE_N / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            if (!valid) continue;
            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 32);
            uint4* d4 = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldc + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
                uint4 u;
                u.x = Op16<0>::pack2(__uint_as_float(v[8 * j + 0]) + ba.x, __uint_as_float(v[8 * j + 1]) + ba.y);
                u.y = Op16<0>::pack2(__uint_as_float(v[8 * j + 2]) + ba.z, __uint_as_float(v[8 * j + 3]) + ba.w);
                u.z = Op16<0>::pack2(__uint_as_float(v[8 * j + 4]) + bb.x, __uint_as_float(v[8 * j + 5]) + bb.y);
                u.w = Op16<0>::pack2(__uint_as_float(v[8 * j + 6]) + bb.z, __uint_as_float(v[8 * j + 7]) + bb.w);
                d4[j] = u;
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // statistics of what LayerNorm will actually read back: the rounded values
                    const float2 f = Op16<0>::unpack2(w[q]);
                    s1 += f.x + f.y;
                    s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
                }
            }
        }
        if (valid) stats[static_cast<int64_t>(n0 / TILE_N) * M + row] = make_float2(s1, s2);
    }
};

// (sum, sum of squares) partials of the n_tiles column tiles of a row -> (rstd, -mean * rstd); biased variance, eps inside
// the square root (nn.LayerNorm, rnn.py:41).  One thread per row.
__global__ void __launch_bounds__(256)
ln_finalize_kernel(const float2* __restrict__ stats, float2* __restrict__ rowstat, int M, int n_tiles, int E, float eps) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
        const float2 p = stats[static_cast<int64_t>(t) * M + row];
        s1 += p.x;
        s2 += p.y;
    }
    const float mu = s1 / static_cast<float>(E);
    const float var = fmaxf(s2 / static_cast<float>(E) - mu * mu, 0.f);
    const float rstd = 1.0f / sqrtf(var + eps);
    rowstat[row] = make_float2(rstd, -mu * rstd);
}

template <int TILE_N, int STAGES>
struct GemmXfCfg {
    static constexpr int kABytes = kTileM * kTileK * 2;
    static constexpr int kBBytes = (TILE_N / 2) * kTileK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kAccStride = GemmCfg<TILE_N>::kAccStride;
    static constexpr int kTmemCols = 2 * kAccStride;
    static constexpr int kBarBytes = (4 * STAGES + 4) * 8 + 16;
    static constexpr int smem_bytes(int K) { return STAGES * kStageBytes + 2 * K * 4 + 1024 + ((kBarBytes + 255) / 256) * 256; }
};

// IDENTITY = true: the transform warps copy the tile through registers unchanged (measures what the extra pipeline hop and
// its shared-memory traffic cost; used by prego_gemm16_nt's diagnostic tile code).
template <int TILE_N, int STAGES, int FMT, class Epi, bool IDENTITY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXfThreads, 1)
gemm_tc2_xf_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epi epi,
                   LnXf xf) {
    static_assert(STAGES == kXfStages, "two transform warps per ring stage");
    using Cfg = GemmXfCfg<TILE_N, STAGES>;
    using Op = Op16<FMT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_gamma = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
    float* s_beta = s_gamma + K;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + K);
    uint64_t* full_bar = bars;                    // [STAGES]  leader: W tile of both CTAs landed
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  per CTA, multicast commit
    uint64_t* a_full = bars + 2 * STAGES;         // [STAGES]  per CTA: own raw A tile landed
    uint64_t* a_ready = bars + 3 * STAGES;        // [STAGES]  leader: A transformed in both CTAs (4 arrivals)
    uint64_t* acc_full = bars + 4 * STAGES;       // [2]
    uint64_t* acc_empty = bars + 4 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int n_tiles = N / TILE_N;
    const int m_tiles = (M + 2 * kTileM - 1) / (2 * kTileM);
    const int total_tiles = n_tiles * m_tiles;
    const int k_blocks = K / kTileK;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
            ptx::mbar_init(&a_full[s], 1);
            ptx::mbar_init(&a_ready[s], 4);  // two transform warps per CTA per k-block
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&acc_full[b], 1);
            ptx::mbar_init(&acc_empty[b], 8);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(tmem_slot, Cfg::kTmemCols);
        ptx::tmem_relinquish2();
    }
    if (!IDENTITY) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            s_gamma[i] = xf.gamma[i];
            s_beta[i] = xf.beta[i];
        }
    }
    __syncthreads();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                const int m0 = (tile / n_tiles) * (2 * kTileM) + static_cast<int>(rank) * kTileM;
                const int n0 = (tile % n_tiles) * TILE_N + static_cast<int>(rank) * (TILE_N / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    ptx::mbar_expect_tx(&a_full[stage], Cfg::kABytes);
                    ptx::tma_load_3d(&tmA, sa, &a_full[stage], kb * kTileK, 0, m0, ptx::kEvictFirst);
                    if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * Cfg::kBBytes);
                    ptx::tma_load_2d_2sm(&tmB, sa + Cfg::kABytes, &full_bar[stage], kb * kTileK, n0, ptx::kEvictLast);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(FMT, 2 * kTileM, TILE_N);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
                const int buf = it & 1;
                ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * Cfg::kAccStride;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::mbar_wait(&a_ready[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::mma_f16_ss_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::mma_commit_2sm(&empty_bar[stage], 3);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit_2sm(&acc_full[buf], 3);
            }
        }
    } else if (warp >= 6) {
        // ------------------------------------------------------------ A-operand transform (in place)
        // Two warps OWN each ring stage (64 rows each) and visit its consecutive phases: the chain wait -> LDS -> math -> STS ->
        // proxy fence -> arrive is several hundred cycles long, so a warp that had to touch every k-block would pace the whole
        // pipeline (measured: 8 warps x 16 rows of every k-block ran the GEMM at 66 % of the plain kernel even with an IDENTITY
        // transform); one visit per warp every STAGES x 512 cycles hides it.  Owning a stage also keeps every mbarrier wait
        // within one phase of the barrier (parity waits alias two phases apart).
        const int w = warp - 6;
        const int stage = w % STAGES;
        const int half = w / STAGES;  // rows half * 64 .. + 63 of the CTA's 128
        const int c = lane & 7;       // logical 16-byte chunk of the 128-byte row = columns c*8 .. c*8+7 of the k-block
        const int rsub = lane >> 3;
        // 128-byte swizzle: chunk c of row r sits at chunk position c ^ (r & 7); 8 consecutive lanes cover one whole row, so
        // every quarter-warp access is one conflict-free 128-byte wavefront.  r = 4 i + rsub: (r & 7) alternates with i.
        const uint32_t off_even = static_cast<uint32_t>(half * 64 + rsub) * 128u + (static_cast<uint32_t>(c ^ rsub) << 4);
        const uint32_t off_odd = static_cast<uint32_t>(half * 64 + rsub + 4) * 128u + (static_cast<uint32_t>(c ^ (rsub + 4)) << 4);
        const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const int my_tiles = cluster_id < total_tiles ? (total_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
        const int total_g = my_tiles * k_blocks;  // k-blocks of all tiles of this CTA, in ring order
        uint32_t phase = 0;
        for (int g = stage; g < total_g; g += STAGES, phase ^= 1) {
            const int tile = cluster_id + (g / k_blocks) * num_clusters;
            const int kb = g % k_blocks;
            const int m0 = (tile / n_tiles) * (2 * kTileM) + static_cast<int>(rank) * kTileM + half * 64;
            float ga[8], be[8];
            if (!IDENTITY) {
                const float4* g4 = reinterpret_cast<const float4*>(s_gamma + kb * kTileK + c * 8);
                const float4* b4 = reinterpret_cast<const float4*>(s_beta + kb * kTileK + c * 8);
                const float4 g0 = g4[0], g1 = g4[1], b0 = b4[0], b1 = b4[1];
                ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
                be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
            }
            ptx::mbar_wait(&a_full[stage], phase);
#pragma unroll 4
            for (int i = 0; i < 16; i += 2) {  // rows 4 i + rsub and 4 (i + 1) + rsub of this warp's 64
                const uint32_t a0 = sa + static_cast<uint32_t>(i >> 1) * 1024u + off_even;
                const uint32_t a1 = sa + static_cast<uint32_t>(i >> 1) * 1024u + off_odd;
                uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
                if (!(xf.dbg & 1)) { v0 = lds128(a0); v1 = lds128(a1); }
                if (!IDENTITY) {
                    const int r0 = m0 + 4 * i + rsub;
                    const float2 s0 = r0 < M ? __ldg(xf.rowstat + r0) : make_float2(0.f, 0.f);
                    const float2 s1 = r0 + 4 < M ? __ldg(xf.rowstat + r0 + 4) : make_float2(0.f, 0.f);
                    const uint32_t in0[4] = {v0.x, v0.y, v0.z, v0.w}, in1[4] = {v1.x, v1.y, v1.z, v1.w};
                    uint32_t o0[4], o1[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 y0 = Op16<0>::unpack2(in0[q]);  // y is stored as fp16 by the projection GEMM
                        const float2 y1 = Op16<0>::unpack2(in1[q]);
                        o0[q] = Op::pack2(fmaxf(fmaf(fmaf(y0.x, s0.x, s0.y), ga[2 * q], be[2 * q]), 0.f),
                                          fmaxf(fmaf(fmaf(y0.y, s0.x, s0.y), ga[2 * q + 1], be[2 * q + 1]), 0.f));
                        o1[q] = Op::pack2(fmaxf(fmaf(fmaf(y1.x, s1.x, s1.y), ga[2 * q], be[2 * q]), 0.f),
                                          fmaxf(fmaf(fmaf(y1.y, s1.x, s1.y), ga[2 * q + 1], be[2 * q + 1]), 0.f));
                    }
                    v0 = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                    v1 = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                }
                if (!(xf.dbg & 3)) { sts128(a0, v0); sts128(a1, v1); }
                else if (v0.x == 0x12345678u && v1.y == 0x9abcdef0u) sts128(a0, v1);  // keep the loads alive
            }
            if (!(xf.dbg & 4)) ptx::fence_proxy_async_smem();  // the generic-proxy stores above -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&a_ready[stage]);
        }
    } else {
        const int quad = warp & 3;
        int it = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
            const int buf = it & 1;
            const int m0 = (tile / n_tiles) * (2 * kTileM) + static_cast<int>(rank) * kTileM;
            const int n0 = (tile % n_tiles) * TILE_N;
            ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + buf * Cfg::kAccStride + (static_cast<uint32_t>(quad * 32) << 16);
            const int row = m0 + quad * 32 + lane;
            epi(taddr, row, n0, row < M);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&acc_empty[buf]);
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace prego
