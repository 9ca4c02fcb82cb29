// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M, N] = A[M, K] * W[N, K]^T          (bf16 operands, fp32 accumulate in TMEM)
//
// A and W are both K-major in global memory (W is exactly the nn.Linear / nn.GRU
// weight layout), staged into shared memory by TMA with the 128-byte swizzle and
// consumed by tcgen05.mma straight from shared memory.  One CTA per SM loops over
// 128 x TILE_N output tiles; the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (192 threads):  warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue
// (warp w owns TMEM lanes 32*(w%4) .. +31, i.e. tile rows of that quadrant; one
// thread per output row).
//
// The epilogue is a functor so the same main loop serves
//   * the input projection  (Linear 4096->2048, rnn.py:40,58)       -> EpiStore
//   * the GRU input gates   (weight_ih_l0, rnn.py:38,61)            -> EpiStore
//   * one GRU time step     (weight_hh_l0 + gate math, rnn.py:61)   -> EpiGruStep
//   * the classifier head   (f_classification + softmax/argmax,
//                            rnn.py:62-69, trainer/eval.py:53)      -> EpiHead
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "ptx.cuh"

namespace prego {

constexpr int kTileM = 128;
constexpr int kTileK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 192;

// 16-byte shared-memory accesses by shared-window address (epilogues / transform warps working on swizzled TMA boxes)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int TILE_N>
struct GemmCfg {
    static constexpr int kABytes = kTileM * kTileK * 2;
    static constexpr int kBBytes = TILE_N * kTileK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // accumulator buffer stride in TMEM columns (power of two >= TILE_N)
    static constexpr int kAccStride = TILE_N <= 32 ? 32 : TILE_N <= 64 ? 64 : TILE_N <= 128 ? 128 : 256;
    static constexpr int kTmemCols = 2 * kAccStride;
    static constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/; }
};

// 16-bit operand formats of tcgen05 kind::f16.  FMT 0 = fp16 (10-bit mantissa, the accuracy of TF32 at
// twice its rate; features / activations / weights of this model are far inside its range and the
// converters saturate), FMT 1 = bf16.
template <int FMT> struct Op16;
template <> struct Op16<0> {
    using T = __half;
    static __device__ __forceinline__ uint32_t pack2(float a, float b) {
        a = fminf(fmaxf(a, -65504.f), 65504.f);
        b = fminf(fmaxf(b, -65504.f), 65504.f);
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ T from_float(float a) { return __float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)); }
    static __device__ __forceinline__ float2 unpack2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};
template <> struct Op16<1> {
    using T = __nv_bfloat16;
    static __device__ __forceinline__ uint32_t pack2(float a, float b) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ T from_float(float a) { return __float2bfloat16_rn(a); }
    static __device__ __forceinline__ float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
};

template <int TILE_N, int STAGES, int FMT, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               int a_c1, Epi epi) {
    using Cfg = GemmCfg<TILE_N>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * STAGES;       // [2]       MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tiles = N / TILE_N;
    const int m_tiles = (M + kTileM - 1) / kTileM;
    const int total_tiles = n_tiles * m_tiles;
    const int k_blocks = K / kTileK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&acc_full[b], 1);
            ptx::mbar_init(&acc_empty[b], 4);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // broadcast through a shuffle so the compiler KNOWS the address is warp-uniform (tcgen05 operands live in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * kTileM;
                const int n0 = (tile % n_tiles) * TILE_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kABytes;
                    ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    ptx::tma_load_3d(&tmA, sa, &full_bar[stage], kb * kTileK, a_c1, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d(&tmB, sb, &full_bar[stage], kb * kTileK, n0, ptx::kEvictLast);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(FMT, kTileM, TILE_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t acc_parity = (it >> 1) & 1;
                ptx::mbar_wait(&acc_empty[buf], acc_parity ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * Cfg::kAccStride;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k) {
                        // +32 bytes per 16-element K slice inside the 128B swizzle row (>>4 encoded)
                        ptx::mma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit(&acc_full[buf]);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t acc_parity = (it >> 1) & 1;
            const int m0 = (tile / n_tiles) * kTileM;
            const int n0 = (tile % n_tiles) * TILE_N;
            ptx::mbar_wait(&acc_full[buf], acc_parity);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + buf * Cfg::kAccStride + (static_cast<uint32_t>(quad * 32) << 16);
            const int row = m0 + quad * 32 + lane;
            epi(taddr, row, n0, row < M);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------ CTA-pair variant
// Same contract as gemm_tc_kernel, executed by clusters of two CTAs (cta_group::2): one 256 x TILE_N
// tile per pair, each CTA stages its own 128 rows of A but only HALF of the W tile, and the leader
// issues one M = 256 MMA for both.  Per CTA a k-block costs 16 KB + TILE_N*64 B instead of
// 16 KB + TILE_N*128 B, so the same shared memory holds more stages: the single-CTA kernel is bound
// by (bytes in flight) / (TMA latency ~1 us), not by the tensor pipe (profiles/r01_ncu_full_summary.txt).
template <int TILE_N>
struct Gemm2Cfg {
    static constexpr int kABytes = kTileM * kTileK * 2;
    static constexpr int kBBytes = (TILE_N / 2) * kTileK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kAccStride = GemmCfg<TILE_N>::kAccStride;
    static constexpr int kTmemCols = 2 * kAccStride;
    static constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 + 256; }
};

template <int TILE_N, int STAGES, int FMT, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                int a_c1, Epi epi, const __grid_constant__ CUtensorMap tmA2, int k_split, int rows_per_c1) {
    // A operand, general form: columns [0, k_split) come from tmA, the rest from tmA2 (rgb | flow read in place from the
    // caller's tensors); rows_per_c1 > 0 folds the row index into the middle coordinate (row m -> (a_c1 + m / rows_per_c1,
    // m % rows_per_c1): time-major chunk rows gathered straight from a [B, T, D] tensor).  Plain use: k_split >= K, 0.
    using Cfg = Gemm2Cfg<TILE_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                    // [STAGES]  used in the leader CTA only
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  per CTA, released by the multicast commit
    uint64_t* acc_full = bars + 2 * STAGES;       // [2]       per CTA, multicast commit
    uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]       leader only: 4 epilogue warps x 2 CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int n_tiles = N / TILE_N;
    const int m_tiles = (M + 2 * kTileM - 1) / (2 * kTileM);
    const int total_tiles = n_tiles * m_tiles;
    constexpr int kElemsK = FMT == 2 ? 32 : kTileK;  // elements per 128-byte swizzle row: 64 x 16-bit or 32 x tf32
    const int k_blocks = K / kElemsK;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmA2);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
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
    ptx::tc_fence_before();
    ptx::cluster_sync();  // barriers of both CTAs initialised before any remote signal
    ptx::tc_fence_after();
    // broadcast through a shuffle so the compiler KNOWS the address is warp-uniform (tcgen05 operands live in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                const int m0 = (tile / n_tiles) * (2 * kTileM) + static_cast<int>(rank) * kTileM;
                const int n0 = (tile % n_tiles) * TILE_N + static_cast<int>(rank) * (TILE_N / 2);
                const int c1 = rows_per_c1 > 0 ? a_c1 + m0 / rows_per_c1 : a_c1;
                const int mr = rows_per_c1 > 0 ? m0 % rows_per_c1 : m0;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                    const int kk = kb * kElemsK;
                    if (kk < k_split) ptx::tma_load_3d_2sm(&tmA, sa, &full_bar[stage], kk, c1, mr, ptx::kEvictNormal);
                    else ptx::tma_load_3d_2sm(&tmA2, sa, &full_bar[stage], kk - k_split, c1, mr, ptx::kEvictNormal);
                    ptx::tma_load_2d_2sm(&tmB, sa + Cfg::kABytes, &full_bar[stage], kb * kElemsK, n0, ptx::kEvictLast);
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
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 MMAs of 32 bytes of K each (K = 16 halves or 8 tf32)
                        if constexpr (FMT == 2) ptx::mma_tf32_ss_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else ptx::mma_f16_ss_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::mma_commit_2sm(&empty_bar[stage], 3);  // frees this stage in both CTAs
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit_2sm(&acc_full[buf], 3);  // accumulators of both CTAs complete
            }
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
    ptx::cluster_sync();  // nobody exits (or frees TMEM) while the peer may still signal into this CTA
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------ epilogues

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// out[orow, n] = acc + bias[n].  OUT_FMT: -1 = fp32, 0 = fp16, 1 = bf16.  Row-major with ldc.
// With Tc > 0 the rows are re-ordered from stream-major (m = b*Tc + t) to time-major (t*B + b),
// the layout the recurrence consumes.
template <int TILE_N, int OUT_FMT>
struct EpiStore {
    void* out;
    const float* bias;  // [N] or nullptr
    int64_t ldc;
    int Tc, B;          // Tc == 0: no re-ordering
    int accumulate = 0; // fp32 output only: out += acc (+ bias)
    int relu = 0;       // out = max(acc + bias, 0)  (anticipation layer of MROADA, rnn.py:125-126)
    float scale = 1.0f; // out = acc * scale + bias  (split-fp16 mode: the weights travel pre-scaled by a power of two)

    __device__ __forceinline__ void operator()(uint32_t taddr, int row, int n0, bool valid) const {
        const int64_t orow = Tc > 0 ? static_cast<int64_t>(row % Tc) * B + row / Tc : row;
#pragma unroll 1
        for (int c = 0; c < TILE_N / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            if (!valid) continue;
            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 32);
            float f[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bb = bias != nullptr ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                f[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), scale, bb.x);
                f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), scale, bb.y);
                f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), scale, bb.z);
                f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), scale, bb.w);
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if constexpr (OUT_FMT < 0) {
                float4* d4 = reinterpret_cast<float4*>(static_cast<float*>(out) + orow * ldc + n0 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    if (accumulate) {
                        const float4 p = d4[j];
                        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                    }
                    d4[j] = o;
                }
            } else {
                using O = Op16<OUT_FMT>;
                uint4* d4 = reinterpret_cast<uint4*>(static_cast<typename O::T*>(out) + orow * ldc + n0 + c * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = O::pack2(f[8 * j + 0], f[8 * j + 1]);
                    u.y = O::pack2(f[8 * j + 2], f[8 * j + 3]);
                    u.z = O::pack2(f[8 * j + 4], f[8 * j + 5]);
                    u.w = O::pack2(f[8 * j + 6], f[8 * j + 7]);
                    d4[j] = u;
                }
            }
        }
    }
};

// Classifier head epilogue: logits = acc + bc; probs = softmax(logits); label = first
// index of max(probs) (numpy argmax semantics of trainer/eval.py:53).  TILE_N = padded class
// count; only the first K columns are real.  Input rows are time-major (m = t*B + b); outputs go
// to the caller's [B, T, .] tensors.
template <int TILE_N>
struct EpiHead {
    const float* bias;  // [K]
    float* probs;       // [B, T, K] or nullptr
    float* logits;      // [B, T, K] or nullptr
    int32_t* labels;    // [B, T] or nullptr
    int K, B, T, t0;
    int A = 1;          // anticipation head (rnn.py:125-126): input row = row0 + row = m * A + a, output [B, T, A, .]
    int64_t row0 = 0;

    __device__ __forceinline__ void operator()(uint32_t taddr, int row, int /*n0*/, bool valid) const {
        float v[TILE_N];
#pragma unroll
        for (int c = 0; c < TILE_N / 32; ++c) {
            uint32_t u[32];
            ptx::tmem_ld32(taddr + c * 32, u);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(u[j]);
        }
        if (!valid) return;
        int64_t g;
        if (A == 1) {
            g = static_cast<int64_t>(row % B) * T + t0 + (row / B);
        } else {
            const int64_t r = row0 + row, mrow = r / A;
            g = ((mrow % B) * T + t0 + mrow / B) * A + (r - mrow * A);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < TILE_N; ++j) {
            if (j < K) {
                v[j] += __ldg(bias + j);
                mx = fmaxf(mx, v[j]);
            }
        }
        if (logits != nullptr) {
#pragma unroll
            for (int j = 0; j < TILE_N; ++j)
                if (j < K) logits[g * K + j] = v[j];
        }
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < TILE_N; ++j) {
            if (j < K) {
                v[j] = expf(v[j] - mx);
                sum += v[j];
            }
        }
        float best = -1.0f;
        int arg = 0;
#pragma unroll
        for (int j = 0; j < TILE_N; ++j) {
            if (j < K) {
                v[j] = v[j] / sum;
                if (v[j] > best) {  // strict: first maximum wins
                    best = v[j];
                    arg = j;
                }
            }
        }
        if (probs != nullptr) {
#pragma unroll
            for (int j = 0; j < TILE_N; ++j)
                if (j < K) probs[g * K + j] = v[j];
        }
        if (labels != nullptr) labels[g] = arg;
    }
};

}  // namespace prego
