// One GRU time step for all streams (rnn.py:61), batched regime (B > 16):
//
//   gh = h_{t-1} W_hh'^T          tcgen05 GEMM, 128 streams x (64 hidden units x 3 gates) per tile
//   r = sigma(gi_r + gh_r + b_hr),  z = sigma(gi_z + gh_z + b_hz)
//   n = tanh(gi_n + r * (gh_n + b_hn)),  h_t = (h_{t-1} - n) * z + n       (ATen's order)
//
// fused in one persistent kernel.  Every byte the epilogue touches moves by TMA: the gate
// pre-activations gi[t] (fp16, 3 x [128 x 64] boxes) and the fp32 master state (2 x [128 x 32]
// boxes) are bulk-loaded into swizzled shared memory by a dedicated producer warp while the
// tensor core works on the tile, and the three results -- fp32 state (in place), the 16-bit
// operand copy of h_t for the next step (history slot t+1) and relu(h_t) for the classifier --
// leave through TMA stores.  No epilogue thread issues a global load or store, so the step is
// not exposed to DRAM latency (the first version, with per-thread global accesses, spent 74 %
// of its samples in long-scoreboard stalls: profiles/r01_gru_step_v1_ncu.txt).
//
// Layouts (time-major inside a chunk so one step touches contiguous rows):
//   hseq  [Tc+1, B, H]  16-bit operand history, slot t = h_{t-1}
//   gi    [Tc,   B, 3H] fp16, gate-interleaved columns (tile n: [r(64) | z(64) | n(64)])
//   hrelu [Tc,   B, H]  16-bit relu(h_t)
//   h32   [B, H]        fp32 master state
//
// Warp roles (224 threads): warp 0 = operand TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4), warp 6 = epilogue-operand TMA producer.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace prego {

constexpr int kGruThreads = 224;
constexpr int kGruTileN = 192;
constexpr int kGruStages = 3;
constexpr int kGruStageBytes = kTileM * kTileK * 2 + kGruTileN * kTileK * 2;  // 40960
constexpr int kGruBoxBytes = 128 * 128;                                       // one [128 rows x 128 B] box
constexpr int kGruEpiBytes = 5 * kGruBoxBytes;                                // gi r,z,n + h32 lo,hi
constexpr int kGruSmemBytes = kGruStages * kGruStageBytes + kGruEpiBytes + 256 + 1024;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_step_kernel(const __grid_constant__ CUtensorMap tmHseq,   // 3-D (H, B, Tc+1) 16-bit, box (64, 128, 1)
                const __grid_constant__ CUtensorMap tmW,      // 2-D (H, 3H) 16-bit, box (64, 192)
                const __grid_constant__ CUtensorMap tmGi,     // 3-D (3H, B, Tc) fp16, box (64, 128, 1)
                const __grid_constant__ CUtensorMap tmH32,    // 2-D (H, B) fp32, box (32, 128)
                const __grid_constant__ CUtensorMap tmHrelu,  // 3-D (H, B, Tc) 16-bit, box (64, 128, 1)
                const float* __restrict__ bhh, int B, int H, int t) {
    using Op = Op16<FMT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + kGruStages * kGruStageBytes;  // gi_r, gi_z, gi_n, h_lo, h_hi boxes
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + kGruEpiBytes);
    uint64_t* full_bar = bars;                        // [3]
    uint64_t* empty_bar = bars + kGruStages;          // [3]
    uint64_t* acc_full = bars + 2 * kGruStages;       // [2]
    uint64_t* acc_empty = bars + 2 * kGruStages + 2;  // [2]
    uint64_t* epi_full = bars + 2 * kGruStages + 4;   // [1]
    uint64_t* epi_empty = bars + 2 * kGruStages + 5;  // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGruStages + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = (3 * H) / kGruTileN;
    const int m_tiles = (B + kTileM - 1) / kTileM;
    const int total_tiles = n_tiles * m_tiles;
    const int k_blocks = H / kTileK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmHseq);
        ptx::prefetch_tmap(&tmW);
        ptx::prefetch_tmap(&tmGi);
        ptx::prefetch_tmap(&tmH32);
        ptx::prefetch_tmap(&tmHrelu);
        for (int s = 0; s < kGruStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&acc_full[b], 1);
            ptx::mbar_init(&acc_empty[b], 4);
        }
        ptx::mbar_init(epi_full, 1);
        ptx::mbar_init(epi_empty, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * kTileM;
                const int n0 = (tile % n_tiles) * kGruTileN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kGruStageBytes;
                    ptx::mbar_expect_tx(&full_bar[stage], kGruStageBytes);
                    ptx::tma_load_3d(&tmHseq, sa, &full_bar[stage], kb * kTileK, m0, t, ptx::kEvictNormal);
                    ptx::tma_load_2d(&tmW, sa + kTileM * kTileK * 2, &full_bar[stage], kb * kTileK, n0, ptx::kEvictLast);
                    if (++stage == kGruStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc(FMT, kTileM, kGruTileN);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 256;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * kGruStageBytes);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + kTileM * kTileK * 2);
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)
                        ptx::mma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::mma_commit(&empty_bar[stage]);
                    if (++stage == kGruStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit(&acc_full[buf]);
            }
        }
    } else if (warp == 6) {
        // ------------------------------------------- epilogue-operand producer (gi[t] + fp32 state)
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int m0 = (tile / n_tiles) * kTileM;
                const int nt = tile % n_tiles;
                ptx::mbar_wait(epi_empty, (it & 1) ^ 1);
                ptx::mbar_expect_tx(epi_full, kGruEpiBytes);
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    ptx::tma_load_3d(&tmGi, epi_smem + g * kGruBoxBytes, epi_full, nt * kGruTileN + g * 64, m0, t, ptx::kEvictFirst);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    ptx::tma_load_2d(&tmH32, epi_smem + (3 + j) * kGruBoxBytes, epi_full, nt * 64 + j * 32, m0, ptx::kEvictNormal);
            }
        }
    } else {
        // -------------------------------------------------------------------------- epilogue
        const int quad = warp & 3;
        const int r = quad * 32 + lane;  // row of the tile = TMEM lane
        const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        const uint32_t s_gi = ptx::smem_u32(epi_smem);
        const bool store_thread = (warp == 2 && lane == 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int m0 = (tile / n_tiles) * kTileM;
            const int nt = tile % n_tiles;
            ptx::mbar_wait(epi_full, it & 1);
            ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + buf * 256 + (static_cast<uint32_t>(quad * 32) << 16);
            const float* bh = bhh + nt * kGruTileN;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {  // 8 hidden units per iteration
                uint32_t vr[8], vz[8], vn[8];
                ptx::tmem_ld8(taddr + c * 8, vr);
                ptx::tmem_ld8(taddr + 64 + c * 8, vz);
                ptx::tmem_ld8(taddr + 128 + c * 8, vn);
                const uint32_t a_gi = s_gi + row_off + ((static_cast<uint32_t>(c) ^ sw) << 4);
                const uint4 qr = lds128(a_gi);
                const uint4 qz = lds128(a_gi + kGruBoxBytes);
                const uint4 qn = lds128(a_gi + 2 * kGruBoxBytes);
                const uint32_t hbox = s_gi + (3 + (c >> 2)) * kGruBoxBytes + row_off;
                const uint32_t a_h0 = hbox + ((static_cast<uint32_t>((c & 3) * 2) ^ sw) << 4);
                const uint32_t a_h1 = hbox + ((static_cast<uint32_t>((c & 3) * 2 + 1) ^ sw) << 4);
                const uint4 h0 = lds128(a_h0);
                const uint4 h1 = lds128(a_h1);
                float br[8], bz[8], bn[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(bh + c * 8) + q);
                    const float4 y = __ldg(reinterpret_cast<const float4*>(bh + 64 + c * 8) + q);
                    const float4 w = __ldg(reinterpret_cast<const float4*>(bh + 128 + c * 8) + q);
                    br[4 * q] = x.x; br[4 * q + 1] = x.y; br[4 * q + 2] = x.z; br[4 * q + 3] = x.w;
                    bz[4 * q] = y.x; bz[4 * q + 1] = y.y; bz[4 * q + 2] = y.z; bz[4 * q + 3] = y.w;
                    bn[4 * q] = w.x; bn[4 * q + 1] = w.y; bn[4 * q + 2] = w.z; bn[4 * q + 3] = w.w;
                }
                const uint32_t gr[4] = {qr.x, qr.y, qr.z, qr.w}, gz[4] = {qz.x, qz.y, qz.z, qz.w},
                               gn[4] = {qn.x, qn.y, qn.z, qn.w};
                const uint32_t hp[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                ptx::tmem_ld_wait_dep8(vr);
                ptx::tmem_ld_wait_dep8(vz);
                ptx::tmem_ld_wait_dep8(vn);
                float hn[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fr = __half22float2(*reinterpret_cast<const __half2*>(&gr[j]));
                    const float2 fz = __half22float2(*reinterpret_cast<const __half2*>(&gz[j]));
                    const float2 fn = __half22float2(*reinterpret_cast<const __half2*>(&gn[j]));
                    const float gir[2] = {fr.x, fr.y}, giz[2] = {fz.x, fz.y}, gin[2] = {fn.x, fn.y};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = 2 * j + e;
                        const float rr = fast_sigmoid(gir[e] + (__uint_as_float(vr[i]) + br[i]));
                        const float zz = fast_sigmoid(giz[e] + (__uint_as_float(vz[i]) + bz[i]));
                        const float nn = fast_tanh(gin[e] + rr * (__uint_as_float(vn[i]) + bn[i]));
                        hn[i] = (__uint_as_float(hp[i]) - nn) * zz + nn;
                    }
                }
                uint4 o0, o1, os, orl;
                o0.x = __float_as_uint(hn[0]); o0.y = __float_as_uint(hn[1]); o0.z = __float_as_uint(hn[2]); o0.w = __float_as_uint(hn[3]);
                o1.x = __float_as_uint(hn[4]); o1.y = __float_as_uint(hn[5]); o1.z = __float_as_uint(hn[6]); o1.w = __float_as_uint(hn[7]);
                os.x = Op::pack2(hn[0], hn[1]); os.y = Op::pack2(hn[2], hn[3]);
                os.z = Op::pack2(hn[4], hn[5]); os.w = Op::pack2(hn[6], hn[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) hn[i] = fmaxf(hn[i], 0.0f);
                orl.x = Op::pack2(hn[0], hn[1]); orl.y = Op::pack2(hn[2], hn[3]);
                orl.z = Op::pack2(hn[4], hn[5]); orl.w = Op::pack2(hn[6], hn[7]);
                sts128(a_h0, o0);                      // fp32 state, in place
                sts128(a_h1, o1);
                sts128(a_gi, os);                      // operand copy of h_t  (over the consumed gi_r chunk)
                sts128(a_gi + kGruBoxBytes, orl);      // relu(h_t)            (over the consumed gi_z chunk)
            }
            // accumulator buffer is free again
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
            // results: shared memory -> global by TMA
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1, 128);
            if (store_thread) {
                ptx::tma_store_2d(&tmH32, epi_smem + 3 * kGruBoxBytes, nt * 64, m0);
                ptx::tma_store_2d(&tmH32, epi_smem + 4 * kGruBoxBytes, nt * 64 + 32, m0);
                ptx::tma_store_3d(&tmHseq, epi_smem, nt * 64, m0, t + 1);
                ptx::tma_store_3d(&tmHrelu, epi_smem + kGruBoxBytes, nt * 64, m0, t);
                ptx::tma_store_commit();
                ptx::tma_store_wait_read();
                ptx::mbar_arrive(epi_empty);  // operand buffers may be refilled
            }
        }
        if (store_thread) ptx::tma_store_wait_all();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace prego
