// GRU recurrence for all streams (rnn.py:61), batched regime (B > 16).  Per time step t:
//
//   gh = h_{t-1} W_hh'^T          tcgen05 GEMM, 256 streams x (64 hidden units x 3 gates) per CTA-pair tile
//   r = sigma(gi_r + gh_r + b_hr),  z = sigma(gi_z + gh_z + b_hz)
//   n = tanh(gi_n + r * (gh_n + b_hn)),  h_t = (h_{t-1} - n) * z + n       (ATen's order)
//
// ONE persistent launch covers a whole range of time steps [t_begin, t_end): the work items (t, m, n) --
// step, 256-stream block, 64-unit block -- are numbered step-major and dealt round-robin to the CTA pairs, and
// the only synchronisation between steps is a DATAFLOW dependency: item (t, m, .) may read h_{t-1} of stream
// block m once the 16 items (t-1, m, *) have published it.  Each CTA bumps done[t][m] (release) after its
// results are globally visible; consumers spin on it with acquire loads.  There is no grid-wide barrier, no
// per-step launch / prologue / tail, and pairs that finish step t early simply start on step t+1.
// (One launch per step cost 14.8 us of fixed overhead per step on top of ~6 us of work per wave:
// scripts/sweep_recurrence.py, profiles/r01_recurrence_sweep.txt.)  With sync == nullptr the kernel runs a
// single step without any inter-CTA protocol (used when the pairs cannot all be co-resident).
//
// The epilogue never waits on DRAM: gi[t] of a tile (3 fp16 boxes of [128 x 64]) is bulk-loaded by TMA into
// swizzled shared memory behind the previous tile's stores; the 16-bit operand copy of h_t (history slot t+1,
// the next step's MMA operand) and relu(h_t) (the classifier's operand) are written over the consumed gi boxes
// and leave through TMA stores; the fp32 master state moves through registers.
// The GEMM runs on CTA PAIRS (cta_group::2): each CTA stages its own 128 rows of h_{t-1} and half of the W_hh'
// tile (28 KB per k-block), 6-deep pipeline.
//
// Layouts (time-major inside a chunk so one step touches contiguous rows):
//   hseq  [2, B, H]     16-bit operand copies of the state, ping-pong: step t reads slot t & 1 (h_{t-1}) and writes slot (t + 1) & 1.
//                       Two slots suffice: the items (t, m, .) that overwrite the slot holding h_{t-2} of stream block m only start once
//                       done[t-1][m] says every item (t-1, m, .) -- the readers of h_{t-2} -- has published.  16 MB at B = 4096: the
//                       operand copy lives in L2 and never costs DRAM writes (a [Tc+1, B, H] history did: 0.54 GB per 64 steps)
//   gi    [Tc,   B, 3H] fp16, gate-interleaved columns (tile n: [r(64) | z(64) | n(64)]); includes b_ih and the
//                       r / z parts of b_hh (pre-summed; b_hn must stay inside r * (W_hn h + b_hn))
//   hrelu [Tc,   B, H]  16-bit relu(h_t)
//   h32t  fp32 master state in the epilogue's own TILED order [B/128][H/64][16][128][4] (row block, unit block,
//         4-unit chunk, row, unit): a warp's 32 rows x 16 B are 512 contiguous bytes, so the per-thread state
//         loads / stores are perfectly coalesced (plain [B, H] rows cost 32 sectors per instruction and made the
//         epilogue LSU-bound).  Converted from / to [B, H] once per chunk (simt_kernels.cuh).
//   done  [Tc, m_tiles] uint32 dependency counters (zeroed by the host before the launch)
//
// Warp roles (640 threads): warps 0 and 19 = operand TMA producers (even / odd k-blocks), warp 1 = TMEM alloc + MMA issuer,
// warps 2..17 = epilogue (TMEM lane quadrant = warp % 4, hidden-unit group of 16 = (warp - 2) / 4; the
// epilogue is latency-bound -- MUFU chains, TMEM loads -- so it is spread over 16 warps), warp 18 = gi
// loader + result storer + publisher.  Results are staged in their own two boxes so the gi boxes can be
// refilled as soon as the epilogue has read them (the TMA stores drain in the background).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace prego {

constexpr int kGruEpiWarps = 16;
constexpr int kGruThreads = (kGruEpiWarps + 4) * 32;  // 640
constexpr int kGruIoWarp = kGruEpiWarps + 2;          // 18
constexpr int kGruProd2Warp = kGruEpiWarps + 3;       // 19: second operand producer (odd k-blocks)
constexpr int kGruTileN = 192;
constexpr int kGruStages = 5;
constexpr int kGruABytes = kTileM * kTileK * 2;                               // 16384: own 128 rows of h
constexpr int kGruStageBytes = kGruABytes + (kGruTileN / 2) * kTileK * 2;     // + half of the W tile = 28672
constexpr int kGruBoxBytes = 128 * 128;                                       // one [128 rows x 128 B] box
constexpr int kGruGiBytes = 3 * kGruBoxBytes;                                 // gi r, z, n of one tile
constexpr int kGruOutBytes = 2 * kGruBoxBytes;                                // operand copy of h_t, relu(h_t)
constexpr int kGruSmemBytes = kGruStages * kGruStageBytes + kGruGiBytes + kGruOutBytes + 256 + 1024;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }


// Dependency wait: spin (acquire) until done[idx] >= target, then order later async-proxy (TMA) reads after it.
// Bounded: a missing peer sets the error flag instead of hanging the GPU.
__device__ __forceinline__ void dep_wait(const uint32_t* ctr, uint32_t target, int* err_flag) {
    uint32_t v;
    long long spins = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        if (++spins > (1ll << 26)) {
            *err_flag = 2;
            break;
        }
    } while (true);
    asm volatile("fence.proxy.async;" ::: "memory");
}

struct GruSeqArgs {
    const float* bhh;   // [3H] packed order
    float* h32t;        // fp32 master state, tiled order (in/out), rows padded to a multiple of 256 (the pair tile: no row mask in the epilogue)
    uint32_t* done;     // [Tc][m_tiles] dependency counters, or nullptr (single step, no protocol)
    int* err_flag;
    int B, H;
    int t_begin, t_end;
    long long* stats = nullptr;  // diagnostics: per-CTA wait-cycle counters [grid][16] (PREGO_GRU_STATS)
    int dbg = 0;        // diagnostics only (PREGO_GRU_DBG): 1 = no gate math, 2 = no dependency waits, 4 = no gi loads, 8 = no result stores / publish, 16 = epilogue handshakes only, 32 / 64 = no W / h operand loads, 256 = no result stores (publish kept), 512 = no publish (stores kept; combine with 2), 1024 / 2048 / 4096 = publish without its __threadfence / proxy fence / bulk-store wait
};

// DIAG = true compiles the diagnostic knobs (a.dbg) and the per-role wait-cycle counters (a.stats) in; the product
// instantiation (DIAG = false) carries neither.
template <int FMT, bool DIAG = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGruThreads, 1)
gru_seq_kernel(const __grid_constant__ CUtensorMap tmHseq,   // 3-D (H, B, 2) 16-bit, box (64, 128, 1)
               const __grid_constant__ CUtensorMap tmW,      // 2-D (H, 3H) 16-bit, box (64, 96): half a tile
               const __grid_constant__ CUtensorMap tmGi,     // 3-D (3H, B, Tc) fp16, box (64, 128, 1)
               const __grid_constant__ CUtensorMap tmHrelu,  // 3-D (H, B, Tc) 16-bit, box (64, 128, 1)
               const GruSeqArgs a) {
    using Op = Op16<FMT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* gi_smem = smem + kGruStages * kGruStageBytes;  // 3 boxes
    uint8_t* out_smem = gi_smem + kGruGiBytes;              // 2 boxes
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_smem + kGruOutBytes);
    uint64_t* full_bar = bars;                        // [S]  operand stage landed (leader CTA's copy is used)
    uint64_t* empty_bar = bars + kGruStages;          // [S]  operand stage consumed (multicast commit, per CTA)
    uint64_t* acc_full = bars + 2 * kGruStages;       // [2]  accumulator complete (multicast commit, per CTA)
    uint64_t* acc_empty = bars + 2 * kGruStages + 2;  // [2]  accumulator drained (leader: 8 warps x 2 CTAs)
    uint64_t* gi_full = bars + 2 * kGruStages + 4;    // [1]  gi boxes landed
    uint64_t* epi_done = bars + 2 * kGruStages + 5;   // [1]  results staged (all epilogue warps)
    uint64_t* out_free = bars + 2 * kGruStages + 6;   // [1]  staged results have been read by the TMA stores
    uint64_t* gi_free = bars + 2 * kGruStages + 7;    // [1]  gi boxes copied to registers (all epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGruStages + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int B = a.B, H = a.H;
    const int n_tiles = (3 * H) / kGruTileN;
    const int m_tiles = (B + 2 * kTileM - 1) / (2 * kTileM);
    const int per_step = n_tiles * m_tiles;
    const int k_blocks = H / kTileK;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int row_base = static_cast<int>(rank) * kTileM;  // this CTA's 128 rows inside the 256-row pair tile
    const int item_begin = a.t_begin * per_step + cluster_id;
    const int item_end = a.t_end * per_step;
    const uint32_t dep_target = 2u * static_cast<uint32_t>(n_tiles);  // both CTAs of all n-tiles of a stream block
    const int dbg = DIAG ? a.dbg : 0;
    auto now = [] { return DIAG ? clock64() : 0ll; };
    long long* const stats = DIAG ? a.stats : nullptr;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmHseq);
        ptx::prefetch_tmap(&tmW);
        ptx::prefetch_tmap(&tmGi);
        ptx::prefetch_tmap(&tmHrelu);
        for (int s = 0; s < kGruStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&acc_full[b], 1);
            ptx::mbar_init(&acc_empty[b], 2 * kGruEpiWarps);
        }
        ptx::mbar_init(gi_full, 1);
        ptx::mbar_init(epi_done, kGruEpiWarps);
        ptx::mbar_init(out_free, 1);
        ptx::mbar_init(gi_free, kGruEpiWarps);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(tmem_slot, 512);
        ptx::tmem_relinquish2();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    // broadcast through a shuffle so the compiler KNOWS the address is warp-uniform (tcgen05 operands live in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0 || warp == kGruProd2Warp) {
        // ---------------------------------------------------------------- operand producers
        // One thread needs ~400 cycles per k-block (mbarrier try_wait ~90, expect_tx, two UTMALDG with their uniform-register
        // set-up): more than the 384 cycles the tensor pipe takes for the k-block's four 256 x 192 x 16 MMAs, so a single
        // producer can never run ahead and the MMA issuer waits on `full` (measured: 5 500 of 9 550 cycles per item,
        // profiles/r02_recurrence_diag.txt).  Two producer threads take the even / odd k-blocks of the same stage ring.
        if (ptx::elect_one()) {
            const int par = warp == 0 ? 0 : 1;
            int stage = par, g0 = 0;  // g0: k-blocks issued by both producers before this item (thread `par` takes the global k-blocks g = par mod 2)
            uint32_t phase = 0;
            long long w_dep = 0, w_empty = 0, t_all = now();
            for (int item = item_begin; item < item_end; item += num_clusters) {
                const int t = item / per_step, rem = item % per_step;
                const int mt = rem / n_tiles;
                const int m0 = mt * (2 * kTileM) + row_base;
                const int n0 = (rem % n_tiles) * kGruTileN + static_cast<int>(rank) * (kGruTileN / 2);
                long long c0 = now();
                if (a.done != nullptr && t > a.t_begin && !(dbg & 2)) dep_wait(a.done + (t - 1) * m_tiles + mt, dep_target, a.err_flag);
                w_dep += now() - c0;
                for (int kb = par ^ (g0 & 1); kb < k_blocks; kb += 2) {
                    c0 = now();
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    w_empty += now() - c0;
                    uint8_t* sa = smem + stage * kGruStageBytes;
                    const uint32_t tx = ((dbg & 64) ? 0u : 2u * kGruABytes) + ((dbg & 32) ? 0u : 2u * (kGruStageBytes - kGruABytes));
                    if (leader) { if (tx) ptx::mbar_expect_tx(&full_bar[stage], tx); else ptx::mbar_arrive(&full_bar[stage]); }
                    if (!(dbg & 64)) ptx::tma_load_3d_2sm(&tmHseq, sa, &full_bar[stage], kb * kTileK, m0, t & 1, ptx::kEvictNormal);
                    if (!(dbg & 32)) ptx::tma_load_2d_2sm(&tmW, sa + kGruABytes, &full_bar[stage], kb * kTileK, n0, ptx::kEvictLast);
                    stage += 2;  // this thread's k-blocks are every other slot of the ring
                    if (stage >= kGruStages) {
                        stage -= kGruStages;
                        phase ^= 1;
                    }
                }
                g0 += k_blocks;
            }
            if (stats != nullptr && par == 0) {
                long long* st = stats + blockIdx.x * 16;
                st[0] = now() - t_all; st[1] = w_dep; st[2] = w_empty;
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------------- MMA issuer
        if (leader && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(FMT, 2 * kTileM, kGruTileN);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            long long w_acc = 0, w_full = 0, t_all = now(), c0;
            for (int item = item_begin; item < item_end; item += num_clusters, ++it) {
                const int buf = it & 1;
                c0 = now();
                ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                w_acc += now() - c0;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * 256;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    c0 = now();
                    ptx::mbar_wait(&full_bar[stage], phase);
                    w_full += now() - c0;
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * kGruStageBytes);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + kGruABytes);
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)
                        ptx::mma_f16_ss_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::mma_commit_2sm(&empty_bar[stage], 3);
                    if (++stage == kGruStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit_2sm(&acc_full[buf], 3);
            }
            if (stats != nullptr) {
                long long* st = stats + blockIdx.x * 16;
                st[3] = now() - t_all; st[4] = w_acc; st[5] = w_full; st[6] = it;
            }
        }
    } else if (warp == kGruIoWarp) {
        // ------------------- gi loader + result storer + publisher (TMA both ways)
        if (ptx::elect_one()) {
            int it = 0;
            int pm0 = 0, pnt = 0, pt = 0, pmt = 0;
            auto store_and_publish = [&](int i_prev) {
                // results of item i_prev are staged (epi_done already observed): store, free the staging, publish
                if (!(dbg & (8 | 256))) {
                ptx::tma_store_3d(&tmHseq, out_smem, pnt * 64, pm0, (pt + 1) & 1);
                ptx::tma_store_3d(&tmHrelu, out_smem + kGruBoxBytes, pnt * 64, pm0, pt);
                ptx::tma_store_commit();
                ptx::tma_store_wait_read();
                }
                ptx::mbar_arrive(out_free);
                if (a.done != nullptr && !(dbg & (8 | 512))) {
                    if (!(dbg & 4096)) ptx::tma_store_wait_all();
                    if (!(dbg & 2048)) asm volatile("fence.proxy.async;" ::: "memory");
                    if (!(dbg & 1024)) __threadfence();  // also carries the epilogue warps' fp32 state stores (observed through epi_done)
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.done + pt * m_tiles + pmt) : "memory");
                }
                (void)i_prev;
            };
            for (int item = item_begin; item < item_end; item += num_clusters, ++it) {
                const int t = item / per_step, rem = item % per_step;
                const int mt = rem / n_tiles, nt = rem % n_tiles;
                const int m0 = mt * (2 * kTileM) + row_base;
                if (it >= 1) ptx::mbar_wait(gi_free, (it - 1) & 1);  // the epilogue holds the previous gi in registers
                // gi[t] itself has no dependency (GEMM2 finished before the launch): refill right away
                if (dbg & 4) {
                    ptx::mbar_arrive(gi_full);
                } else {
                ptx::mbar_expect_tx(gi_full, kGruGiBytes);
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    ptx::tma_load_3d(&tmGi, gi_smem + g * kGruBoxBytes, gi_full, nt * kGruTileN + g * 64, m0, t, ptx::kEvictFirst);
                }
                {   // gi streams from HBM (GEMM2's 1.6 GB output): warm L2 for this pair's NEXT item now
                    const int nitem = item + num_clusters;
                    if (nitem < item_end && !(dbg & 4)) {
                        const int t2 = nitem / per_step, rem2 = nitem % per_step;
                        const int m2 = (rem2 / n_tiles) * (2 * kTileM) + row_base, nt2 = rem2 % n_tiles;
#pragma unroll
                        for (int g = 0; g < 3; ++g) ptx::tma_prefetch_3d(&tmGi, nt2 * kGruTileN + g * 64, m2, t2);
                    }
                }
                if (it >= 1) {
                    ptx::mbar_wait(epi_done, (it - 1) & 1);
                    store_and_publish(it - 1);
                }
                pm0 = m0; pnt = nt; pt = t; pmt = mt;
            }
            if (it >= 1) {
                ptx::mbar_wait(epi_done, (it - 1) & 1);
                store_and_publish(it - 1);
            }
            ptx::tma_store_wait_all();
        }
    } else {
        // -------------------------------------------------------------------------- epilogue
        const int quad = warp & 3;
        const int ugrp = (warp - 2) >> 2;  // which 16 hidden units of the tile
        const int r = quad * 32 + lane;    // row of the tile = TMEM lane
        const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        const uint32_t s_box = ptx::smem_u32(gi_smem) + row_off;
        const uint32_t s_out = ptx::smem_u32(out_smem) + row_off;
        int it = 0;
        long long w_gi = 0, w_accf = 0, w_outf = 0, t_all = now(), c0;
        for (int item = item_begin; item < item_end; item += num_clusters, ++it) {
            const int buf = it & 1;
            const int rem = item % per_step;
            const int m0 = (rem / n_tiles) * (2 * kTileM) + row_base;
            const int nt = rem % n_tiles;
            const int row = m0 + r;
            // fp32 state of step t-1 for this tile.  It may be loaded as soon as the item's dependency is known to be
            // satisfied: always (per-step launches), or when a non-blocking look at the counter says so -- the usual
            // case, the dependency is a whole step old -- so the load hides behind the accumulator wait.  Otherwise
            // it is loaded after acc_full: the operand producer waited for the dependency before loading h_{t-1},
            // and the accumulator cannot complete before those loads.  L1 is bypassed (another SM wrote the data).
            float4* hptr = reinterpret_cast<float4*>(a.h32t) +
                           ((static_cast<int64_t>(row >> 7) * (H / 64) + nt) * 16 + ugrp * 4) * 128 + (row & 127);
            float4 hcur[4];
            c0 = now();
            ptx::mbar_wait(gi_full, it & 1);
            w_gi += now() - c0;
            if (dbg & 16) {  // handshakes only
                ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
                ptx::tc_fence_after();
                ptx::mbar_wait(out_free, (it & 1) ^ 1);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(gi_free);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive_leader(&acc_empty[buf]);
                    ptx::mbar_arrive(epi_done);
                }
                continue;
            }
            bool early = true;
            if (a.done != nullptr) {
                const int t = item / per_step;
                if (t > a.t_begin) {
                    uint32_t v = 0;
                    if (lane == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.done + (t - 1) * m_tiles + rem / n_tiles) : "memory");
                    early = __shfl_sync(0xffffffffu, v, 0) >= dep_target;
                }
            }
            if (early) {
#pragma unroll
                for (int i = 0; i < 4; ++i) hcur[i] = __ldcg(hptr + i * 128);
            }
            c0 = now();
            ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
            w_accf += now() - c0;
            ptx::tc_fence_after();
            if (!early) {
#pragma unroll
                for (int i = 0; i < 4; ++i) hcur[i] = __ldcg(hptr + i * 128);
            }
            c0 = now();
            ptx::mbar_wait(out_free, (it & 1) ^ 1);  // the previous item's staged results have left
            w_outf += now() - c0;
            const uint32_t taddr = tmem_base + buf * 256 + (static_cast<uint32_t>(quad * 32) << 16);
            const float* bh = a.bhh + nt * kGruTileN;
            // gi of this thread's 16 units -> registers, then hand the boxes back so the next item's gi streams in
            // behind this item's gate math
            uint4 gq[2][3];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t a_gi = s_box + ((static_cast<uint32_t>(ugrp * 2 + i) ^ sw) << 4);
                gq[i][0] = lds128(a_gi);
                gq[i][1] = lds128(a_gi + kGruBoxBytes);
                gq[i][2] = lds128(a_gi + 2 * kGruBoxBytes);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(gi_free);
#pragma unroll
            for (int i = 0; i < 2; ++i) {  // 8 hidden units per iteration
                const int c = ugrp * 2 + i;
                uint32_t vr[8], vz[8], vn[8];
                ptx::tmem_ld8(taddr + c * 8, vr);
                ptx::tmem_ld8(taddr + 64 + c * 8, vz);
                ptx::tmem_ld8(taddr + 128 + c * 8, vn);
                const uint4 qr = gq[i][0], qz = gq[i][1], qn = gq[i][2];
                float bn[8];  // b_hr / b_hz are pre-summed into gi by the input-gate GEMM; b_hn stays inside r * (.)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(bh + 128 + c * 8) + q);
                    bn[4 * q] = w.x; bn[4 * q + 1] = w.y; bn[4 * q + 2] = w.z; bn[4 * q + 3] = w.w;
                }
                const uint32_t gr[4] = {qr.x, qr.y, qr.z, qr.w}, gz[4] = {qz.x, qz.y, qz.z, qz.w},
                               gn[4] = {qn.x, qn.y, qn.z, qn.w};
                const float hp[8] = {hcur[2 * i].x, hcur[2 * i].y, hcur[2 * i].z, hcur[2 * i].w,
                                     hcur[2 * i + 1].x, hcur[2 * i + 1].y, hcur[2 * i + 1].z, hcur[2 * i + 1].w};
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
                        const int u = 2 * j + e;
                        if (dbg & 1) { hn[u] = hp[u] + gir[e] + giz[e] + gin[e] + __uint_as_float(vr[u]) + __uint_as_float(vz[u]) + __uint_as_float(vn[u]); continue; }
                        const float rr = fast_sigmoid(gir[e] + __uint_as_float(vr[u]));
                        const float zz = fast_sigmoid(giz[e] + __uint_as_float(vz[u]));
                        const float nn = fast_tanh(gin[e] + rr * (__uint_as_float(vn[u]) + bn[u]));
                        hn[u] = (hp[u] - nn) * zz + nn;
                    }
                }
                // fp32 master state straight from registers (padding rows are allocated, no mask needed)
                hptr[(2 * i) * 128] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                hptr[(2 * i + 1) * 128] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                uint4 os, orl;
                os.x = Op::pack2(hn[0], hn[1]); os.y = Op::pack2(hn[2], hn[3]);
                os.z = Op::pack2(hn[4], hn[5]); os.w = Op::pack2(hn[6], hn[7]);
#pragma unroll
                for (int u = 0; u < 8; ++u) hn[u] = fmaxf(hn[u], 0.0f);
                orl.x = Op::pack2(hn[0], hn[1]); orl.y = Op::pack2(hn[2], hn[3]);
                orl.z = Op::pack2(hn[4], hn[5]); orl.w = Op::pack2(hn[6], hn[7]);
                const uint32_t a_out = s_out + ((static_cast<uint32_t>(c) ^ sw) << 4);
                sts128(a_out, os);                  // operand copy of h_t
                sts128(a_out + kGruBoxBytes, orl);  // relu(h_t)
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();  // make the st.shared results visible to the TMA store
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive_leader(&acc_empty[buf]);
                ptx::mbar_arrive(epi_done);  // release: the publisher's gpu-scope fence is cumulative over these stores
            }
        }
        if (stats != nullptr && warp == 2 && lane == 0) {
            long long* st = stats + blockIdx.x * 16;
            st[8] = now() - t_all; st[9] = w_gi; st[10] = w_accf; st[11] = w_outf;
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

}  // namespace prego
