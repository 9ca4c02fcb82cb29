// Persistent single-/few-stream GRU recurrence (rnn.py:61, the T-step hot loop) for the
// latency regime (B <= 4 streams per pass).
//
// * weight_hh_l0 (3H x H fp32, 12.6 MB for H = 1024) is partitioned over H/8 CTAs and stays
//   resident in shared memory for the whole sequence (96 KB per CTA): no weight byte is
//   re-read from HBM/L2 after the prologue.
// * One warp owns one hidden unit: three length-H dot products per stream (conflict-free
//   smem reads, warp-shuffle reduction), then the fused sigmoid/tanh/state update.
// * The per-step all-to-all exchange of h needs no separate grid barrier: every h value is
//   published as one 8-byte word {fp32 value, step tag}; consumers poll the tagged words
//   straight from L2 (LL-protocol style), so a time step costs one store->load round trip.
//   Exchange slots are double-buffered by step parity.
// * Arithmetic is exact fp32 in ATen's order:  h' = (h - n) * z + n.
//
// Must be launched with cudaLaunchCooperativeKernel (co-residency of all CTAs).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace prego {

constexpr int kLatUnitsPerCta = 8;
constexpr int kLatThreads = kLatUnitsPerCta * 32;

struct GruLatencyArgs {
    const float* whh;      // [3H, H] fp32, packed gate-interleaved row order
    const float* bhh;      // [3H] packed order
    const float* gi;       // fp32 gate pre-activations, packed column order (b_ih folded in);
                           // row of (stream b, step t) = b * row_sb + t * row_st
    const float* h_in;     // [B, H] state before the first step of this launch
    float* h_out;          // [B, H] state after the last step (must not alias h_in)
    void* hrelu;           // relu(h_t), same row indexing; fp32 / fp16 / bf16 (out_fmt -1 / 0 / 1)
    uint2* xchg;           // [2][NB][H] tagged exchange words
    int* err_flag;         // set to 1 on spin timeout
    int H, Tc;
    int b0;                // first stream of this pass
    int nb;                // streams in this pass (<= NB)
    uint32_t tag_base;     // tags used: tag_base + 1 .. tag_base + Tc
    int out_fmt;
    int64_t row_sb, row_st;
    // training forward (rnn.py:61 in train mode): gates and states saved for BPTT, time-major [T][sv_B][H]
    // (sv_h: [T + 1][sv_B][H], slot t + 1 = h_t); all NULL in inference
    float *sv_r = nullptr, *sv_z = nullptr, *sv_n = nullptr, *sv_ghn = nullptr, *sv_h = nullptr;
    int sv_B = 0;
};

// REGW = true (H == 1024): the warp's three weight rows live in REGISTERS (96 per lane) -- the per-step
// smem traffic drops from 96 KB of weights to the 4 KB/stream of h; REGW = false keeps them in smem (any H).
// G > 1: G independent groups of NB streams interleaved in one launch (streams b0 + g NB ..): a group's exchange is in
// flight while the CTA computes the other group's step, so the store -> load round trip leaves the critical path.
template <int NB, bool REGW, int G = 1>
__global__ void __launch_bounds__(kLatThreads, 1)
gru_latency_kernel(GruLatencyArgs a) {
    extern __shared__ float smem_f[];
    const int H = REGW ? 1024 : a.H;
    float* wsm = smem_f;                                              // [3][8][H]   (REGW: unused, size 0)
    float* hbuf_all = smem_f + (REGW ? 0 : 3 * kLatUnitsPerCta * H);  // [G][2][NB][H]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u = blockIdx.x * kLatUnitsPerCta + warp;      // hidden unit of this warp
    const int pcol = (u / 64) * 192 + (u % 64);             // packed column of gate r; z: +64, n: +128

    // Prologue: this CTA's 24 weight rows -> registers (REGW) or shared memory.
    float wreg[3][32];
    if constexpr (REGW) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float* wrow = a.whh + static_cast<int64_t>((u / 64) * 192 + g * 64 + (u % 64)) * H;
#pragma unroll
            for (int i = 0; i < 32; ++i) wreg[g][i] = __ldg(wrow + lane + 32 * i);
        }
    }
    for (int idx = tid; !REGW && idx < 3 * kLatUnitsPerCta * (H / 4); idx += kLatThreads) {
        const int row = idx / (H / 4), c4 = idx % (H / 4);
        const int g = row / kLatUnitsPerCta, w = row % kLatUnitsPerCta;
        const int uu = blockIdx.x * kLatUnitsPerCta + w;
        const int prow = (uu / 64) * 192 + g * 64 + (uu % 64);
        reinterpret_cast<float4*>(wsm)[idx] = __ldg(reinterpret_cast<const float4*>(a.whh + static_cast<int64_t>(prow) * H) + c4);
    }
    float bh[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) bh[g] = __ldg(a.bhh + pcol + 64 * g);

    for (int t = 0; t < a.Tc; ++t) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int nb_g = a.nb - g * NB < NB ? a.nb - g * NB : NB;  // streams of this group (CTA-uniform)
        if (G > 1 && nb_g <= 0) continue;
        const int b0_g = a.b0 + g * NB;
        float* hbuf = hbuf_all + g * 2 * NB * H;
        uint2* xchg_g = a.xchg + static_cast<int64_t>(g) * 2 * NB * H;
        // gate pre-activations of this step (independent of h: issue before polling)
        float gv[3] = {0.f, 0.f, 0.f};
        if (lane < nb_g) {
            const float* gp = a.gi + ((b0_g + lane) * a.row_sb + t * a.row_st) * (3 * H) + pcol;
            gv[0] = __ldcs(gp);
            gv[1] = __ldcs(gp + 64);
            gv[2] = __ldcs(gp + 128);
        }
        float* hb = hbuf + (t & 1) * NB * H;
        int timed_out = 0;
        if (t == 0) {
            for (int idx = tid; idx < NB * H; idx += kLatThreads) {
                const int s = idx / H, k = idx % H;
                hb[idx] = (s < nb_g) ? a.h_in[static_cast<int64_t>(b0_g + s) * H + k] : 0.f;
            }
        } else {
            const uint32_t want = a.tag_base + static_cast<uint32_t>(t);
            const uint2* xs = xchg_g + ((t - 1) & 1) * NB * H;
            // all of this thread's words are requested in one batch per poll round: the step costs ~one L2 round trip
            constexpr int PW = REGW ? NB * 1024 / kLatThreads : 4;  // words per thread per batch (NB = 8: 32 words = 64 registers)
            for (int idx = tid; idx < NB * H; idx += PW * kLatThreads) {
                uint2 v[PW];
                long long spins = 0;
                bool done;
                do {
                    done = true;
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        const int ii = idx + j * kLatThreads;
                        const bool live = ii < NB * H && (ii / H) < nb_g;
                        v[j] = live ? ptx::ld_volatile_u64(xs + ii) : make_uint2(0u, want);
                    }
#pragma unroll
                    for (int j = 0; j < PW; ++j) done = done && (v[j].y == want);
                    if (!done && ++spins > (1ll << 22)) {  // ~seconds: a peer CTA is missing
                        *a.err_flag = 1;
                        timed_out = 1;
                        done = true;
                    }
                } while (!done);
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    const int ii = idx + j * kLatThreads;
                    if (ii < NB * H) hb[ii] = __uint_as_float(v[j].x);
                }
            }
        }
        if (__syncthreads_or(timed_out)) return;  // never hang the GPU: bail out CTA-uniformly

        float acc[3][NB];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int s = 0; s < NB; ++s) acc[g][s] = 0.f;
        const float* w0 = wsm + (0 * kLatUnitsPerCta + warp) * H;
        const float* w1 = wsm + (1 * kLatUnitsPerCta + warp) * H;
        const float* w2 = wsm + (2 * kLatUnitsPerCta + warp) * H;
        if constexpr (REGW) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
#pragma unroll
                for (int s = 0; s < NB; ++s) {
                    const float hv = hb[s * H + lane + 32 * i];
                    acc[0][s] = fmaf(wreg[0][i], hv, acc[0][s]);
                    acc[1][s] = fmaf(wreg[1][i], hv, acc[1][s]);
                    acc[2][s] = fmaf(wreg[2][i], hv, acc[2][s]);
                }
            }
        } else {
#pragma unroll 8
            for (int k = lane; k < H; k += 32) {
                const float wr = w0[k], wz = w1[k], wn = w2[k];
#pragma unroll
                for (int s = 0; s < NB; ++s) {
                    const float hv = hb[s * H + k];
                    acc[0][s] = fmaf(wr, hv, acc[0][s]);
                    acc[1][s] = fmaf(wz, hv, acc[1][s]);
                    acc[2][s] = fmaf(wn, hv, acc[2][s]);
                }
            }
        }
        // lane s finishes stream s
        float ghr = 0.f, ghz = 0.f, ghn = 0.f;
        if constexpr (NB == 8) {
            // 24 sums over 32 lanes by a halving butterfly (31 shuffles instead of 120): value index gate * 8 + stream, padded
            // to 32; afterwards lane l holds the total of index l
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = i < 24 ? acc[i / 8][i % 8] : 0.f;
#pragma unroll
            for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float send = hi ? v[i] : v[i + n];
                    const float keep = hi ? v[i + n] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            ghr = v[0];
            ghz = __shfl_sync(0xffffffffu, v[0], (lane & 7) + 8);
            ghn = __shfl_sync(0xffffffffu, v[0], (lane & 7) + 16);
        } else {
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int s = 0; s < NB; ++s)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[g][s] += __shfl_xor_sync(0xffffffffu, acc[g][s], o);
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                if (lane == s) {
                    ghr = acc[0][s];
                    ghz = acc[1][s];
                    ghn = acc[2][s];
                }
            }
        }
        if (lane < nb_g) {
            const float hp = hb[lane * H + u];
            const float r = sigmoid_f(gv[0] + (ghr + bh[0]));
            const float z = sigmoid_f(gv[1] + (ghz + bh[1]));
            const float n = tanhf(gv[2] + r * (ghn + bh[2]));
            const float hn = (hp - n) * z + n;
            ptx::st_volatile_u64(xchg_g + ((t & 1) * NB + lane) * H + u, __float_as_uint(hn),
                                 a.tag_base + static_cast<uint32_t>(t) + 1u);
            const int64_t orow = (b0_g + lane) * a.row_sb + t * a.row_st;
            if (a.out_fmt < 0)
                reinterpret_cast<float*>(a.hrelu)[orow * H + u] = fmaxf(hn, 0.f);
            else if (a.out_fmt == 0)
                reinterpret_cast<__half*>(a.hrelu)[orow * H + u] = __float2half_rn(fmaxf(hn, 0.f));
            else
                reinterpret_cast<__nv_bfloat16*>(a.hrelu)[orow * H + u] = __float2bfloat16_rn(fmaxf(hn, 0.f));
            if (t == a.Tc - 1) a.h_out[static_cast<int64_t>(b0_g + lane) * H + u] = hn;
            if (a.sv_r != nullptr) {
                const int64_t si = (static_cast<int64_t>(t) * a.sv_B + b0_g + lane) * H + u;
                a.sv_r[si] = r;
                a.sv_z[si] = z;
                a.sv_n[si] = n;
                a.sv_ghn[si] = ghn + bh[2];
                a.sv_h[si + static_cast<int64_t>(a.sv_B) * H] = hn;
            }
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent BPTT through the GRU recurrence (the backward of rnn.py:61), same structure run in reverse: one warp
// owns hidden unit u, keeps COLUMN u of W_hh' (3H fp32 = 96 registers per lane) for the whole sequence and carries
// d h[:, u] in registers.  Per step t (descending):
//   dh   = dh_carry + dhrelu_t * [h_t > 0]
//   dn~  = dh (1 - z) (1 - n^2);  dz~ = dh (h_{t-1} - n) z (1 - z);  dr~ = dn~ ghn r (1 - r)
//   dgi_t = (dr~, dz~, dn~),  dgh_t = (dr~, dz~, dn~ r)            -> global (inputs of the weight-gradient GEMMs)
//   dh_carry' = dh z + sum_p dgh_t[p] W_hh'[p, u]                   (all-to-all of dgh_t: tagged 8-byte words, as above)
struct GruBpttArgs {
    const float* whhT;     // [H, 3H] fp32: W_hh' transposed (row u = column u of the packed matrix)
    const float *r, *z, *n, *ghn;  // [T][B][H]
    const float* hall;     // [T + 1][B][H]
    const float* dhrelu;   // [T][B][H]
    float *dgi, *dgh;      // [T][B][3H] packed columns
    uint4* xchg;           // [2][NB][H] tagged exchange words {dr~, dz~, dn~ r, tag}: one aligned 16-byte store each
    int* err_flag;
    int T, B, b0, nb;
    uint32_t tag_base;     // tags used: tag_base + 1 .. tag_base + T
};

template <int NB>
__global__ void __launch_bounds__(kLatThreads, 1)
gru_bptt_kernel(GruBpttArgs a) {
    constexpr int H = 1024, H3 = 3 * H;
    extern __shared__ float smem_f[];  // [2][NB][3H]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u = blockIdx.x * kLatUnitsPerCta + warp;
    const int pcol = (u / 64) * 192 + (u % 64);
    float wcol[96];
#pragma unroll
    for (int i = 0; i < 96; ++i) wcol[i] = __ldg(a.whhT + static_cast<int64_t>(u) * H3 + lane + 32 * i);
    float dh_carry = 0.f;  // lane s: d h_t[stream b0 + s, u] flowing in from step t + 1
    for (int step = 0; step < a.T; ++step) {
        const int t = a.T - 1 - step;
        float dgate_n_r = 0.f, dhz = 0.f;
        if (lane < a.nb) {
            const int64_t si = (static_cast<int64_t>(t) * a.B + a.b0 + lane) * H + u;
            const float r = __ldcs(a.r + si), z = __ldcs(a.z + si), n = __ldcs(a.n + si), ghn = __ldcs(a.ghn + si);
            const float h_t = __ldcs(a.hall + si + static_cast<int64_t>(a.B) * H), h_prev = __ldcs(a.hall + si);
            const float dh = dh_carry + (h_t > 0.f ? __ldcs(a.dhrelu + si) : 0.f);
            const float dn = dh * (1.0f - z);
            const float dz = dh * (h_prev - n);
            const float dan = dn * (1.0f - n * n);
            const float dar = dan * ghn * r * (1.0f - r);
            const float daz = dz * z * (1.0f - z);
            dgate_n_r = dan * r;
            dhz = dh * z;
            const int64_t gi = (static_cast<int64_t>(t) * a.B + a.b0 + lane) * H3 + pcol;
            a.dgi[gi] = dar; a.dgi[gi + 64] = daz; a.dgi[gi + 128] = dan;
            a.dgh[gi] = dar; a.dgh[gi + 64] = daz; a.dgh[gi + 128] = dgate_n_r;
            ptx::st_volatile_u128(a.xchg + (static_cast<int64_t>(step & 1) * NB + lane) * H + u,
                                  make_uint4(__float_as_uint(dar), __float_as_uint(daz), __float_as_uint(dgate_n_r),
                                             a.tag_base + static_cast<uint32_t>(step) + 1u));
        }
        if (t == 0) break;  // d h_{-1} is not needed (h0 is a constant, rnn.py:49)
        // gather dgh_t of all units
        float* db = smem_f + (step & 1) * NB * H3;
        const uint32_t want = a.tag_base + static_cast<uint32_t>(step) + 1u;
        const uint4* xs = a.xchg + static_cast<int64_t>(step & 1) * NB * H;
        int timed_out = 0;
        constexpr int PW = NB * H / kLatThreads;  // words per thread, all requested in one batch per poll round
        {
            uint4 v[PW];
            long long spins = 0;
            bool done;
            do {
                done = true;
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    const int ii = tid + j * kLatThreads;
                    v[j] = (ii / H) < a.nb ? ptx::ld_volatile_u128(xs + ii) : make_uint4(0u, 0u, 0u, want);
                }
#pragma unroll
                for (int j = 0; j < PW; ++j) done = done && (v[j].w == want);
                if (!done && ++spins > (1ll << 22)) {
                    *a.err_flag = 2;
                    timed_out = 1;
                    done = true;
                }
            } while (!done);
#pragma unroll
            for (int j = 0; j < PW; ++j) {
                const int ii = tid + j * kLatThreads;
                const int sidx = ii / H, uu = ii % H;
                float* d = db + sidx * H3 + (uu / 64) * 192 + (uu % 64);
                d[0] = __uint_as_float(v[j].x);
                d[64] = __uint_as_float(v[j].y);
                d[128] = __uint_as_float(v[j].z);
            }
        }
        if (__syncthreads_or(timed_out)) return;
        float acc[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) acc[s] = 0.f;
#pragma unroll
        for (int i = 0; i < 96; ++i)
#pragma unroll
            for (int s = 0; s < NB; ++s) acc[s] = fmaf(wcol[i], db[s * H3 + lane + 32 * i], acc[s]);
#pragma unroll
        for (int s = 0; s < NB; ++s)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
        float mine = 0.f;
#pragma unroll
        for (int s = 0; s < NB; ++s)
            if (lane == s) mine = acc[s];
        dh_carry = dhz + mine;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same BPTT step with the contraction split over the THREADS of the CTA instead of over the lanes of one warp: thread
// `tid` keeps W_hh'[p, u] for its 12 rows p = 1024 j + 4 tid + c and ALL 8 units of the CTA (96 registers), so every
// d gh value read from shared memory feeds 8 FMAs (the warp-per-unit mapping above reads one value per FMA and is bound by
// shared-memory bandwidth: 96 KB per warp per step and stream).  The 8 x NB partial sums per thread are reduced with a
// halving butterfly (V - V/32 shuffles for V values) and one pass through shared memory across the 8 warps.
// G groups of NB streams (b0 + g NB ..) are interleaved: a group publishes its d gh words right after its own product, then
// the CTA gathers and contracts the OTHER group's step, so the store -> load round trip of the exchange and the loads of the
// saved gates are off the critical path.  NB = 4 or 8, G = 1 or 2.
template <int NB, int G = 1>
__global__ void __launch_bounds__(kLatThreads, 1)
gru_bptt2_kernel(GruBpttArgs a) {
    constexpr int H = 1024, H3 = 3 * H, U = kLatUnitsPerCta, V = U * NB;
    static_assert(V % 32 == 0 && kLatThreads * 4 == H, "thread t owns rows 4t..4t+3 of each gate block");
    extern __shared__ float smem_f[];
    float* db = smem_f;             // [NB][3H] d gh of the step being contracted, packed column order
    float* red = smem_f + NB * H3;  // [8 warps][V] per-warp partial sums
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u0 = blockIdx.x * U, u = u0 + warp;
    const int pcol = (u / 64) * 192 + (u % 64);
    float w[3][4][U];
#pragma unroll
    for (int k = 0; k < U; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(a.whhT + static_cast<int64_t>(u0 + k) * H3 + 1024 * j + 4 * tid));
            w[j][0][k] = q.x; w[j][1][k] = q.y; w[j][2][k] = q.z; w[j][3][k] = q.w;
        }
    float pr = 0.f, pz = 0.f, pn = 0.f, pghn = 0.f, ph_t = 0.f, ph_prev = 0.f, pdhr = 0.f;  // saved values of the step about to run
    auto fetch = [&](int g, int t) {
        const int64_t si = (static_cast<int64_t>(t) * a.B + a.b0 + g * NB + lane) * H + u;
        pr = __ldcs(a.r + si); pz = __ldcs(a.z + si); pn = __ldcs(a.n + si); pghn = __ldcs(a.ghn + si);
        ph_t = __ldcs(a.hall + si + static_cast<int64_t>(a.B) * H); ph_prev = __ldcs(a.hall + si);
        pdhr = __ldcs(a.dhrelu + si);
    };
    float dh_carry[G], dhz[G];
#pragma unroll
    for (int g = 0; g < G; ++g) dh_carry[g] = dhz[g] = 0.f;
    // gate derivatives of (group g, step) from the fetched values: d gi / d gh to global, the exchange words published
    auto gates = [&](int g, int step) {
        const int t = a.T - 1 - step;
        const float r = pr, z = pz, n = pn, ghn = pghn;
        const float dh = dh_carry[g] + (ph_t > 0.f ? pdhr : 0.f);
        const float dn = dh * (1.0f - z);
        const float dz = dh * (ph_prev - n);
        const float dan = dn * (1.0f - n * n);
        const float dar = dan * ghn * r * (1.0f - r);
        const float daz = dz * z * (1.0f - z);
        const float danr = dan * r;
        dhz[g] = dh * z;
        ptx::st_volatile_u128(a.xchg + ((static_cast<int64_t>(g) * 2 + (step & 1)) * NB + lane) * H + u,
                              make_uint4(__float_as_uint(dar), __float_as_uint(daz), __float_as_uint(danr),
                                         a.tag_base + static_cast<uint32_t>(step) + 1u));
        const int64_t gi = (static_cast<int64_t>(t) * a.B + a.b0 + g * NB + lane) * H3 + pcol;
        a.dgi[gi] = dar; a.dgi[gi + 64] = daz; a.dgi[gi + 128] = dan;
        a.dgh[gi] = dar; a.dgh[gi + 64] = daz; a.dgh[gi + 128] = danr;
    };
    int nbg[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        nbg[g] = a.nb - g * NB < NB ? a.nb - g * NB : NB;  // CTA-uniform; <= 0: group absent
        if (lane < nbg[g]) {
            fetch(g, a.T - 1);
            gates(g, 0);
        }
    }
    for (int step = 0; step + 1 < a.T; ++step) {  // d h_{-1} is not needed (h0 is a constant, rnn.py:49)
        const int t = a.T - 1 - step;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (G > 1 && nbg[g] <= 0) continue;
            const bool live = lane < nbg[g];
            if (live) fetch(g, t - 1);
            // gather d gh_t of all units of this group into shared memory
            const uint32_t want = a.tag_base + static_cast<uint32_t>(step) + 1u;
            const uint4* xs = a.xchg + (static_cast<int64_t>(g) * 2 + (step & 1)) * NB * H;
            int timed_out = 0;
            constexpr int WORDS = NB * H / kLatThreads;  // words per thread
            constexpr int PW = WORDS < 16 ? WORDS : 16;  // requested in one batch per poll round
#pragma unroll 1
            for (int c0 = 0; c0 < WORDS; c0 += PW) {
                uint4 v[PW];
                long long spins = 0;
                bool done;
                do {
                    done = true;
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        const int ii = tid + (c0 + j) * kLatThreads;
                        v[j] = (ii / H) < nbg[g] ? ptx::ld_volatile_u128(xs + ii) : make_uint4(0u, 0u, 0u, want);
                    }
#pragma unroll
                    for (int j = 0; j < PW; ++j) done = done && (v[j].w == want);
                    if (!done && ++spins > (1ll << 22)) {
                        *a.err_flag = 2;
                        timed_out = 1;
                        done = true;
                    }
                } while (!done);
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    const int ii = tid + (c0 + j) * kLatThreads;
                    const int sidx = ii / H, uu = ii % H;
                    float* d = db + sidx * H3 + (uu / 64) * 192 + (uu % 64);
                    d[0] = __uint_as_float(v[j].x);
                    d[64] = __uint_as_float(v[j].y);
                    d[128] = __uint_as_float(v[j].z);
                }
            }
            if (__syncthreads_or(timed_out)) return;
            // partial products of this thread's 12 rows for the CTA's 8 units
            float acc[V];  // index k * NB + s
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                if (s < nbg[g]) {
                    const float* ds = db + s * H3 + 4 * tid;
                    const float4 d0 = *reinterpret_cast<const float4*>(ds);
                    const float4 d1 = *reinterpret_cast<const float4*>(ds + 1024);
                    const float4 d2 = *reinterpret_cast<const float4*>(ds + 2048);
#pragma unroll
                    for (int k = 0; k < U; ++k) {
                        float x = w[0][0][k] * d0.x;
                        x = fmaf(w[0][1][k], d0.y, x); x = fmaf(w[0][2][k], d0.z, x); x = fmaf(w[0][3][k], d0.w, x);
                        x = fmaf(w[1][0][k], d1.x, x); x = fmaf(w[1][1][k], d1.y, x); x = fmaf(w[1][2][k], d1.z, x); x = fmaf(w[1][3][k], d1.w, x);
                        x = fmaf(w[2][0][k], d2.x, x); x = fmaf(w[2][1][k], d2.y, x); x = fmaf(w[2][2][k], d2.z, x); x = fmaf(w[2][3][k], d2.w, x);
                        acc[k * NB + s] = x;
                    }
                }
            }
            // halving butterfly over the 32 lanes: afterwards lane l holds the warp totals of indices l * V/32 .. + V/32 - 1
#pragma unroll
            for (int off = 16, n = V / 2; off >= 1; off >>= 1, n >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float send = hi ? acc[i] : acc[i + n];
                    const float keep = hi ? acc[i + n] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
#pragma unroll
            for (int i = 0; i < V / 32; ++i) red[warp * V + lane * (V / 32) + i] = acc[i];
            __syncthreads();
            if (live) {
                float mine = 0.f;
#pragma unroll
                for (int ww = 0; ww < kLatUnitsPerCta; ++ww) mine += red[ww * V + warp * NB + lane];
                dh_carry[g] = dhz[g] + mine;
                gates(g, step + 1);  // published now; gathered after the other group's step
            }
        }
    }
}

}  // namespace prego
