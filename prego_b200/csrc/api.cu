// C ABI of prego_b200 (see include/prego_b200.h): model handle, weight packing, the forward
// pipeline (time-chunked, carried GRU state) and the aggregation entry points.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <mutex>
#include <new>
#include <tuple>
#include <vector>

#include "../../include/prego_b200.h"
#include "aggregate.cuh"
#include "gemm_tc.cuh"
#include "gemm_xf.cuh"
#include "gru_latency.cuh"
#include "gru_step.cuh"
#include "metrics.cuh"
#include "online_fused.cuh"
#include "online_kernels.cuh"
#include "simt_kernels.cuh"
#include "train_kernels.cuh"

using namespace prego;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            (void)cudaGetLastError(); /* do not leave it for the next launch check to trip over */  \
            return fail(PREGO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

#define LAUNCH_CHECK(name)                                                                          \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess)                                                                      \
            return fail(PREGO_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define RC_TRY(expr)                     \
    do {                                 \
        int _rc = (expr);                \
        if (_rc != PREGO_OK) return _rc; \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;

int get_encode_fn() {
    if (g_encode != nullptr) return PREGO_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
        return fail(PREGO_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return PREGO_OK;
}

enum DType { kF16 = 0, kBF16 = 1, kF32 = 2 };

// Generic tiled map with the 128-byte swizzle (inner box extent is always 128 bytes).
// dims / box: innermost first; strides_bytes: for dims 1..rank-1.
int make_tmap(CUtensorMap* tm, DType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
    RC_TRY(get_encode_fn());
    cuuint64_t d[3], s[2];
    cuuint32_t b[3], e[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        if (i > 0) s[i - 1] = strides_bytes[i - 1];
    }
    const CUtensorMapDataType cdt = dt == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                  : dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = g_encode(tm, cdt, rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PREGO_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
    return PREGO_OK;
}

// A operand of the generic GEMM: row-major [rows, k] 16-bit, viewed as (k, 1, rows), box (64, 1, 128).
int make_tmap_a(CUtensorMap* tm, DType dt, const void* base, uint64_t k, uint64_t rows) {
    const uint64_t dims[3] = {k, 1, rows}, str[2] = {k * 2, k * 2};
    const uint32_t box[3] = {kTileK, 1, kTileM};
    return make_tmap(tm, dt, 3, base, dims, str, box);
}
// W operand: row-major [n, k] 16-bit, box (64, tile_n).
int make_tmap_w(CUtensorMap* tm, DType dt, const void* base, uint64_t k, uint64_t n, uint32_t tile_n) {
    const uint64_t dims[2] = {k, n}, str[1] = {k * 2};
    const uint32_t box[2] = {kTileK, tile_n};
    return make_tmap(tm, dt, 2, base, dims, str, box);
}
// fp32 (tf32) operands of the CTA-pair GEMM: 32 elements = 128 bytes per swizzle row.
int make_tmap_a32(CUtensorMap* tm, const void* base, uint64_t k, uint64_t rows, uint64_t ld_elems) {
    const uint64_t dims[3] = {k, 1, rows}, str[2] = {ld_elems * 4, ld_elems * 4};
    const uint32_t box[3] = {32, 1, kTileM};
    return make_tmap(tm, kF32, 3, base, dims, str, box);
}
int make_tmap_w32(CUtensorMap* tm, const void* base, uint64_t k, uint64_t n, uint64_t ld_elems, uint32_t box_rows) {
    const uint64_t dims[2] = {k, n}, str[1] = {ld_elems * 4};
    const uint32_t box[2] = {32, box_rows};
    return make_tmap(tm, kF32, 2, base, dims, str, box);
}

// time-major activation [slots, B, cols] (elem_bytes each), box (128 B worth of cols, 128 rows, 1).
int make_tmap_tm(CUtensorMap* tm, DType dt, const void* base, uint64_t cols, uint64_t B, uint64_t slots, int elem_bytes) {
    const uint64_t dims[3] = {cols, B, slots}, str[2] = {cols * elem_bytes, B * cols * elem_bytes};
    const uint32_t box[3] = {128u / elem_bytes, kTileM, 1};
    return make_tmap(tm, dt, 3, base, dims, str, box);
}

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
inline int grid_for(int64_t work_items, int threads, int sm_count) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, DEVICE): a model created on a second GPU of
// the same process needs its own opt-in.  Remembers the largest size set per (function, current device).
int ensure_dyn_smem(const void* fn, int bytes) {
    static std::mutex mu;
    static std::vector<std::tuple<const void*, int, int>> done;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : done) {
        if (std::get<0>(e) == fn && std::get<1>(e) == dev) {
            if (std::get<2>(e) >= bytes) return PREGO_OK;
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            std::get<2>(e) = bytes;
            return PREGO_OK;
        }
    }
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.emplace_back(fn, dev, bytes > 48 * 1024 ? bytes : 48 * 1024);
    return PREGO_OK;
}

constexpr int kProfMaxEvents = 8192;
constexpr int kLatencyMaxB = 16;  // up to this many streams the recurrence runs on the persistent SIMT kernel

}  // namespace

struct prego_model {
    prego_dims_t d;
    int device = 0;
    int sm_count = 0;
    int din = 0;
    int kpad = 0;  // padded class count of the tcgen05 head (96 or 128), 0 if unsupported
    bool loaded = false;
    uint32_t packed = 0;  // PREGO_PACK_* formats that match the last loaded weights
    // fp32 (exact path + SIMT recurrence)
    float *w1_f32 = nullptr, *b1 = nullptr, *ln_g = nullptr, *ln_b = nullptr;
    float *wih_f32p = nullptr, *whh_f32p = nullptr, *bih_p = nullptr, *bhh_p = nullptr;
    float* bgi_p = nullptr;  // b_ih' + (r, z parts of b_hh'): bias of the input-gate GEMM on the batched 16-bit path
    float *wc_f32 = nullptr, *bc = nullptr;
    float* wct_f32 = nullptr;         // classifier transposed [H, K] (per-frame kernel)
    // MROADA anticipation layer (rnn.py:108-110): A = anticipation_length, weight [A*H, H], bias [A*H]; 0 = not loaded
    int ant_len = 0;
    float *wa_f32 = nullptr, *ba = nullptr;
    void* wa_16[2] = {nullptr, nullptr};
    float* online_scratch = nullptr;  // y | LayerNorm partials | logit partials | counters of the per-frame kernel
    uint4* online_stream[2] = {nullptr, nullptr};  // weights in the per-frame kernel's load order ([0] fp16, [1] bf16)
    // 16-bit operands of the tcgen05 path, [0] = fp16, [1] = bf16
    void *w1_16[2] = {nullptr, nullptr}, *wih_16p[2] = {nullptr, nullptr}, *whh_16p[2] = {nullptr, nullptr},
         *wc_16p[2] = {nullptr, nullptr};
    // split-fp16 (PREGO_PREC_F16X3) weights: rows [hi | lo | hi] of scale * W (fp16, [N, 3 K]), gate-interleaved like the 16-bit copies
    __half *w1_x3 = nullptr, *wih_x3 = nullptr, *whh_x3 = nullptr;
    float inv_scale_x3[3] = {1.f, 1.f, 1.f};  // 1 / scale of w1, wih, whh (powers of two)
    unsigned* absmax = nullptr;               // [3] scratch of the scale search
    // latency-kernel exchange
    uint2* xchg = nullptr;
    uint4* xchg_bwd = nullptr;  // [2 groups][2][8][H] exchange words of the persistent BPTT kernels
    int* err_flag = nullptr;
    uint32_t tag_base = 0;
    int64_t coop_fallbacks = 0;  // persistent recurrence launched WITHOUT the cooperative attribute (occupancy-checked)
    // side stream: stages the features of the next time chunk (HBM-bound) under the current chunk's GEMMs
    cudaStream_t side = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    // optional phase profiling (CUDA events on the launching stream)
    bool prof = false;
    int prof_n = 0;
    cudaEvent_t prof_ev[kProfMaxEvents] = {};
    int8_t prof_phase[kProfMaxEvents] = {};
    int64_t prof_launches[PREGO_NUM_PHASES] = {};
};

namespace {

// Record "phase `phase` ended here" (phase < 0: start marker).
inline void prof_mark(prego_model* m, cudaStream_t s, int phase, int launches) {
    if (!m->prof) return;
    if (phase >= 0) m->prof_launches[phase] += launches;
    if (m->prof_n >= kProfMaxEvents) return;
    if (m->prof_ev[m->prof_n] == nullptr) {
        if (cudaEventCreate(&m->prof_ev[m->prof_n]) != cudaSuccess) return;
    }
    cudaEventRecord(m->prof_ev[m->prof_n], s);
    m->prof_phase[m->prof_n] = static_cast<int8_t>(phase);
    ++m->prof_n;
}

struct Plan {
    int64_t h32_a, h32_b;  // [B, H] fp32 state ping-pong
    int64_t xb;            // 16-bit [Mc, Din]                      (16-bit modes)
    int64_t xb2;           // second staging buffer: features of chunk c+1 are staged while chunk c computes (0 = none)
    int64_t ye;            // fp16 y -> 16-bit e in place, or fp32  [Mc, E]
    int64_t gi;            // [Mc, 3H] fp16 (batched 16-bit recurrence) or fp32
    int64_t hseq;          // 16-bit [2, B, H]: operand copies of h_{t-1} / h_t, ping-pong (batched 16-bit recurrence)
    int64_t hrelu;         // 16-bit or fp32 [Mc, H]
    int64_t gh;            // fp32 [B, 3H]                          (fp32 batched recurrence)
    int64_t logits;        // fp32 [Mc, K]                          (fp32 head)
    int64_t sync;          // uint32 [Tc, ceil(B/256)] dependency counters  (batched 16-bit recurrence)
    int64_t online;        // fp32 scratch of the per-frame online path: 8 x (E + 3H + H)
    int64_t h32t;          // fp32 state in the recurrence's tiled order, rows padded to 128 (batched 16-bit recurrence)
    int64_t x3;            // fp16 [Mc, 3 max(Din, E)] | [B, 3 H]: split operands of the PREGO_PREC_F16X3 GEMMs
    int64_t lnstat;        // float2 [E / 256][Mc] partial (sum, sum of squares) + float2 [Mc] (rstd, -mean rstd): fused-LayerNorm path
    int64_t total;
};

bool use_ln_fused();

Plan make_plan(const prego_dims_t& d, int64_t B, int64_t Tc, int prec, bool double_xb = true) {
    Plan p{};
    const int64_t Mc = B * Tc, H = d.hidden_dim, E = d.embed_dim, Din = d.d_rgb + d.d_flow, K = d.num_classes;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        const int64_t o = off;
        off += align_up(bytes, 1024);
        return o;
    };
    p.h32_a = take(B * H * 4);
    p.h32_b = take(B * H * 4);
    const bool h16 = prec != PREGO_PREC_FP32 && prec != PREGO_PREC_F16X3;
    const bool batched = B > kLatencyMaxB;
    p.xb = h16 ? take(Mc * Din * 2) : 0;
    p.xb2 = (h16 && double_xb && B > kLatencyMaxB) ? take(Mc * Din * 2) : 0;
    p.ye = take(Mc * E * (h16 ? 2 : 4));
    p.gi = take(Mc * 3 * H * ((h16 && batched) ? 2 : 4));
    p.hseq = (h16 && batched) ? take(B * 2 * H * 2) : 0;
    p.hrelu = take(Mc * H * (h16 ? 2 : 4));
    p.gh = (!h16 && batched) ? take(B * 3 * H * 4) : 0;
    p.logits = h16 ? 0 : take(Mc * K * 4);
    p.online = h16 ? take(kOnlineMaxRows * (E + 3 * H + H) * 4) : 0;
    p.sync = (h16 && batched) ? take(Tc * ((B + 255) / 256) * 4) : 0;
    // a CTA pair owns 256 rows and reads / writes the state of all of them without a mask: pad to the PAIR tile (padding to 128 let
    // the second CTA of the last pair run past the buffer whenever B % 256 is in (0, 128]; found by compute-sanitizer in round 2)
    p.h32t = (h16 && batched) ? take((B + 255) / 256 * 256 * H * 4) : 0;
    p.lnstat = (h16 && use_ln_fused()) ? take(Mc * (E / 256 + 1) * 8) : 0;
    p.x3 = prec == PREGO_PREC_F16X3 ? take(std::max(Mc * 3 * std::max(Din, E), B * 3 * H) * 2) : 0;
    p.total = off;
    return p;
}

template <int TILE_N, int STAGES, int FMT, class Epi>
int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int a_c1, const Epi& epi,
                   int sm_count, cudaStream_t stream, const char* name) {
    using Cfg = GemmCfg<TILE_N>;
    auto kfn = gemm_tc_kernel<TILE_N, STAGES, FMT, Epi>;
    const int smem = Cfg::smem_bytes(STAGES);
    RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kfn), smem));
    const int tiles = (N / TILE_N) * ((M + kTileM - 1) / kTileM);
    const int grid = tiles < sm_count ? tiles : sm_count;
    kfn<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, M, N, K, a_c1, epi);
    LAUNCH_CHECK(name);
    return PREGO_OK;
}

// CTA-pair (cta_group::2) launch of the same GEMM contract; tmB must be encoded with box rows TILE_N / 2.
template <int TILE_N, int STAGES, int FMT, class Epi>
int launch_gemm_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int a_c1, const Epi& epi,
                    int sm_count, cudaStream_t stream, const char* name, const CUtensorMap* tmA2 = nullptr, int k_split = 0,
                    int rows_per_c1 = 0) {
    using Cfg = Gemm2Cfg<TILE_N>;
    auto kfn = gemm_tc2_kernel<TILE_N, STAGES, FMT, Epi>;
    const int smem = Cfg::smem_bytes(STAGES);
    RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kfn), smem));
    const int tiles = (N / TILE_N) * ((M + 2 * kTileM - 1) / (2 * kTileM));
    int grid = 2 * tiles < sm_count ? 2 * tiles : (sm_count & ~1);
    {   // experiment knob: cap the number of CTA pairs (e.g. a multiple of N / TILE_N so tiles sharing an A block stay in one wave)
        static int cap = -1;
        if (cap < 0) {
            const char* e = getenv("PREGO_GEMM_PAIRS");
            cap = e != nullptr ? atoi(e) : 0;
        }
        if (cap > 0 && 2 * cap < grid) grid = 2 * cap;
    }
    if (tmA2 == nullptr) kfn<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, M, N, K, a_c1, epi, tmA, K, 0);
    else kfn<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, M, N, K, a_c1, epi, *tmA2, k_split, rows_per_c1);
    LAUNCH_CHECK(name);
    return PREGO_OK;
}


// CTA-pair GEMM with the in-place A transform (gemm_xf.cuh); tmB encoded with box rows TILE_N / 2.
template <int TILE_N, int STAGES, int FMT, class Epi, bool IDENTITY>
int launch_gemm_xf(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epi& epi, const LnXf& xf, int sm_count,
                   cudaStream_t stream, const char* name) {
    using Cfg = GemmXfCfg<TILE_N, STAGES>;
    auto kfn = gemm_tc2_xf_kernel<TILE_N, STAGES, FMT, Epi, IDENTITY>;
    const int smem = Cfg::smem_bytes(K);
    if (smem > 227 * 1024) return fail(PREGO_ERR_INVALID, "%s: K = %d needs %d bytes of shared memory", name, K, smem);
    RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kfn), smem));
    const int tiles = (N / TILE_N) * ((M + 2 * kTileM - 1) / (2 * kTileM));
    const int grid = 2 * tiles < sm_count ? 2 * tiles : (sm_count & ~1);
    kfn<<<grid, kXfThreads, smem, stream>>>(tmA, tmB, M, N, K, epi, xf);
    LAUNCH_CHECK(name);
    return PREGO_OK;
}

// C[M, N] (fp32, ldc) (+)= A[M, K] (lda) * W[N, K]^T (ldw) + bias, TF32 operands on CTA pairs.  N % 256 == 0, K % 32 == 0.
// tile_n = 64: narrow tiles for the per-time-step products of the training recurrence (M = B rows only: 256-wide tiles
// would leave most of the GPU idle).
int gemm_tf32_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M,
                 int N, int K, int accumulate, int sm_count, cudaStream_t s, int tile_n = 256) {
    if (N % 256 != 0 || K % 32 != 0 || lda % 4 != 0 || ldw % 4 != 0) return fail(PREGO_ERR_INVALID, "gemm_tf32_nt: need N %% 256 == 0, K %% 32 == 0 (got N=%d K=%d)", N, K);
    CUtensorMap tmA, tmB;
    RC_TRY(make_tmap_a32(&tmA, A, K, M, lda));
    if (tile_n == 64) {
        RC_TRY(make_tmap_w32(&tmB, W, K, N, ldw, 32));
        EpiStore<64, -1> epi{C, bias, ldc, 0, 0, accumulate};
        return launch_gemm_tc2<64, 8, 2>(tmA, tmB, M, N, K, 0, epi, sm_count, s, "gemm_tf32 n64 (2cta)");
    }
    if (tile_n != 256) return fail(PREGO_ERR_INVALID, "gemm_tf32_nt: tile_n must be 256 or 64 (got %d)", tile_n);
    RC_TRY(make_tmap_w32(&tmB, W, K, N, ldw, 128));
    EpiStore<256, -1> epi{C, bias, ldc, 0, 0, accumulate};
    return launch_gemm_tc2<256, 6, 2>(tmA, tmB, M, N, K, 0, epi, sm_count, s, "gemm_tf32 (2cta)");
}

bool use_persistent_gru() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PREGO_GRU_PERSISTENT");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// PREGO_STAGE_OVERLAP: 0 = stage every chunk on the main stream; 1 (default) = stage chunk c + 1 on the side stream
// under chunk c's GEMM1; 2 = under chunk c's recurrence (tensor/epilogue-bound with HBM to spare).
int overlap_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PREGO_STAGE_OVERLAP");
        v = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return v;
}
bool use_overlap() { return overlap_mode() != 0; }

// PREGO_LN_FUSED=1: LayerNorm + ReLU folded into the input-gate GEMM's A operand (gemm_xf.cuh) instead of the separate
// layernorm_relu_16 pass.  OFF by default: measured slower (profiles/r02_ln_fusion.txt); kept selectable so the comparison can be re-run.
bool use_ln_fused() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PREGO_LN_FUSED");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

bool use_2cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PREGO_GEMM_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

bool use_online_fused() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PREGO_ONLINE_FUSED");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// The one-launch-per-frame kernel covers the shipped shapes (H 1024, E 2048, D 2048 | 4096) on a full-size part.
bool online_fused_ok(const prego_model* m) {
    return use_online_fused() && m->d.hidden_dim == 1024 && m->d.embed_dim == 2048 && (m->din == 2048 || m->din == 4096) &&
           m->d.d_rgb % 64 == 0 && m->sm_count >= 128 && m->sm_count <= kFusedPartStride;
}

void* online_fused_fn(int fmt, int din) {
    if (fmt == 0) return din == 4096 ? (void*)online_fused_kernel<0, 4, 1> : (void*)online_fused_kernel<0, 2, 1>;
    return din == 4096 ? (void*)online_fused_kernel<1, 4, 1> : (void*)online_fused_kernel<1, 2, 1>;
}

int online_fused_prepare(void* fn, size_t smem) {
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return PREGO_OK;
}

int check_model(const prego_model* m, bool need_weights) {
    if (m == nullptr) return fail(PREGO_ERR_INVALID, "model handle is NULL");
    if (need_weights && !m->loaded) return fail(PREGO_ERR_STATE, "weights not loaded: call prego_model_load_weights first");
    return PREGO_OK;
}

template <int NB, bool REGW, int G = 1>
int launch_latency_impl(const GruLatencyArgs& a, int H, cudaStream_t stream) {
    auto kfn = gru_latency_kernel<NB, REGW, G>;
    const size_t smem = ((REGW ? 0 : 3 * kLatUnitsPerCta * H) + G * 2 * NB * H) * sizeof(float);
    RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kfn), (int)smem));
    GruLatencyArgs args = a;
    void* params[] = {&args};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kfn, dim3(H / kLatUnitsPerCta), dim3(kLatThreads), params, smem, stream));
    return PREGO_OK;
}

template <int NB>
int launch_latency(const GruLatencyArgs& a, int H, cudaStream_t stream) {
    return H == 1024 ? launch_latency_impl<NB, true>(a, H, stream) : launch_latency_impl<NB, false>(a, H, stream);
}

// Few-stream recurrence on the persistent SIMT kernel (passes of <= 4 streams).
int run_latency_recurrence(prego_model* m, const float* gi, float*& h_cur, float*& h_alt, void* hrelu, int64_t B, int tc,
                           int out_fmt, int64_t row_sb, int64_t row_st, cudaStream_t s) {
    const int H = m->d.hidden_dim;
    for (int b0 = 0; b0 < B; b0 += 4) {
        const int nb = (int)(B - b0 < 4 ? B - b0 : 4);
        GruLatencyArgs la{m->whh_f32p, m->bhh_p, gi, h_cur, h_alt, hrelu, m->xchg, m->err_flag, H, tc, b0, nb, m->tag_base,
                          out_fmt, row_sb, row_st};
        m->tag_base += static_cast<uint32_t>(tc);
        if (nb == 1) RC_TRY(launch_latency<1>(la, H, s));
        else if (nb == 2) RC_TRY(launch_latency<2>(la, H, s));
        else RC_TRY(launch_latency<4>(la, H, s));
    }
    float* tmp = h_cur;
    h_cur = h_alt;
    h_alt = tmp;
    return PREGO_OK;
}

// Width of the projection input actually read: the flow half is skipped when the caller declares it all-zero.
inline int eff_dflow(const prego_model* m, const prego_forward_args_t* a) { return a->flow_is_zero ? 0 : m->d.d_flow; }

// Feature staging of one time chunk of the 16-bit tensor-core path: time-major operand rows [Mc, d_rgb + eff_dflow].
template <int FMT>
int stage_chunk(prego_model* m, const prego_forward_args_t* a, void* xb, int64_t t0, int tc, int blocks_per_sm, cudaStream_t s) {
    using OpT = typename Op16<FMT>::T;
    const int64_t Mc = a->B * tc;
    const int Df = eff_dflow(m, a), D = m->d.d_rgb + Df;
    int grid = grid_for(Mc * (D / 8), 256, m->sm_count);
    if (blocks_per_sm > 0 && grid > m->sm_count * blocks_per_sm) grid = m->sm_count * blocks_per_sm;
    if (a->feature_dtype == PREGO_FEAT_16)
        stage_features_16from16<<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(a->rgb), static_cast<const uint16_t*>(a->flow),
                                                      static_cast<uint16_t*>(xb), Mc, m->d.d_rgb, Df, (int)a->B, (int)a->T, (int)t0);
    else
        stage_features_16<FMT><<<grid, 256, 0, s>>>(static_cast<const float*>(a->rgb), static_cast<const float*>(a->flow),
                                                    reinterpret_cast<OpT*>(xb), Mc, m->d.d_rgb, Df, (int)a->B, (int)a->T, (int)t0);
    LAUNCH_CHECK("stage_features_16");
    return PREGO_OK;
}

// 16-bit features of a [B, T, D] tensor read in place by the projection GEMM: (k, t, b) view, box (64, 1, 128).
int make_tmap_feat(CUtensorMap* tm, DType dt, const void* base, uint64_t D, uint64_t T, uint64_t B) {
    const uint64_t dims[3] = {D, T, B}, str[2] = {D * 2, T * D * 2};
    const uint32_t box[3] = {kTileK, 1, kTileM};
    return make_tmap(tm, dt, 3, base, dims, str, box);
}

// ci = chunk index; with overlap the features of chunk ci were staged into xb[ci & 1] ahead of time on the side
// stream, and this call stages chunk ci + 1 (starting at t_next, tc_next frames) behind its own GEMM1.
// ---- MROADA anticipation head on one time chunk (rnn.py:125-126,133): ant = relu(relu(h) Wa^T + ba) viewed as
// [rows * A, H], then the SAME classifier + softmax / argmax.  Produced in row slabs that fit the caller's buffer.
inline int64_t ant_row_bytes(const prego_model* m, int prec) {
    const int64_t A = m->ant_len, H = m->d.hidden_dim, K = m->d.num_classes;
    return (prec == PREGO_PREC_FP32 || prec == PREGO_PREC_F16X3) ? A * H * 4 + A * K * 4 : A * H * 2;
}
inline int64_t ant_slab_rows(const prego_model* m, const prego_anticipation_args_t* ant, int prec, int64_t Mc) {
    int64_t rows = static_cast<int64_t>(ant->workspace_bytes) / ant_row_bytes(m, prec);
    const int64_t cap = (int64_t(1) << 22) / m->ant_len;  // slab * A rows per head launch (grid.y / int limits)
    if (rows > cap) rows = cap;
    if (rows >= 128) rows = rows / 128 * 128;
    return rows < Mc ? rows : Mc;
}

template <int FMT>
int anticipation_16(prego_model* m, const prego_forward_args_t* a, const prego_anticipation_args_t* ant, const void* hrelu,
                    int64_t Mc, int64_t t0, cudaStream_t s) {
    using OpT = typename Op16<FMT>::T;
    const DType dt = FMT == 0 ? kF16 : kBF16;
    const int H = m->d.hidden_dim, K = m->d.num_classes, A = m->ant_len;
    const int64_t slab = ant_slab_rows(m, ant, a->precision, Mc);
    OpT* ant16 = static_cast<OpT*>(ant->workspace);
    CUtensorMap tmA, tmB, tmH, tmC;
    RC_TRY(make_tmap_w(&tmB, dt, m->wa_16[FMT], H, (uint64_t)A * H, use_2cta() ? 128 : 256));
    RC_TRY(make_tmap_w(&tmC, dt, m->wc_16p[FMT], H, m->kpad, m->kpad));
    for (int64_t r0 = 0; r0 < Mc; r0 += slab) {
        const int ms = static_cast<int>(Mc - r0 < slab ? Mc - r0 : slab);
        RC_TRY(make_tmap_a(&tmA, dt, static_cast<const OpT*>(hrelu) + r0 * H, H, ms));
        EpiStore<256, FMT> epi{ant16, m->ba, (int64_t)A * H, 0, 0, 0, 1};
        if (use_2cta()) RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, ms, A * H, H, 0, epi, m->sm_count, s, "anticipation layer (2cta)")));
        else RC_TRY((launch_gemm_tc<256, 4, FMT>(tmA, tmB, ms, A * H, H, 0, epi, m->sm_count, s, "anticipation layer")));
        RC_TRY(make_tmap_a(&tmH, dt, ant16, H, (uint64_t)ms * A));
        if (m->kpad == 96)
            RC_TRY((launch_gemm_tc<96, 6, FMT>(tmH, tmC, ms * A, 96, H, 0, EpiHead<96>{m->bc, ant->probs, ant->logits, ant->labels, K, (int)a->B, (int)a->T, (int)t0, A, r0 * A}, m->sm_count, s, "anticipation head96")));
        else
            RC_TRY((launch_gemm_tc<128, 6, FMT>(tmH, tmC, ms * A, 128, H, 0, EpiHead<128>{m->bc, ant->probs, ant->logits, ant->labels, K, (int)a->B, (int)a->T, (int)t0, A, r0 * A}, m->sm_count, s, "anticipation head128")));
    }
    return PREGO_OK;
}

// exact fp32: hr32 rows are stream-major inside the chunk (m = b*tc + t)
int anticipation_f32(prego_model* m, const prego_forward_args_t* a, const prego_anticipation_args_t* ant, const float* hr32,
                     int64_t Mc, int tc, int64_t t0, cudaStream_t s) {
    const int H = m->d.hidden_dim, K = m->d.num_classes, A = m->ant_len;
    const int64_t slab = ant_slab_rows(m, ant, a->precision, Mc);
    float* ant32 = static_cast<float*>(ant->workspace);
    float* lg = ant32 + slab * A * H;
    for (int64_t r0 = 0; r0 < Mc; r0 += slab) {
        const int ms = static_cast<int>(Mc - r0 < slab ? Mc - r0 : slab);
        SgemmA Aa{hr32 + r0 * H, nullptr, H, H, 0, 0, tc, (int)a->T, (int)t0, 0, 1};
        sgemm_nt_f32<<<dim3((A * H + 127) / 128, (ms + 127) / 128), 256, 0, s>>>(Aa, m->wa_f32, m->ba, ant32, ms, A * H, H, (int64_t)A * H);
        SgemmA Ab{ant32, nullptr, H, H, 0, 0, tc, (int)a->T, (int)t0};
        sgemm_nt_f32<<<dim3((K + 127) / 128, (ms * A + 127) / 128), 256, 0, s>>>(Ab, m->wc_f32, m->bc, lg, ms * A, K, H, K);
        softmax_argmax_f32<<<grid_for((int64_t)ms * A * 32, 256, m->sm_count), 256, 0, s>>>(lg, ant->probs, ant->logits, ant->labels, (int64_t)ms * A, K, tc,
                                                                                          (int)a->T, (int)t0, A, r0 * A);
        LAUNCH_CHECK("fp32 anticipation head");
    }
    return PREGO_OK;
}

template <int FMT>
int chunk_16(prego_model* m, const prego_forward_args_t* a, const Plan& p, uint8_t* ws, float*& h_cur, float*& h_alt,
             int64_t t0, int tc, cudaStream_t s, int ci = 0, bool overlap = false, int64_t t_next = 0, int tc_next = 0,
             const prego_anticipation_args_t* ant = nullptr) {
    using OpT = typename Op16<FMT>::T;
    const DType dt = FMT == 0 ? kF16 : kBF16;
    const prego_dims_t& d = m->d;
    const int64_t B = a->B, T = a->T, Mc = B * tc;
    const int Mi = static_cast<int>(Mc);
    const int H = d.hidden_dim, E = d.embed_dim, K = d.num_classes, Din = m->din;
    const int Dfe = eff_dflow(m, a), Dine = d.d_rgb + Dfe;  // projection columns actually read
    const bool batched = B > kLatencyMaxB;
    // 16-bit features + whole 128-row tiles per time step: GEMM1's TMA gathers the time-major rows from the caller's
    // tensors, no staging pass and no xb traffic
    const bool direct = a->feature_dtype == PREGO_FEAT_16 && B % kTileM == 0 && use_2cta();
    OpT* xb = reinterpret_cast<OpT*>(ws + ((overlap && (ci & 1)) ? p.xb2 : p.xb));
    void* ye = ws + p.ye;
    void* gi = ws + p.gi;
    OpT* hseq = reinterpret_cast<OpT*>(ws + p.hseq);
    OpT* hrelu = reinterpret_cast<OpT*>(ws + p.hrelu);
    CUtensorMap tmA, tmB;
    const bool fused_ln = use_ln_fused() && use_2cta() && p.lnstat != 0 && E == 2048;
    float2* ln_part = reinterpret_cast<float2*>(ws + p.lnstat);
    float2* ln_row = ln_part + static_cast<int64_t>(E / 256) * Mc;

    // 1. stage features: concat + operand rounding (replaces torch.cat, rnn.py:53)
    if (direct) {
    } else if (!overlap) {
        RC_TRY(stage_chunk<FMT>(m, a, xb, t0, tc, 0, s));
    } else {
        CUDA_TRY(cudaStreamWaitEvent(s, m->ev_ready[ci & 1], 0));  // staged on the side stream (or by the prologue)
    }
    prof_mark(m, s, PREGO_PHASE_STAGE, 1);
    // 2. y = x W1^T + b1  (fp16 out; rows stay time-major, m = t*B + b)
    if (direct) {
        CUtensorMap tmA2;
        const bool both = d.d_rgb > 0 && Dfe > 0;
        if (d.d_rgb > 0) RC_TRY(make_tmap_feat(&tmA, dt, a->rgb, d.d_rgb, T, B));
        else RC_TRY(make_tmap_feat(&tmA, dt, a->flow, d.d_flow, T, B));
        if (both) RC_TRY(make_tmap_feat(&tmA2, dt, a->flow, d.d_flow, T, B));
        RC_TRY(make_tmap_w(&tmB, dt, m->w1_16[FMT], Din, E, 128));
        if (fused_ln)
            RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, E, Dine, (int)t0, EpiStoreStats<256>{reinterpret_cast<__half*>(ye), m->b1, E, ln_part, Mi}, m->sm_count, s,
                                                 "gemm1 (2cta, features in place, LN statistics)", both ? &tmA2 : &tmA, both ? d.d_rgb : Dine, (int)B)));
        else
        RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, E, Dine, (int)t0, EpiStore<256, 0>{ye, m->b1, E, 0, 0}, m->sm_count, s,
                                             "gemm1 (2cta, features in place)", both ? &tmA2 : &tmA, both ? d.d_rgb : Dine, (int)B)));
    } else if (use_2cta()) {
        RC_TRY(make_tmap_a(&tmA, dt, xb, Dine, Mc));
        RC_TRY(make_tmap_w(&tmB, dt, m->w1_16[FMT], Din, E, 128));
        if (fused_ln) RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, E, Dine, 0, EpiStoreStats<256>{reinterpret_cast<__half*>(ye), m->b1, E, ln_part, Mi}, m->sm_count, s, "gemm1 (2cta, LN statistics)")));
        else RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, E, Dine, 0, EpiStore<256, 0>{ye, m->b1, E, 0, 0}, m->sm_count, s, "gemm1 (2cta)")));
    } else {
        RC_TRY(make_tmap_a(&tmA, dt, xb, Dine, Mc));
        RC_TRY(make_tmap_w(&tmB, dt, m->w1_16[FMT], Din, E, 256));
        RC_TRY((launch_gemm_tc<256, 4, FMT>(tmA, tmB, Mi, E, Dine, 0, EpiStore<256, 0>{ye, m->b1, E, 0, 0}, m->sm_count, s, "gemm1")));
    }
    auto stage_next = [&]() -> int {
        // next chunk's features -> the other buffer, on the side stream
        CUDA_TRY(cudaStreamWaitEvent(m->side, m->ev_free[(ci + 1) & 1], 0));
        RC_TRY(stage_chunk<FMT>(m, a, ws + (((ci + 1) & 1) ? p.xb2 : p.xb), t_next, tc_next, 4, m->side));
        CUDA_TRY(cudaEventRecord(m->ev_ready[(ci + 1) & 1], m->side));
        return PREGO_OK;
    };
    if (overlap) {
        CUDA_TRY(cudaEventRecord(m->ev_free[ci & 1], s));  // GEMM1 has consumed this staging buffer
        if (tc_next > 0 && overlap_mode() == 1) {
            // issued AFTER GEMM1 so the GEMM's CTAs are placed first; a capped grid leaves room for them to co-reside
            // (HBM-bound next to tensor-bound work)
            RC_TRY(stage_next());
        }
    }
    prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
    // 3. e = relu(LN(y))  (in place, operand format) -- or, fused: only the row statistics are finalised here and y is
    //    normalised inside the input-gate GEMM's A operand
    if (fused_ln) {
        ln_finalize_kernel<<<(unsigned)((Mc + 255) / 256), 256, 0, s>>>(ln_part, ln_row, Mi, E / 256, E, 1e-5f);
        LAUNCH_CHECK("ln_finalize_kernel");
    } else {
        layernorm_relu_16<2048, FMT><<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(
            reinterpret_cast<const __half*>(ye), reinterpret_cast<OpT*>(ye), m->ln_g, m->ln_b, Mc, 1e-5f);
        LAUNCH_CHECK("layernorm_relu_16");
    }
    prof_mark(m, s, PREGO_PHASE_LAYERNORM, 1);
    // 4. gi = e W_ih'^T + b_ih'  (gate-interleaved columns, time-major rows)
    RC_TRY(make_tmap_a(&tmA, fused_ln ? kF16 : dt, ye, E, Mc));
    if (fused_ln) {
        RC_TRY(make_tmap_w(&tmB, dt, m->wih_16p[FMT], E, 3 * H, 128));
        const LnXf xf{ln_row, m->ln_g, m->ln_b};
        if (batched)
            RC_TRY((launch_gemm_xf<256, 6, FMT, EpiStore<256, 0>, false>(tmA, tmB, Mi, 3 * H, E, EpiStore<256, 0>{gi, m->bgi_p, 3 * H, 0, 0}, xf, m->sm_count, s, "gemm2 (LayerNorm fused)")));
        else
            RC_TRY((launch_gemm_xf<256, 6, FMT, EpiStore<256, -1>, false>(tmA, tmB, Mi, 3 * H, E, EpiStore<256, -1>{gi, m->bih_p, 3 * H, 0, 0}, xf, m->sm_count, s, "gemm2 (LayerNorm fused)")));
    } else if (use_2cta()) {
        RC_TRY(make_tmap_w(&tmB, dt, m->wih_16p[FMT], E, 3 * H, 128));
        if (batched)
            RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, 3 * H, E, 0, EpiStore<256, 0>{gi, m->bgi_p, 3 * H, 0, 0}, m->sm_count, s, "gemm2 (2cta)")));
        else
            RC_TRY((launch_gemm_tc2<256, 6, FMT>(tmA, tmB, Mi, 3 * H, E, 0, EpiStore<256, -1>{gi, m->bih_p, 3 * H, 0, 0}, m->sm_count, s, "gemm2 (2cta)")));
    } else {
        RC_TRY(make_tmap_w(&tmB, dt, m->wih_16p[FMT], E, 3 * H, 192));
        if (batched)
            RC_TRY((launch_gemm_tc<192, 5, FMT>(tmA, tmB, Mi, 3 * H, E, 0, EpiStore<192, 0>{gi, m->bgi_p, 3 * H, 0, 0}, m->sm_count, s, "gemm2")));
        else
            RC_TRY((launch_gemm_tc<192, 5, FMT>(tmA, tmB, Mi, 3 * H, E, 0, EpiStore<192, -1>{gi, m->bih_p, 3 * H, 0, 0}, m->sm_count, s, "gemm2")));
    }
    prof_mark(m, s, PREGO_PHASE_GEMM2, 1);
    if (overlap && tc_next > 0 && overlap_mode() == 2) {
        CUDA_TRY(cudaEventRecord(m->ev_begin, s));  // GEMM2 done -> the side stream may start staging under the recurrence
        CUDA_TRY(cudaStreamWaitEvent(m->side, m->ev_begin, 0));
        RC_TRY(stage_next());
    }

    // 5. recurrence
    if (batched) {
        f32_to_16<FMT><<<grid_for(B * H, 256, m->sm_count), 256, 0, s>>>(h_cur, hseq, B * H);  // slot 0 = carried state
        float* h32t = reinterpret_cast<float*>(ws + p.h32t);
        h32_retile<<<grid_for((B + 127) / 128 * 128 * (H / 4), 256, m->sm_count), 256, 0, s>>>(h_cur, h32t, (int)B, H, 1);
        LAUNCH_CHECK("f32_to_16 (hseq slot 0) / h32_retile");
        CUtensorMap tmHseq, tmW, tmGi, tmHrelu;
        RC_TRY(make_tmap_tm(&tmHseq, dt, hseq, H, B, 2, 2));
        RC_TRY(make_tmap_w(&tmW, dt, m->whh_16p[FMT], H, 3 * H, kGruTileN / 2));
        RC_TRY(make_tmap_tm(&tmGi, kF16, gi, 3 * H, B, tc, 2));
        RC_TRY(make_tmap_tm(&tmHrelu, dt, hrelu, H, B, tc, 2));
        // PREGO_GRU_DBG / PREGO_GRU_STATS (diagnostics, scripts/diag_recurrence.py) select the instrumented instantiation
        const bool diag = getenv("PREGO_GRU_DBG") != nullptr || getenv("PREGO_GRU_STATS") != nullptr;
        auto kfn = diag ? gru_seq_kernel<FMT, true> : gru_seq_kernel<FMT, false>;
        RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kfn), kGruSmemBytes));
        const int m_tiles = (int)((B + 2 * kTileM - 1) / (2 * kTileM));
        const int per_step = (3 * H / kGruTileN) * m_tiles;  // CTA-pair tiles per time step
        const int max_grid = m->sm_count & ~1;
        int launches = 0;
        if (use_persistent_gru()) {
            // one launch for the whole chunk: dataflow dependencies between steps (done counters)
            uint32_t* done = reinterpret_cast<uint32_t*>(ws + p.sync);
            CUDA_TRY(cudaMemsetAsync(done, 0, (size_t)tc * m_tiles * 4, s));
            const int64_t items = (int64_t)per_step * tc;
            const int grid = 2 * items < max_grid ? (int)(2 * items) : max_grid;
            GruSeqArgs ga{m->bhh_p, h32t, done, m->err_flag, (int)B, H, 0, tc};
            if (const char* e = getenv("PREGO_GRU_DBG")) ga.dbg = atoi(e);  // diagnostics: results are garbage with any bit set
            if (const char* e = getenv("PREGO_GRU_STATS")) ga.stats = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));  // device pointer [grid][16]
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(kGruThreads);
            cfg.dynamicSmemBytes = kGruSmemBytes;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;  // all CTA pairs co-resident (the dependency spins rely on it)
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t le = cudaLaunchKernelEx(&cfg, kfn, tmHseq, tmW, tmGi, tmHrelu, ga);
            if (le != cudaSuccess) {
                // Cooperative attribute refused (cluster launches on some drivers).  The dependency spins need every pair
                // resident: fall back to a plain launch ONLY when the occupancy calculator says the whole grid fits at
                // once, otherwise fail loudly -- never spin on CTAs that may not be scheduled.
                (void)cudaGetLastError();
                cfg.numAttrs = 0;
                int max_clusters = 0;
                cudaLaunchAttribute cattr[1];
                cattr[0].id = cudaLaunchAttributeClusterDimension;
                cattr[0].val.clusterDim.x = 2; cattr[0].val.clusterDim.y = 1; cattr[0].val.clusterDim.z = 1;
                cudaLaunchConfig_t occ = cfg;
                occ.attrs = cattr;
                occ.numAttrs = 1;
                CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &occ));
                if (2 * max_clusters < grid)
                    return fail(PREGO_ERR_CUDA, "persistent recurrence: cooperative launch refused (%s) and only %d of %d CTA pairs can be co-resident",
                                cudaGetErrorString(le), max_clusters, grid / 2);
                CUDA_TRY(cudaLaunchKernelEx(&cfg, kfn, tmHseq, tmW, tmGi, tmHrelu, ga));
                m->coop_fallbacks += 1;
            }
            launches = 2;
        } else {
            const int grid = 2 * per_step < max_grid ? 2 * per_step : max_grid;
            for (int t = 0; t < tc; ++t) {
                GruSeqArgs ga{m->bhh_p, h32t, nullptr, m->err_flag, (int)B, H, t, t + 1};
                kfn<<<grid, kGruThreads, kGruSmemBytes, s>>>(tmHseq, tmW, tmGi, tmHrelu, ga);
            }
            launches = tc + 1;
        }
        h32_retile<<<grid_for((B + 127) / 128 * 128 * (H / 4), 256, m->sm_count), 256, 0, s>>>(h_cur, h32t, (int)B, H, 0);
        LAUNCH_CHECK("gru_seq_kernel");
        launches += 2;
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, launches);
    } else {
        RC_TRY(run_latency_recurrence(m, reinterpret_cast<const float*>(gi), h_cur, h_alt, hrelu, B, tc, FMT, 1, B, s));
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, (int)((B + 3) / 4));
    }

    // 6. head: logits, softmax, argmax (rows are time-major)
    RC_TRY(make_tmap_a(&tmA, dt, hrelu, H, Mc));
    RC_TRY(make_tmap_w(&tmB, dt, m->wc_16p[FMT], H, m->kpad, m->kpad));
    if (m->kpad == 96)
        RC_TRY((launch_gemm_tc<96, 6, FMT>(tmA, tmB, Mi, 96, H, 0, EpiHead<96>{m->bc, a->probs, a->logits, a->labels, K, (int)B, (int)T, (int)t0}, m->sm_count, s, "head96")));
    else
        RC_TRY((launch_gemm_tc<128, 6, FMT>(tmA, tmB, Mi, 128, H, 0, EpiHead<128>{m->bc, a->probs, a->logits, a->labels, K, (int)B, (int)T, (int)t0}, m->sm_count, s, "head128")));
    if (ant != nullptr) RC_TRY(anticipation_16<FMT>(m, a, ant, hrelu, Mc, t0, s));
    prof_mark(m, s, PREGO_PHASE_HEAD, 1);
    return PREGO_OK;
}

// Strict per-frame online step (T == 1, B <= 8) on the GEMV kernels of online_kernels.cuh.
template <int FMT, int R>
int online_step_r(prego_model* m, const prego_forward_args_t* a, const Plan& p, uint8_t* ws, float*& h_cur, float*& h_alt,
                  cudaStream_t s) {
    using OpT = typename Op16<FMT>::T;
    const prego_dims_t& d = m->d;
    const int rows = (int)a->B, H = d.hidden_dim, E = d.embed_dim, K = d.num_classes, Din = m->din;
    float* y = reinterpret_cast<float*>(ws + p.online);
    float* gi = y + kOnlineMaxRows * E;
    float* hrelu = gi + kOnlineMaxRows * 3 * H;
    const int smem1 = R * Din * 2;
    RC_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(online_proj1<FMT, R>), smem1));  // models may differ in Din
    const int wpb = kOnlineThreads / 32;
    online_proj1<FMT, R><<<(E + wpb - 1) / wpb, kOnlineThreads, smem1, s>>>(static_cast<const float*>(a->rgb), static_cast<const float*>(a->flow), reinterpret_cast<const OpT*>(m->w1_16[FMT]), m->b1, y,
                                                                            rows, d.d_rgb, d.d_flow, E, a->T, 0);
    prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
    online_proj2<FMT, R><<<(3 * H + wpb - 1) / wpb, kOnlineThreads, R * E * 2, s>>>(y, m->ln_g, m->ln_b, reinterpret_cast<const OpT*>(m->wih_16p[FMT]),
                                                                                     m->bih_p, gi, rows, E, 3 * H, 1e-5f);
    prof_mark(m, s, PREGO_PHASE_GEMM2, 1);
    online_gru<FMT, R><<<(H + wpb - 1) / wpb, kOnlineThreads, R * H * 2, s>>>(gi, reinterpret_cast<const OpT*>(m->whh_16p[FMT]), m->bhh_p, h_cur, h_alt,
                                                                               hrelu, rows, H);
    prof_mark(m, s, PREGO_PHASE_RECURRENCE, 1);
    online_head<<<rows, kOnlineThreads, 0, s>>>(hrelu, m->wc_f32, m->bc, a->probs, a->logits, a->labels, H, K, a->T, 0);
    LAUNCH_CHECK("online step");
    prof_mark(m, s, PREGO_PHASE_HEAD, 1);
    float* tmp = h_cur; h_cur = h_alt; h_alt = tmp;
    return PREGO_OK;
}

// Carves one online context's scratch (online_fused_scratch_floats) into the kernel's pointers.
void online_fused_carve(float* scratch, int E, int K, OnlineFusedArgs* fa) {
    fa->y = scratch;
    fa->stats = reinterpret_cast<float2*>(scratch + (size_t)kFusedMaxRows * E);
    fa->gpart = scratch + (size_t)kFusedMaxRows * (E + 2 * kFusedPartStride);
    fa->sync = reinterpret_cast<unsigned*>(fa->gpart + (size_t)kFusedMaxRows * K * kFusedPartStride);
}

// One cooperative launch per frame (online_fused.cuh); the state is updated in place.
int online_step_fused(prego_model* m, const prego_forward_args_t* a, float* h, int fmt, cudaStream_t s) {
    const prego_dims_t& d = m->d;
    OnlineFusedArgs fa{};
    fa.rgb = static_cast<const float*>(a->rgb); fa.flow = static_cast<const float*>(a->flow);
    fa.wstream = m->online_stream[fmt];
    fa.b1 = m->b1; fa.ln_g = m->ln_g; fa.ln_b = m->ln_b; fa.bih = m->bih_p; fa.bhh = m->bhh_p; fa.wct = m->wct_f32; fa.bc = m->bc;
    online_fused_carve(m->online_scratch, d.embed_dim, d.num_classes, &fa);
    fa.h = h; fa.probs = a->probs; fa.logits = a->logits; fa.labels = a->labels; fa.err_flag = m->err_flag; fa.trace = nullptr; fa.host_seq = nullptr; fa.seq = 0; fa.fence_outputs = 0;
    fa.rows = (int)a->B; fa.Dr = d.d_rgb; fa.Df = d.d_flow; fa.E = d.embed_dim; fa.H = d.hidden_dim; fa.K = d.num_classes;
    fa.T = a->T; fa.t0 = 0; fa.eps = 1e-5f;
    void* fn = online_fused_fn(fmt, m->din);
    const size_t smem = online_fused_smem((int)a->B, m->din, d.hidden_dim);
    RC_TRY(online_fused_prepare(fn, online_fused_smem(kFusedMaxRows, m->din, d.hidden_dim)));
    void* params[] = {&fa};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(m->sm_count), dim3(kFusedThreads), params, smem, s));
    prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
    return PREGO_OK;
}

template <int FMT>
int online_step(prego_model* m, const prego_forward_args_t* a, const Plan& p, uint8_t* ws, float*& h_cur, float*& h_alt,
                cudaStream_t s) {
    const int B = (int)a->B;
    if (online_fused_ok(m)) return online_step_fused(m, a, h_cur, FMT, s);
    if (B == 1) return online_step_r<FMT, 1>(m, a, p, ws, h_cur, h_alt, s);
    if (B == 2) return online_step_r<FMT, 2>(m, a, p, ws, h_cur, h_alt, s);
    if (B <= 4) return online_step_r<FMT, 4>(m, a, p, ws, h_cur, h_alt, s);
    return online_step_r<FMT, 8>(m, a, p, ws, h_cur, h_alt, s);
}

// One time chunk of the exact-fp32 CUDA-core path (stream-major rows throughout).
int chunk_f32(prego_model* m, const prego_forward_args_t* a, const Plan& p, uint8_t* ws, float*& h_cur, float*& h_alt,
              int64_t t0, int tc, cudaStream_t s, const prego_anticipation_args_t* ant = nullptr) {
    const prego_dims_t& d = m->d;
    const int64_t B = a->B, T = a->T, Mc = B * tc;
    const int Mi = static_cast<int>(Mc);
    const int H = d.hidden_dim, E = d.embed_dim, K = d.num_classes, Din = m->din;
    float* y32 = reinterpret_cast<float*>(ws + p.ye);
    float* gi = reinterpret_cast<float*>(ws + p.gi);
    float* hr32 = reinterpret_cast<float*>(ws + p.hrelu);
    float* gh = reinterpret_cast<float*>(ws + p.gh);
    float* logits_ws = reinterpret_cast<float*>(ws + p.logits);

    const float *rgb32 = static_cast<const float*>(a->rgb), *flow32 = static_cast<const float*>(a->flow);
    SgemmA A1{rgb32, flow32, d.d_rgb, d.d_rgb, d.d_flow, 1, tc, (int)T, (int)t0};
    int K1 = Din;
    if (d.d_rgb == 0) { A1.a0 = flow32; A1.a1 = nullptr; A1.k_split = Din; A1.lda0 = d.d_flow; }
    if (d.d_flow == 0) { A1.a1 = nullptr; A1.k_split = Din; }
    if (a->flow_is_zero && d.d_flow > 0) { A1.a1 = nullptr; A1.k_split = d.d_rgb; A1.ldw = Din; K1 = d.d_rgb; }  // rgb columns of W1 only
    sgemm_nt_f32<<<dim3(E / 128, (Mi + 127) / 128), 256, 0, s>>>(A1, m->w1_f32, m->b1, y32, Mi, E, K1, E);
    LAUNCH_CHECK("sgemm gemm1");
    prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
    layernorm_relu_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(y32, y32, m->ln_g, m->ln_b, Mc, E, 1e-5f);
    LAUNCH_CHECK("layernorm_relu_f32");
    prof_mark(m, s, PREGO_PHASE_LAYERNORM, 1);
    SgemmA A2{y32, nullptr, E, E, 0, 0, tc, (int)T, (int)t0};
    sgemm_nt_f32<<<dim3(3 * H / 128, (Mi + 127) / 128), 256, 0, s>>>(A2, m->wih_f32p, m->bih_p, gi, Mi, 3 * H, E, 3 * H);
    LAUNCH_CHECK("sgemm gemm2");
    prof_mark(m, s, PREGO_PHASE_GEMM2, 1);

    if (B <= kLatencyMaxB) {
        RC_TRY(run_latency_recurrence(m, gi, h_cur, h_alt, hr32, B, tc, -1, tc, 1, s));
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, (int)((B + 3) / 4));
    } else {
        for (int t = 0; t < tc; ++t) {
            SgemmA Ah{h_cur, nullptr, H, H, 0, 0, tc, (int)T, (int)t0};
            sgemm_nt_f32<<<dim3(3 * H / 128, (int)((B + 127) / 128)), 256, 0, s>>>(Ah, m->whh_f32p, m->bhh_p, gh, (int)B, 3 * H, H, 3 * H);
            gru_gates_f32<<<grid_for(B * H, 256, m->sm_count), 256, 0, s>>>(gi, gh, h_cur, hr32, (int)B, H, tc, t);
        }
        LAUNCH_CHECK("fp32 recurrence");
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, 2 * tc);
    }

    SgemmA Ah{hr32, nullptr, H, H, 0, 0, tc, (int)T, (int)t0};
    sgemm_nt_f32<<<dim3((K + 127) / 128, (Mi + 127) / 128), 256, 0, s>>>(Ah, m->wc_f32, m->bc, logits_ws, Mi, K, H, K);
    softmax_argmax_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(logits_ws, a->probs, a->logits, a->labels, Mc, K, tc, (int)T, (int)t0, 1, 0);
    LAUNCH_CHECK("fp32 head");
    if (ant != nullptr) RC_TRY(anticipation_f32(m, a, ant, hr32, Mc, tc, t0, s));
    prof_mark(m, s, PREGO_PHASE_HEAD, 2);
    return PREGO_OK;
}

// One time chunk of the split-fp16 mode (PREGO_PREC_F16X3): the exact-fp32 path's structure (stream-major rows, fp32 LayerNorm /
// gates / state / softmax) with every large GEMM on tcgen05 over [lo | hi | hi] x [hi | lo | hi] operands, small terms first (simt_kernels.cuh).
int chunk_x3(prego_model* m, const prego_forward_args_t* a, const Plan& p, uint8_t* ws, float*& h_cur, float*& h_alt, int64_t t0, int tc,
             cudaStream_t s, const prego_anticipation_args_t* ant = nullptr) {
    const prego_dims_t& d = m->d;
    const int64_t B = a->B, T = a->T, Mc = B * tc;
    const int Mi = static_cast<int>(Mc);
    const int H = d.hidden_dim, E = d.embed_dim, K = d.num_classes, Din = m->din;
    float* y32 = reinterpret_cast<float*>(ws + p.ye);
    float* gi = reinterpret_cast<float*>(ws + p.gi);
    float* hr32 = reinterpret_cast<float*>(ws + p.hrelu);
    float* gh = reinterpret_cast<float*>(ws + p.gh);
    float* logits_ws = reinterpret_cast<float*>(ws + p.logits);
    __half* x3 = reinterpret_cast<__half*>(ws + p.x3);
    CUtensorMap tmA, tmB;

    stage_features_split3<<<grid_for(Mc * (Din / 4), 256, m->sm_count), 256, 0, s>>>(static_cast<const float*>(a->rgb), a->flow_is_zero ? nullptr : static_cast<const float*>(a->flow), x3, Mc,
                                                                                    d.d_rgb, d.d_flow, tc, (int)T, (int)t0);
    LAUNCH_CHECK("stage_features_split3");
    prof_mark(m, s, PREGO_PHASE_STAGE, 1);
    RC_TRY(make_tmap_a(&tmA, kF16, x3, 3 * (uint64_t)Din, Mc));
    RC_TRY(make_tmap_w(&tmB, kF16, m->w1_x3, 3 * (uint64_t)Din, E, 128));
    RC_TRY((launch_gemm_tc2<256, 6, 0>(tmA, tmB, Mi, E, 3 * Din, 0, EpiStore<256, -1>{y32, m->b1, E, 0, 0, 0, 0, m->inv_scale_x3[0]}, m->sm_count, s, "gemm1 (split fp16)")));
    prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
    layernorm_relu_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(y32, y32, m->ln_g, m->ln_b, Mc, E, 1e-5f);
    split3_rows_f32<<<grid_for(Mc * (E / 4), 256, m->sm_count), 256, 0, s>>>(y32, x3, Mc, E);
    LAUNCH_CHECK("layernorm_relu_f32 / split3_rows_f32");
    prof_mark(m, s, PREGO_PHASE_LAYERNORM, 2);
    RC_TRY(make_tmap_a(&tmA, kF16, x3, 3 * (uint64_t)E, Mc));
    RC_TRY(make_tmap_w(&tmB, kF16, m->wih_x3, 3 * (uint64_t)E, 3 * H, 128));
    RC_TRY((launch_gemm_tc2<256, 6, 0>(tmA, tmB, Mi, 3 * H, 3 * E, 0, EpiStore<256, -1>{gi, m->bih_p, 3 * H, 0, 0, 0, 0, m->inv_scale_x3[1]}, m->sm_count, s, "gemm2 (split fp16)")));
    prof_mark(m, s, PREGO_PHASE_GEMM2, 1);

    if (B <= kLatencyMaxB) {
        RC_TRY(run_latency_recurrence(m, gi, h_cur, h_alt, hr32, B, tc, -1, tc, 1, s));
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, (int)((B + 3) / 4));
    } else {
        RC_TRY(make_tmap_a(&tmA, kF16, x3, 3 * (uint64_t)H, B));
        RC_TRY(make_tmap_w(&tmB, kF16, m->whh_x3, 3 * (uint64_t)H, 3 * H, 128));
        for (int t = 0; t < tc; ++t) {
            split3_rows_f32<<<grid_for(B * (H / 4), 256, m->sm_count), 256, 0, s>>>(h_cur, x3, B, H);
            RC_TRY((launch_gemm_tc2<256, 6, 0>(tmA, tmB, (int)B, 3 * H, 3 * H, 0, EpiStore<256, -1>{gh, m->bhh_p, 3 * H, 0, 0, 0, 0, m->inv_scale_x3[2]}, m->sm_count, s, "recurrent gemm (split fp16)")));
            gru_gates_f32<<<grid_for(B * H, 256, m->sm_count), 256, 0, s>>>(gi, gh, h_cur, hr32, (int)B, H, tc, t);
        }
        LAUNCH_CHECK("split fp16 recurrence");
        prof_mark(m, s, PREGO_PHASE_RECURRENCE, 3 * tc);
    }

    SgemmA Ah{hr32, nullptr, H, H, 0, 0, tc, (int)T, (int)t0};
    sgemm_nt_f32<<<dim3((K + 127) / 128, (Mi + 127) / 128), 256, 0, s>>>(Ah, m->wc_f32, m->bc, logits_ws, Mi, K, H, K);
    softmax_argmax_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(logits_ws, a->probs, a->logits, a->labels, Mc, K, tc, (int)T, (int)t0, 1, 0);
    LAUNCH_CHECK("fp32 head");
    if (ant != nullptr) RC_TRY(anticipation_f32(m, a, ant, hr32, Mc, tc, t0, s));
    prof_mark(m, s, PREGO_PHASE_HEAD, 2);
    return PREGO_OK;
}

template <int FMT>
int pack16(prego_model* m, const prego_weights_t* w, cudaStream_t s) {
    using OpT = typename Op16<FMT>::T;
    const int H = m->d.hidden_dim, E = m->d.embed_dim, K = m->d.num_classes, din = m->din;
    const int T = 256;
    auto g = [&](int64_t n) { return grid_for(n, T, m->sm_count); };
    f32_to_16<FMT><<<g((int64_t)E * din), T, 0, s>>>(w->layer1_0_weight, reinterpret_cast<OpT*>(m->w1_16[FMT]), (int64_t)E * din);
    pack_rows_16<FMT><<<g((int64_t)3 * H * E), T, 0, s>>>(w->gru_weight_ih_l0, reinterpret_cast<OpT*>(m->wih_16p[FMT]), 3 * H, 3 * H, E, H, 1);
    pack_rows_16<FMT><<<g((int64_t)3 * H * H), T, 0, s>>>(w->gru_weight_hh_l0, reinterpret_cast<OpT*>(m->whh_16p[FMT]), 3 * H, 3 * H, H, H, 1);
    if (m->kpad)
        pack_rows_16<FMT><<<g((int64_t)m->kpad * H), T, 0, s>>>(w->f_classification_0_weight, reinterpret_cast<OpT*>(m->wc_16p[FMT]), m->kpad, K, H, H, 0);
    if (m->online_stream[FMT] != nullptr)
        online_pack_stream<FMT><<<g((int64_t)m->sm_count * kFusedWarps * online_stream_loads(din / 1024) * 32), T, 0, s>>>(
            reinterpret_cast<const OpT*>(m->w1_16[FMT]), reinterpret_cast<const OpT*>(m->wih_16p[FMT]), reinterpret_cast<const OpT*>(m->whh_16p[FMT]),
            m->online_stream[FMT], m->sm_count, din, E, H);
    LAUNCH_CHECK("16-bit weight packing");
    return PREGO_OK;
}

template <int FMT>
int gemm16_test(const void* A, const void* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K, int tile_n,
                int sms, cudaStream_t s) {
    const DType dt = FMT == 0 ? kF16 : kBF16;
    CUtensorMap tmA, tmB;
    RC_TRY(make_tmap_a(&tmA, dt, A, K, M));
    if (tile_n < 0) {  // CTA-pair kernels: tile_n = -256 / -192
        RC_TRY(make_tmap_w(&tmB, dt, W, K, N, -tile_n / 2));
        if (tile_n == -256) return launch_gemm_tc2<256, 6, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<256, -1>{C, bias, N, 0, 0}, sms, s, "gemm256 (2cta)");
        if (tile_n == -192) return launch_gemm_tc2<192, 7, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<192, -1>{C, bias, N, 0, 0}, sms, s, "gemm192 (2cta)");
        return fail(PREGO_ERR_INVALID, "CTA-pair tile_n must be -256 or -192 (got %d)", tile_n);
    }
    RC_TRY(make_tmap_w(&tmB, dt, W, K, N, tile_n));
    switch (tile_n) {
        case 96: return launch_gemm_tc<96, 6, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<96, -1>{C, bias, N, 0, 0}, sms, s, "gemm96");
        case 128: return launch_gemm_tc<128, 6, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<128, -1>{C, bias, N, 0, 0}, sms, s, "gemm128");
        case 192: return launch_gemm_tc<192, 5, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<192, -1>{C, bias, N, 0, 0}, sms, s, "gemm192");
        case 256: return launch_gemm_tc<256, 4, FMT>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<256, -1>{C, bias, N, 0, 0}, sms, s, "gemm256");
        default: return fail(PREGO_ERR_INVALID, "tile_n must be 96, 128, 192 or 256 (got %d)", tile_n);
    }
}

}  // namespace

extern "C" {

int prego_abi_version(void) { return PREGO_ABI_VERSION; }
const char* prego_last_error(void) { return g_err; }

int prego_model_create(const prego_dims_t* dims, int32_t device, prego_model_t** out) {
    if (dims == nullptr || out == nullptr) return fail(PREGO_ERR_INVALID, "dims/out is NULL");
    const prego_dims_t d = *dims;
    const int din = d.d_rgb + d.d_flow;
    if (d.d_rgb < 0 || d.d_flow < 0 || din <= 0) return fail(PREGO_ERR_INVALID, "need at least one of rgb / flow (rnn.py:23-29)");
    if (din % 64 != 0 || d.d_rgb % 8 != 0) return fail(PREGO_ERR_INVALID, "input feature widths must be multiples of 64 (got %d + %d)", d.d_rgb, d.d_flow);
    if (d.embed_dim != 2048) return fail(PREGO_ERR_INVALID, "embedding_dim must be 2048 (got %d)", d.embed_dim);
    // the few-stream recurrence runs H / 8 co-resident CTAs (one warp per hidden unit): H <= 1024 keeps that within any B200's SMs
    if (d.hidden_dim <= 0 || d.hidden_dim % 64 != 0 || d.hidden_dim > 1024) return fail(PREGO_ERR_INVALID, "hidden_dim must be a multiple of 64, <= 1024 (got %d; the reference ships 1024)", d.hidden_dim);
    if (d.num_classes <= 0 || d.num_classes > 1024) return fail(PREGO_ERR_INVALID, "num_classes must be in [1, 1024] (got %d)", d.num_classes);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PREGO_ERR_INVALID, "prego_b200 is built for sm_100a (B200) only; device %d is sm_%d%d", device, prop.major, prop.minor);
    prego_model* m = new (std::nothrow) prego_model();
    if (m == nullptr) return fail(PREGO_ERR_INVALID, "out of host memory");
    m->d = d;
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->din = din;
    m->kpad = d.num_classes <= 96 ? 96 : (d.num_classes <= 128 ? 128 : 0);
    const int64_t H = d.hidden_dim, E = d.embed_dim, K = d.num_classes;
#define ALLOC(ptr, bytes) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes)))
#define ALLOC2(ptr, bytes) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes)))
    ALLOC(m->w1_f32, E * din * 4); ALLOC(m->b1, E * 4); ALLOC(m->ln_g, E * 4); ALLOC(m->ln_b, E * 4);
    ALLOC(m->wih_f32p, 3 * H * E * 4); ALLOC(m->whh_f32p, 3 * H * H * 4); ALLOC(m->bih_p, 3 * H * 4); ALLOC(m->bhh_p, 3 * H * 4); ALLOC(m->bgi_p, 3 * H * 4);
    ALLOC(m->wc_f32, K * H * 4); ALLOC(m->bc, K * 4);
    for (int f = 0; f < 2; ++f) {
        ALLOC(m->w1_16[f], E * din * 2); ALLOC(m->wih_16p[f], 3 * H * E * 2); ALLOC(m->whh_16p[f], 3 * H * H * 2);
        if (m->kpad) ALLOC(m->wc_16p[f], (int64_t)m->kpad * H * 2);
    }
    ALLOC(m->w1_x3, E * 3 * din * 2); ALLOC(m->wih_x3, 3 * H * 3 * E * 2); ALLOC(m->whh_x3, 3 * H * 3 * H * 2); ALLOC(m->absmax, 3 * sizeof(unsigned));
    ALLOC(m->xchg, 2 * 16 * H * sizeof(uint2)); ALLOC(m->err_flag, sizeof(int)); ALLOC(m->wct_f32, K * H * 4);
    ALLOC(m->online_scratch, online_fused_scratch_floats((int)E, (int)K) * 4);
#undef ALLOC
    CUDA_TRY(cudaMemset(m->online_scratch, 0, online_fused_scratch_floats((int)E, (int)K) * 4));
    if (online_fused_ok(m))
        for (int f = 0; f < 2; ++f) ALLOC2(m->online_stream[f], (size_t)m->sm_count * kFusedWarps * online_stream_loads(din / 1024) * 512);
    CUDA_TRY(cudaMemset(m->xchg, 0, 2 * 16 * H * sizeof(uint2)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->xchg_bwd), 2 * 16 * H * sizeof(uint4)));
    CUDA_TRY(cudaMemset(m->xchg_bwd, 0, 2 * 16 * H * sizeof(uint4)));
    CUDA_TRY(cudaMemset(m->err_flag, 0, sizeof(int)));
    CUDA_TRY(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_begin, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&m->ev_ready[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&m->ev_free[i], cudaEventDisableTiming));
    }
    *out = m;
    return PREGO_OK;
}

int prego_model_destroy(prego_model_t* m) {
    if (m == nullptr) return PREGO_OK;
    cudaSetDevice(m->device);
    void* ptrs[] = {m->w1_f32, m->b1, m->ln_g, m->ln_b, m->wih_f32p, m->whh_f32p, m->bih_p, m->bhh_p, m->bgi_p, m->wc_f32, m->bc,
                    m->w1_16[0], m->w1_16[1], m->wih_16p[0], m->wih_16p[1], m->whh_16p[0], m->whh_16p[1], m->wc_16p[0],
                    m->wc_16p[1], m->xchg, m->xchg_bwd, m->err_flag, m->wct_f32, m->online_scratch, m->online_stream[0], m->online_stream[1],
                    m->wa_f32, m->ba, m->wa_16[0], m->wa_16[1], m->w1_x3, m->wih_x3, m->whh_x3, m->absmax};
    for (void* p : ptrs)
        if (p != nullptr) cudaFree(p);
    for (cudaEvent_t e : m->prof_ev)
        if (e != nullptr) cudaEventDestroy(e);
    for (cudaEvent_t e : {m->ev_begin, m->ev_ready[0], m->ev_ready[1], m->ev_free[0], m->ev_free[1]})
        if (e != nullptr) cudaEventDestroy(e);
    if (m->side != nullptr) cudaStreamDestroy(m->side);
    delete m;
    return PREGO_OK;
}

int prego_model_load_weights(prego_model_t* m, const prego_weights_t* w, void* stream_) {
    return prego_model_load_weights_ex(m, w, PREGO_PACK_ALL, stream_);
}

int prego_model_load_weights_ex(prego_model_t* m, const prego_weights_t* w, uint32_t formats, void* stream_) {
    RC_TRY(check_model(m, false));
    if (w == nullptr) return fail(PREGO_ERR_INVALID, "weights is NULL");
    if ((formats & ~static_cast<uint32_t>(PREGO_PACK_ALL)) != 0) return fail(PREGO_ERR_INVALID, "unknown PREGO_PACK_* bits 0x%x", formats);
    formats |= PREGO_PACK_F32;
    const void* all[] = {w->layer1_0_weight, w->layer1_0_bias, w->layer1_1_weight, w->layer1_1_bias, w->gru_weight_ih_l0,
                         w->gru_weight_hh_l0, w->gru_bias_ih_l0, w->gru_bias_hh_l0, w->f_classification_0_weight,
                         w->f_classification_0_bias};
    for (const void* p : all)
        if (p == nullptr) return fail(PREGO_ERR_INVALID, "a weight pointer is NULL (all ten state_dict tensors are required)");
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CUDA_TRY(cudaSetDevice(m->device));
    const int H = m->d.hidden_dim, E = m->d.embed_dim, K = m->d.num_classes, din = m->din;
    const int T = 256;
    auto g = [&](int64_t n) { return grid_for(n, T, m->sm_count); };
    CUDA_TRY(cudaMemcpyAsync(m->w1_f32, w->layer1_0_weight, (size_t)E * din * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->b1, w->layer1_0_bias, (size_t)E * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->ln_g, w->layer1_1_weight, (size_t)E * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->ln_b, w->layer1_1_bias, (size_t)E * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->wc_f32, w->f_classification_0_weight, (size_t)K * H * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->bc, w->f_classification_0_bias, (size_t)K * 4, cudaMemcpyDeviceToDevice, s));
    pack_rows_f32<<<g((int64_t)3 * H * E), T, 0, s>>>(w->gru_weight_ih_l0, m->wih_f32p, 3 * H, E, H, 1);
    pack_rows_f32<<<g((int64_t)3 * H * H), T, 0, s>>>(w->gru_weight_hh_l0, m->whh_f32p, 3 * H, H, H, 1);
    pack_rows_f32<<<g(3 * H), T, 0, s>>>(w->gru_bias_ih_l0, m->bih_p, 3 * H, 1, H, 1);
    pack_rows_f32<<<g(3 * H), T, 0, s>>>(w->gru_bias_hh_l0, m->bhh_p, 3 * H, 1, H, 1);
    presum_gate_bias<<<g(3 * H), T, 0, s>>>(m->bih_p, m->bhh_p, m->bgi_p, 3 * H);
    transpose_f32<<<g((int64_t)K * H), T, 0, s>>>(w->f_classification_0_weight, m->wct_f32, K, H);
    LAUNCH_CHECK("fp32 weight packing");
    if (formats & PREGO_PACK_16) {
        RC_TRY(pack16<0>(m, w, s));
        RC_TRY(pack16<1>(m, w, s));
    }
    if (formats & PREGO_PACK_X3) {   // split-fp16 copies: power-of-two scale per matrix so that max |scale * w| sits near 2^14 (hi far from fp16's overflow, lo in
        // its normal range); one small synchronisation per load_state_dict
        const float* srcs[3] = {w->layer1_0_weight, w->gru_weight_ih_l0, w->gru_weight_hh_l0};
        const int64_t ns[3] = {(int64_t)E * din, (int64_t)3 * H * E, (int64_t)3 * H * H};
        CUDA_TRY(cudaMemsetAsync(m->absmax, 0, 3 * sizeof(unsigned), s));
        for (int i = 0; i < 3; ++i) absmax_f32<<<g(ns[i]), T, 0, s>>>(srcs[i], ns[i], m->absmax + i);
        unsigned bits[3];
        CUDA_TRY(cudaMemcpyAsync(bits, m->absmax, sizeof(bits), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        float scale[3];
        for (int i = 0; i < 3; ++i) {
            float mx;
            memcpy(&mx, &bits[i], 4);
            int e = (mx > 0.f && std::isfinite(mx)) ? static_cast<int>(std::floor(std::log2(16384.0f / mx))) : 0;
            e = std::max(-24, std::min(24, e));
            scale[i] = std::ldexp(1.0f, e);
            m->inv_scale_x3[i] = std::ldexp(1.0f, -e);
        }
        pack_rows_split3<<<g(ns[0]), T, 0, s>>>(w->layer1_0_weight, m->w1_x3, E, din, H, 0, scale[0]);
        pack_rows_split3<<<g(ns[1]), T, 0, s>>>(w->gru_weight_ih_l0, m->wih_x3, 3 * H, E, H, 1, scale[1]);
        pack_rows_split3<<<g(ns[2]), T, 0, s>>>(w->gru_weight_hh_l0, m->whh_x3, 3 * H, H, H, 1, scale[2]);
        LAUNCH_CHECK("split-fp16 weight packing");
    }
    m->loaded = true;
    m->packed = formats;
    return PREGO_OK;
}

size_t prego_workspace_bytes(const prego_model_t* m, int64_t B, int64_t chunk_T, int32_t precision) {
    if (m == nullptr || B <= 0 || chunk_T <= 0) return 0;
    return static_cast<size_t>(make_plan(m->d, B, chunk_T, precision).total);
}

static int forward_impl(prego_model_t* m, const prego_forward_args_t* a, const prego_anticipation_args_t* ant, void* stream_) {
    RC_TRY(check_model(m, true));
    if (a == nullptr) return fail(PREGO_ERR_INVALID, "args is NULL");
    const prego_dims_t& d = m->d;
    const int64_t B = a->B, T = a->T;
    if (B <= 0 || T <= 0) return fail(PREGO_ERR_INVALID, "B and T must be positive (got B=%lld, T=%lld)", (long long)B, (long long)T);
    if (a->feature_dtype != PREGO_FEAT_F32 && a->feature_dtype != PREGO_FEAT_16) return fail(PREGO_ERR_INVALID, "unknown feature_dtype %d", a->feature_dtype);
    if (a->flow_is_zero && d.d_rgb == 0) return fail(PREGO_ERR_INVALID, "flow_is_zero on a flow-only model leaves no input");
    if ((d.d_rgb > 0 && a->rgb == nullptr) || (d.d_flow > 0 && !a->flow_is_zero && a->flow == nullptr)) return fail(PREGO_ERR_INVALID, "rgb / flow pointer is NULL");
    const bool exact = a->precision == PREGO_PREC_FP32 || a->precision == PREGO_PREC_F16X3;
    if (a->feature_dtype == PREGO_FEAT_16 && exact)
        return fail(PREGO_ERR_INVALID, "16-bit features (PREGO_FEAT_16) need PREGO_PREC_F16 or PREGO_PREC_BF16; the fp32-class paths read fp32 features");
    if (a->precision != PREGO_PREC_BF16 && a->precision != PREGO_PREC_FP32 && a->precision != PREGO_PREC_F16 && a->precision != PREGO_PREC_F16X3)
        return fail(PREGO_ERR_INVALID, "unknown precision %d", a->precision);
    if (a->precision == PREGO_PREC_F16X3 && ((3 * d.hidden_dim) % 256 != 0 || m->din % 64 != 0))
        return fail(PREGO_ERR_INVALID, "PREGO_PREC_F16X3 needs 3 * hidden_dim %% 256 == 0 (got hidden_dim %d); use PREGO_PREC_FP32", d.hidden_dim);
    const bool h16 = !exact;
    const uint32_t need = a->precision == PREGO_PREC_F16X3 ? PREGO_PACK_X3 : h16 ? PREGO_PACK_16 : PREGO_PACK_F32;
    if ((m->packed & need) == 0)
        return fail(PREGO_ERR_STATE, "the operand format of precision %d was not packed by the last prego_model_load_weights_ex (formats 0x%x)", a->precision, m->packed);
    if (h16 && m->kpad == 0) return fail(PREGO_ERR_INVALID, "16-bit paths support num_classes <= 128 (got %d); use PREGO_PREC_FP32", d.num_classes);
    const int64_t Tc = (a->chunk_T > 0 && a->chunk_T < T) ? a->chunk_T : T;
    if (B * Tc >= (int64_t(1) << 31) / 4) return fail(PREGO_ERR_INVALID, "B * chunk_T = %lld is too large for one pass; lower chunk_T", (long long)(B * Tc));
    // double-buffered feature staging only when there is a next chunk to stage and the caller's workspace has room
    Plan p = make_plan(d, B, Tc, a->precision, true);
    const bool overlap = h16 && T > Tc && B > kLatencyMaxB && p.xb2 != 0 && a->workspace_bytes >= (size_t)p.total && use_overlap() &&
                         a->feature_dtype == PREGO_FEAT_F32;
    if (!overlap) p = make_plan(d, B, Tc, a->precision, false);
    if (a->workspace == nullptr || a->workspace_bytes < (size_t)p.total)
        return fail(PREGO_ERR_WORKSPACE, "workspace too small: need %lld bytes, got %zu", (long long)p.total, a->workspace_bytes);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 1023) != 0) return fail(PREGO_ERR_INVALID, "workspace must be 1024-byte aligned");

    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CUDA_TRY(cudaSetDevice(m->device));
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    const int H = d.hidden_dim;
    float* h_cur = reinterpret_cast<float*>(ws + p.h32_a);
    float* h_alt = reinterpret_cast<float*>(ws + p.h32_b);

    if (ant != nullptr) {
        if (m->ant_len <= 0) return fail(PREGO_ERR_STATE, "anticipation weights not loaded (prego_model_load_anticipation)");
        if (ant->workspace == nullptr || (reinterpret_cast<uintptr_t>(ant->workspace) & 1023) != 0)
            return fail(PREGO_ERR_INVALID, "anticipation workspace must be non-NULL and 1024-byte aligned");
        if (h16 && (static_cast<int64_t>(m->ant_len) * H) % 256 != 0)
            return fail(PREGO_ERR_INVALID, "16-bit anticipation path needs anticipation_length * hidden_dim %% 256 == 0; use PREGO_PREC_FP32");
        const int64_t need_rows = B * Tc < 128 ? B * Tc : 128;
        if (ant_slab_rows(m, ant, a->precision, B * Tc) < need_rows)
            return fail(PREGO_ERR_WORKSPACE, "anticipation workspace too small: need %lld bytes for a %lld-row slab, got %zu",
                        (long long)(need_rows * ant_row_bytes(m, a->precision)), (long long)need_rows, ant->workspace_bytes);
    }
    const bool online = ant == nullptr && h16 && T == 1 && B <= kOnlineMaxRows && d.d_rgb % 2 == 0 && m->din % 8 == 0 &&
                        a->feature_dtype == PREGO_FEAT_F32 && !a->flow_is_zero;
    if (online && a->h_state != nullptr) {
        h_cur = a->h_state;  // per-frame path: read the caller's state in place (one copy back instead of two)
    } else if (a->h_state != nullptr) {
        CUDA_TRY(cudaMemcpyAsync(h_cur, a->h_state, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, s));
    } else {
        CUDA_TRY(cudaMemsetAsync(h_cur, 0, (size_t)B * H * 4, s));
    }

    if (overlap) {
        // prologue: chunk 0 is staged on the main stream; both staging buffers start out free
        const int tc0 = static_cast<int>(T < Tc ? T : Tc);
        if (a->precision == PREGO_PREC_F16) RC_TRY(stage_chunk<0>(m, a, ws + p.xb, 0, tc0, 0, s));
        else RC_TRY(stage_chunk<1>(m, a, ws + p.xb, 0, tc0, 0, s));
        CUDA_TRY(cudaEventRecord(m->ev_ready[0], s));
        CUDA_TRY(cudaEventRecord(m->ev_free[1], s));  // also orders the side stream after everything before this call
    }
    int ci = 0;
    for (int64_t t0 = 0; t0 < T; t0 += Tc, ++ci) {
        const int tc = static_cast<int>(T - t0 < Tc ? T - t0 : Tc);
        const int64_t t_next = t0 + Tc;
        const int tc_next = t_next < T ? static_cast<int>(T - t_next < Tc ? T - t_next : Tc) : 0;
        prof_mark(m, s, -1, 0);
        if (online) {
            if (a->precision == PREGO_PREC_F16) RC_TRY(online_step<0>(m, a, p, ws, h_cur, h_alt, s));
            else RC_TRY(online_step<1>(m, a, p, ws, h_cur, h_alt, s));
            continue;
        }
        if (a->precision == PREGO_PREC_F16) RC_TRY(chunk_16<0>(m, a, p, ws, h_cur, h_alt, t0, tc, s, ci, overlap, t_next, tc_next, ant));
        else if (a->precision == PREGO_PREC_BF16) RC_TRY(chunk_16<1>(m, a, p, ws, h_cur, h_alt, t0, tc, s, ci, overlap, t_next, tc_next, ant));
        else if (a->precision == PREGO_PREC_F16X3) RC_TRY(chunk_x3(m, a, p, ws, h_cur, h_alt, t0, tc, s, ant));
        else RC_TRY(chunk_f32(m, a, p, ws, h_cur, h_alt, t0, tc, s, ant));
    }
    if (a->h_state != nullptr && h_cur != a->h_state)
        CUDA_TRY(cudaMemcpyAsync(a->h_state, h_cur, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, s));
    return PREGO_OK;
}

int prego_forward(prego_model_t* m, const prego_forward_args_t* a, void* stream_) { return forward_impl(m, a, nullptr, stream_); }

int prego_forward_anticipation(prego_model_t* m, const prego_forward_args_t* a, const prego_anticipation_args_t* ant, void* stream_) {
    if (ant == nullptr) return fail(PREGO_ERR_INVALID, "anticipation args is NULL");
    return forward_impl(m, a, ant, stream_);
}

int prego_model_load_anticipation(prego_model_t* m, int32_t A, const float* w, const float* b, void* stream_) {
    RC_TRY(check_model(m, false));
    if (A <= 0 || A > 1024) return fail(PREGO_ERR_INVALID, "anticipation_length must be in [1, 1024] (got %d)", A);
    if (w == nullptr || b == nullptr) return fail(PREGO_ERR_INVALID, "anticipation weight / bias pointer is NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CUDA_TRY(cudaSetDevice(m->device));
    const int64_t H = m->d.hidden_dim, n = (int64_t)A * H * H;
    if (A != m->ant_len) {
        CUDA_TRY(cudaStreamSynchronize(s));  // nothing may still read the old copies
        for (void* p : {(void*)m->wa_f32, (void*)m->ba, m->wa_16[0], m->wa_16[1]})
            if (p != nullptr) cudaFree(p);
        m->wa_f32 = m->ba = nullptr;
        m->wa_16[0] = m->wa_16[1] = nullptr;
        m->ant_len = 0;
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->wa_f32), n * 4));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->ba), (int64_t)A * H * 4));
        CUDA_TRY(cudaMalloc(&m->wa_16[0], n * 2));
        CUDA_TRY(cudaMalloc(&m->wa_16[1], n * 2));
    }
    CUDA_TRY(cudaMemcpyAsync(m->wa_f32, w, n * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m->ba, b, (int64_t)A * H * 4, cudaMemcpyDeviceToDevice, s));
    f32_to_16<0><<<grid_for(n, 256, m->sm_count), 256, 0, s>>>(w, static_cast<__half*>(m->wa_16[0]), n);
    f32_to_16<1><<<grid_for(n, 256, m->sm_count), 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(m->wa_16[1]), n);
    LAUNCH_CHECK("anticipation weight packing");
    m->ant_len = A;
    return PREGO_OK;
}

size_t prego_anticipation_workspace_bytes(const prego_model_t* m, int64_t slab_rows, int32_t precision) {
    if (m == nullptr || m->ant_len <= 0 || slab_rows <= 0) return 0;
    return static_cast<size_t>(align_up(slab_rows * ant_row_bytes(m, precision), 1024));
}

size_t prego_ap_workspace_bytes(int64_t N, int32_t K) {
    if (N <= 0 || K <= 0) return 0;
    const int64_t S = ap_num_slices(N, K);
    // two key buffers | digit counts [K][256][S] | slice partials {positives, last threshold} | float64 slice sums
    return static_cast<size_t>(2 * align_up(N * K * 4, 1024) + align_up((int64_t)K * 256 * S * 4, 1024) + 2 * align_up((int64_t)K * S * 8, 1024));
}

int prego_perframe_ap(const float* scores, const float* targets, const int32_t* target_labels, int64_t N, int32_t K, double* ap,
                      int64_t* num_pos, void* workspace, size_t workspace_bytes, int32_t* err_flag, void* stream) {
    if (scores == nullptr || (targets == nullptr && target_labels == nullptr) || ap == nullptr || num_pos == nullptr || err_flag == nullptr)
        return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (N <= 0 || N >= (int64_t(1) << 31) || K <= 0 || K > 65535) return fail(PREGO_ERR_INVALID, "need 0 < N < 2^31 frames and 0 < K <= 65535 classes (got N=%lld, K=%d)", (long long)N, K);
    if (workspace == nullptr || workspace_bytes < prego_ap_workspace_bytes(N, K))
        return fail(PREGO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", prego_ap_workspace_bytes(N, K), workspace_bytes);
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return fail(PREGO_ERR_INVALID, "workspace must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int S = ap_num_slices(N, K);
    const int64_t L = ap_slice_len(N, S);
    uint8_t* w = static_cast<uint8_t*>(workspace);
    uint32_t* keys_a = reinterpret_cast<uint32_t*>(w);
    uint32_t* keys_b = reinterpret_cast<uint32_t*>(w += align_up(N * K * 4, 1024));
    uint32_t* hist = reinterpret_cast<uint32_t*>(w += align_up(N * K * 4, 1024));
    uint2* part = reinterpret_cast<uint2*>(w += align_up((int64_t)K * 256 * S * 4, 1024));
    double* acc_part = reinterpret_cast<double*>(w += align_up((int64_t)K * S * 8, 1024));
    int64_t gx = (N + 31) / 32;
    if (gx > 148 * 16) gx = 148 * 16;
    ap_build_keys<<<(unsigned)gx, 256, 0, s>>>(scores, targets, target_labels, N, K, keys_a, err_flag);
    uint32_t *in = keys_a, *out = keys_b;
    for (int pass = 0; pass < 4; ++pass) {  // keys are < 2^31: four 8-bit digits; the order ends up back in keys_a
        ap_hist_kernel<<<K * S, kApThreads, 0, s>>>(in, N, S, L, 8 * pass, hist);
        ap_offsets_kernel<<<K, 256, 0, s>>>(hist, S);
        ap_scatter_kernel<<<K * S, kApThreads, 0, s>>>(in, out, N, S, L, 8 * pass, hist);
        uint32_t* t = in;
        in = out;
        out = t;
    }
    ap_scan_local<<<K * S, kApThreads, 0, s>>>(in, N, S, L, part);
    ap_scan_final<<<K * S, kApThreads, 0, s>>>(in, N, S, L, part, acc_part);
    ap_reduce<<<(K + 127) / 128, 128, 0, s>>>(part, acc_part, K, S, ap, num_pos);
    LAUNCH_CHECK("perframe_ap kernels");
    return PREGO_OK;
}

int prego_device_error(prego_model_t* m, int32_t* out) {
    RC_TRY(check_model(m, false));
    if (out == nullptr) return fail(PREGO_ERR_INVALID, "out is NULL");
    int v = 0;
    CUDA_TRY(cudaMemcpy(&v, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));  // synchronises with prior work
    *out = v;
    if (v != 0) CUDA_TRY(cudaMemset(m->err_flag, 0, sizeof(int)));
    return PREGO_OK;
}

int prego_recurrence_fallbacks(prego_model_t* m, int64_t* out) {
    RC_TRY(check_model(m, false));
    if (out == nullptr) return fail(PREGO_ERR_INVALID, "out is NULL");
    *out = m->coop_fallbacks;
    return PREGO_OK;
}

int prego_profile_begin(prego_model_t* m) {
    RC_TRY(check_model(m, false));
    m->prof = true;
    m->prof_n = 0;
    for (auto& l : m->prof_launches) l = 0;
    return PREGO_OK;
}

int prego_profile_end(prego_model_t* m, double* phase_ms, int64_t* phase_launches) {
    RC_TRY(check_model(m, false));
    if (phase_ms == nullptr || phase_launches == nullptr) return fail(PREGO_ERR_INVALID, "NULL output pointer");
    m->prof = false;
    for (int i = 0; i < PREGO_NUM_PHASES; ++i) {
        phase_ms[i] = 0.0;
        phase_launches[i] = m->prof_launches[i];
    }
    if (m->prof_n > 0) CUDA_TRY(cudaEventSynchronize(m->prof_ev[m->prof_n - 1]));
    for (int i = 1; i < m->prof_n; ++i) {
        const int ph = m->prof_phase[i];
        if (ph < 0) continue;
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, m->prof_ev[i - 1], m->prof_ev[i]));
        phase_ms[ph] += ms;
    }
    m->prof_n = 0;
    return PREGO_OK;
}

int prego_window_mode(const int32_t* labels, const int64_t* offsets, const int64_t* win_offsets, int32_t B,
                      int64_t total_windows, int32_t window, int32_t num_labels, int32_t* modes, int32_t* err_flag,
                      void* stream) {
    if (labels == nullptr || offsets == nullptr || win_offsets == nullptr || modes == nullptr || err_flag == nullptr)
        return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (B <= 0 || window <= 0) return fail(PREGO_ERR_INVALID, "B and window must be positive");
    if (num_labels <= 0 || num_labels > kAggMaxLabels) return fail(PREGO_ERR_INVALID, "num_labels must be in [1, %d] (got %d)", kAggMaxLabels, num_labels);
    if (total_windows <= 0) return PREGO_OK;
    const int warps = kAggThreads / 32;
    int64_t blocks = (total_windows + warps - 1) / warps;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)warps * num_labels * sizeof(int);
    window_mode_kernel<<<(int)blocks, kAggThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        labels, offsets, win_offsets, B, window, num_labels, modes, err_flag);
    LAUNCH_CHECK("window_mode_kernel");
    return PREGO_OK;
}

int prego_rle(const int32_t* seq, const int64_t* seg_offsets, const int64_t* final_len, int32_t B, int64_t scale,
              int32_t* out_vals, int64_t* out_changes, int32_t* counts, void* stream) {
    if (seq == nullptr || seg_offsets == nullptr || final_len == nullptr || out_vals == nullptr || out_changes == nullptr || counts == nullptr)
        return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (B <= 0) return fail(PREGO_ERR_INVALID, "B must be positive");
    rle_kernel<<<B, kAggThreads, 0, static_cast<cudaStream_t>(stream)>>>(seq, seg_offsets, final_len, scale, out_vals, out_changes, counts);
    LAUNCH_CHECK("rle_kernel");
    return PREGO_OK;
}

int prego_gemm16_nt(const void* A, const void* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                    int32_t tile_n, int32_t precision, void* stream) {
    if (A == nullptr || W == nullptr || bias == nullptr || C == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (M <= 0 || N <= 0 || K <= 0 || K % kTileK != 0 || tile_n == 0 || N % (tile_n < 0 ? -tile_n : tile_n) != 0) return fail(PREGO_ERR_INVALID, "need K %% 64 == 0 and N %% tile_n == 0");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == PREGO_PREC_F16) return gemm16_test<0>(A, W, bias, C, M, N, K, tile_n, sms, s);
    if (precision == PREGO_PREC_BF16) return gemm16_test<1>(A, W, bias, C, M, N, K, tile_n, sms, s);
    return fail(PREGO_ERR_INVALID, "precision must be PREGO_PREC_F16 or PREGO_PREC_BF16");
}

// Test / diagnostic entry points of the fused LayerNorm path (gemm_xf.cuh), N % 256 == 0, K % 64 == 0:
//  prego_gemm16_ln_nt:    C[M, N] fp32 = relu((Y * a_r + b_r) * gamma + beta) W^T + bias, Y fp16 [M, K], rowstat = (a_r, b_r) per row;
//                         rowstat == NULL runs the IDENTITY transform (A passes through the extra pipeline hop unchanged).
//  prego_gemm16_stats_nt: Y[M, N] fp16 = A W^T + bias plus the LayerNorm row statistics of the rounded Y (partials + finalised).
int prego_gemm16_ln_nt(const void* Y, const float* rowstat, const float* gamma, const float* beta, const void* W, const float* bias,
                       float* C, int64_t M, int64_t N, int64_t K, int32_t precision, void* stream) {
    if (Y == nullptr || W == nullptr || bias == nullptr || C == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (rowstat != nullptr && (gamma == nullptr || beta == nullptr)) return fail(PREGO_ERR_INVALID, "gamma / beta is NULL");
    if (M <= 0 || N <= 0 || K <= 0 || K % kTileK != 0 || N % 256 != 0) return fail(PREGO_ERR_INVALID, "need K %% 64 == 0 and N %% 256 == 0");
    if (precision != PREGO_PREC_F16 && precision != PREGO_PREC_BF16) return fail(PREGO_ERR_INVALID, "precision must be PREGO_PREC_F16 or PREGO_PREC_BF16");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool f16 = precision == PREGO_PREC_F16;
    CUtensorMap tmA, tmB;
    // the raw A tile is fp16 for the LayerNorm transform (y is stored as fp16); the identity transform moves 16-bit words as they are
    RC_TRY(make_tmap_a(&tmA, (rowstat != nullptr || f16) ? kF16 : kBF16, Y, K, M));
    RC_TRY(make_tmap_w(&tmB, f16 ? kF16 : kBF16, W, K, N, 128));
    LnXf xf{reinterpret_cast<const float2*>(rowstat), gamma, beta};
    if (const char* e = getenv("PREGO_XF_DBG")) xf.dbg = atoi(e);
    const EpiStore<256, -1> epi{C, bias, N, 0, 0};
    if (rowstat == nullptr) {
        if (f16) return launch_gemm_xf<256, 6, 0, EpiStore<256, -1>, true>(tmA, tmB, (int)M, (int)N, (int)K, epi, xf, sms, s, "gemm_xf identity");
        return launch_gemm_xf<256, 6, 1, EpiStore<256, -1>, true>(tmA, tmB, (int)M, (int)N, (int)K, epi, xf, sms, s, "gemm_xf identity");
    }
    if (f16) return launch_gemm_xf<256, 6, 0, EpiStore<256, -1>, false>(tmA, tmB, (int)M, (int)N, (int)K, epi, xf, sms, s, "gemm_xf LayerNorm");
    return launch_gemm_xf<256, 6, 1, EpiStore<256, -1>, false>(tmA, tmB, (int)M, (int)N, (int)K, epi, xf, sms, s, "gemm_xf LayerNorm");
}

int prego_gemm16_stats_nt(const void* A, const void* W, const float* bias, void* Y, float* stats, float* rowstat, int64_t M, int64_t N,
                          int64_t K, int32_t precision, float eps, void* stream) {
    if (A == nullptr || W == nullptr || bias == nullptr || Y == nullptr || stats == nullptr || rowstat == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (M <= 0 || N <= 0 || K <= 0 || K % kTileK != 0 || N % 256 != 0) return fail(PREGO_ERR_INVALID, "need K %% 64 == 0 and N %% 256 == 0");
    if (precision != PREGO_PREC_F16 && precision != PREGO_PREC_BF16) return fail(PREGO_ERR_INVALID, "precision must be PREGO_PREC_F16 or PREGO_PREC_BF16");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const DType dt = precision == PREGO_PREC_F16 ? kF16 : kBF16;
    CUtensorMap tmA, tmB;
    RC_TRY(make_tmap_a(&tmA, dt, A, K, M));
    RC_TRY(make_tmap_w(&tmB, dt, W, K, N, 128));
    const EpiStoreStats<256> epi{static_cast<__half*>(Y), bias, N, reinterpret_cast<float2*>(stats), (int)M};
    if (precision == PREGO_PREC_F16) RC_TRY((launch_gemm_tc2<256, 6, 0>(tmA, tmB, (int)M, (int)N, (int)K, 0, epi, sms, s, "gemm + LN statistics")));
    else RC_TRY((launch_gemm_tc2<256, 6, 1>(tmA, tmB, (int)M, (int)N, (int)K, 0, epi, sms, s, "gemm + LN statistics")));
    ln_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float2*>(stats), reinterpret_cast<float2*>(rowstat), (int)M, (int)(N / 256), (int)N, eps);
    LAUNCH_CHECK("ln_finalize_kernel");
    return PREGO_OK;
}

int prego_gemm_tf32_nt(const float* A, const float* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                       int32_t accumulate, void* stream) {
    if (A == nullptr || W == nullptr || C == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (M <= 0 || N <= 0 || K <= 0) return fail(PREGO_ERR_INVALID, "bad shape");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // accumulate bit 8 (0x100) selects the narrow 64-column tiles (test hook for the training recurrence's per-step products)
    return gemm_tf32_nt(A, K, W, K, bias, C, N, (int)M, (int)N, (int)K, accumulate & 1, sms, static_cast<cudaStream_t>(stream),
                        (accumulate & 0x100) ? 64 : 256);
}

int prego_gemm_f32_nt(const float* A, const float* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                      void* stream) {
    if (A == nullptr || W == nullptr || C == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (M <= 0 || N <= 0 || K <= 0 || K % 16 != 0) return fail(PREGO_ERR_INVALID, "need K %% 16 == 0");
    SgemmA a{A, nullptr, (int)K, K, 0, 0, 1, 1, 0};
    sgemm_nt_f32<<<dim3((unsigned)((N + 127) / 128), (unsigned)((M + 127) / 128)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, W, bias, C, (int)M, (int)N, (int)K, N);
    LAUNCH_CHECK("sgemm_nt_f32");
    return PREGO_OK;
}

}  // extern "C"

#include "train_api.inc"
#include "online_api.inc"
