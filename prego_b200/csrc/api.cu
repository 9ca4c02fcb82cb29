// C ABI of prego_b200 (see include/prego_b200.h): model handle, weight packing, the forward
// pipeline (time-chunked, carried GRU state) and the aggregation entry points.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/prego_b200.h"
#include "aggregate.cuh"
#include "gemm_tc.cuh"
#include "gru_latency.cuh"
#include "simt_kernels.cuh"

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
        if (_e != cudaSuccess)                                                                      \
            return fail(PREGO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                          \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess)                                                                      \
            return fail(PREGO_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
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

// bf16 K-major operand map, 128B swizzle.  3-D view (k, inner, outer): box (64, 1, box_rows).
int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t k, uint64_t inner, uint64_t outer,
                   uint64_t inner_stride_elems, uint64_t outer_stride_elems, uint32_t box_rows) {
    int rc = get_encode_fn();
    if (rc != PREGO_OK) return rc;
    cuuint64_t dims[3] = {k, inner, outer};
    cuuint64_t strides[2] = {inner_stride_elems * 2, outer_stride_elems * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(kTileK), 1, box_rows};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PREGO_ERR_CUDA, "cuTensorMapEncodeTiled (3d) failed with CUresult %d", (int)r);
    return PREGO_OK;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t k, uint64_t rows, uint32_t box_rows) {
    int rc = get_encode_fn();
    if (rc != PREGO_OK) return rc;
    cuuint64_t dims[2] = {k, rows};
    cuuint64_t strides[1] = {k * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PREGO_ERR_CUDA, "cuTensorMapEncodeTiled (2d) failed with CUresult %d", (int)r);
    return PREGO_OK;
}

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
inline int grid_for(int64_t work_items, int threads, int sm_count) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
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
    // fp32 (exact path + SIMT recurrence)
    float *w1_f32 = nullptr, *b1 = nullptr, *ln_g = nullptr, *ln_b = nullptr;
    float *wih_f32p = nullptr, *whh_f32p = nullptr, *bih_p = nullptr, *bhh_p = nullptr;
    float *wc_f32 = nullptr, *bc = nullptr;
    // bf16 operands of the tcgen05 path
    __nv_bfloat16 *w1_bf = nullptr, *wih_bfp = nullptr, *whh_bfp = nullptr, *wc_bfp = nullptr;
    // latency-kernel exchange
    uint2* xchg = nullptr;
    int* err_flag = nullptr;
    uint32_t tag_base = 0;
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
}  // namespace

namespace {

struct Plan {
    int64_t h32_a, h32_b;     // [B, H] fp32 state ping-pong
    int64_t xb;               // bf16 [Mc, Din]
    int64_t ye;               // bf16 or fp32 [Mc, E]  (y, normalised in place to e)
    int64_t gi;               // fp32 [Mc, 3H]
    int64_t hseq;             // bf16 [B, Tc+1, H]   (tensor-core recurrence only)
    int64_t hrelu;            // bf16 or fp32 [Mc, H]
    int64_t gh;               // fp32 [B, 3H]        (fp32 batched recurrence only)
    int64_t logits;           // fp32 [Mc, K]        (fp32 head only)
    int64_t total;
};

Plan make_plan(const prego_dims_t& d, int64_t B, int64_t Tc, int prec) {
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
    const bool bf = prec == PREGO_PREC_BF16;
    p.xb = bf ? take(Mc * Din * 2) : 0;
    p.ye = take(Mc * E * (bf ? 2 : 4));
    p.gi = take(Mc * 3 * H * 4);
    p.hseq = (bf && B > kLatencyMaxB) ? take(B * (Tc + 1) * H * 2) : 0;
    p.hrelu = take(Mc * H * (bf ? 2 : 4));
    p.gh = (!bf && B > kLatencyMaxB) ? take(B * 3 * H * 4) : 0;
    p.logits = bf ? 0 : take(Mc * K * 4);
    p.total = off;
    return p;
}

template <int TILE_N, int STAGES, class Epi>
int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int a_c1, const Epi& epi,
                   int sm_count, cudaStream_t stream, const char* name) {
    using Cfg = GemmCfg<TILE_N>;
    auto kfn = gemm_tc_kernel<TILE_N, STAGES, Epi>;
    static bool attr_set = false;  // per instantiation
    const int smem = Cfg::smem_bytes(STAGES);
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int tiles = (N / TILE_N) * ((M + kTileM - 1) / kTileM);
    const int grid = tiles < sm_count ? tiles : sm_count;
    kfn<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, M, N, K, a_c1, epi);
    LAUNCH_CHECK(name);
    return PREGO_OK;
}

int check_model(const prego_model* m, bool need_weights) {
    if (m == nullptr) return fail(PREGO_ERR_INVALID, "model handle is NULL");
    if (need_weights && !m->loaded) return fail(PREGO_ERR_STATE, "weights not loaded: call prego_model_load_weights first");
    return PREGO_OK;
}

template <int NB>
int launch_latency(const GruLatencyArgs& a, int H, cudaStream_t stream) {
    auto kfn = gru_latency_kernel<NB>;
    const size_t smem = (3 * kLatUnitsPerCta * H + 2 * NB * H) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    GruLatencyArgs args = a;
    void* params[] = {&args};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kfn, dim3(H / kLatUnitsPerCta), dim3(kLatThreads), params, smem, stream));
    return PREGO_OK;
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
    if (d.hidden_dim <= 0 || d.hidden_dim % 64 != 0 || d.hidden_dim > 2048) return fail(PREGO_ERR_INVALID, "hidden_dim must be a multiple of 64, <= 2048 (got %d)", d.hidden_dim);
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
    ALLOC(m->w1_f32, E * din * 4); ALLOC(m->b1, E * 4); ALLOC(m->ln_g, E * 4); ALLOC(m->ln_b, E * 4);
    ALLOC(m->wih_f32p, 3 * H * E * 4); ALLOC(m->whh_f32p, 3 * H * H * 4); ALLOC(m->bih_p, 3 * H * 4); ALLOC(m->bhh_p, 3 * H * 4);
    ALLOC(m->wc_f32, K * H * 4); ALLOC(m->bc, K * 4);
    ALLOC(m->w1_bf, E * din * 2); ALLOC(m->wih_bfp, 3 * H * E * 2); ALLOC(m->whh_bfp, 3 * H * H * 2);
    if (m->kpad) ALLOC(m->wc_bfp, (int64_t)m->kpad * H * 2);
    ALLOC(m->xchg, 2 * 4 * H * sizeof(uint2)); ALLOC(m->err_flag, sizeof(int));
#undef ALLOC
    CUDA_TRY(cudaMemset(m->xchg, 0, 2 * 4 * H * sizeof(uint2)));
    CUDA_TRY(cudaMemset(m->err_flag, 0, sizeof(int)));
    *out = m;
    return PREGO_OK;
}

int prego_model_destroy(prego_model_t* m) {
    if (m == nullptr) return PREGO_OK;
    cudaSetDevice(m->device);
    void* ptrs[] = {m->w1_f32, m->b1, m->ln_g, m->ln_b, m->wih_f32p, m->whh_f32p, m->bih_p, m->bhh_p, m->wc_f32, m->bc,
                    m->w1_bf, m->wih_bfp, m->whh_bfp, m->wc_bfp, m->xchg, m->err_flag};
    for (void* p : ptrs)
        if (p != nullptr) cudaFree(p);
    for (cudaEvent_t e : m->prof_ev)
        if (e != nullptr) cudaEventDestroy(e);
    delete m;
    return PREGO_OK;
}

int prego_model_load_weights(prego_model_t* m, const prego_weights_t* w, void* stream_) {
    int rc = check_model(m, false);
    if (rc != PREGO_OK) return rc;
    if (w == nullptr) return fail(PREGO_ERR_INVALID, "weights is NULL");
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
    f32_to_bf16<<<g((int64_t)E * din), T, 0, s>>>(w->layer1_0_weight, m->w1_bf, (int64_t)E * din);
    pack_rows_bf16<<<g((int64_t)3 * H * E), T, 0, s>>>(w->gru_weight_ih_l0, m->wih_bfp, 3 * H, 3 * H, E, H, 1);
    pack_rows_bf16<<<g((int64_t)3 * H * H), T, 0, s>>>(w->gru_weight_hh_l0, m->whh_bfp, 3 * H, 3 * H, H, H, 1);
    if (m->kpad)
        pack_rows_bf16<<<g((int64_t)m->kpad * H), T, 0, s>>>(w->f_classification_0_weight, m->wc_bfp, m->kpad, K, H, H, 0);
    LAUNCH_CHECK("weight packing");
    m->loaded = true;
    return PREGO_OK;
}

size_t prego_workspace_bytes(const prego_model_t* m, int64_t B, int64_t chunk_T, int32_t precision) {
    if (m == nullptr || B <= 0 || chunk_T <= 0) return 0;
    return static_cast<size_t>(make_plan(m->d, B, chunk_T, precision).total);
}

int prego_forward(prego_model_t* m, const prego_forward_args_t* a, void* stream_) {
    int rc = check_model(m, true);
    if (rc != PREGO_OK) return rc;
    if (a == nullptr) return fail(PREGO_ERR_INVALID, "args is NULL");
    const prego_dims_t& d = m->d;
    const int64_t B = a->B, T = a->T;
    if (B <= 0 || T <= 0) return fail(PREGO_ERR_INVALID, "B and T must be positive (got B=%lld, T=%lld)", (long long)B, (long long)T);
    if ((d.d_rgb > 0 && a->rgb == nullptr) || (d.d_flow > 0 && a->flow == nullptr)) return fail(PREGO_ERR_INVALID, "rgb / flow pointer is NULL");
    if (a->precision != PREGO_PREC_BF16 && a->precision != PREGO_PREC_FP32) return fail(PREGO_ERR_INVALID, "unknown precision %d", a->precision);
    const bool bf = a->precision == PREGO_PREC_BF16;
    if (bf && m->kpad == 0) return fail(PREGO_ERR_INVALID, "bf16 path supports num_classes <= 128 (got %d); use PREGO_PREC_FP32", d.num_classes);
    const int64_t Tc = (a->chunk_T > 0 && a->chunk_T < T) ? a->chunk_T : T;
    if (B * Tc >= (int64_t(1) << 31) / 4) return fail(PREGO_ERR_INVALID, "B * chunk_T = %lld is too large for one pass; lower chunk_T", (long long)(B * Tc));
    const Plan p = make_plan(d, B, Tc, a->precision);
    if (a->workspace == nullptr || a->workspace_bytes < (size_t)p.total)
        return fail(PREGO_ERR_WORKSPACE, "workspace too small: need %lld bytes, got %zu", (long long)p.total, a->workspace_bytes);
    if ((reinterpret_cast<uintptr_t>(a->workspace) & 1023) != 0) return fail(PREGO_ERR_INVALID, "workspace must be 1024-byte aligned");

    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CUDA_TRY(cudaSetDevice(m->device));
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    const int H = d.hidden_dim, E = d.embed_dim, K = d.num_classes, Din = m->din;
    float* h_cur = reinterpret_cast<float*>(ws + p.h32_a);
    float* h_alt = reinterpret_cast<float*>(ws + p.h32_b);
    __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws + p.xb);
    void* ye = ws + p.ye;
    float* gi = reinterpret_cast<float*>(ws + p.gi);
    __nv_bfloat16* hseq = reinterpret_cast<__nv_bfloat16*>(ws + p.hseq);
    void* hrelu = ws + p.hrelu;
    float* gh = reinterpret_cast<float*>(ws + p.gh);
    float* logits_ws = reinterpret_cast<float*>(ws + p.logits);
    const bool tensor_rec = B > kLatencyMaxB;

    if (a->h_state != nullptr)
        CUDA_TRY(cudaMemcpyAsync(h_cur, a->h_state, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, s));
    else
        CUDA_TRY(cudaMemsetAsync(h_cur, 0, (size_t)B * H * 4, s));

    for (int64_t t0 = 0; t0 < T; t0 += Tc) {
        const int tc = static_cast<int>(T - t0 < Tc ? T - t0 : Tc);
        const int64_t Mc = B * tc;
        const int Mi = static_cast<int>(Mc);
        prof_mark(m, s, -1, 0);
        if (bf) {
            // 1. stage features: concat + bf16
            stage_features_bf16<<<grid_for(Mc * (Din / 8), 256, m->sm_count), 256, 0, s>>>(a->rgb, a->flow, xb, Mc, d.d_rgb, d.d_flow, tc, (int)T, (int)t0);
            LAUNCH_CHECK("stage_features_bf16");
            prof_mark(m, s, PREGO_PHASE_STAGE, 1);
            // 2. y = x W1^T + b1   (bf16 out)
            CUtensorMap tmA, tmB;
            if ((rc = make_tmap_bf16(&tmA, xb, Din, 1, Mc, Din, Din, kTileM)) != PREGO_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmB, m->w1_bf, Din, E, 256)) != PREGO_OK) return rc;
            EpiStore<256, __nv_bfloat16> ep1{reinterpret_cast<__nv_bfloat16*>(ye), m->b1, E};
            if ((rc = launch_gemm_tc<256, 4>(tmA, tmB, Mi, E, Din, 0, ep1, m->sm_count, s, "gemm1")) != PREGO_OK) return rc;
            prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
            // 3. e = relu(LN(y))   (in place)
            layernorm_relu_bf16<2048><<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(
                reinterpret_cast<const __nv_bfloat16*>(ye), reinterpret_cast<__nv_bfloat16*>(ye), m->ln_g, m->ln_b, Mc, 1e-5f);
            LAUNCH_CHECK("layernorm_relu_bf16");
            prof_mark(m, s, PREGO_PHASE_LAYERNORM, 1);
            // 4. gi = e W_ih'^T + b_ih'  (fp32 out, gate-interleaved columns)
            if ((rc = make_tmap_bf16(&tmA, ye, E, 1, Mc, E, E, kTileM)) != PREGO_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmB, m->wih_bfp, E, 3 * H, 192)) != PREGO_OK) return rc;
            EpiStore<192, float> ep2{gi, m->bih_p, 3 * H};
            if ((rc = launch_gemm_tc<192, 5>(tmA, tmB, Mi, 3 * H, E, 0, ep2, m->sm_count, s, "gemm2")) != PREGO_OK) return rc;
            prof_mark(m, s, PREGO_PHASE_GEMM2, 1);
        } else {
            SgemmA A1{a->rgb, a->flow, d.d_rgb, d.d_rgb, d.d_flow, 1, tc, (int)T, (int)t0};
            if (d.d_rgb == 0) { A1.a0 = a->flow; A1.a1 = nullptr; A1.k_split = Din; A1.lda0 = d.d_flow; }
            if (d.d_flow == 0) { A1.a1 = nullptr; A1.k_split = Din; }
            float* y32 = reinterpret_cast<float*>(ye);
            sgemm_nt_f32<<<dim3(E / 128, (Mi + 127) / 128), 256, 0, s>>>(A1, m->w1_f32, m->b1, y32, Mi, E, Din, E);
            LAUNCH_CHECK("sgemm gemm1");
            prof_mark(m, s, PREGO_PHASE_GEMM1, 1);
            layernorm_relu_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(y32, y32, m->ln_g, m->ln_b, Mc, E, 1e-5f);
            LAUNCH_CHECK("layernorm_relu_f32");
            prof_mark(m, s, PREGO_PHASE_LAYERNORM, 1);
            SgemmA A2{y32, nullptr, E, E, 0, 0, tc, (int)T, (int)t0};
            sgemm_nt_f32<<<dim3(3 * H / 128, (Mi + 127) / 128), 256, 0, s>>>(A2, m->wih_f32p, m->bih_p, gi, Mi, 3 * H, E, 3 * H);
            LAUNCH_CHECK("sgemm gemm2");
            prof_mark(m, s, PREGO_PHASE_GEMM2, 1);
        }

        // 5. recurrence over the tc steps of this chunk
        if (!tensor_rec) {
            for (int b0 = 0; b0 < B; b0 += 4) {
                const int nb = (int)(B - b0 < 4 ? B - b0 : 4);
                GruLatencyArgs la{m->whh_f32p, m->bhh_p, gi, h_cur, h_alt, hrelu, m->xchg, m->err_flag, H, tc, b0, nb, m->tag_base, bf ? 0 : 1};
                m->tag_base += static_cast<uint32_t>(tc);
                if (nb == 1) rc = launch_latency<1>(la, H, s);
                else if (nb == 2) rc = launch_latency<2>(la, H, s);
                else rc = launch_latency<4>(la, H, s);
                if (rc != PREGO_OK) return rc;
            }
            float* tmp = h_cur; h_cur = h_alt; h_alt = tmp;
        } else if (bf) {
            // slot 0 of the bf16 state history = bf16(carried state); step t reads slot t, writes slot t+1
            init_hseq_slot0<<<grid_for(B * H, 256, m->sm_count), 256, 0, s>>>(h_cur, hseq, B, H, tc + 1);
            LAUNCH_CHECK("init_hseq_slot0");
            CUtensorMap tmA, tmB;
            if ((rc = make_tmap_bf16(&tmA, hseq, H, tc + 1, B, H, (uint64_t)(tc + 1) * H, kTileM)) != PREGO_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmB, m->whh_bfp, H, 3 * H, 192)) != PREGO_OK) return rc;
            for (int t = 0; t < tc; ++t) {
                EpiGruStep eg{gi, m->bhh_p, h_cur, hseq, reinterpret_cast<__nv_bfloat16*>(hrelu), t, tc, H};
                if ((rc = launch_gemm_tc<192, 5>(tmA, tmB, (int)B, 3 * H, H, t, eg, m->sm_count, s, "gru_step")) != PREGO_OK) return rc;
            }
        } else {
            float* hr32 = reinterpret_cast<float*>(hrelu);
            for (int t = 0; t < tc; ++t) {
                SgemmA Ah{h_cur, nullptr, H, H, 0, 0, tc, (int)T, (int)t0};
                sgemm_nt_f32<<<dim3(3 * H / 128, (int)((B + 127) / 128)), 256, 0, s>>>(Ah, m->whh_f32p, m->bhh_p, gh, (int)B, 3 * H, H, 3 * H);
                gru_gates_f32<<<grid_for(B * H, 256, m->sm_count), 256, 0, s>>>(gi, gh, h_cur, hr32, (int)B, H, tc, t);
            }
            LAUNCH_CHECK("fp32 recurrence");
        }

        prof_mark(m, s, PREGO_PHASE_RECURRENCE, tensor_rec ? (bf ? tc + 1 : 2 * tc) : (int)((B + 3) / 4));
        // 6. head
        if (bf) {
            CUtensorMap tmA, tmB;
            if ((rc = make_tmap_bf16(&tmA, hrelu, H, 1, Mc, H, H, kTileM)) != PREGO_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmB, m->wc_bfp, H, m->kpad, m->kpad)) != PREGO_OK) return rc;
            if (m->kpad == 96) {
                EpiHead<96> eh{m->bc, a->probs, a->logits, a->labels, K, tc, (int)T, (int)t0};
                rc = launch_gemm_tc<96, 6>(tmA, tmB, Mi, 96, H, 0, eh, m->sm_count, s, "head96");
            } else {
                EpiHead<128> eh{m->bc, a->probs, a->logits, a->labels, K, tc, (int)T, (int)t0};
                rc = launch_gemm_tc<128, 6>(tmA, tmB, Mi, 128, H, 0, eh, m->sm_count, s, "head128");
            }
            if (rc != PREGO_OK) return rc;
        } else {
            SgemmA Ah{reinterpret_cast<const float*>(hrelu), nullptr, H, H, 0, 0, tc, (int)T, (int)t0};
            sgemm_nt_f32<<<dim3((K + 127) / 128, (Mi + 127) / 128), 256, 0, s>>>(Ah, m->wc_f32, m->bc, logits_ws, Mi, K, H, K);
            softmax_argmax_f32<<<grid_for(Mc * 32, 256, m->sm_count), 256, 0, s>>>(logits_ws, a->probs, a->logits, a->labels, Mc, K, tc, (int)T, (int)t0);
            LAUNCH_CHECK("fp32 head");
        }
        prof_mark(m, s, PREGO_PHASE_HEAD, bf ? 1 : 2);
    }
    if (a->h_state != nullptr)
        CUDA_TRY(cudaMemcpyAsync(a->h_state, h_cur, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, s));
    return PREGO_OK;
}

int prego_profile_begin(prego_model_t* m) {
    int rc = check_model(m, false);
    if (rc != PREGO_OK) return rc;
    m->prof = true;
    m->prof_n = 0;
    for (auto& l : m->prof_launches) l = 0;
    return PREGO_OK;
}

int prego_profile_end(prego_model_t* m, double* phase_ms, int64_t* phase_launches) {
    int rc = check_model(m, false);
    if (rc != PREGO_OK) return rc;
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

int prego_gemm_bf16_nt(const void* A, const void* W, const float* bias, float* C, int64_t M, int64_t N, int64_t K,
                       int32_t tile_n, void* stream) {
    if (A == nullptr || W == nullptr || bias == nullptr || C == nullptr) return fail(PREGO_ERR_INVALID, "NULL pointer argument");
    if (M <= 0 || N <= 0 || K <= 0 || K % kTileK != 0 || N % tile_n != 0) return fail(PREGO_ERR_INVALID, "need K %% 64 == 0 and N %% tile_n == 0");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUtensorMap tmA, tmB;
    int rc;
    if ((rc = make_tmap_bf16(&tmA, A, K, 1, M, K, K, kTileM)) != PREGO_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tmB, W, K, N, tile_n)) != PREGO_OK) return rc;
    switch (tile_n) {
        case 96: return launch_gemm_tc<96, 6>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<96, float>{C, bias, N}, sms, s, "gemm96");
        case 128: return launch_gemm_tc<128, 6>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<128, float>{C, bias, N}, sms, s, "gemm128");
        case 192: return launch_gemm_tc<192, 5>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<192, float>{C, bias, N}, sms, s, "gemm192");
        case 256: return launch_gemm_tc<256, 4>(tmA, tmB, (int)M, (int)N, (int)K, 0, EpiStore<256, float>{C, bias, N}, sms, s, "gemm256");
        default: return fail(PREGO_ERR_INVALID, "tile_n must be 96, 128, 192 or 256 (got %d)", tile_n);
    }
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
