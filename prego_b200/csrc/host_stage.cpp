// Host half of the feature ingest (reference: datasets/dataset.py:120-132 hands the model fp32 host
// tensors).  End to end the path is bound by the host->device link (16 KiB per frame as fp32), and the first thing the
// device does with a feature is round it to the 16-bit operand format of the projection GEMM.  Doing that rounding on the
// host, with the device path's exact rule, halves the bytes on the link and leaves the results bit-identical
// (PREGO_FEAT_16).  This file is data-format staging only: no part of the model is computed on the CPU.
//
// Rounding rule = Op16<FMT>::from_float of csrc/gemm_tc.cuh: fp16: clamp to +-65504 (NaN -> -65504, as fmaxf/fminf do on
// the device), round to nearest even; bf16: round to nearest even, NaN -> 0x7FFF.
#include <immintrin.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/prego_b200.h"

namespace {

inline uint16_t f32_to_f16_scalar(float f) {
    // fmaxf(NaN, x) = x on the device, so NaN ends up at the lower clamp
    if (!(f == f)) f = -65504.f;
    f = std::min(std::max(f, -65504.f), 65504.f);
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x < 0x33000001u) return static_cast<uint16_t>(sign);  // rounds to zero (<= 2^-25)
    if (x < 0x38800000u) {                                    // subnormal half
        const int shift = 126 - static_cast<int>(x >> 23);    // 14..24
        const uint32_t mant = (x & 0x7FFFFFu) | 0x800000u;
        const uint32_t q = mant >> shift, rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
        return static_cast<uint16_t>(sign | (q + ((rem > half || (rem == half && (q & 1u))) ? 1u : 0u)));
    }
    const uint32_t rounded = x + 0xFFFu + ((x >> 13) & 1u);   // round to nearest even on the 13 dropped bits
    return static_cast<uint16_t>(sign | ((rounded - 0x38000000u) >> 13));
}

inline uint16_t f32_to_bf16_scalar(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFFu;
    return static_cast<uint16_t>((x + 0x7FFFu + ((x >> 16) & 1u)) >> 16);
}

void round_scalar(const float* src, uint16_t* dst, int64_t n, int fmt) {
    if (fmt == 0)
        for (int64_t i = 0; i < n; ++i) dst[i] = f32_to_f16_scalar(src[i]);
    else
        for (int64_t i = 0; i < n; ++i) dst[i] = f32_to_bf16_scalar(src[i]);
}

__attribute__((target("avx2,f16c"))) void round_avx2(const float* src, uint16_t* dst, int64_t n, int fmt) {
    const bool nt = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;  // streaming stores: the staging buffer is only read by the DMA engine
    int64_t i = 0;
    if (fmt == 0) {
        const __m256 lo = _mm256_set1_ps(-65504.f), hi = _mm256_set1_ps(65504.f);
        for (; i + 8 <= n; i += 8) {
            __m256 v = _mm256_loadu_ps(src + i);
            v = _mm256_min_ps(_mm256_max_ps(v, lo), hi);  // max(NaN, lo) = lo: second operand, like fmaxf on the device
            const __m128i h = _mm256_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
            if (nt) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), h);
            else _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h);
        }
    } else {
        const __m256i bias = _mm256_set1_epi32(0x7FFF), one = _mm256_set1_epi32(1), qnan = _mm256_set1_epi32(0x7FFF);
        for (; i + 8 <= n; i += 8) {
            const __m256 v = _mm256_loadu_ps(src + i);
            const __m256i u = _mm256_castps_si256(v);
            __m256i r = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(u, bias), _mm256_and_si256(_mm256_srli_epi32(u, 16), one)), 16);
            const __m256i isnan = _mm256_castps_si256(_mm256_cmp_ps(v, v, _CMP_UNORD_Q));
            r = _mm256_blendv_epi8(r, qnan, isnan);
            const __m256i p = _mm256_packus_epi32(r, r);  // per 128-bit lane: [r0..r3 r0..r3 | r4..r7 r4..r7]
            const __m128i h = _mm_unpacklo_epi64(_mm256_castsi256_si128(p), _mm256_extracti128_si256(p, 1));
            if (nt) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), h);
            else _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h);
        }
    }
    if (nt) _mm_sfence();
    round_scalar(src + i, dst + i, n - i, fmt);
}

__attribute__((target("avx512f,avx512bw"))) void round_avx512(const float* src, uint16_t* dst, int64_t n, int fmt) {
    const bool nt = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    int64_t i = 0;
    if (fmt == 0) {
        const __m512 lo = _mm512_set1_ps(-65504.f), hi = _mm512_set1_ps(65504.f);
        for (; i + 16 <= n; i += 16) {
            __m512 v = _mm512_loadu_ps(src + i);
            v = _mm512_min_ps(_mm512_max_ps(v, lo), hi);
            const __m256i h = _mm512_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
            if (nt) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), h);
            else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), h);
        }
    } else {
        const __m512i bias = _mm512_set1_epi32(0x7FFF), one = _mm512_set1_epi32(1), qnan = _mm512_set1_epi32(0x7FFF);
        for (; i + 16 <= n; i += 16) {
            const __m512 v = _mm512_loadu_ps(src + i);
            const __m512i u = _mm512_castps_si512(v);
            __m512i r = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(u, bias), _mm512_and_si512(_mm512_srli_epi32(u, 16), one)), 16);
            r = _mm512_mask_mov_epi32(r, _mm512_cmp_ps_mask(v, v, _CMP_UNORD_Q), qnan);
            const __m256i h = _mm512_cvtepi32_epi16(r);
            if (nt) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), h);
            else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), h);
        }
    }
    if (nt) _mm_sfence();
    round_scalar(src + i, dst + i, n - i, fmt);
}

using RoundFn = void (*)(const float*, uint16_t*, int64_t, int);

RoundFn pick_impl() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return round_avx512;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c")) return round_avx2;
    return round_scalar;
}

}  // namespace

extern "C" int prego_host_round_features(const float* src, void* dst, int64_t n, int32_t precision, int32_t num_threads) {
    if (src == nullptr || dst == nullptr || n < 0) return PREGO_ERR_INVALID;
    if (precision != PREGO_PREC_F16 && precision != PREGO_PREC_BF16) return PREGO_ERR_INVALID;
    static const RoundFn fn = pick_impl();
    const int fmt = precision == PREGO_PREC_F16 ? 0 : 1;
    uint16_t* out = static_cast<uint16_t*>(dst);
    int64_t nt = num_threads > 0 ? num_threads : 1;
    const int64_t min_chunk = 1 << 16;  // below 256 KiB of input a thread costs more than it saves
    if (nt > (n + min_chunk - 1) / min_chunk) nt = (n + min_chunk - 1) / min_chunk;
    if (nt <= 1) {
        fn(src, out, n, fmt);
        return PREGO_OK;
    }
    const int64_t chunk = ((n + nt - 1) / nt + 63) / 64 * 64;  // 64-element boundaries keep the streaming stores aligned
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(nt));
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        pool.emplace_back(fn, src + b, out + b, e - b, fmt);
    }
    for (auto& th : pool) th.join();
    return PREGO_OK;
}

extern "C" int prego_host_round_impl(void) {
    static const RoundFn fn = pick_impl();
    return fn == round_avx512 ? 2 : (fn == round_avx2 ? 1 : 0);
}
