// Host half of the feature ingest (reference: datasets/dataset.py:120-132 hands the model fp32 host
// tensors).  End to end the path is bound by the host->device link (16 KiB per frame as fp32), and the first thing the
// device does with a feature is round it to the 16-bit operand format of the projection GEMM.  Doing that rounding on the
// host, with the device path's exact rule, halves the bytes on the link and leaves the results bit-identical
// (PREGO_FEAT_16).  This file is data-format staging only: no part of the model is computed on the CPU.
//
// Rounding rule = Op16<FMT>::from_float of csrc/gemm_tc.cuh: fp16: clamp to +-65504 (NaN -> -65504, as fmaxf/fminf do on
// the device), round to nearest even; bf16: round to nearest even, NaN -> 0x7FFF.
#include <cuda_runtime_api.h>
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <new>
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

// fmt: bit 0 = 0 fp16 / 1 bf16; kCachedStores: plain stores (the destination is about to be read out of the cache by the
// DMA engine) instead of streaming stores (the destination is a large staging buffer)
constexpr int kCachedStores = 256;

void round_scalar(const float* src, uint16_t* dst, int64_t n, int fmt) {
    if ((fmt & 1) == 0)
        for (int64_t i = 0; i < n; ++i) dst[i] = f32_to_f16_scalar(src[i]);
    else
        for (int64_t i = 0; i < n; ++i) dst[i] = f32_to_bf16_scalar(src[i]);
}

__attribute__((target("avx2,f16c"))) void round_avx2(const float* src, uint16_t* dst, int64_t n, int fmt) {
    const bool nt = (fmt & kCachedStores) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    fmt &= 1;  // streaming stores: the staging buffer is only read by the DMA engine
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
    const bool nt = (fmt & kCachedStores) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    fmt &= 1;
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

// 1 when every value is +-0.0 (the reference loader's flow dummy is np.zeros, datasets/dataset.py:63-69), else 0.
// Threads stop early once any of them has seen a non-zero value.
extern "C" int prego_host_all_zero(const float* src, int64_t n, int32_t num_threads) {
    if (src == nullptr || n < 0) return -1;
    std::atomic<int> nonzero{0};
    auto scan = [&](int64_t b, int64_t e) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(src);
        for (int64_t i = b; i < e && !nonzero.load(std::memory_order_relaxed);) {
            const int64_t stop = std::min(e, i + 16384);
            uint32_t acc = 0;
            for (; i < stop; ++i) acc |= p[i];
            if (acc & 0x7FFFFFFFu) nonzero.store(1, std::memory_order_relaxed);
        }
    };
    int64_t nt = num_threads > 0 ? num_threads : 1;
    const int64_t min_chunk = 1 << 18;
    if (nt > (n + min_chunk - 1) / min_chunk) nt = (n + min_chunk - 1) / min_chunk;
    if (nt <= 1) {
        scan(0, n);
    } else {
        const int64_t chunk = (n + nt - 1) / nt;
        std::vector<std::thread> pool;
        for (int64_t t = 0; t < nt; ++t) {
            const int64_t b = t * chunk, e = std::min(n, b + chunk);
            if (b < e) pool.emplace_back(scan, b, e);
        }
        for (auto& th : pool) th.join();
    }
    return nonzero.load() ? 0 : 1;
}

extern "C" int prego_host_round_impl(void) {
    static const RoundFn fn = pick_impl();
    return fn == round_avx512 ? 2 : (fn == round_avx2 ? 1 : 0);
}


// ---- Ring stager: round a large fp32 host tensor and copy it to the device through a SMALL pinned ring.
// Rounding into a full-size staging buffer costs the host's memory system 4 B read + 2 B written + 2 B read again by
// the DMA engine per value, and that memory system -- not the link, not the cores -- is what bounds the end-to-end
// rate.  With a ring of a few slots of a few MiB the rounded values are written with plain stores and are still in the
// last-level cache when the DMA engine reads them, so DRAM only sees the 4 B fp32 read; the slots are large because
// the copy engine loses ~30 % on sub-MiB copies (measured).  A persistent pool works through the tensor in small
// chunks claimed from one atomic counter (a descheduled thread just takes fewer chunks: no barrier, no fixed shares);
// whoever rounds the last chunk of a slot hands the slot to cudaMemcpyAsync, whoever claims the first chunk of a slot
// waits for the copy that last read it.  The pool spins while a run is in flight and sleeps between runs.
struct prego_host_stager {
    static constexpr int64_t kChunk = 32768;  // values per claim (128 KiB of fp32 in, 64 KiB out)
    int threads = 0, slots = 0, device = 0;
    int64_t slot_elems = 0, chunks_per_slot = 0;
    uint16_t* ring = nullptr;                     // [slots][slot_elems]
    std::vector<cudaEvent_t> ev;                  // [slots] recorded after the slot's copy
    std::vector<std::atomic<int64_t>> issued;     // [slots] 1 + last slot instance whose copy was enqueued
    std::vector<std::atomic<int>> filled;         // [slots] chunks rounded into the current instance
    std::atomic<int64_t> free_upto{-1};           // slot instances <= this may be written
    std::atomic<int64_t> next_chunk{0};
    int64_t inst_base = 0;                        // slot instances used by earlier runs (ring position carries over)
    RoundFn fn = nullptr;
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> gen{0};
    std::atomic<int> done{0}, sleepers{0}, failed{0};
    std::atomic<bool> stop{false};
    const float* job_src = nullptr;
    uint16_t* job_dst = nullptr;
    int64_t job_n = 0, job_chunks = 0;
    int job_fmt = 0;
    cudaStream_t job_stream = nullptr;
    std::mutex call_mu;  // one run() at a time

    prego_host_stager(int nslots) : issued(nslots), filled(nslots) {}

    void work() {
        for (;;) {
            const int64_t c = next_chunk.fetch_add(1, std::memory_order_relaxed);
            if (c >= job_chunks || failed.load(std::memory_order_relaxed)) return;
            const int64_t li = c / chunks_per_slot;          // slot instance within this run
            const int64_t gi = inst_base + li;               // ... and since the stager was created
            const int slot = static_cast<int>(gi % slots);
            if (c % chunks_per_slot == 0) {
                // first claim of the instance: the copy out of this slot's previous instance must have been enqueued
                // (by whoever finished it) and completed
                if (gi >= slots) {
                    while (issued[slot].load(std::memory_order_acquire) != gi - slots + 1 && !failed.load(std::memory_order_relaxed)) _mm_pause();
                    if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { (void)cudaGetLastError(); failed.store(1); }
                }
                filled[slot].store(0, std::memory_order_relaxed);
                while (free_upto.load(std::memory_order_acquire) != gi - 1 && !failed.load(std::memory_order_relaxed)) _mm_pause();  // instances open in order
                free_upto.store(gi, std::memory_order_release);
            } else {
                while (free_upto.load(std::memory_order_acquire) < gi && !failed.load(std::memory_order_relaxed)) _mm_pause();
            }
            const int64_t off = c * kChunk, m = std::min(kChunk, job_n - off);
            const int64_t in_slot = off - li * slot_elems;
            uint16_t* r = ring + static_cast<int64_t>(slot) * slot_elems;
            fn(job_src + off, r + in_slot, m, job_fmt);
            const int64_t slot_first = li * chunks_per_slot;
            const int slot_chunks = static_cast<int>(std::min(chunks_per_slot, job_chunks - slot_first));
            if (filled[slot].fetch_add(1, std::memory_order_acq_rel) + 1 == slot_chunks) {
                // last chunk of the instance: hand the slot to the copy engine
                const int64_t e0 = li * slot_elems, em = std::min(slot_elems, job_n - e0);
                if (cudaMemcpyAsync(job_dst + e0, r, static_cast<size_t>(em) * 2, cudaMemcpyHostToDevice, job_stream) != cudaSuccess ||
                    cudaEventRecord(ev[slot], job_stream) != cudaSuccess) {
                    (void)cudaGetLastError();
                    failed.store(1);
                }
                issued[slot].store(gi + 1, std::memory_order_release);
            }
        }
    }

    void worker() {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
                if (++spins < 20000) {
                    _mm_pause();
                } else {  // idle between batches: sleep instead of burning a core
                    std::unique_lock<std::mutex> lk(mu);
                    sleepers.fetch_add(1);
                    cv.wait(lk, [&] { return gen.load(std::memory_order_acquire) != seen || stop.load(); });
                    sleepers.fetch_sub(1);
                }
            }
            if (stop.load()) return;
            seen = gen.load(std::memory_order_acquire);
            work();
            done.fetch_add(1, std::memory_order_release);
        }
    }
};

extern "C" int prego_host_stager_create(int32_t num_threads, int32_t ring_slots, int64_t slot_bytes, prego_host_stager_t** out) {
    if (out == nullptr || num_threads <= 0 || num_threads > 1024 || ring_slots < 2 || ring_slots > 64 || slot_bytes < 2 * prego_host_stager::kChunk)
        return PREGO_ERR_INVALID;
    prego_host_stager* s = new (std::nothrow) prego_host_stager(ring_slots);
    if (s == nullptr) return PREGO_ERR_INVALID;
    s->threads = num_threads;
    s->slots = ring_slots;
    s->chunks_per_slot = slot_bytes / 2 / prego_host_stager::kChunk;
    s->slot_elems = s->chunks_per_slot * prego_host_stager::kChunk;
    s->fn = pick_impl();
    if (cudaGetDevice(&s->device) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&s->ring), static_cast<size_t>(ring_slots) * s->slot_elems * 2, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        delete s;
        return PREGO_ERR_CUDA;
    }
    s->ev.resize(ring_slots);
    for (int i = 0; i < ring_slots; ++i) {
        s->issued[i].store(0);
        s->filled[i].store(0);
        if (cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            for (int j = 0; j < i; ++j) cudaEventDestroy(s->ev[j]);
            cudaFreeHost(s->ring);
            delete s;
            return PREGO_ERR_CUDA;
        }
    }
    for (int t = 0; t + 1 < num_threads; ++t) s->pool.emplace_back(&prego_host_stager::worker, s);  // + the caller = num_threads
    *out = s;
    return PREGO_OK;
}

extern "C" int prego_host_stager_destroy(prego_host_stager_t* s) {
    if (s == nullptr) return PREGO_OK;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop.store(true);
        s->cv.notify_all();
    }
    for (auto& th : s->pool) th.join();
    for (size_t i = 0; i < s->ev.size(); ++i) {
        if (s->issued[i].load() > 0) cudaEventSynchronize(s->ev[i]);
        cudaEventDestroy(s->ev[i]);
    }
    cudaFreeHost(s->ring);
    delete s;
    return PREGO_OK;
}

extern "C" int prego_host_stager_run(prego_host_stager_t* s, const float* src, void* dst_device, int64_t n, int32_t precision, void* stream) {
    if (s == nullptr || src == nullptr || dst_device == nullptr || n < 0) return PREGO_ERR_INVALID;
    if (precision != PREGO_PREC_F16 && precision != PREGO_PREC_BF16) return PREGO_ERR_INVALID;
    if (n == 0) return PREGO_OK;
    std::lock_guard<std::mutex> call(s->call_mu);
    s->job_src = src;
    s->job_dst = static_cast<uint16_t*>(dst_device);
    s->job_n = n;
    s->job_chunks = (n + prego_host_stager::kChunk - 1) / prego_host_stager::kChunk;
    s->job_fmt = (precision == PREGO_PREC_F16 ? 0 : 1) | kCachedStores;
    s->job_stream = static_cast<cudaStream_t>(stream);
    s->failed.store(0);
    s->next_chunk.store(0);
    s->free_upto.store(s->inst_base - 1);
    s->done.store(0, std::memory_order_relaxed);
    s->gen.fetch_add(1, std::memory_order_release);
    if (s->sleepers.load() > 0) {
        std::lock_guard<std::mutex> lk(s->mu);
        s->cv.notify_all();
    }
    s->work();  // the calling thread works too instead of spinning on a core the pool could use
    while (s->done.load(std::memory_order_acquire) < s->threads - 1) _mm_pause();
    s->inst_base += (s->job_chunks + s->chunks_per_slot - 1) / s->chunks_per_slot;
    return s->failed.load() ? PREGO_ERR_CUDA : PREGO_OK;
}
