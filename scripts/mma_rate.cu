// Microbenchmark (diagnostic, not part of the library): how long does ONE tcgen05.mma.cta_group::2.kind::f16 of shape
// M = 256 (128 per CTA) x N x K = 16 take, back to back, with the operands already in shared memory?
// Every CTA pair of the chip issues R MMAs on a fixed (zeroed) operand stage, commits once, and reports SM cycles and
// nanoseconds per instruction.  Compares N = 64 ... 256: the recurrence's 192-column tiles against the 256-column
// projection tiles.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I prego_b200/csrc -o /tmp/mma_rate scripts/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace prego;

__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int N, int NSLICE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_rate_kernel(int reps, long long* cyc_out, long long* ns_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // NSLICE stages of [128 rows x 128 B] A + [N/2 rows x 128 B] B, all zero
    constexpr int kStage = 16384 + (N / 2) * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NSLICE * kStage);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    for (int i = threadIdx.x; i < NSLICE * kStage / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool leader = ptx::cluster_ctarank() == 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(tmem_slot, 512);
        ptx::tmem_relinquish2();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler: UTCHMMA takes it from a uniform register without a waterfall loop
    if (warp == 0 && lane == 0 && leader) {
        constexpr uint32_t idesc = ptx::make_idesc(0, 256, N);
        const long long c0 = clock64();
        const uint64_t t0 = gtime();
        for (int r = 0; r < reps; ++r) {
            const uint32_t sa = ptx::smem_u32(smem + (r % NSLICE) * kStage);
            const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ptx::mma_f16_ss_2sm(tmem_base + ((r >> 4) & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (r & 15) | k ? 1u : 0u);
        }
        ptx::mma_commit_2sm(&bar[0], 1);
        ptx::mbar_wait(&bar[0], 0);
        const long long c1 = clock64();
        const uint64_t t1 = gtime();
        cyc_out[blockIdx.x >> 1] = c1 - c0;
        ns_out[blockIdx.x >> 1] = (long long)(t1 - t0);
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}


// The projection / recurrence kernels' main-loop STRUCTURE without any data movement: a producer thread and an MMA
// thread hand NSTAGE "stages" back and forth through full / empty mbarriers (tcgen05.commit multicast frees a stage),
// KB k-blocks of 4 MMAs per "tile", accumulators double-buffered without an epilogue.  Shows what the handshake
// alone costs per MMA for a given N and ring depth.
template <int N, int NSTAGE, int KB, bool PEER_ARRIVES, int MPI = 4, bool FENCE = true, bool COMMIT1 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
ring_kernel(int tiles, long long* cyc_out, long long* ns_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kStage = 16384 + (N / 2) * 128;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * kStage);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* fin = empty_bar + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fin + 2);
    for (int i = threadIdx.x; i < NSTAGE * kStage / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool leader = ptx::cluster_ctarank() == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            ptx::mbar_init(&full_bar[s], PEER_ARRIVES ? 2 : 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(&fin[0], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc2(tmem_slot, 512);
        ptx::tmem_relinquish2();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler: UTCHMMA takes it from a uniform register without a waterfall loop
    if (warp == 0 && lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < tiles; ++t)
            for (int kb = 0; kb < KB; ++kb) {
                if (COMMIT1 && !leader) break;
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (leader) ptx::mbar_arrive(&full_bar[stage]);
                else if (PEER_ARRIVES) ptx::mbar_arrive_leader(&full_bar[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
    } else if (warp == 1 && leader && ptx::elect_one()) {
        constexpr uint32_t idesc = ptx::make_idesc(0, 256, N);
        int stage = 0;
        uint32_t phase = 0;
        const long long c0 = clock64();
        const uint64_t t0 = gtime();
        for (int t = 0; t < tiles; ++t) {
            const uint32_t tmem_d = tmem_base + (t & 1) * 256;
            for (int kb = 0; kb < KB; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                if (FENCE) ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * kStage);
                const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + 16384);
#pragma unroll
                for (int k = 0; k < MPI; ++k) ptx::mma_f16_ss_2sm(tmem_d, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), idesc, (kb | k) != 0 ? 1u : 0u);
                if (COMMIT1) ptx::mma_commit(&empty_bar[stage]);   // non-multicast commit: frees the stage in the leader only (peer producer idles)
                else ptx::mma_commit_2sm(&empty_bar[stage], 3);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
        ptx::mma_commit_2sm(&fin[0], 1);
        ptx::mbar_wait(&fin[0], 0);
        const long long c1 = clock64();
        const uint64_t t1 = gtime();
        cyc_out[blockIdx.x >> 1] = c1 - c0;
        ns_out[blockIdx.x >> 1] = (long long)(t1 - t0);
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

template <int N, int NSTAGE, int KB, bool PEER, int MPI = 4, bool FENCE = true, bool COMMIT1 = false>
void run_ring(int tiles) {
    const int smem = NSTAGE * (16384 + (N / 2) * 128) + 1024 + 256;
    auto k = ring_kernel<N, NSTAGE, KB, PEER, MPI, FENCE, COMMIT1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long *cyc, *ns;
    cudaMalloc(&cyc, 74 * 8);
    cudaMalloc(&ns, 74 * 8);
    for (int it = 0; it < 3; ++it) {
        k<<<148, 128, smem>>>(tiles, cyc, ns);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ring N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    }
    long long hc[74], hn[74];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    cudaMemcpy(hn, ns, sizeof(hn), cudaMemcpyDeviceToHost);
    double c = 0, n = 0;
    for (int i = 0; i < 74; ++i) { c += hc[i]; n += hn[i]; }
    c /= 74; n /= 74;
    const double mmas = (double)MPI * KB * tiles;
    printf("ring N=%3d stages=%d k-blocks/tile=%2d peer_arrives=%d mma/iter=%d fence=%d commit1=%d: %6.1f cycles / MMA (ideal %d) = %6.1f cycles / iteration, clock %.0f MHz\n", N, NSTAGE, KB, (int)PEER, MPI, (int)FENCE, (int)COMMIT1, c / mmas, N / 2, c / mmas * MPI, c / n * 1e3);
    cudaFree(cyc);
    cudaFree(ns);
}

template <int N>
void run(int reps) {
    constexpr int NSLICE = 4;
    const int smem = NSLICE * (16384 + (N / 2) * 128) + 1024 + 64;
    auto k = mma_rate_kernel<N, NSLICE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long *cyc, *ns;
    cudaMalloc(&cyc, 74 * 8);
    cudaMalloc(&ns, 74 * 8);
    for (int it = 0; it < 3; ++it) {
        k<<<148, 128, smem>>>(reps, cyc, ns);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("N=%d: %s\n", N, cudaGetErrorString(e));
            return;
        }
    }
    long long hc[74], hn[74];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    cudaMemcpy(hn, ns, sizeof(hn), cudaMemcpyDeviceToHost);
    double c = 0, n = 0;
    for (int i = 0; i < 74; ++i) { c += hc[i]; n += hn[i]; }
    c /= 74; n /= 74;
    const double mmas = 4.0 * reps;
    printf("M=256 N=%3d K=16 cta_group::2: %7.1f cycles / MMA, %6.1f ns / MMA (SM clock %.0f MHz) -> %7.1f TFLOP/s chip-wide; ideal %d cycles\n", N, c / mmas,
           n / mmas, c / n * 1e3, 74 * mmas * 2.0 * 256 * N * 16 / n / 1e3, N / 2);
    cudaFree(cyc);
    cudaFree(ns);
}

int main() {
    const int reps = 20000;
    run<256>(reps);
    run<192>(reps);
    run<128>(reps);
    run<64>(reps);
    run<192>(reps);
    run<256>(reps);
    run_ring<256, 6, 32, false>(400);
    run_ring<192, 5, 16, false>(800);
    run_ring<192, 5, 16, false, 4, false>(800);
    run_ring<192, 5, 16, false, 4, true, true>(800);
    run_ring<192, 5, 8, false, 8>(800);
    run_ring<192, 5, 16, false, 2>(800);
    run_ring<192, 5, 16, false, 1>(800);
    run_ring<128, 6, 16, false, 8>(800);
    run_ring<64, 6, 16, false, 8>(800);
    run_ring<256, 6, 32, false, 2>(400);
    run_ring<256, 6, 32, false, 1>(400);
    return 0;
}
