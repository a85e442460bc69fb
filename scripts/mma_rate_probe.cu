// Hardware probe (not product code): cycles per tcgen05.mma (kind::f16, bf16, M=128, SS mode, SWIZZLE_128B K-major operands already
// resident in shared memory) for N = 64 / 128 / 256, alone and with the other warps streaming through shared memory
// (the role TMA writes and the epilogue staging play in the real kernel).  One CTA per SM, all 148 SMs busy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe scripts/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    for (uint32_t i = 0; i < (1u << 24) && !ok; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    if (!ok) __trap();
}
__device__ __forceinline__ void mbar_spin(uint64_t* b, uint32_t par) {      // non-blocking test_wait poll
    uint32_t ok = 0;
    for (uint32_t i = 0; i < (1u << 26) && !ok; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    if (!ok) __trap();
}
template <bool kSpin> __device__ __forceinline__ void mbar_w(uint64_t* b, uint32_t par) { if (kSpin) mbar_spin(b, par); else mbar_wait(b, par); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct P { long long* cycles; unsigned idesc; int n, kblocks, traffic, stages, commit_every, handshake; float* sink; };

// smem: stages x (A 16 KB + B 32 KB) + scratch 32 KB for the traffic warps
template <bool kCommit, bool kFence, int kRing, int kMma>
__global__ void __launch_bounds__(288, 1) probe(const P p) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = 16384 + 32768;
    uint8_t* scratch = sm + p.stages * stage_bytes;
    uint64_t* done = reinterpret_cast<uint64_t*>(scratch + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 2);
    uint64_t* full = done + 8;
    uint64_t* empty = done + 16;
    volatile int* stop = reinterpret_cast<volatile int*>(done + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (p.stages * stage_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + (i & 0xff);
    if (threadIdx.x == 0) { mbar_init(done, 1); mbar_init(done + 1, 1); mbar_init(done + 2, 1); for (int i = 0; i < 8; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); } *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (lane == 0) {
            // lean issue loop: 4 stages unrolled, descriptors precomputed, no integer division, accumulate flag constant
            uint64_t ad[4], bd[4];
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                const uint32_t a = smem_u32(sm + s2 * stage_bytes);
                ad[s2] = desc(a, 1024);
                bd[s2] = desc(a + 16384, 1024);
            }
            const uint32_t bar1 = smem_u32(done + 1);
            const long long t0 = clock64();
            if (p.handshake == 3) {
                // real handshake, but the barrier of stage s+1 is polled BEFORE the MMAs of stage s are issued (result consumed after)
                uint32_t phase = 0;
                mbar_wait(full + 0, 0);
                for (int kb = 0; kb < p.kblocks; kb += 4) {
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) {
                        const int nx = (s2 + 1) & 3;
                        const uint32_t nphase = (s2 == 3) ? (phase ^ 1) : phase;
                        uint32_t ok;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(full + nx)), "r"(nphase) : "memory");
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                         ::"r"(tmem), "l"(ad[s2] + 2 * k), "l"(bd[s2] + 2 * k), "r"(p.idesc), "r"(1u) : "memory");
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(empty + s2)) : "memory");
                        if (!ok && (kb + s2 + 1 < p.kblocks)) mbar_wait(full + nx, nphase);
                    }
                    phase ^= 1;
                }
            } else if (p.handshake == 2) {
                // always-satisfied wait (a fresh barrier's "previous" phase, parity 1) + commit to a barrier nobody reads
                for (int kb = 0; kb < p.kblocks; kb += 4) {
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) {
                        mbar_wait(done + 2, 1);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                         ::"r"(tmem), "l"(ad[s2] + 2 * k), "l"(bd[s2] + 2 * k), "r"(p.idesc), "r"(1u) : "memory");
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
                    }
                }
            } else if (p.handshake) {
                uint32_t phase = 0;
                for (int kb = 0; kb < p.kblocks; kb += kRing) {
#pragma unroll
                    for (int s3 = 0; s3 < kRing; ++s3) {
                        const int s2 = s3 & 3;
                        mbar_w<false>(full + s3, phase);
                        if (kFence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int k = 0; k < kMma; ++k)
                            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                         ::"r"(tmem), "l"(ad[s2] + 2 * (k & 3)), "l"(bd[s2] + 2 * (k & 3)), "r"(p.idesc), "r"(1u) : "memory");
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(empty + s3)) : "memory");
                    }
                    phase ^= 1;
                }
            } else
            for (int kb = 0; kb < p.kblocks; kb += 4) {
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                     ::"r"(tmem), "l"(ad[s2] + 2 * k), "l"(bd[s2] + 2 * k), "r"(p.idesc), "r"(1u) : "memory");
                    if (kCommit) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
            mbar_wait(done, 0);
            const long long t1 = clock64();
            p.cycles[blockIdx.x] = t1 - t0;
            *stop = 1;
        }
    } else if ((p.handshake == 1 || p.handshake == 3) && warp == 1) {
        if (lane == 0) {
            uint32_t phase = 0;
            for (int kb = 0; kb < p.kblocks; kb += kRing) {
#pragma unroll
                for (int s2 = 0; s2 < kRing; ++s2) {
                    mbar_w<false>(empty + s2, phase ^ 1);
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + s2)) : "memory");
                }
                phase ^= 1;
            }
        }
    } else if (p.traffic > 0 && warp <= p.traffic) {
        // streaming shared-memory traffic: each warp copies 16-byte vectors scratch -> registers -> scratch (1 load + 1 store wavefront each)
        float4 acc = make_float4(0, 0, 0, 0);
        float4* sc = reinterpret_cast<float4*>(scratch);
        int it = 0;
        while (!*stop) {
#pragma unroll 8
            for (int j = 0; j < 8; ++j) {
                const int idx = ((it * 8 + j) * 32 + lane + warp * 256) & 2047;
                float4 v = sc[idx];
                acc.x += v.x;
                sc[(idx + 1024) & 2047] = acc;
            }
            ++it;
        }
        if (acc.x == 123.456f) p.sink[0] = acc.x;
        if (lane == 0) p.sink[1 + blockIdx.x * 8 + warp - 1] = (float)it;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* dc; float* sink;
    cudaMalloc(&dc, 148 * 8); cudaMalloc(&sink, 4 * (2 + 148 * 8));
    const int stages = 4;
    const int smem = stages * (16384 + 32768) + 32768 + 1024 + 256;
    auto launch = [&](int mode, const P& p) {
#define L(C, S, R, M) { cudaFuncSetAttribute(probe<C, S, R, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); probe<C, S, R, M><<<148, 288, smem>>>(p); }
        switch (mode) {
            case 0: L(false, false, 4, 4); break;      // no commits
            case 7: L(true, false, 4, 4); break;       // handshake with the next stage's barrier polled ahead of this stage's MMAs
            case 6: L(true, false, 4, 4); break;       // always-true wait + commit, no producer
            case 2: L(true, true, 4, 4); break;        // handshake, fence after each full wait
            case 3: L(true, false, 4, 4); break;       // handshake, no fence
            case 4: L(true, false, 8, 4); break;       // no fence, 8 stages
            case 5: L(true, false, 4, 2); break;       // no fence, 2 MMAs per stage
        }
#undef L
    };
    std::vector<long long> h(148);
    std::vector<float> hs(2 + 148 * 8);
    for (int n : {64, 128, 256}) {
        for (int ce : {0, 3, 7}) {
            const int traffic = 0;
            P p; memset(&p, 0, sizeof(p));
            p.commit_every = ce; p.handshake = (ce == 6) ? 2 : (ce == 7) ? 3 : (ce >= 2);
            p.cycles = dc; p.n = n; p.kblocks = 720; p.traffic = traffic; p.stages = stages; p.sink = sink;
            p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((unsigned)n >> 3) << 17) | ((128u >> 4) << 24);
            cudaMemset(sink, 0, 4 * (2 + 148 * 8));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            launch(ce, p);           // warm-up
            cudaEventRecord(e0);
            launch(ce, p);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("n=%d traffic=%d: CUDA error %s\n", n, traffic, cudaGetErrorString(e)); return 2; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(hs.data(), sink, 4 * (2 + 148 * 8), cudaMemcpyDeviceToHost);
            double avg = 0; for (long long c : h) avg += (double)c; avg /= 148;
            const double mmas = (ce == 5 ? 2.0 : 4.0) * p.kblocks;
            double its = 0; for (int w = 0; w < traffic; ++w) its += hs[1 + w];
            // per traffic iteration: 8 x (32 lanes x 16 B load + 16 B store) = 8 KB of shared-memory traffic per warp
            const double lsu_bytes_per_clk = its * 8192.0 / avg;
            const double tflops = 148.0 * mmas * 2.0 * 128 * n * 16 / (ms * 1e-3) / 1e12;
            printf("N=%3d mode=%d (0 free; handshake 2: fence, 3: no fence, 4: no fence x8 stages, 5: no fence 2 MMA/stage, 6: always-true wait + commit) : %.1f cycles/MMA (ideal %d) | smem operand read %.0f B/clk + LSU %.0f B/clk | %.0f TFLOP/s chip, %.3f ms, %.2f GHz\n",
                   n, ce, avg / mmas, n / 2, (128.0 * 32 + n * 32.0) / (avg / mmas), lsu_bytes_per_clk, tflops, ms, avg / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
