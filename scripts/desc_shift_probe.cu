// Hardware probe (not product code): does a K-major SWIZZLE_128B UMMA descriptor accept a start address that is shifted by
// whole 128-byte rows inside a TMA-written tile, with SBO = 2048 B (row pitch 16) and which base_offset does it need?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o desc_shift_probe scripts/desc_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    for (uint32_t i = 0; i < (1u << 22) && !ok; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    if (!ok) __trap();
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

struct P { CUtensorMap tmA, tmB; float* out; unsigned idesc; int shift_rows, sbo, base_off, a_rows; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ P p) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                       // 288 rows x 128 B = 36 KB
    uint8_t* sB = sm + 40960;               // 128 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 40960 + 16384);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, p.a_rows * 128 + 128 * 128);
        tma4(sA, &p.tmA, bar, 0, 0, 0, 0);
        tma4(sB, &p.tmB, bar, 0, 0, 0, 0);
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA) + p.shift_rows * 128;
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = desc(a0, p.sbo, p.base_off) + 2 * k, bd = desc(smem_u32(sB), 1024, 0) + 2 * k;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(p.idesc), "r"(k > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    const int RA = 288, RB = 128, K = 64;
    std::vector<__nv_bfloat16> hA(RA * K), hB(RB * K);
    std::vector<float> fA(RA * K), fB(RB * K);
    srand(1);
    for (int i = 0; i < RA * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < RB * K; ++i) { float v = (rand() % 13 - 6) / 8.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, RA * K * 2); cudaMalloc(&dB, RB * K * 2); cudaMalloc(&dO, 128 * 128 * 4);
    cudaMemcpy(dA, hA.data(), RA * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), RB * K * 2, cudaMemcpyHostToDevice);
    P p; memset(&p, 0, sizeof(p));
    CUtensorMap tmA16, tmA10;
    {   // A as an NHWC tensor [1, 18, pitch, 64]: box (64, pitch, 18, 1) -> 18*pitch smem rows, row = y*pitch + x
        cuuint32_t e[4] = {1, 1, 1, 1};
        for (int pitch : {16, 10}) {
            cuuint64_t d[4] = {64, (cuuint64_t)pitch, 18, 1}, s[3] = {128, (cuuint64_t)128 * pitch, (cuuint64_t)128 * pitch * 18};
            cuuint32_t b[4] = {64, (cuuint32_t)pitch, 18, 1};
            CUresult r = enc(pitch == 16 ? &tmA16 : &tmA10, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode A failed %d\n", (int)r); return 1; }
        }
        CUresult r;
        cuuint64_t d2[4] = {64, 128, 1, 1}, s2[3] = {128, 128 * 128, 128 * 128}; cuuint32_t b2[4] = {64, 128, 1, 1};
        r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dB, d2, s2, b2, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    p.out = dO;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const int smem = 40960 + 16384 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hO(128 * 128);
    // expected: output row r = (g = r/8, i = r%8) reads smem row shift + g*pitch + i, pitch = sbo/128
    for (int sbo : {1024, 2048, 1280}) {
        p.tmA = (sbo == 1280) ? tmA10 : tmA16;
        for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) {
            const int pitch = sbo / 128, shift = dy * pitch + dx;
            if (shift + 15 * pitch + 8 > ((sbo == 1280) ? 180 : RA)) continue;
            for (int bo : {0, shift & 7}) {
                p.shift_rows = shift; p.sbo = sbo; p.base_off = bo; p.a_rows = (sbo == 1280) ? 180 : 288;
                probe<<<1, 128, smem>>>(p);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sbo=%d dy=%d dx=%d base_off=%d: CUDA error %s\n", sbo, dy, dx, bo, cudaGetErrorString(e)); return 2; }
                cudaMemcpy(hO.data(), dO, 128 * 128 * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0;
                for (int r = 0; r < 128; ++r) for (int n = 0; n < 128; ++n) {
                    const int ar = shift + (r / 8) * pitch + (r % 8);
                    double acc = 0; for (int k = 0; k < K; ++k) acc += (double)fA[ar * K + k] * fB[n * K + k];
                    double err = fabs(acc - hO[r * 128 + n]); if (err > maxerr) maxerr = err;
                }
                printf("sbo=%4d dy=%d dx=%d shift_rows=%2d base_off=%d : max_err=%.4f %s\n", sbo, dy, dx, shift, bo, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
                if ((shift & 7) == 0) break;
            }
        }
    }
    return 0;
}
