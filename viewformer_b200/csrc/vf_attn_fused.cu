// Fused block-causal attention on tcgen05 tensor cores (sm_100a):  O = softmax(mask(Q K^T)) V  per (batch, head).
//
// Replaces viewformer/models/branching_attention.py:41-61 (compute_causal_block_attention: a view attends to all tokens
// of its own and of every earlier view; logits are NOT scaled by 1/sqrt(dh); masked logits are -1e4 in the reference,
// whose exp underflows to exactly 0 in fp32, so masked keys are simply skipped here) for the single-stream forward.
//
// One CTA = 128 queries (two 64-token views) of one (batch, head).  S = Q K^T (128x128, fp32) lives in TMEM
// (double-buffered), the softmax warps own one query row per thread (no shuffles), P is written as bf16 into a
// 128B-swizzled K-major smem tile and fed back as the A operand of O += P V (O: 128x64 fp32 in TMEM).
// Two passes over the visible key tiles avoid any accumulator rescaling: pass 1 finds the exact row maxima
// (QK^T only), pass 2 recomputes S, exponentiates against the final maximum and accumulates P V and the row sums.
// Fully masked key tiles are never loaded.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 softmax / epilogue.
#include "vf_common.cuh"
#include <cuda.h>

namespace {

constexpr int QT = 128;            // queries per CTA
constexpr int KT = 128;            // keys per tile
constexpr int DH = 64;             // head dim (one 128-byte swizzle row)
constexpr int KSTAGES = 3;         // K tile ring
constexpr int VSTAGES = 2;         // V^T tile ring
constexpr int Q_BYTES = QT * 128;              // 16 KB
constexpr int K_BYTES = KT * 128;              // 16 KB
constexpr int V_BYTES = 2 * DH * 128;          // two [64 dh rows x 64 keys] atoms = 16 KB
constexpr int P_BYTES = 2 * QT * 128;          // two [128 q rows x 64 keys] atoms = 32 KB
constexpr int ATTN_THREADS = 192;
constexpr int TMEM_COLS = 512;                 // S0 [0,128) S1 [128,256) O [256,320)

struct AttnParams {
    CUtensorMap tmQ, tmK, tmV;
    int S, H, d, block, n_qtiles, qt0;      // query tiles qt0 .. qt0 + n_qtiles - 1 are computed (qt0 > 0: KV-cache query mode)
    __nv_bfloat16* out;
    unsigned idesc_s, idesc_o;
};

// one elected lane; the compiler knows a single thread is active in the guarded region (plain R2UR for tcgen05 / TMA operands)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return;
    printf("vf_attn: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(ATTN_THREADS, 1) attn_block_causal_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;
    uint8_t* sV = sK + KSTAGES * K_BYTES;
    uint8_t* sP = sV + VSTAGES * V_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;             // KSTAGES
    uint64_t* k_empty = k_full + KSTAGES;    // KSTAGES
    uint64_t* v_full = k_empty + KSTAGES;    // VSTAGES
    uint64_t* v_empty = v_full + VSTAGES;    // VSTAGES
    uint64_t* s_full = v_empty + VSTAGES;    // 2
    uint64_t* s_empty = s_full + 2;          // 2
    uint64_t* p_full = s_empty + 2;          // 2
    uint64_t* p_empty = p_full + 2;          // 2
    uint64_t* o_full = p_empty + 2;          // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = p.qt0 + blockIdx.x % p.n_qtiles;
    const int bh = blockIdx.x / p.n_qtiles;
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = qt * QT;
    // keys visible to the tile's last valid query: views <= view(last query)
    const int last_q = min(q0 + QT, p.S) - 1;
    const int kv_lim = min(p.S, (last_q / p.block + 1) * p.block);
    const int n_kt = (kv_lim + KT - 1) / KT;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmV)) : "memory");
    }
    if (threadIdx.x == 32) {
        mbar_init(q_full, 1);
        for (int i = 0; i < KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);       // one arrive per softmax warp
            mbar_init(&p_full[i], 4);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_o = tmem + 256;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(q_full, Q_BYTES);
            tma_load_4d(sQ, &p.tmQ, q_full, 0, q0, h, b);
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int j = 0; j < n_kt; ++j) {
                    mbar_wait(&k_empty[ks], kph ^ 1);
                    mbar_expect_tx(&k_full[ks], K_BYTES);
                    tma_load_4d(sK + ks * K_BYTES, &p.tmK, &k_full[ks], 0, j * KT, h, b);
                    if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
                    if (pass == 1) {
                        mbar_wait(&v_empty[vs], vph ^ 1);
                        mbar_expect_tx(&v_full[vs], V_BYTES);
                        uint8_t* dst = sV + vs * V_BYTES;
                        tma_load_4d(dst, &p.tmV, &v_full[vs], j * KT, h * DH, b, 0);                    // keys [0,64) of the tile
                        tma_load_4d(dst + DH * 128, &p.tmV, &v_full[vs], j * KT + 64, h * DH, b, 0);    // keys [64,128)
                        if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            mbar_wait(q_full, 0);
            tc_fence_after();
            const uint64_t qdesc = sw128_desc(smem_u32(sQ));
            int ks = 0, vs = 0, sb = 0, pb = 0;
            uint32_t kph = 0, vph = 0, sph = 0, pph = 0;
            auto issue_s = [&]() {       // S[sb] = Q K^T for the next K tile in the ring
                mbar_wait(&k_full[ks], kph);
                mbar_wait(&s_empty[sb], sph ^ 1);
                tc_fence_after();
                const uint64_t kdesc = sw128_desc(smem_u32(sK + ks * K_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem + sb * 128, qdesc + 2 * k, kdesc + 2 * k, p.idesc_s, k > 0);
                tc_commit(&k_empty[ks]);
                tc_commit(&s_full[sb]);
                if (++ks == KSTAGES) { ks = 0; kph ^= 1; }
                if (++sb == 2) { sb = 0; sph ^= 1; }
            };
            // pass 1: row maxima only
            for (int j = 0; j < n_kt; ++j) issue_s();
            // pass 2: S_{j+1} is issued before P_j V_j so the softmax of tile j+1 overlaps the PV MMAs of tile j
            issue_s();
            for (int j = 0; j < n_kt; ++j) {
                if (j + 1 < n_kt) issue_s();
                mbar_wait(&p_full[pb], pph);
                mbar_wait(&v_full[vs], vph);
                tc_fence_after();
                const uint32_t pa = smem_u32(sP + pb * P_BYTES), va = smem_u32(sV + vs * V_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {      // K = 128 keys = 2 atoms x 4 steps of 16
                    const uint64_t adesc = sw128_desc(pa + (k >> 2) * (QT * 128)) + 2 * (k & 3);
                    const uint64_t bdesc = sw128_desc(va + (k >> 2) * (DH * 128)) + 2 * (k & 3);
                    umma_bf16(tmem_o, adesc, bdesc, p.idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit(&p_empty[pb]);
                tc_commit(&v_empty[vs]);
                if (++pb == 2) { pb = 0; pph ^= 1; }
                if (++vs == VSTAGES) { vs = 0; vph ^= 1; }
            }
            tc_commit(o_full);
        }
    } else {
        // ===================== softmax / epilogue: thread = query row =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int qpos = q0 + row;
        const int vis = min(p.S, (min(qpos, p.S - 1) / p.block + 1) * p.block);     // keys [0, vis) are visible to this row
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int sb = 0, pb = 0;
        uint32_t sph = 0, pph = 0;
        float m = -INFINITY;
        // ---- pass 1: exact row maximum over the visible keys
        for (int j = 0; j < n_kt; ++j) {
            mbar_wait(&s_full[sb], sph);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < KT; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + lane_base + sb * 128 + c0, r);
                const int kbase = j * KT + c0;
                if (kbase + 32 <= vis) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (kbase + i < vis) m = fmaxf(m, __uint_as_float(r[i]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            if (++sb == 2) { sb = 0; sph ^= 1; }
        }
        // ---- pass 2: P = exp(S - m) (bf16, swizzled K-major smem tile), row sums
        float l = 0.f;
        for (int j = 0; j < n_kt; ++j) {
            mbar_wait(&s_full[sb], sph);
            tc_fence_after();
            mbar_wait(&p_empty[pb], pph ^ 1);
            uint8_t* pt = sP + pb * P_BYTES;
#pragma unroll 1
            for (int c0 = 0; c0 < KT; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + lane_base + sb * 128 + c0, r);
                const int kbase = j * KT + c0;
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0 = (kbase + i < vis) ? __expf(__uint_as_float(r[i]) - m) : 0.f;
                    float e1 = (kbase + i + 1 < vis) ? __expf(__uint_as_float(r[i + 1]) - m) : 0.f;
                    l += e0 + e1;
                    __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
                    packed[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
                }
                // 32 keys = 64 bytes = four 16-byte chunks of this row; 128B swizzle: chunk' = chunk ^ (row & 7)
                const int atom = c0 >> 6;                    // which 64-key atom
                const int chunk0 = (c0 & 63) >> 3;           // first 16-byte chunk inside the atom's 128-byte row
                uint8_t* rowp = pt + atom * (QT * 128) + row * 128;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int phys = (chunk0 + cc) ^ (row & 7);
                    *reinterpret_cast<uint4*>(rowp + phys * 16) =
                        make_uint4(packed[cc * 4], packed[cc * 4 + 1], packed[cc * 4 + 2], packed[cc * 4 + 3]);
                }
            }
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[sb]);
                mbar_arrive(&p_full[pb]);
            }
            if (++sb == 2) { sb = 0; sph ^= 1; }
            if (++pb == 2) { pb = 0; pph ^= 1; }
        }
        // ---- epilogue: O / l -> bf16
        mbar_wait(o_full, 0);
        tc_fence_after();
        const float inv = 1.0f / l;
        __nv_bfloat16* orow = p.out + ((long long)b * p.S + qpos) * p.d + h * DH;
#pragma unroll 1
        for (int c0 = 0; c0 < DH; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_o + lane_base + c0, r);
            if (qpos < p.S) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 u;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                    u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
                    u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
                    *reinterpret_cast<uint4*>(orow + c0 + i) = u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
int tmap_bf16(CUtensorMap* tm, const void* base, const uint64_t dims[4], const uint64_t strides[3], const uint32_t box[4]) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { vf_set_error("vf_attn: cuTensorMapEncodeTiled unavailable"); return VF_ERR_CUDA; }
    cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vf_set_error("vf_attn: cuTensorMapEncodeTiled failed (%d)", (int)r); return VF_ERR_CUDA; }
    return VF_OK;
}
unsigned idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24); }

}  // namespace

extern "C" int vf_attn_block_causal_tail(const void* qk, const void* vt, int B, int S, int H, int d, int block, int first_query, void* out,
                                         vf_stream_t s);
extern "C" int vf_attn_block_causal(const void* qk, const void* vt, int B, int S, int H, int d, int block, void* out, vf_stream_t s) {
    return vf_attn_block_causal_tail(qk, vt, B, S, H, d, block, 0, out, s);
}

// Only the query rows >= first_query (rounded down to a 128-row tile) are computed: with the context's q|k rows and V^T columns kept from
// a prefill and the query view appended behind them, this is the KV-cache decode step (BASELINE config 5) — the same fused kernel, no
// score matrix in HBM.  Rows of `out` below the first computed tile are left untouched.
extern "C" int vf_attn_block_causal_tail(const void* qk, const void* vt, int B, int S, int H, int d, int block, int first_query, void* out,
                                         vf_stream_t s) {
    VF_CHECK_ARG(qk && vt && out, "vf_attn_block_causal: null pointer");
    VF_CHECK_ARG(first_query >= 0 && first_query < S, "vf_attn_block_causal: first_query out of range");
    VF_CHECK_ARG(H > 0 && d == H * DH, "vf_attn_block_causal: head dim must be 64 (d=%d H=%d)", d, H);
    VF_CHECK_ARG(block > 0 && S % block == 0 && S % 8 == 0, "vf_attn_block_causal: S=%d must be a multiple of block=%d and of 8", S, block);
    if (B == 0 || S == 0) return VF_OK;
    AttnParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.S = S; prm.H = H; prm.d = d; prm.block = block;
    prm.qt0 = first_query / QT;
    prm.n_qtiles = (S + QT - 1) / QT - prm.qt0;
    prm.out = reinterpret_cast<__nv_bfloat16*>(out);
    prm.idesc_s = idesc_bf16(128, 128);
    prm.idesc_o = idesc_bf16(128, 64);
    int rc;
    const uint64_t row = (uint64_t)2 * d * 2;                  // bytes per qk row
    {   // Q / K: [B, S, 2d] viewed as (dh, S, H, B); K is the second half of every row
        const uint64_t dims[4] = {(uint64_t)DH, (uint64_t)S, (uint64_t)H, (uint64_t)B};
        const uint64_t str[3] = {row, (uint64_t)DH * 2, row * S};
        const uint32_t box[4] = {(uint32_t)DH, (uint32_t)QT, 1, 1};
        if ((rc = tmap_bf16(&prm.tmQ, qk, dims, str, box)) != VF_OK) return rc;
        if ((rc = tmap_bf16(&prm.tmK, reinterpret_cast<const __nv_bfloat16*>(qk) + d, dims, str, box)) != VF_OK) return rc;
    }
    {   // V^T: [B, d, S] viewed as (S, d, B, 1); one box = 64 keys x 64 dh rows
        const uint64_t dims[4] = {(uint64_t)S, (uint64_t)d, (uint64_t)B, 1};
        const uint64_t str[3] = {(uint64_t)S * 2, (uint64_t)S * 2 * d, (uint64_t)S * 2 * d * B};
        const uint32_t box[4] = {64, (uint32_t)DH, 1, 1};
        if ((rc = tmap_bf16(&prm.tmV, vt, dims, str, box)) != VF_OK) return rc;
    }
    constexpr int smem = Q_BYTES + KSTAGES * K_BYTES + VSTAGES * V_BYTES + 2 * P_BYTES + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_block_causal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { vf_set_error("vf_attn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VF_ERR_CUDA; }
        configured = true;
    }
    const long long ctas = (long long)B * H * prm.n_qtiles;
    VF_CHECK_ARG(ctas < (1ll << 31), "vf_attn_block_causal: grid too large");
    attn_block_causal_kernel<<<(unsigned)ctas, ATTN_THREADS, smem, vf_s(s)>>>(prm);
    VF_CHECK_LAUNCH("vf_attn_block_causal");
    return VF_OK;
}
