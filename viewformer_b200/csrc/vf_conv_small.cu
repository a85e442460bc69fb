// Exact fp32 3x3 convolutions for the two layers whose channel counts are too small for a tensor-core tile:
//   conv_in   3 -> 128 @128x128  (viewformer/models/vqgan_th.py:159-163): output-write bound
//   conv_out  128 -> 3 @128x128  (vqgan_th.py:285-289): input-read bound
// Both are stride-1, pad-1, NHWC.
#include "vf_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Small Cin (<= 4): block = 32 lanes (pixels along x) x (Cout/16) warps; each thread computes 2 pixels
// (x, x+32) x 16 output channels.  Weights [9*Cin][Cout] live in shared memory and are read as broadcast float4.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ w_kn,
                                                                  const float* __restrict__ bias, int N, int H, int W, int Cout,
                                                                  float* __restrict__ y) {
    extern __shared__ float ws[];                       // [9*CIN][Cout]
    constexpr int K = 9 * CIN;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) ws[i] = w_kn[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;      // g: group of 16 output channels
    const int xb = blockIdx.x * 64;
    const int yy = blockIdx.y, n = blockIdx.z;
    float in[2][K];
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        const int xx = xb + lane + px * 32;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
            for (int c = 0; c < CIN; ++c) in[px][t * CIN + c] = ok ? __ldg(x + (((int64_t)n * H + iy) * W + ix) * CIN + c) : 0.f;
        }
    }
    float acc[2][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float b = bias ? __ldg(bias + g * 16 + j) : 0.f;
        acc[0][j] = b;
        acc[1][j] = b;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(&ws[k * Cout + g * 16 + j4 * 4]);
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                acc[px][j4 * 4 + 0] = fmaf(in[px][k], w4.x, acc[px][j4 * 4 + 0]);
                acc[px][j4 * 4 + 1] = fmaf(in[px][k], w4.y, acc[px][j4 * 4 + 1]);
                acc[px][j4 * 4 + 2] = fmaf(in[px][k], w4.z, acc[px][j4 * 4 + 2]);
                acc[px][j4 * 4 + 3] = fmaf(in[px][k], w4.w, acc[px][j4 * 4 + 3]);
            }
        }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        const int xx = xb + lane + px * 32;
        if (xx >= W) continue;
        float* o = y + (((int64_t)n * H + yy) * W + xx) * Cout + g * 16;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(o + j4 * 4) = make_float4(acc[px][j4 * 4], acc[px][j4 * 4 + 1], acc[px][j4 * 4 + 2], acc[px][j4 * 4 + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
// Cin = 3, Cout = 128: one warp walks a run of pixels; lane l owns output channels 4l..4l+3 and keeps its 27 x 4 weights in
// registers; the 27 inputs of a pixel are warp-uniform (broadcast loads); every store is one coalesced 512-byte row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv3x3_cin3_cout128_kernel(const float* __restrict__ x, const float* __restrict__ w_kn,
                                                                   const float* __restrict__ bias, int N, int H, int W, int px_per_warp,
                                                                   float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int runs_per_row = (W + px_per_warp - 1) / px_per_warp;
    const int run = (int)(warp_global % runs_per_row);
    const long long row = warp_global / runs_per_row;           // n*H + y
    if (row >= (long long)N * H) return;
    const int yy = (int)(row % H), n = (int)(row / H);
    float4 wr[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w_kn + k * 128) + lane);
    const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int x0 = run * px_per_warp, x1 = min(W, x0 + px_per_warp);
    for (int xx = x0; xx < x1; xx += 2) {                       // two pixels per iteration for ILP
        float4 a0 = b4, a1 = b4;
        const bool two = xx + 1 < x1;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
            const bool oky = iy >= 0 && iy < H;
            const float* p0 = x + (((long long)n * H + iy) * W + ix) * 3;
            const bool ok0 = oky && ix >= 0 && ix < W, ok1 = oky && two && ix + 1 >= 0 && ix + 1 < W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v0 = ok0 ? __ldg(p0 + c) : 0.f;      // warp-uniform address: one broadcast transaction
                const float v1 = ok1 ? __ldg(p0 + 3 + c) : 0.f;
                const float4 w4 = wr[t * 3 + c];
                a0.x = fmaf(v0, w4.x, a0.x); a0.y = fmaf(v0, w4.y, a0.y); a0.z = fmaf(v0, w4.z, a0.z); a0.w = fmaf(v0, w4.w, a0.w);
                a1.x = fmaf(v1, w4.x, a1.x); a1.y = fmaf(v1, w4.y, a1.y); a1.z = fmaf(v1, w4.z, a1.z); a1.w = fmaf(v1, w4.w, a1.w);
            }
        }
        float* o = y + (((long long)n * H + yy) * W + xx) * 128 + lane * 4;
        *reinterpret_cast<float4*>(o) = a0;
        if (two) *reinterpret_cast<float4*>(o + 128) = a1;
    }
}

// ------------------------------------------------------------------------------------------------
// Small Cout (<= 4), Cin == 128: one warp walks a run of pixels along x; lane l owns input channels 4l..4l+3 and keeps
// its 9 x 4 x COUT weights in registers; per pixel 9 coalesced 512-byte loads, COUT warp reductions.
// ------------------------------------------------------------------------------------------------
template <typename InT> __device__ __forceinline__ float4 ld4(const InT* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <int COUT, typename InT>
__global__ void __launch_bounds__(256) conv3x3_small_cout_kernel(const InT* __restrict__ x, const float* __restrict__ w_kn,
                                                                 const float* __restrict__ bias, int N, int H, int W,
                                                                 int px_per_warp, float* __restrict__ y) {
    constexpr int CIN = 128;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int runs_per_row = (W + px_per_warp - 1) / px_per_warp;
    const int run = warp_global % runs_per_row;
    const int row = warp_global / runs_per_row;           // n*H + y
    if (row >= N * H) return;
    const int yy = row % H, n = row / H;
    float wr[9][4][COUT];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int o = 0; o < COUT; ++o) wr[t][c][o] = __ldg(w_kn + (int64_t)(t * CIN + lane * 4 + c) * COUT + o);
    float bs[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) bs[o] = bias ? __ldg(bias + o) : 0.f;
    const int x0 = run * px_per_warp, x1 = min(W, x0 + px_per_warp);
    for (int xx = x0; xx < x1; ++xx) {
        float acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;      // warp-uniform
            const float4 v = ld4<InT>(x + (((int64_t)n * H + iy) * W + ix) * CIN + lane * 4);
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                acc[o] = fmaf(v.x, wr[t][0][o], acc[o]);
                acc[o] = fmaf(v.y, wr[t][1][o], acc[o]);
                acc[o] = fmaf(v.z, wr[t][2][o], acc[o]);
                acc[o] = fmaf(v.w, wr[t][3][o], acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = warp_sum(acc[o]);
        if (lane < COUT) {
            float r = acc[0], b = bs[0];
#pragma unroll
            for (int o = 1; o < COUT; ++o) {
                r = (lane == o) ? acc[o] : r;
                b = (lane == o) ? bs[o] : b;
            }
            y[(((int64_t)n * H + yy) * W + xx) * COUT + lane] = r + b;
        }
    }
}

}  // namespace

extern "C" int vf_conv3x3_small_cin(const float* x, const float* w_kn, const float* bias, int N, int H, int W, int Cin, int Cout,
                                    float* y, vf_stream_t s) {
    VF_CHECK_ARG(x && w_kn && y, "vf_conv3x3_small_cin: null pointer");
    VF_CHECK_ARG(Cin == 3 && Cout % 16 == 0 && Cout <= 128, "vf_conv3x3_small_cin: supports Cin=3, Cout%%16==0, Cout<=128 (got %d->%d)", Cin, Cout);
    if (N == 0) return VF_OK;
    if (false && Cout == 128) {   // measured 2.75 ms vs 1.6 ms for the smem-weight kernel below at 288x128x128: latency-bound, kept for reference
        const int px_per_warp = W >= 64 ? 64 : W;
        const long long warps = (long long)N * H * ((W + px_per_warp - 1) / px_per_warp);
        conv3x3_cin3_cout128_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, vf_s(s)>>>(x, w_kn, bias, N, H, W, px_per_warp, y);
        VF_CHECK_LAUNCH("vf_conv3x3_small_cin");
        return VF_OK;
    }
    dim3 grid((W + 63) / 64, H, N);
    VF_CHECK_ARG(H <= 65535 && N <= 65535, "vf_conv3x3_small_cin: grid too large");
    conv3x3_small_cin_kernel<3><<<grid, 32 * (Cout / 16), sizeof(float) * 27 * Cout, vf_s(s)>>>(x, w_kn, bias, N, H, W, Cout, y);
    VF_CHECK_LAUNCH("vf_conv3x3_small_cin");
    return VF_OK;
}

extern "C" int vf_conv3x3_small_cout(const void* x, int x_dtype, const float* w_kn, const float* bias, int N, int H, int W, int Cin,
                                     int Cout, float* y, vf_stream_t s) {
    VF_CHECK_ARG(x && w_kn && y, "vf_conv3x3_small_cout: null pointer");
    VF_CHECK_ARG(Cin == 128 && Cout == 3, "vf_conv3x3_small_cout: supports 128->3 (got %d->%d)", Cin, Cout);
    if (N == 0) return VF_OK;
    const int px_per_warp = W >= 32 ? 32 : W;
    const long long warps = (long long)N * H * ((W + px_per_warp - 1) / px_per_warp);
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    if (x_dtype == VF_F32)
        conv3x3_small_cout_kernel<3, float><<<blocks, 256, 0, vf_s(s)>>>((const float*)x, w_kn, bias, N, H, W, px_per_warp, y);
    else
        conv3x3_small_cout_kernel<3, __nv_bfloat16><<<blocks, 256, 0, vf_s(s)>>>((const __nv_bfloat16*)x, w_kn, bias, N, H, W,
                                                                                px_per_warp, y);
    VF_CHECK_LAUNCH("vf_conv3x3_small_cout");
    return VF_OK;
}
