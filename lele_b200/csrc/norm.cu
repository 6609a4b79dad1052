// norm.cu -- LayerNorm / Softmax / BatchNorm / RMSNorm (src/kernels/norm.rs + avx/norm.rs).
// HBM-bound row kernels: one warp per row, 128-bit loads, warp-shuffle reductions.
#include "common.cuh"

// LayerNorm (norm.rs:226 -> avx/norm.rs:10-133): mean = sum*(1/n); var = sumsq*(1/n) - mean^2 (not
// clamped); inv = 1/sqrt(var+eps); y = fma((x-mean)*inv, gamma, beta) on the SIMD body (first
// n/8*8 columns), plain mul+add on the scalar tail.  Sum / sum-of-squares follow the AVX2
// accumulator order exactly (lb_avx_order_reduce), so the result is bit-identical to the x86
// reference arithmetic.  One warp per row, coalesced 128-byte warp loads (element j -> lane j%32);
// the second pass re-reads the row from L1/L2.  Optional fused per-slice min/max of the output
// (order-preserving keys) feeds the dynamic quantiser of the next int8 linear.
// kRegs > 0: the row's full 32-element blocks live in registers (n <= 32*kRegs): one global read.
template <int kRegs>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  long long outer, int n, float eps, float* __restrict__ out,
                  unsigned* __restrict__ minmax_keys /*[n_slices][2] or NULL*/, int rows_per_slice) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* xr = x + row * n;
    float* o = out + row * n;
    const float inv_n = __fdiv_rn(1.0f, (float)n);
    const int simd_end = (n / 8) * 8, n32 = n & ~31;
    float v[kRegs > 0 ? kRegs : 1];
    float ps = 0.0f, pq = 0.0f;
    if (kRegs > 0) {
#pragma unroll
        for (int i = 0; i < kRegs; ++i) {
            const int j = lane + 32 * i;
            v[i] = j < n32 ? __ldg(xr + j) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < kRegs; ++i)
            if (lane + 32 * i < n32) { ps = __fadd_rn(ps, v[i]); pq = __fmaf_rn(v[i], v[i], pq); }
    } else {
        for (int j = lane; j < n32; j += 32) { float t = xr[j]; ps = __fadd_rn(ps, t); pq = __fmaf_rn(t, t, pq); }
    }
    const float s = lb_avx_order_finish(ps, n, lane, [&](float acc, int j) { return __fadd_rn(acc, xr[j]); },
                                        [&](float acc, int j) { return __fadd_rn(acc, xr[j]); });
    const float sq = lb_avx_order_finish(pq, n, lane, [&](float acc, int j) { float t = xr[j]; return __fmaf_rn(t, t, acc); },
                                         [&](float acc, int j) { float t = xr[j]; return __fadd_rn(acc, __fmul_rn(t, t)); });
    const float mean = __fmul_rn(s, inv_n);
    const float var = __fsub_rn(__fmul_rn(sq, inv_n), __fmul_rn(mean, mean));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
    float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
    auto emit = [&](int j, float xv) {
        float g = gamma ? __ldg(gamma + j) : 1.0f, b = beta ? __ldg(beta + j) : 0.0f;
        float sc = __fmul_rn(__fsub_rn(xv, mean), inv);
        float r = j < simd_end ? __fmaf_rn(sc, g, b) : __fadd_rn(__fmul_rn(sc, g), b);
        vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
        o[j] = r;
    };
    if (kRegs > 0) {
#pragma unroll
        for (int i = 0; i < kRegs; ++i) { const int j = lane + 32 * i; if (j < n32) emit(j, v[i]); }
        for (int j = n32 + lane; j < n; j += 32) emit(j, xr[j]);
    } else {
        for (int j = lane; j < n; j += 32) emit(j, xr[j]);
    }
    if (minmax_keys) {
        vmin = lb_warp_min(vmin); vmax = lb_warp_max(vmax);
        if (lane == 0) {
            lb_mm_update(minmax_keys, row / rows_per_slice, vmin, vmax);
        }
    }
}

// Specialisation for the row widths of the SenseVoice encoder (N = 512, 560; N % 8 == 0): compile-time
// trip counts, gamma/beta held in registers and reused over RPW consecutive rows per warp, no per-element
// predication.  Same arithmetic and summation order as layer_norm_kernel.
template <int N, int RPW>
__global__ void __launch_bounds__(256)
layer_norm_fixed_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        long long outer, float eps, float* __restrict__ out, unsigned* __restrict__ minmax_keys, int rows_per_slice) {
    constexpr int NB = N / 32, REM = N % 32, REM8 = REM / 8;
    static_assert(N % 8 == 0, "fixed LayerNorm kernel needs N % 8 == 0");
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (row0 >= outer) return;
    float g[NB], bt[NB], gr = 1.0f, br = 0.0f;
#pragma unroll
    for (int i = 0; i < NB; ++i) { g[i] = __ldg(gamma + 32 * i + lane); bt[i] = __ldg(beta + 32 * i + lane); }
    if (REM > 0 && lane < REM) { gr = __ldg(gamma + NB * 32 + lane); br = __ldg(beta + NB * 32 + lane); }
    const float inv_n = __fdiv_rn(1.0f, (float)N);
    const long long slice_a = minmax_keys ? row0 / rows_per_slice : 0;
    float mnA = 3.402823466e+38f, mxA = -3.402823466e+38f, mnB = 3.402823466e+38f, mxB = -3.402823466e+38f;
#pragma unroll 1
    for (int rr = 0; rr < RPW; ++rr) {
        const long long row = row0 + rr;
        if (row >= outer) break;
        const float* xr = x + row * N;
        float* o = out + row * N;
        float v[NB], rem[REM8 > 0 ? REM8 : 1], xrem = 0.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) v[i] = __ldg(xr + 32 * i + lane);
#pragma unroll
        for (int q = 0; q < REM8; ++q) rem[q] = __ldg(xr + NB * 32 + 8 * q + (lane & 7));
        if (REM > 0 && lane < REM) xrem = __ldg(xr + NB * 32 + lane);
        float ps = 0.0f, pq = 0.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) { ps = __fadd_rn(ps, v[i]); pq = __fmaf_rn(v[i], v[i], pq); }
        ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 8));  pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, 8));
        ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 16)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, 16));
#pragma unroll
        for (int q = 0; q < REM8; ++q) { ps = __fadd_rn(ps, rem[q]); pq = __fmaf_rn(rem[q], rem[q], pq); }
        ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 4)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, 4));
        ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 2)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, 2));
        ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 1)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, 1));
        ps = __shfl_sync(0xffffffffu, ps, 0); pq = __shfl_sync(0xffffffffu, pq, 0);
        const float mean = __fmul_rn(ps, inv_n);
        const float var = __fsub_rn(__fmul_rn(pq, inv_n), __fmul_rn(mean, mean));
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
        float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float r = __fmaf_rn(__fmul_rn(__fsub_rn(v[i], mean), inv), g[i], bt[i]);
            vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
            o[32 * i + lane] = r;
        }
        if (REM > 0 && lane < REM) {
            const float r = __fmaf_rn(__fmul_rn(__fsub_rn(xrem, mean), inv), gr, br);
            vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
            o[NB * 32 + lane] = r;
        }
        if (minmax_keys) {
            if (row / rows_per_slice == slice_a) { mnA = fminf(mnA, vmin); mxA = fmaxf(mxA, vmax); }
            else { mnB = fminf(mnB, vmin); mxB = fmaxf(mxB, vmax); }
        }
    }
    if (minmax_keys) {
        mnA = lb_warp_min(mnA); mxA = lb_warp_max(mxA); mnB = lb_warp_min(mnB); mxB = lb_warp_max(mxB);
        if (lane == 0) {
            if (mnA <= mxA) lb_mm_update(minmax_keys, slice_a, mnA, mxA);
            if (mnB <= mxB) lb_mm_update(minmax_keys, slice_a + 1, mnB, mxB);
        }
    }
}

int lb_layer_norm_minmax(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer,
                         int n, float eps, float* out, unsigned* minmax_keys, int rows_per_slice) {
    if (outer == 0) return LELE_B200_OK;
    const int warps = 8;
    constexpr int RPW = 4;
    if (gamma && beta && (n == 512 || n == 560) && (!minmax_keys || rows_per_slice >= RPW)) {
        const int grid = lb_ceil_div(outer, warps * RPW);
        if (n == 512) layer_norm_fixed_kernel<512, RPW><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, eps, out, minmax_keys, rows_per_slice);
        else layer_norm_fixed_kernel<560, RPW><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, eps, out, minmax_keys, rows_per_slice);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    const int grid = lb_ceil_div(outer, warps);
    if (n <= 32 * 18) layer_norm_kernel<18><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, n, eps, out, minmax_keys, rows_per_slice);
    else if (n <= 32 * 64) layer_norm_kernel<64><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, n, eps, out, minmax_keys, rows_per_slice);
    else layer_norm_kernel<0><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, n, eps, out, minmax_keys, rows_per_slice);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_layer_norm(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta,
                                    long long outer, int n, float eps, float* out) {
    LB_REQUIRE(ctx && x && out && n > 0 && outer >= 0, "layer_norm: bad arguments");
    return lb_layer_norm_minmax(ctx, x, gamma, beta, outer, n, eps, out, nullptr, 1);
}

// Softmax over the last axis (norm.rs:8-224 inner_size==1 -> avx/norm.rs:139-229):
// max, exp(x-max) with the polynomial exp on the SIMD body / libm expf on the n%8 tail,
// multiply by 1/sum.
__global__ void __launch_bounds__(256)
softmax_kernel(const float* __restrict__ x, long long outer, int n, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* xr = x + row * n;
    float* o = out + row * n;
    const int simd_end = (n / 8) * 8;
    float mx = -3.402823466e+38f;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, xr[j]);
    mx = lb_warp_max(mx);
    for (int j = lane; j < n; j += 32) {
        float d = __fsub_rn(xr[j], mx);
        o[j] = j < simd_end ? lb_cephes_expf(d) : expf(d);
    }
    __syncwarp();
    // sum in the AVX2 accumulator order (avx/norm.rs:169-205)
    const float sum = lb_avx_order_reduce(n, lane, [&](float acc, int j) { return __fadd_rn(acc, o[j]); },
                                          [&](float acc, int j) { return __fadd_rn(acc, o[j]); });
    const float inv = __fdiv_rn(1.0f, sum);
    for (int j = lane; j < n; j += 32) o[j] = __fmul_rn(o[j], inv);
}

extern "C" int lele_b200_softmax(lele_b200_ctx* ctx, const float* x, long long outer, int n, float* out) {
    LB_REQUIRE(ctx && x && out && n > 0 && outer >= 0, "softmax: bad arguments");
    if (outer == 0) return LELE_B200_OK;
    softmax_kernel<<<lb_ceil_div(outer, 8), 256, 0, ctx->stream>>>(x, outer, n, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// BatchNorm (norm.rs:313-419): y = x*s + (bias - mean*s), s = scale / sqrt(var+eps)
__global__ void batch_norm_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
                                  const float* __restrict__ mean, const float* __restrict__ var, int c, long long inner,
                                  float eps, long long total, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)((i / inner) % c);
        float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[ch], eps)));
        float s = __fmul_rn(scale[ch], inv);
        float sh = __fsub_rn(bias[ch], __fmul_rn(mean[ch], s));
        out[i] = __fadd_rn(__fmul_rn(x[i], s), sh);
    }
}
extern "C" int lele_b200_batch_norm(lele_b200_ctx* ctx, const float* x, const float* scale, const float* bias,
                                    const float* mean, const float* var, int nb, int c, long long inner, float eps,
                                    float* out) {
    LB_REQUIRE(ctx && x && scale && bias && mean && var && out, "batch_norm: NULL argument");
    long long total = (long long)nb * c * inner;
    if (total == 0) return LELE_B200_OK;
    batch_norm_kernel<<<min(lb_ceil_div(total, 256), 148 * 16), 256, 0, ctx->stream>>>(x, scale, bias, mean, var, c, inner, eps, total, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// RMSNorm (norm.rs:420-506): x * w / sqrt(mean(x^2) + eps)
__global__ void __launch_bounds__(256)
rms_norm_kernel(const float* __restrict__ x, const float* __restrict__ w, long long outer, int n, float eps,
                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* xr = x + row * n;
    float sq = 0.0f;
    for (int j = lane; j < n; j += 32) sq = fmaf(xr[j], xr[j], sq);
    sq = lb_warp_sum(sq);
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(sq, (float)n), eps)));
    for (int j = lane; j < n; j += 32) out[row * n + j] = __fmul_rn(__fmul_rn(xr[j], inv), w ? w[j] : 1.0f);
}
extern "C" int lele_b200_rms_norm(lele_b200_ctx* ctx, const float* x, const float* w, long long outer, int n, float eps,
                                  float* out) {
    LB_REQUIRE(ctx && x && out && n > 0, "rms_norm: bad arguments");
    if (outer == 0) return LELE_B200_OK;
    rms_norm_kernel<<<lb_ceil_div(outer, 8), 256, 0, ctx->stream>>>(x, w, outer, n, eps, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
