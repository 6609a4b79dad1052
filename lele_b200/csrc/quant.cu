// quant.cu -- DynamicQuantizeLinear / MatMulInteger / fused quantised linear
// (src/kernels/quantization.rs:8-169,1628 + avx/quantization.rs:102-330,832-921).
//
//  per slice (= one clip's [m,k] activation tensor, the reference's "whole input tensor"):
//    min/max -> scale = max(amax-amin,1e-5)/255, zp = clamp(round(-amin/scale),0,255)
//    a_q = clamp(rint(fma(x, 1/scale, zp)),0,255)        (SIMD body; row tail k%8: mul+add, round half away)
//  then the exact integer GEMM + epilogue of gemm_i8_tc.cu.
#include "gemm_i8_tc.cuh"
#include <stdlib.h>


// ---------------------------------------------------------------------------
// per-slice min/max (HBM-bound reduction: 128-bit loads, shuffle, one atomic pair per CTA)
// ---------------------------------------------------------------------------
__global__ void minmax_init_kernel(unsigned* keys, int n_slices) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slices * LB_MM_SLOTS) { keys[2 * i] = LB_KEY_MIN_INIT; keys[2 * i + 1] = LB_KEY_MAX_INIT; }
}

__global__ void __launch_bounds__(256)
slice_minmax_kernel(const float* __restrict__ x, long long slice_len, unsigned* __restrict__ keys) {
    const int slice = blockIdx.y;
    const float* xs = x + (long long)slice * slice_len;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)xs) & 15) == 0) {
        const long long nv = slice_len >> 2;
        for (long long q = i; q < nv; q += stride) {
            float4 v = __ldg(reinterpret_cast<const float4*>(xs) + q);
            mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
            mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        for (long long q = (nv << 2) + i; q < slice_len; q += stride) { float v = xs[q]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    } else {
        for (long long q = i; q < slice_len; q += stride) { float v = xs[q]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    }
    mn = lb_warp_min(mn); mx = lb_warp_max(mx);
    __shared__ float smn[8], smx[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < 8 ? smn[lane] : 3.402823466e+38f;
        mx = lane < 8 ? smx[lane] : -3.402823466e+38f;
        mn = lb_warp_min(mn); mx = lb_warp_max(mx);
        if (lane == 0) lb_mm_update(keys, slice, mn, mx);
    }
}

// scale / zero point from the min/max keys (avx/quantization.rs:135-140)
__device__ __forceinline__ void dq_params(const unsigned* keys, int slice, float& scale, float& zp, float& inv) {
    float mn, mx;
    lb_mm_read(keys, slice, mn, mx);
    float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);
    float range = fmaxf(__fsub_rn(amax, amin), 1e-5f);
    scale = __fdiv_rn(range, 255.0f);
    zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
    inv = __fdiv_rn(1.0f, scale);
}
__device__ __forceinline__ float dq_one(float v, float inv, float zp, bool simd) {
    float r = simd ? rintf(__fmaf_rn(v, inv, zp)) : roundf(__fadd_rn(__fmul_rn(v, inv), zp));
    return fminf(fmaxf(r, 0.0f), 255.0f);
}

// one warp per row: f32 [M,K] -> u8 [M,K] + row sums + per-row (scale, zp) for the GEMM epilogue
// (dq_to_u8_rowsums_avx2, avx/quantization.rs:102-221).  kAligned: K % 8 == 0 (every SenseVoice shape), i.e.
// the whole row is "SIMD body" (fma + round-half-even) and no per-element tail predicate is needed.
__device__ __forceinline__ unsigned dq_body(float v, float inv, float zp) {
    // round-half-even + clamp to [0, 255]: cvt.rni.u32.f32 saturates negatives (and NaN) to 0, so one conversion and
    // one integer min give exactly clamp(rint(fma(v, inv, zp)), 0, 255)
    return lb_q8(__fmaf_rn(v, inv, zp));
}
template <bool kAligned>
__global__ void __launch_bounds__(256)
quantize_rows_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys, long long M, int rows_per_slice, int K,
                     uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale,
                     int32_t* __restrict__ row_zp) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int slice = (int)(row / rows_per_slice);
    float scale, zp, inv;
    dq_params(keys, slice, scale, zp, inv);
    const float* xr = x + row * K;
    uint8_t* ar = a_u8 + row * K;
    int sum = 0;
    if (kAligned) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        unsigned* a4 = reinterpret_cast<unsigned*>(ar);
        const int nv = K >> 2;
#pragma unroll 4
        for (int c = lane; c < nv; c += 32) {
            const float4 v = __ldg(x4 + c);
            const unsigned q0 = dq_body(v.x, inv, zp), q1 = dq_body(v.y, inv, zp), q2 = dq_body(v.z, inv, zp), q3 = dq_body(v.w, inv, zp);
            sum += (int)(q0 + q1 + q2 + q3);
            a4[c] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
        }
    } else {
        const int k_simd = (K / 8) * 8;
        for (int j = lane; j < K; j += 32) {
            unsigned q = (unsigned)dq_one(xr[j], inv, zp, j < k_simd);
            sum += (int)q;
            ar[j] = (uint8_t)q;
        }
    }
    sum = lb_warp_sum_i(sum);
    if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
}

// Register-resident variant for the SenseVoice widths (K = 128 * KV): the whole row is loaded before the clip's
// quantisation parameters are fetched, so the two L2 round trips overlap.
template <int KV>
__global__ void __launch_bounds__(256)
quantize_rows_reg_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys, long long M, int rows_per_slice,
                         uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale,
                         int32_t* __restrict__ row_zp) {
    constexpr int K = KV * 128;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    lb_pdl_launch_dependents();
    lb_pdl_wait();
    if (row >= M) return;
    const float4* x4 = reinterpret_cast<const float4*>(x + row * K);
    float4 v[KV];
#pragma unroll
    for (int j = 0; j < KV; ++j) v[j] = __ldg(x4 + lane + 32 * j);
    float scale, zp, inv;
    dq_params(keys, (int)(row / rows_per_slice), scale, zp, inv);
    unsigned* a4 = reinterpret_cast<unsigned*>(a_u8 + row * K);
    int sum = 0;
#pragma unroll
    for (int j = 0; j < KV; ++j) {
        const unsigned q0 = dq_body(v[j].x, inv, zp), q1 = dq_body(v[j].y, inv, zp), q2 = dq_body(v[j].z, inv, zp), q3 = dq_body(v[j].w, inv, zp);
        sum += (int)(q0 + q1 + q2 + q3);
        a4[lane + 32 * j] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
    }
    sum = lb_warp_sum_i(sum);
    if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
}

int lb_minmax_init(lele_b200_ctx* ctx, unsigned* keys, int n_slices) {
    minmax_init_kernel<<<lb_ceil_div((long long)n_slices * LB_MM_SLOTS, 128), 128, 0, ctx->stream>>>(keys, n_slices);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
int lb_slice_minmax(lele_b200_ctx* ctx, const float* x, int n_slices, long long slice_len, unsigned* keys) {
    int bx = (int)((slice_len / 4 + 255) / 256);
    int cap = (ctx->num_sms * 8 + n_slices - 1) / n_slices;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    slice_minmax_kernel<<<dim3(bx, n_slices), 256, 0, ctx->stream>>>(x, slice_len, keys);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
int lb_quantize_rows(lele_b200_ctx* ctx, const float* x, const unsigned* keys, long long M, int rows_per_slice, int K,
                     uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp) {
    const bool al = ((((uintptr_t)x) & 15) == 0) && ((((uintptr_t)a_u8) & 3) == 0);
    if (al && K == 512)
        LB_CHECK_CUDA(lb_launch_pdl(quantize_rows_reg_kernel<4>, dim3(lb_ceil_div(M, 8)), dim3(256), 0, ctx->stream, 1, x, keys, M, rows_per_slice, a_u8, rowsum, row_scale, row_zp));
    else if (al && K == 2048)
        LB_CHECK_CUDA(lb_launch_pdl(quantize_rows_reg_kernel<16>, dim3(lb_ceil_div(M, 8)), dim3(256), 0, ctx->stream, 1, x, keys, M, rows_per_slice, a_u8, rowsum, row_scale, row_zp));
    else if (K % 8 == 0 && al)
        quantize_rows_kernel<true><<<lb_ceil_div(M, 8), 256, 0, ctx->stream>>>(x, keys, M, rows_per_slice, K, a_u8, rowsum, row_scale, row_zp);
    else
        quantize_rows_kernel<false><<<lb_ceil_div(M, 8), 256, 0, ctx->stream>>>(x, keys, M, rows_per_slice, K, a_u8, rowsum, row_scale, row_zp);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// ---------------------------------------------------------------------------
// dynamic_quantize_linear operator: q as f32, scale [n_slices], zp [n_slices]
// (avx/quantization.rs:832-921: SIMD body = first len/8*8 elements of the flat tensor)
// ---------------------------------------------------------------------------
__global__ void dql_kernel(const float* __restrict__ x, const unsigned* __restrict__ keys, long long slice_len,
                           float* __restrict__ q, float* __restrict__ scale_out, float* __restrict__ zp_out) {
    const int slice = blockIdx.y;
    float scale, zp, inv;
    dq_params(keys, slice, scale, zp, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scale_out[slice] = scale; zp_out[slice] = zp; }
    const long long simd_end = (slice_len / 8) * 8;
    const float* xs = x + (long long)slice * slice_len;
    float* qs = q + (long long)slice * slice_len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slice_len; i += (long long)gridDim.x * blockDim.x)
        qs[i] = dq_one(xs[i], inv, zp, i < simd_end);
}
__global__ void dql_empty_kernel(float* scale_out, float* zp_out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { scale_out[i] = 1.0f; zp_out[i] = 0.0f; }   // empty input -> (1.0, 0.0) avx/quantization.rs:845-851
}

extern "C" int lele_b200_dynamic_quantize_linear(lele_b200_ctx* ctx, const float* x, int n_slices, long long slice_len,
                                                 float* q, float* scale, float* zp) {
    LB_REQUIRE(ctx && scale && zp && n_slices >= 0 && slice_len >= 0, "dynamic_quantize_linear: bad arguments");
    LB_ENTER(ctx);
    if (n_slices == 0) return LELE_B200_OK;
    if (slice_len == 0) {
        dql_empty_kernel<<<lb_ceil_div(n_slices, 128), 128, 0, ctx->stream>>>(scale, zp, n_slices);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    LB_REQUIRE(x && q, "dynamic_quantize_linear: NULL tensor");
    void* sc;
    int rc = lb_scratch(ctx, sizeof(unsigned) * 2 * LB_MM_SLOTS * n_slices, &sc);
    if (rc) return rc;
    unsigned* keys = (unsigned*)sc;
    if ((rc = lb_minmax_init(ctx, keys, n_slices))) return rc;
    if ((rc = lb_slice_minmax(ctx, x, n_slices, slice_len, keys))) return rc;
    int bx = (int)((slice_len + 255) / 256);
    if (bx > 1024) bx = 1024;
    dql_kernel<<<dim3(bx, n_slices), 256, 0, ctx->stream>>>(x, keys, slice_len, q, scale, zp);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// ---------------------------------------------------------------------------
// mat_mul_integer*: f32-coded integers, exact i32 accumulation (quantization.rs:1137-1236)
// generic operator (the unfused fallback of the codegen patterns); CUDA cores.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int f32_as_u8(float v) {  // Rust `as u8`: saturating, NaN -> 0
    return (int)fminf(fmaxf(v, 0.0f), 255.0f);
}
__global__ void __launch_bounds__(256)
mmi_kernel(const float* __restrict__ a, const float* __restrict__ b_all, long long a_bs, long long b_bs, int m, int k, int n, int zpa, int zpb,
           const float* __restrict__ scale, int scale_len, const float* __restrict__ bias, int relu, float* __restrict__ out) {
    // 16x16 output tile per CTA, k staged through shared memory in chunks of 16
    __shared__ int sa[16][17], sb[16][17];
    const int bi = blockIdx.z;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    const float* ab = a + (long long)bi * a_bs;                   // a stride of 0 broadcasts the single matrix (quantization.rs:1172-1173)
    const float* b = b_all + (long long)bi * b_bs;
    int acc = 0;
    for (int k0 = 0; k0 < k; k0 += 16) {
        sa[ty][tx] = (row < m && k0 + tx < k) ? f32_as_u8(ab[(long long)row * k + k0 + tx]) - zpa : 0;
        sb[ty][tx] = (k0 + ty < k && col < n) ? f32_as_u8(b[(long long)(k0 + ty) * n + col]) - zpb : 0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc += sa[ty][kk] * sb[kk][tx];
        __syncthreads();
    }
    if (row < m && col < n) {
        float v = (float)acc;
        if (scale) v = __fmul_rn(v, scale_len == 1 ? scale[0] : scale[col]);
        if (bias) v = __fadd_rn(v, bias[col]);
        if (relu) v = fmaxf(v, 0.0f);
        out[((long long)bi * m + row) * n + col] = v;
    }
}
extern "C" int lele_b200_mat_mul_integer_batched(lele_b200_ctx* ctx, const float* a, const float* b, int batch_a, int batch_b, int m, int k,
                                                 int n, float a_zp, float b_zp, const float* scale, int scale_len, const float* bias,
                                                 int relu, float* out) {
    LB_REQUIRE(ctx && a && b && out, "mat_mul_integer: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(batch_a >= 0 && batch_b >= 0 && m >= 0 && k >= 0 && n >= 0, "mat_mul_integer: negative dims");
    LB_REQUIRE(batch_a == batch_b || batch_a == 1 || batch_b == 1, "mat_mul_integer: batch %d vs %d (equal, or one side 1; quantization.rs:1157-1173)", batch_a, batch_b);
    LB_REQUIRE(!scale || scale_len == 1 || scale_len == n, "mat_mul_integer: scale length %d is neither 1 nor n=%d", scale_len, n);
    const int batch = batch_a > batch_b ? batch_a : batch_b;
    if (batch == 0 || m == 0 || n == 0) return LELE_B200_OK;
    dim3 grid(lb_ceil_div(n, 16), lb_ceil_div(m, 16), batch);
    mmi_kernel<<<grid, 256, 0, ctx->stream>>>(a, b, batch_a == 1 ? 0ll : (long long)m * k, batch_b == 1 ? 0ll : (long long)k * n, m, k, n, (int)a_zp,
                                              (int)b_zp, scale, scale_len, bias, relu, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_mat_mul_integer(lele_b200_ctx* ctx, const float* a, const float* b, int batch, int m, int k, int n,
                                         float a_zp, float b_zp, const float* scale, int scale_len, const float* bias,
                                         int relu, float* out) {
    return lele_b200_mat_mul_integer_batched(ctx, a, b, batch, 1, m, k, n, a_zp, b_zp, scale, scale_len, bias, relu, out);
}

// ---------------------------------------------------------------------------
// prepare_weights: u8 [k,n] -> K-major [n,k] + column sums (+ padded per-column vectors)
// ---------------------------------------------------------------------------
// `flip` = 0x80 stores w ^ 0x80, i.e. the s8 value w - 128 (the zero point 128 folded into the operand), and sums that
__global__ void prep_weights_kernel(const uint8_t* __restrict__ w, int k, int n, uint8_t* __restrict__ wt, int32_t* __restrict__ colsum, int flip) {
    // 32x32 tile transpose through shared memory; column sums accumulated with atomics per tile
    __shared__ uint8_t tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int kk = k0 + r, j = n0 + tx;
        tile[r][tx] = (kk < k && j < n) ? w[(long long)kk * n + j] : 0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int j = n0 + r, kk = k0 + tx;
        if (j < n && kk < k) wt[(long long)j * k + kk] = (uint8_t)(tile[tx][r] ^ flip);
    }
    if (ty == 0) {
        int s = 0;
        for (int r = 0; r < 32; ++r) s += tile[r][tx];
        if (flip) s -= 128 * min(32, k - k0);
        if (n0 + tx < n) atomicAdd(colsum + n0 + tx, s);
    }
}
__global__ void prep_vectors_kernel(const float* __restrict__ w_scale, int w_scale_len, const float* __restrict__ bias, int n, int n_pad,
                                    float* __restrict__ ws_out, float* __restrict__ bias_out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    ws_out[j] = j < n ? (w_scale_len == 1 ? w_scale[0] : w_scale[j]) : 0.0f;
    bias_out[j] = (j < n && bias) ? bias[j] : 0.0f;
}

extern "C" int lele_b200_prepare_weights(lele_b200_ctx* ctx, const uint8_t* w, int k, int n, const float* w_scale,
                                         int w_scale_len, int w_zp, const float* bias, lele_b200_qweights** out) {
    LB_REQUIRE(ctx && w && w_scale && out, "prepare_weights: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(k > 0 && n > 0, "prepare_weights: empty weight");
    LB_REQUIRE(w_scale_len == 1 || w_scale_len == n, "prepare_weights: weight_scale length %d is neither 1 nor n=%d", w_scale_len, n);
    LB_REQUIRE(w_zp >= 0 && w_zp <= 255, "prepare_weights: u8 zero point out of range (x86 handles u8 weights only, SURVEY appendix A)");
    lele_b200_qweights* q = new lele_b200_qweights();
    q->k = k; q->n = n; q->n_pad = (n + 255) / 256 * 256; q->w_zp = w_zp; q->has_bias = bias ? 1 : 0;
    // zero point 128 (lele's u8 weights): keep w - 128 as s8 -- the tensor core takes u8 x s8, so the
    // epilogue's zero-point correction loses its per-row term (w_zp * rowsum) and one integer add per output element
    q->w_signed = (w_zp == 128 && !lb_env_flag("LELE_B200_W_UNSIGNED", 0)) ? 1 : 0;
    LB_CHECK_CUDA(cudaMalloc(&q->wt, (size_t)n * k));
    LB_CHECK_CUDA(cudaMalloc(&q->colsum, sizeof(int32_t) * q->n_pad));
    LB_CHECK_CUDA(cudaMalloc(&q->w_scale, sizeof(float) * q->n_pad));
    LB_CHECK_CUDA(cudaMalloc(&q->bias, sizeof(float) * q->n_pad));
    LB_CHECK_CUDA(cudaMemsetAsync(q->colsum, 0, sizeof(int32_t) * q->n_pad, ctx->stream));
    prep_weights_kernel<<<dim3(lb_ceil_div(n, 32), lb_ceil_div(k, 32)), 256, 0, ctx->stream>>>(w, k, n, q->wt, q->colsum, q->w_signed ? 0x80 : 0);
    LB_LAUNCH_CHECK(ctx);
    prep_vectors_kernel<<<lb_ceil_div(q->n_pad, 256), 256, 0, ctx->stream>>>(w_scale, w_scale_len, bias, n, q->n_pad, q->w_scale, q->bias);
    LB_LAUNCH_CHECK(ctx);
    *out = q;
    return LELE_B200_OK;
}
extern "C" int lele_b200_qweights_destroy(lele_b200_ctx* ctx, lele_b200_qweights* w) {
    if (!w) return LELE_B200_OK;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(w->wt); cudaFree(w->colsum); cudaFree(w->w_scale); cudaFree(w->bias);
    delete w;
    return LELE_B200_OK;
}

// ---------------------------------------------------------------------------
// CUDA-core u8 GEMM with the same epilogue: shapes the TMA path cannot take (K % 16 != 0)
// and an on-device cross-check of the tcgen05 kernel (LELE_B200_FORCE_SIMT=1).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_i8_simt_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ Wt, int M, int N, int K, LbI8Epilogue ep) {
    __shared__ int sa[16][17], sb[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    int acc = 0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        sa[ty][tx] = (row < M && k0 + tx < K) ? A[(long long)row * K + k0 + tx] : 0;
        int wc = blockIdx.x * 16 + ty;
        const int wraw = (wc < N && k0 + tx < K) ? Wt[(long long)wc * K + k0 + tx] : 0;
        sb[tx][ty] = ep.w_signed ? (int)(int8_t)wraw : wraw;   // sb[kk][col]; signed storage: the byte w ^ 0x80 read as s8 is w - 128
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc += sa[ty][kk] * sb[kk][tx];
        __syncthreads();
    }
    if (row < M && col < N) {
        int zpa = ep.row_zp[row];
        const int wz = ep.w_signed ? ep.w_zp - 128 : ep.w_zp;
        int acci = acc + (wz ? K * zpa * wz - wz * ep.rowsum[row] : 0) - zpa * ep.colsum[col];   // s8 weights: no row term (rowsum may be NULL)
        float v = __fmul_rn((float)acci, __fmul_rn(ep.row_scale[row], ep.w_scale[col]));
        if (ep.has_bias) v = __fadd_rn(v, ep.bias[col]);
        if (ep.relu) v = fmaxf(v, 0.0f);
        long long o = (long long)row * N + col;
        if (ep.add1) v = __fadd_rn(v, ep.add1[o]);
        if (ep.add2) v = __fadd_rn(ep.add2[o], v);
        if (ep.minmax_keys) {
            lb_mm_update(ep.minmax_keys, row / ep.rows_per_slice, v, v);
        }
        if (ep.argmax_keys) atomicMax(ep.argmax_keys + row, ((unsigned long long)lb_fkey_argmax(v) << 32) | (unsigned)col);
        if (ep.out) ep.out[o] = v;
    }
}

static bool force_simt() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LELE_B200_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int lb_gemm_i8(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K, const LbI8Epilogue& ep) {
    if (K % 16 == 0 && !force_simt()) return lb_gemm_i8_tc(ctx, A, Wt, M, N, K, ep);
    gemm_i8_simt_kernel<<<dim3(lb_ceil_div(N, 16), lb_ceil_div(M, 16)), 256, 0, ctx->stream>>>(A, Wt, M, N, K, ep);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// Shared by the operator and the SenseVoice runner: quantise rows of x (per-slice keys must
// already hold min/max) and run the GEMM.  Scratch layout is carved by the caller.
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
size_t lb_quant_scratch_bytes(long long M, int K) {
    return align_up((size_t)M * K, 256) + 3 * align_up(sizeof(int32_t) * (size_t)M, 256);
}
LbQuantScratch lb_quant_scratch_carve(void* base, long long M, int K) {
    LbQuantScratch s;
    uint8_t* p = (uint8_t*)base;
    s.a_u8 = p; p += align_up((size_t)M * K, 256);
    s.rowsum = (int32_t*)p; p += align_up(sizeof(int32_t) * (size_t)M, 256);
    s.row_scale = (float*)p; p += align_up(sizeof(float) * (size_t)M, 256);
    s.row_zp = (int32_t*)p;
    return s;
}
void lb_fill_weight_fields(LbI8Epilogue& ep, const lele_b200_qweights* w, const LbQuantScratch& s) {
    ep.rowsum = s.rowsum; ep.row_scale = s.row_scale; ep.row_zp = s.row_zp;
    ep.colsum = w->colsum; ep.w_scale = w->w_scale; ep.bias = w->bias; ep.w_zp = w->w_zp; ep.w_signed = w->w_signed; ep.has_bias = w->has_bias;
}
static int lb_quantized_linear(lele_b200_ctx* ctx, const float* x, const unsigned* keys, long long M, int rows_per_slice,
                               const lele_b200_qweights* w, const LbQuantScratch& s, LbI8Epilogue ep) {
    int rc = lb_quantize_rows(ctx, x, keys, M, rows_per_slice, w->k, s.a_u8, s.rowsum, s.row_scale, s.row_zp);
    if (rc) return rc;
    lb_fill_weight_fields(ep, w, s);
    return lb_gemm_i8(ctx, s.a_u8, w->wt, (int)M, w->n, w->k, ep);
}

extern "C" int lele_b200_fused_quantized_linear(lele_b200_ctx* ctx, const float* x, int n_slices, int m,
                                                const lele_b200_qweights* w, int relu, float* out) {
    LB_REQUIRE(ctx && x && w && out, "fused_quantized_linear: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_slices >= 0 && m >= 0, "fused_quantized_linear: negative dims");
    const long long M = (long long)n_slices * m;
    if (M == 0) return LELE_B200_OK;
    LB_REQUIRE(M < (1ll << 31), "fused_quantized_linear: too many rows");
    size_t keys_bytes = align_up(sizeof(unsigned) * 2 * LB_MM_SLOTS * (size_t)n_slices, 256);
    void* sc;
    int rc = lb_scratch(ctx, keys_bytes + lb_quant_scratch_bytes(M, w->k), &sc);
    if (rc) return rc;
    unsigned* keys = (unsigned*)sc;
    LbQuantScratch qs = lb_quant_scratch_carve((uint8_t*)sc + keys_bytes, M, w->k);
    if ((rc = lb_minmax_init(ctx, keys, n_slices))) return rc;
    if ((rc = lb_slice_minmax(ctx, x, n_slices, (long long)m * w->k, keys))) return rc;
    LbI8Epilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.relu = relu; ep.out = out; ep.rows_per_slice = m;
    return lb_quantized_linear(ctx, x, keys, M, m, w, qs, ep);
}
