// norm.cu -- LayerNorm / Softmax / BatchNorm / RMSNorm (src/kernels/norm.rs + avx/norm.rs).
// HBM-bound row kernels: one warp per row, 128-bit loads, warp-shuffle reductions.
#include "common.cuh"
#include <cooperative_groups.h>

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
// STORE = false: the normalised rows are not written; only (mean, 1/std) per row and the per-slice min/max of
// the would-be output, for ln_quantize_kernel to re-derive and quantise the rows in one more read of x.
template <int N, int RPW, bool STORE>
__global__ void __launch_bounds__(256)
layer_norm_fixed_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        long long outer, float eps, float* __restrict__ out, unsigned* __restrict__ minmax_keys, int rows_per_slice,
                        float2* __restrict__ stats) {
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
        float* o = STORE ? out + row * N : nullptr;
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
        if (!STORE && lane == 0) stats[row] = make_float2(mean, inv);
        float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float r = __fmaf_rn(__fmul_rn(__fsub_rn(v[i], mean), inv), g[i], bt[i]);
            vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
            if (STORE) o[32 * i + lane] = r;
        }
        if (REM > 0 && lane < REM) {
            const float r = __fmaf_rn(__fmul_rn(__fsub_rn(xrem, mean), inv), gr, br);
            vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
            if (STORE) o[NB * 32 + lane] = r;
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
        if (n == 512) layer_norm_fixed_kernel<512, RPW, true><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, eps, out, minmax_keys, rows_per_slice, nullptr);
        else layer_norm_fixed_kernel<560, RPW, true><<<grid, warps * 32, 0, ctx->stream>>>(x, gamma, beta, outer, eps, out, minmax_keys, rows_per_slice, nullptr);
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

// ---- LayerNorm fused with the dynamic quantiser that consumes it (SenseVoice encoder, N = 512) ----------------
// The reference materialises LayerNorm's f32 output and re-reads it to quantise (norm.rs:227 ->
// avx/quantization.rs:102).  Here pass 1 (layer_norm_fixed_kernel<.., STORE=false>) leaves only (mean, 1/std)
// per row and the per-clip min/max of the normalised values; pass 2 re-derives each value with the identical
// operation sequence (so it is bit-identical to what pass 1 measured) and emits u8 + row sums directly.
// pass 1 for N = 512: RPW rows per warp, all row loads issued before anything depends on them
template <int RPW>
__global__ void __launch_bounds__(256)
ln_stats512_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long outer, float eps,
                   float2* __restrict__ stats, unsigned* __restrict__ minmax_keys, int rows_per_slice) {
    constexpr int N = 512, NB = 16;
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (row0 >= outer) return;
    float v[RPW][NB];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const long long row = row0 + rr < outer ? row0 + rr : outer - 1;
#pragma unroll
        for (int i = 0; i < NB; ++i) v[rr][i] = __ldg(x + row * N + 32 * i + lane);
    }
    float g[NB], bt[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { g[i] = __ldg(gamma + 32 * i + lane); bt[i] = __ldg(beta + 32 * i + lane); }
    const float inv_n = __fdiv_rn(1.0f, (float)N);
    const long long slice_a = row0 / rows_per_slice;
    float mnA = 3.402823466e+38f, mxA = -3.402823466e+38f, mnB = 3.402823466e+38f, mxB = -3.402823466e+38f;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const long long row = row0 + rr;
        if (row < outer) {
            float ps = 0.0f, pq = 0.0f;
#pragma unroll
            for (int i = 0; i < NB; ++i) { ps = __fadd_rn(ps, v[rr][i]); pq = __fmaf_rn(v[rr][i], v[rr][i], pq); }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, o)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, o)); }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, o)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, o)); }
            ps = __shfl_sync(0xffffffffu, ps, 0); pq = __shfl_sync(0xffffffffu, pq, 0);
            const float mean = __fmul_rn(ps, inv_n);
            const float var = __fsub_rn(__fmul_rn(pq, inv_n), __fmul_rn(mean, mean));
            const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
            if (lane == 0) stats[row] = make_float2(mean, inv);
            float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float r = __fmaf_rn(__fmul_rn(__fsub_rn(v[rr][i], mean), inv), g[i], bt[i]);
                vmin = fminf(vmin, r); vmax = fmaxf(vmax, r);
            }
            if (row / rows_per_slice == slice_a) { mnA = fminf(mnA, vmin); mxA = fmaxf(mxA, vmax); }
            else { mnB = fminf(mnB, vmin); mxB = fmaxf(mxB, vmax); }
        }
    }
    mnA = lb_warp_min(mnA); mxA = lb_warp_max(mxA); mnB = lb_warp_min(mnB); mxB = lb_warp_max(mxB);
    if (lane == 0) {
        if (mnA <= mxA) lb_mm_update(minmax_keys, slice_a, mnA, mxA);
        if (mnB <= mxB) lb_mm_update(minmax_keys, slice_a + 1, mnB, mxB);
    }
}

__device__ __forceinline__ void lnq_params(const unsigned* keys, long long slice, float& scale, float& zp, float& inv) {
    float mn, mx;
    lb_mm_read(keys, slice, mn, mx);
    const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);
    const float range = fmaxf(__fsub_rn(amax, amin), 1e-5f);
    scale = __fdiv_rn(range, 255.0f);
    zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
    inv = __fdiv_rn(1.0f, scale);
}
__device__ __forceinline__ unsigned lnq_one(float x, float mean, float rstd, float g, float b, float inv, float zp) {
    const float r = __fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), g, b);
    return lb_q8(__fmaf_rn(r, inv, zp));   // == clamp(rint(.), 0, 255)
}
template <int N, int RPW>
__global__ void __launch_bounds__(256)
ln_quantize_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float2* __restrict__ stats, const unsigned* __restrict__ keys, long long M, int rows_per_slice,
                   uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale, int32_t* __restrict__ row_zp) {
    static_assert(N % 128 == 0, "one float4 per lane per 128 columns");
    constexpr int NV = N / 128;
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (row0 >= M) return;
    // every global load is issued before the first dependent instruction
    float4 v[RPW][NV];
    float2 st[RPW];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const long long row = row0 + rr < M ? row0 + rr : M - 1;
        const float4* x4 = reinterpret_cast<const float4*>(x + row * N);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[rr][j] = __ldg(x4 + lane + 32 * j);
        st[rr] = __ldg(stats + row);
    }
    float4 g[NV], bt[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        g[j] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
        bt[j] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
    }
    const long long slice_a = row0 / rows_per_slice;
    float scA, zpA, inA, scB = 0.0f, zpB = 0.0f, inB = 0.0f;
    lnq_params(keys, slice_a, scA, zpA, inA);
    const long long last = (row0 + RPW - 1 < M ? row0 + RPW - 1 : M - 1);
    if (last / rows_per_slice != slice_a) lnq_params(keys, slice_a + 1, scB, zpB, inB);
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const long long row = row0 + rr;
        if (row < M) {
            const bool a = row / rows_per_slice == slice_a;
            const float scale = a ? scA : scB, zp = a ? zpA : zpB, inv = a ? inA : inB;
            unsigned* a4 = reinterpret_cast<unsigned*>(a_u8 + row * N);
            int sum = 0;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const unsigned q0 = lnq_one(v[rr][j].x, st[rr].x, st[rr].y, g[j].x, bt[j].x, inv, zp), q1 = lnq_one(v[rr][j].y, st[rr].x, st[rr].y, g[j].y, bt[j].y, inv, zp);
                const unsigned q2 = lnq_one(v[rr][j].z, st[rr].x, st[rr].y, g[j].z, bt[j].z, inv, zp), q3 = lnq_one(v[rr][j].w, st[rr].x, st[rr].y, g[j].w, bt[j].w, inv, zp);
                sum += (int)(q0 + q1 + q2 + q3);
                a4[lane + 32 * j] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
            }
            sum = lb_warp_sum_i(sum);
            if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
        }
    }
}

// ---- single-pass variant: one thread-block cluster per clip keeps the clip's normalised rows in shared memory ----
// CS CTAs (one SM each) split the clip's rows; each normalises its rows into shared memory (x is read exactly once),
// the per-CTA min/max are exchanged through distributed shared memory, then every CTA quantises its resident rows.
// No f32 intermediate, no statistics buffer, no global atomics.  Same arithmetic as the two-pass kernels.
constexpr int LNQ_CHUNK = 8;         // rows per bulk copy / mbarrier
constexpr int LNQ_MAX_CHUNKS = 16;   // <= 128 rows per CTA
// LNQ_CS CTAs per clip (the cluster), LNQ_WARPS warps per CTA, MINB co-resident CTAs per SM: 8 x 8 warps x 3 keeps ~105 rows
// in flight per SM in three independent CTAs (their copy / normalise / exchange / quantise phases overlap) instead of one CTA
// holding 69 rows and running its phases back to back.
template <int LNQ_CS, int LNQ_WARPS, int MINB>
__global__ void __launch_bounds__(LNQ_WARPS * 32, MINB)
ln_quant_cluster_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int T, float eps,
                        uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale, int32_t* __restrict__ row_zp,
                        unsigned* __restrict__ keys_out, int dbg) {
    namespace cg = cooperative_groups;
    constexpr int N = 512, NB = 16;
    long long t_dbg[6]; t_dbg[0] = clock64();
    extern __shared__ __align__(16) float lnq_rows[];            // [R][512] normalised rows of this CTA
    __shared__ float red[2][32];
    __shared__ float peer_mm[LNQ_CS][2];                          // (min, max) of every CTA of the cluster, pushed by its owner
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int clip = blockIdx.y;
    const int R = (T + LNQ_CS - 1) / LNQ_CS;
    const int r_begin = rank * R, r_end = min(T, r_begin + R);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xc = x + (long long)clip * T * N;
    // The CTA's rows are one contiguous block of x: fetch it with bulk async copies (LNQ_CHUNK rows per mbarrier) so
    // every byte is in flight at once; warps normalise rows in place as their chunk lands.
    __shared__ __align__(8) unsigned long long chunk_bar[LNQ_MAX_CHUNKS];
    const int n_rows = r_end - r_begin;
    const int n_chunks = (n_rows + LNQ_CHUNK - 1) / LNQ_CHUNK;
    if (threadIdx.x == 0) {
        for (int c = 0; c < n_chunks; ++c)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&chunk_bar[c])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    lb_pdl_launch_dependents();
    lb_pdl_wait();                 // PDL: x is the previous kernel's output
    if (threadIdx.x == 0) {
        for (int c = 0; c < n_chunks; ++c) {
            const int rows = min(LNQ_CHUNK, n_rows - c * LNQ_CHUNK);
            const uint32_t bytes = (uint32_t)rows * N * 4;
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&chunk_bar[c]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(lnq_rows + (size_t)c * LNQ_CHUNK * N)),
                           "l"(xc + (long long)(r_begin + c * LNQ_CHUNK) * N), "r"(bytes), "r"(bar) : "memory");
        }
    }
    float g[NB], bt[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { g[i] = __ldg(gamma + 32 * i + lane); bt[i] = __ldg(beta + 32 * i + lane); }
    const float inv_n = __fdiv_rn(1.0f, (float)N);
    float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
    t_dbg[1] = clock64();
    for (int r = warp; r < n_rows; r += LNQ_WARPS) {
        {   // wait for the chunk holding row r (single use: parity 0)
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&chunk_bar[r / LNQ_CHUNK]);
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar) : "memory");
        }
        float* o = lnq_rows + (size_t)r * N;
        float v[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) v[i] = o[32 * i + lane];
        float ps = 0.0f, pq = 0.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) { ps = __fadd_rn(ps, v[i]); pq = __fmaf_rn(v[i], v[i], pq); }
#pragma unroll
        for (int of = 8; of <= 16; of <<= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
#pragma unroll
        for (int of = 4; of >= 1; of >>= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
        ps = __shfl_sync(0xffffffffu, ps, 0); pq = __shfl_sync(0xffffffffu, pq, 0);
        const float mean = __fmul_rn(ps, inv_n);
        const float var = __fsub_rn(__fmul_rn(pq, inv_n), __fmul_rn(mean, mean));
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float y = __fmaf_rn(__fmul_rn(__fsub_rn(v[i], mean), inv), g[i], bt[i]);
            vmin = fminf(vmin, y); vmax = fmaxf(vmax, y);
            o[32 * i + lane] = y;
        }
    }
    t_dbg[2] = clock64();
    vmin = lb_warp_min(vmin); vmax = lb_warp_max(vmax);
    if (lane == 0) { red[0][warp] = vmin; red[1][warp] = vmax; }
    __syncthreads();
    if (warp == 0) {
        float a = lane < LNQ_WARPS ? red[0][lane] : 3.402823466e+38f, b = lane < LNQ_WARPS ? red[1][lane] : -3.402823466e+38f;
        a = lb_warp_min(a); b = lb_warp_max(b);
        if (lane < LNQ_CS) {                                      // push into every CTA of the cluster (lane = destination rank):
            float* dst = cluster.map_shared_rank(&peer_mm[rank][0], lane);   // after the barrier each CTA reads only its own
            dst[0] = a; dst[1] = b;                               // shared memory, so no CTA has to outlive a peer's reads
        }
    }
    cluster.sync();                                               // every CTA's (min, max) has landed everywhere
    t_dbg[3] = clock64();
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
#pragma unroll
    for (int p = 0; p < LNQ_CS; ++p) { mn = fminf(mn, peer_mm[p][0]); mx = fmaxf(mx, peer_mm[p][1]); }
    if (keys_out && rank == 0 && threadIdx.x == 0) {              // diagnostics / tests: slot 0 of the clip's key set
        keys_out[(size_t)clip * LB_MM_SLOTS * 2] = lb_fkey(mn); keys_out[(size_t)clip * LB_MM_SLOTS * 2 + 1] = lb_fkey(mx);
    }
    const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);
    const float range = fmaxf(__fsub_rn(amax, amin), 1e-5f);
    const float scale = __fdiv_rn(range, 255.0f);
    const float zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
    const float inv = __fdiv_rn(1.0f, scale);
    __syncwarp();
    for (int r = warp; r < n_rows; r += LNQ_WARPS) {              // the same warp normalised this row
        const float4* y4 = reinterpret_cast<const float4*>(lnq_rows + (size_t)r * N);
        const long long row = (long long)clip * T + r_begin + r;
        unsigned* a4 = reinterpret_cast<unsigned*>(a_u8 + row * N);
        int sum = 0;
#pragma unroll
        for (int j = 0; j < N / 128; ++j) {
            const float4 y = y4[lane + 32 * j];
            a4[lane + 32 * j] = lb_q8x4(__fmaf_rn(y.x, inv, zp), __fmaf_rn(y.y, inv, zp), __fmaf_rn(y.z, inv, zp), __fmaf_rn(y.w, inv, zp), sum);
        }
        sum = lb_warp_sum_i(sum);
        if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
    }
    t_dbg[4] = clock64();
    if (dbg && threadIdx.x == 0 && (blockIdx.y == 3 || blockIdx.y == 40) && blockIdx.x < 2)
        printf("LNQDBG blk (%d,%d) start %lld | setup %lld rows %lld xchg %lld quant %lld end %lld\n", blockIdx.x, blockIdx.y, t_dbg[0] % 100000000ll, t_dbg[1] - t_dbg[0],
               t_dbg[2] - t_dbg[0], t_dbg[3] - t_dbg[0], t_dbg[4] - t_dbg[0], clock64() - t_dbg[0]);
}


// ---- register-row variant: the clip's rows do not fit the chip's shared memory in one wave (64 clips x 275 rows x 2 KB =
// 36 MB vs 33.6 MB), so with shared memory alone 512 CTAs run as 444 + 68 -- two waves, the second almost empty.  Here every
// warp keeps RR of its rows in REGISTERS (loaded with plain coalesced loads) and only the rest goes through the bulk-copy /
// shared-memory path: 27 instead of 35 rows of shared memory per CTA -> 4 CTAs per SM -> all 512 CTAs resident at once.
// Same arithmetic as ln_quant_cluster_kernel (bit-identical outputs).
template <int LNQ_CS, int LNQ_WARPS, int MINB, int RR>
__global__ void __launch_bounds__(LNQ_WARPS * 32, MINB)
ln_quant_cluster_reg_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int T, float eps,
                            uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale, int32_t* __restrict__ row_zp,
                            unsigned* __restrict__ keys_out) {
    namespace cg = cooperative_groups;
    constexpr int N = 512, NB = 16;
    extern __shared__ __align__(16) float lnq_rows[];            // [n_smem][512] normalised rows (the CTA's rows after the register rows)
    __shared__ float red[2][32];
    __shared__ float peer_mm[LNQ_CS][2];
    __shared__ __align__(8) unsigned long long chunk_bar[LNQ_MAX_CHUNKS];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int clip = blockIdx.y;
    const int R = (T + LNQ_CS - 1) / LNQ_CS;
    const int r_begin = rank * R, r_end = min(T, r_begin + R);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xc = x + (long long)clip * T * N;
    const int n_rows = max(r_end - r_begin, 0);
    const int n_reg = min(LNQ_WARPS * RR, n_rows);              // rows [0, n_reg) of the CTA live in registers (row r -> warp r % NW)
    const int n_smem = n_rows - n_reg;
    const int n_chunks = (n_smem + LNQ_CHUNK - 1) / LNQ_CHUNK;
    if (threadIdx.x == 0) {
        for (int c = 0; c < n_chunks; ++c)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&chunk_bar[c])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    lb_pdl_launch_dependents();
    lb_pdl_wait();                 // PDL: x is the previous kernel's output
    if (threadIdx.x == 0) {
        for (int c = 0; c < n_chunks; ++c) {
            const int rows = min(LNQ_CHUNK, n_smem - c * LNQ_CHUNK);
            const uint32_t bytes = (uint32_t)rows * N * 4;
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&chunk_bar[c]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(lnq_rows + (size_t)c * LNQ_CHUNK * N)),
                           "l"(xc + (long long)(r_begin + n_reg + c * LNQ_CHUNK) * N), "r"(bytes), "r"(bar) : "memory");
        }
    }
    float yr[RR][NB];
#pragma unroll
    for (int i = 0; i < RR; ++i) {                               // register rows: plain coalesced loads, all in flight
        const int r = warp + i * LNQ_WARPS;
        if (r < n_reg) {
            const float* src = xc + (long long)(r_begin + r) * N;
#pragma unroll
            for (int j = 0; j < NB; ++j) yr[i][j] = __ldg(src + 32 * j + lane);
        }
    }
    float g[NB], bt[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { g[i] = __ldg(gamma + 32 * i + lane); bt[i] = __ldg(beta + 32 * i + lane); }
    const float inv_n = __fdiv_rn(1.0f, (float)N);
    float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
    auto normalise = [&](float (&v)[NB]) {                       // v <- LayerNorm(v) (same operation order as ln_quant_cluster_kernel)
        float ps = 0.0f, pq = 0.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) { ps = __fadd_rn(ps, v[i]); pq = __fmaf_rn(v[i], v[i], pq); }
#pragma unroll
        for (int of = 8; of <= 16; of <<= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
#pragma unroll
        for (int of = 4; of >= 1; of >>= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
        ps = __shfl_sync(0xffffffffu, ps, 0); pq = __shfl_sync(0xffffffffu, pq, 0);
        const float mean = __fmul_rn(ps, inv_n);
        const float var = __fsub_rn(__fmul_rn(pq, inv_n), __fmul_rn(mean, mean));
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float y = __fmaf_rn(__fmul_rn(__fsub_rn(v[i], mean), inv), g[i], bt[i]);
            vmin = fminf(vmin, y); vmax = fmaxf(vmax, y);
            v[i] = y;
        }
    };
#pragma unroll
    for (int i = 0; i < RR; ++i)
        if (warp + i * LNQ_WARPS < n_reg) normalise(yr[i]);
    for (int r = warp; r < n_smem; r += LNQ_WARPS) {
        {   // wait for the chunk holding shared-memory row r (single use: parity 0)
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&chunk_bar[r / LNQ_CHUNK]);
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar) : "memory");
        }
        float* o = lnq_rows + (size_t)r * N;
        float v[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) v[i] = o[32 * i + lane];
        normalise(v);
#pragma unroll
        for (int i = 0; i < NB; ++i) o[32 * i + lane] = v[i];
    }
    vmin = lb_warp_min(vmin); vmax = lb_warp_max(vmax);
    if (lane == 0) { red[0][warp] = vmin; red[1][warp] = vmax; }
    __syncthreads();
    if (warp == 0) {
        float a = lane < LNQ_WARPS ? red[0][lane] : 3.402823466e+38f, b = lane < LNQ_WARPS ? red[1][lane] : -3.402823466e+38f;
        a = lb_warp_min(a); b = lb_warp_max(b);
        if (lane < LNQ_CS) {
            float* dst = cluster.map_shared_rank(&peer_mm[rank][0], lane);
            dst[0] = a; dst[1] = b;
        }
    }
    cluster.sync();
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
#pragma unroll
    for (int p = 0; p < LNQ_CS; ++p) { mn = fminf(mn, peer_mm[p][0]); mx = fmaxf(mx, peer_mm[p][1]); }
    if (keys_out && rank == 0 && threadIdx.x == 0) {
        keys_out[(size_t)clip * LB_MM_SLOTS * 2] = lb_fkey(mn); keys_out[(size_t)clip * LB_MM_SLOTS * 2 + 1] = lb_fkey(mx);
    }
    const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);
    const float range = fmaxf(__fsub_rn(amax, amin), 1e-5f);
    const float scale = __fdiv_rn(range, 255.0f);
    const float zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
    const float inv = __fdiv_rn(1.0f, scale);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < RR; ++i) {                               // register rows: lane holds elements 32 j + lane -> byte stores, 32 B per warp store
        const int r = warp + i * LNQ_WARPS;
        if (r < n_reg) {
            const long long row = (long long)clip * T + r_begin + r;
            uint8_t* dst = a_u8 + row * N + lane;
            int sum = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const unsigned q = lb_q8(__fmaf_rn(yr[i][j], inv, zp));
                sum += (int)q;
                dst[32 * j] = (uint8_t)q;
            }
            sum = lb_warp_sum_i(sum);
            if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
        }
    }
    for (int r = warp; r < n_smem; r += LNQ_WARPS) {
        const float4* y4 = reinterpret_cast<const float4*>(lnq_rows + (size_t)r * N);
        const long long row = (long long)clip * T + r_begin + n_reg + r;
        unsigned* a4 = reinterpret_cast<unsigned*>(a_u8 + row * N);
        int sum = 0;
#pragma unroll
        for (int j = 0; j < N / 128; ++j) {
            const float4 y = y4[lane + 32 * j];
            a4[lane + 32 * j] = lb_q8x4(__fmaf_rn(y.x, inv, zp), __fmaf_rn(y.y, inv, zp), __fmaf_rn(y.z, inv, zp), __fmaf_rn(y.w, inv, zp), sum);
        }
        sum = lb_warp_sum_i(sum);
        if (lane == 0) { rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
    }
}

static int lnq_cluster_size() {
    static int cs = 0;
    if (!cs) { const char* e = getenv("LELE_B200_LNQ_CS"); cs = (e && e[0] == '4') ? 4 : 8; }
    return cs;
}
bool lb_layer_norm_quantize_cluster_supported(int n, int T) {
    return n == 512 && T >= 1 && (T + lnq_cluster_size() - 1) / lnq_cluster_size() <= 100;   // 100 rows x 2 KB = 200 KB of shared memory, <= LNQ_MAX_CHUNKS chunks
}
// x [clips*T, 512] -> u8 rows + row sums + per-row (scale, zp); one launch, x read once
int lb_layer_norm_quantize_cluster(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, int clips, int T, float eps,
                                   uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp, unsigned* keys_out) {
    LB_REQUIRE(lb_layer_norm_quantize_cluster_supported(512, T) && gamma && beta, "layer_norm_quantize_cluster: unsupported shape");
    if (clips == 0) return LELE_B200_OK;
    const int CS = lnq_cluster_size();
    static const int dbg = getenv("LELE_B200_LNQ_DBG") ? 1 : 0;
    // register-row variant (CS = 8): 4 warps x 2 register rows, the other rows in shared memory -> 4 CTAs per SM, one wave
    if (CS == 8 && lb_env_flag("LELE_B200_LNQ_REG", 0)) {   // opt-in: the kernel itself is faster (18.6 vs 22 us) but the replayed step is not (PDL overlap), so the default stays
        const int R = (T + CS - 1) / CS;
        const size_t smem_r = (size_t)(R > 8 ? R - 8 : 0) * 512 * 4;
        if (smem_r <= 56 * 1024 && (R > 8 ? R - 8 : 0) <= LNQ_CHUNK * LNQ_MAX_CHUNKS) {
            { int rc_a = lb_func_smem(ctx, (const void*)ln_quant_cluster_reg_kernel<8, 4, 4, 2>, smem_r ? smem_r : 16); if (rc_a) return rc_a; }
            LB_CHECK_CUDA(lb_launch_pdl(ln_quant_cluster_reg_kernel<8, 4, 4, 2>, dim3(CS, clips, 1), dim3(4 * 32), smem_r, ctx->stream, CS, x, gamma, beta, T, eps, a_u8, rowsum, row_scale, row_zp, keys_out));
            LB_LAUNCH_CHECK(ctx);
            return LELE_B200_OK;
        }
    }
    const size_t smem = (size_t)((T + CS - 1) / CS) * 512 * 4;
    {
        int rc_a = CS == 4 ? lb_func_smem(ctx, (const void*)ln_quant_cluster_kernel<4, 24, 1>, smem) : lb_func_smem(ctx, (const void*)ln_quant_cluster_kernel<8, 8, 3>, smem);
        if (rc_a) return rc_a;
    }
    if (CS == 4) LB_CHECK_CUDA(lb_launch_pdl(ln_quant_cluster_kernel<4, 24, 1>, dim3(CS, clips, 1), dim3(24 * 32), smem, ctx->stream, CS, x, gamma, beta, T, eps, a_u8, rowsum, row_scale, row_zp, keys_out, dbg));
    else LB_CHECK_CUDA(lb_launch_pdl(ln_quant_cluster_kernel<8, 8, 3>, dim3(CS, clips, 1), dim3(8 * 32), smem, ctx->stream, CS, x, gamma, beta, T, eps, a_u8, rowsum, row_scale, row_zp, keys_out, dbg));
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}


// ---- streaming variant (round 2): ONE persistent CTA per SM, no clusters, no wave quantisation --------------------------------
// The cluster kernel above holds a clip's rows on chip until the clip's min / max is known, which costs a second (almost empty) wave
// and lock-step load -> normalise -> exchange -> quantise phases (22 us for 45 MB).  Here the rows STREAM: a CTA walks 16-row units
// (unit u -> CTA u % grid, 4 bulk-copy buffers in flight), and each round runs
//     A(i):   normalise unit i in place in shared memory (one row per warp, the reference's accumulator order), reduce the unit's
//             per-clip min / max (warp -> shared-memory atomics -> one pair of global atomics per CTA and clip) and ARRIVE on the clip;
//     B(i-1): once the counters of the row's clip show every unit that holds rows of it (acquire load; the other CTAs finished their
//             A(i-1) a whole round ago), derive (scale, zp) and quantise unit i-1 out of shared memory.
// Same arithmetic, operation for operation, as ln_quant_cluster_kernel (bit-identical outputs).  Every CTA must be resident (grid <=
// #SMs; one CTA per SM by shared memory), so the runner uses it only when the forward owns the device (not inside a clip lane).
constexpr int LQS_U = 16;                 // rows per unit == warps per CTA
constexpr int LQS_NBUF = 6;               // 192 KB: units i-2, i-1 (normalised, waiting for their clips), i (being normalised), i+1 .. i+3 (in flight)
constexpr int LQS_LAG = 2;                // a unit is quantised LQS_LAG rounds after it was normalised: the clip's parameters have long arrived
constexpr int LQS_SMEM = LQS_NBUF * LQS_U * 512 * 4;

__global__ void __launch_bounds__(LQS_U * 32 + 64, 1)
ln_quant_stream_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long M, int T, float eps,
                       uint8_t* __restrict__ a_u8, int32_t* __restrict__ rowsum, float* __restrict__ row_scale, int32_t* __restrict__ row_zp,
                       uint4* __restrict__ records, int rec_stride, int dbg) {
    constexpr int N = 512, NB = 16, U = LQS_U;
    long long w_full = 0, t_a = 0, w_cnt = 0, t_b = 0, w_free = 0; const long long t_begin = clock64();
    extern __shared__ __align__(128) float lqs_rows[];           // [NBUF][U][512]
    __shared__ __align__(8) unsigned long long full_bar[LQS_NBUF];
    __shared__ unsigned slot_keys[2][2][2];                      // [round parity][clip slot of the unit][min, max]
    // warp 17 is the bulk-copy PRODUCER, warp 16 the PUBLISHER: the global atomics + fence + arrival of a round cost ~1-2 us of latency, which must not sit in a
    // compute warp between two block barriers.  slots_full[p]: 16 compute warps -> publisher; slots_free[p]: publisher -> compute warps.
    __shared__ __align__(8) unsigned long long slots_full[2], slots_free[2], buf_empty[LQS_NBUF];   // buf_empty[b]: 16 compute warps quantised the unit in buffer b
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_units = (M + U - 1) / U;
    const int n_my = (int)((n_units - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x);   // units blockIdx.x + i * gridDim.x
    if (threadIdx.x == 0) {
        for (int b = 0; b < LQS_NBUF; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full_bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 4; ++i) { slot_keys[i >> 1][i & 1][0] = LB_KEY_MIN_INIT; slot_keys[i >> 1][i & 1][1] = LB_KEY_MAX_INIT; }
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slots_full[b])), "r"(LQS_U));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slots_free[b])));
        }
        for (int b = 0; b < LQS_NBUF; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&buf_empty[b])), "r"(LQS_U));
    }
    __syncthreads();
    lb_pdl_launch_dependents();
    lb_pdl_wait();                 // PDL: x is the previous kernel's output
    auto mbar_wait_par = [&](unsigned long long* b, uint32_t par) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(b);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    };
    auto mbar_arrive1 = [&](unsigned long long* b) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
    };
    auto issue = [&](int i) {      // producer lane: bulk copy of the CTA's i-th unit into buffer i % NBUF
        if (i >= n_my) return;
        const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * U;
        const uint32_t bytes = (uint32_t)min((long long)U, M - r0) * N * 4;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[i % LQS_NBUF]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the buffer's last generic-proxy accesses precede this copy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(lqs_rows + (size_t)(i % LQS_NBUF) * U * N)), "l"(x + r0 * N), "r"(bytes), "r"(bar) : "memory");
    };
    if (warp == LQS_U + 1) {
        // ===================== producer warp: keeps LQS_NBUF units in flight =====================
        if (lane == 0) {
            for (int i = 0; i < LQS_NBUF; ++i) issue(i);
            for (int i = LQS_NBUF; i < n_my; ++i) {              // unit i - NBUF has been quantised by every warp: refill its buffer
                mbar_wait_par(&buf_empty[i % LQS_NBUF], (uint32_t)(i / LQS_NBUF - 1) & 1u);
                issue(i);
            }
        }
        return;
    }
    if (warp == LQS_U) {
        // ===================== publisher warp =====================
        for (int i = 0; i < n_my; ++i) {
            const int p = i & 1;
            mbar_wait_par(&slots_full[p], (uint32_t)(i >> 1) & 1u);          // every compute warp's keys of round i are in the slots
            if (lane < 2) {
                const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * U;
                const long long r_end = min(r0 + U, M);
                const int clip = (int)(r0 / T) + lane;
                unsigned* sk = slot_keys[p][lane];
                const unsigned mn = sk[0], mx = sk[1];
                sk[0] = LB_KEY_MIN_INIT; sk[1] = LB_KEY_MAX_INIT;
                if ((long long)clip * T < r_end) {
                    // the unit holds rows of this clip: publish its (min, max) as ONE self-validating 16-byte record {min key, 1, max key, 1}
                    // in the clip's record row -- no atomics, no fence, no counter: the flag travels with the data (each 8-byte half
                    // carries its own flag), and a quantising warp polls the whole row with a single load per lane
                    const long long unit = (long long)blockIdx.x + (long long)i * gridDim.x;
                    const int j = (int)(unit - ((long long)clip * T) / U);
                    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(records + (size_t)clip * rec_stride + j), "r"(mn), "r"(1u), "r"(mx), "r"(1u) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive1(&slots_free[p]);                     // the slots may take round i + 2
        }
        return;
    }
    float g[NB], bt[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { g[i] = __ldg(gamma + 32 * i + lane); bt[i] = __ldg(beta + 32 * i + lane); }
    const float inv_n = __fdiv_rn(1.0f, (float)N);

    // records of the clip of the warp's row in unit i, requested early (consumed by quantise_unit(i) after the next normalisation)
    uint4 pf_rec = make_uint4(LB_KEY_MIN_INIT, 1u, LB_KEY_MAX_INIT, 1u); bool pf_valid = false;
    auto prefetch_records = [&](int i) {
        pf_valid = false;
        const long long row = ((long long)blockIdx.x + (long long)i * gridDim.x) * U + warp;
        if (i < 0 || i >= n_my || row >= M) return;
        const int clip = (int)(row / T);
        const long long c0 = (long long)clip * T;
        const int n_rec = (int)((c0 + T - 1) / U - c0 / U + 1);
        pf_rec = make_uint4(LB_KEY_MIN_INIT, 1u, LB_KEY_MAX_INIT, 1u);
        if (lane < n_rec) asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pf_rec.x), "=r"(pf_rec.y), "=r"(pf_rec.z), "=r"(pf_rec.w)
                                       : "l"(records + (size_t)clip * rec_stride + lane) : "memory");
        pf_valid = true;
    };
    // B(i): quantise the warp's row of the CTA's i-th unit out of shared memory
    auto quantise_unit = [&](int i) {
        const long long row = ((long long)blockIdx.x + (long long)i * gridDim.x) * U + warp;
        if (row >= M) return;
        const int clip = (int)(row / T);
        const long long t_b0 = dbg ? clock64() : 0;
        // the clip's min / max = reduction over the records of every unit that holds rows of it (lane = record); the first 32 records
        // were requested before this round's normalisation (pf_rec), so their L2 round trip is already behind us
        const long long c0 = (long long)clip * T;
        const int n_rec = (int)((c0 + T - 1) / U - c0 / U + 1);
        unsigned kmn = LB_KEY_MIN_INIT, kmx = LB_KEY_MAX_INIT;
        for (int jb = 0; jb < n_rec; jb += 32) {
            const uint4* rp = records + (size_t)clip * rec_stride + jb + lane;
            const bool mine = jb + lane < n_rec;
            uint4 rc = make_uint4(LB_KEY_MIN_INIT, 1u, LB_KEY_MAX_INIT, 1u);
            if (jb == 0 && pf_valid) rc = pf_rec;
            else if (mine) asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rc.x), "=r"(rc.y), "=r"(rc.z), "=r"(rc.w) : "l"(rp) : "memory");
            if (!__all_sync(0xffffffffu, rc.y == 1u && rc.w == 1u)) {
                const long long t0 = clock64();
                do {
                    __nanosleep(32);
                    if (mine) asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rc.x), "=r"(rc.y), "=r"(rc.z), "=r"(rc.w) : "l"(rp) : "memory");
                    if (clock64() - t0 > 4000000000ll) { if (lane == 0) printf("lele_b200 ln_quant_stream: clip %d never completed (block %d)\n", clip, blockIdx.x); __trap(); }
                } while (!__all_sync(0xffffffffu, rc.y == 1u && rc.w == 1u));
            }
            kmn = min(kmn, __reduce_min_sync(0xffffffffu, rc.x)); kmx = max(kmx, __reduce_max_sync(0xffffffffu, rc.z));
        }
        if (dbg) w_cnt += clock64() - t_b0;
        const float mn = lb_fkey_inv(kmn), mx = lb_fkey_inv(kmx);
        const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);
        const float range = fmaxf(__fsub_rn(amax, amin), 1e-5f);
        const float scale = __fdiv_rn(range, 255.0f);
        const float zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
        const float inv = __fdiv_rn(1.0f, scale);
        const float4* y4 = reinterpret_cast<const float4*>(lqs_rows + ((size_t)(i % LQS_NBUF) * U + warp) * N);
        unsigned* a4 = reinterpret_cast<unsigned*>(a_u8 + row * N);
        int sum = 0;
#pragma unroll
        for (int j = 0; j < N / 128; ++j) {
            const float4 y = y4[lane + 32 * j];
            a4[lane + 32 * j] = lb_q8x4(__fmaf_rn(y.x, inv, zp), __fmaf_rn(y.y, inv, zp), __fmaf_rn(y.z, inv, zp), __fmaf_rn(y.w, inv, zp), sum);
        }
        if (rowsum) sum = lb_warp_sum_i(sum);
        if (lane == 0) { if (rowsum) rowsum[row] = sum; row_scale[row] = scale; row_zp[row] = (int)zp; }
        if (dbg) t_b += clock64() - t_b0;
    };

    for (int i = 0; i < n_my; ++i) {
        // ---- A(i): normalise the unit in place, reduce its per-clip min / max ----
        const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * U;
        const long long row = r0 + warp;
        const int clip0 = (int)(r0 / T);
        prefetch_records(i - LQS_LAG);
        const long long t_a0 = dbg ? clock64() : 0;
        {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[i % LQS_NBUF]);
            const uint32_t par = (uint32_t)(i / LQS_NBUF) & 1u;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar), "r"(par) : "memory");
        }
        if (dbg) w_full += clock64() - t_a0;
        if (row < M) {
            float* o = lqs_rows + ((size_t)(i % LQS_NBUF) * U + warp) * N;
            float v[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) v[j] = o[32 * j + lane];
            float ps = 0.0f, pq = 0.0f;
#pragma unroll
            for (int j = 0; j < NB; ++j) { ps = __fadd_rn(ps, v[j]); pq = __fmaf_rn(v[j], v[j], pq); }
#pragma unroll
            for (int of = 8; of <= 16; of <<= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
#pragma unroll
            for (int of = 4; of >= 1; of >>= 1) { ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, of)); pq = __fadd_rn(pq, __shfl_xor_sync(0xffffffffu, pq, of)); }
            ps = __shfl_sync(0xffffffffu, ps, 0); pq = __shfl_sync(0xffffffffu, pq, 0);
            const float mean = __fmul_rn(ps, inv_n);
            const float var = __fsub_rn(__fmul_rn(pq, inv_n), __fmul_rn(mean, mean));
            const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
            float vmin = 3.402823466e+38f, vmax = -3.402823466e+38f;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const float y = __fmaf_rn(__fmul_rn(__fsub_rn(v[j], mean), inv), g[j], bt[j]);
                vmin = fminf(vmin, y); vmax = fmaxf(vmax, y);
                o[32 * j + lane] = y;
            }
            const unsigned kvmin = __reduce_min_sync(0xffffffffu, lb_fkey(vmin)), kvmax = __reduce_max_sync(0xffffffffu, lb_fkey(vmax));
            if (lane == 0) {
                const long long t_f0 = dbg ? clock64() : 0;
                if (i >= 2) mbar_wait_par(&slots_free[i & 1], (uint32_t)((i >> 1) - 1) & 1u);   // the publisher has drained round i - 2
                if (dbg) w_free += clock64() - t_f0;
                unsigned* sk = slot_keys[i & 1][(int)(row / T) - clip0];
                atomicMin(sk, kvmin); atomicMax(sk + 1, kvmax);
            }
        }
        __syncwarp();                                            // the row is normalised in place (the same warp quantises it next round)
        if (lane == 0) mbar_arrive1(&slots_full[i & 1]);         // release: this warp's keys (if any) are in the slots
        if (dbg) t_a += clock64() - t_a0;
        // ---- B(i-LAG): row `warp` of an earlier unit (no block barrier anywhere: a row belongs to one warp from copy to store) ----
        if (i >= LQS_LAG) {
            quantise_unit(i - LQS_LAG);
            __syncwarp();
            if (lane == 0) mbar_arrive1(&buf_empty[(i - LQS_LAG) % LQS_NBUF]);   // release: this warp is done with the buffer
        }
    }
    for (int i = max(n_my - LQS_LAG, 0); i < n_my; ++i) { pf_valid = false; quantise_unit(i); }
    if (dbg && lane == 0 && (warp == 0 || warp == 15) && (blockIdx.x == 0 || blockIdx.x == 77))
        printf("LQSDBG blk %d warp %d: total %lld | A %lld (wait copy %lld, wait publisher %lld) | B %lld (wait clip %lld) | units %d\n", blockIdx.x, warp, clock64() - t_begin,
               t_a, w_full, w_free, t_b, w_cnt, n_my);
}

bool lb_layer_norm_quantize_stream_supported(lele_b200_ctx* ctx, int n, int T) { return n == 512 && T >= LQS_U && ctx->num_sms >= 1; }
int lb_layer_norm_quantize_stream_records(int T) { return (T + LQS_U - 1) / LQS_U + 1; }   // 16-byte records per clip (units that can hold rows of one clip)
// x [clips*T, 512] -> u8 rows (+ row sums when rowsum != NULL) + per-row (scale, zp); records = [clips][rec_stride] ZEROED 16-byte records
// (rec_stride >= lb_layer_norm_quantize_stream_records(T)) through which the CTAs exchange the per-clip min / max
int lb_layer_norm_quantize_stream(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, int clips, int T, float eps,
                                  uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp, void* records, int rec_stride) {
    LB_REQUIRE(lb_layer_norm_quantize_stream_supported(ctx, 512, T) && gamma && beta && records && (((uintptr_t)records) & 15) == 0 && rec_stride >= lb_layer_norm_quantize_stream_records(T), "layer_norm_quantize_stream: unsupported shape");
    if (clips == 0) return LELE_B200_OK;
    const long long M = (long long)clips * T;
    const long long n_units = (M + LQS_U - 1) / LQS_U;
    const int grid = (int)(n_units < ctx->num_sms ? n_units : ctx->num_sms);
    { int rc_a = lb_func_smem(ctx, (const void*)ln_quant_stream_kernel, LQS_SMEM); if (rc_a) return rc_a; }
    LB_CHECK_CUDA(lb_launch_pdl(ln_quant_stream_kernel, dim3(grid), dim3(LQS_U * 32 + 64), (size_t)LQS_SMEM, ctx->stream, 1, x, gamma, beta, M, T, eps, a_u8, rowsum, row_scale,
                                row_zp, reinterpret_cast<uint4*>(records), rec_stride, (int)(getenv("LELE_B200_LNQ_DBG") ? 1 : 0)));
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

bool lb_layer_norm_quantize_supported(int n, int rows_per_slice) { return n == 512 && rows_per_slice >= 4; }

// pass 1: statistics + min/max keys (keys must be initialised by the caller)
int lb_layer_norm_stats(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer, int n, float eps,
                        float* stats, unsigned* minmax_keys, int rows_per_slice) {
    LB_REQUIRE(lb_layer_norm_quantize_supported(n, rows_per_slice) && gamma && beta && minmax_keys, "layer_norm_stats: unsupported shape");
    if (outer == 0) return LELE_B200_OK;
    ln_stats512_kernel<2><<<lb_ceil_div(outer, 8 * 2), 256, 0, ctx->stream>>>(x, gamma, beta, outer, eps, reinterpret_cast<float2*>(stats), minmax_keys,
                                                                              rows_per_slice);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
// pass 2: u8 rows + row sums + per-row (scale, zero point)
int lb_layer_norm_quantize(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer, int n,
                           const float* stats, const unsigned* keys, int rows_per_slice, uint8_t* a_u8, int32_t* rowsum,
                           float* row_scale, int32_t* row_zp) {
    LB_REQUIRE(lb_layer_norm_quantize_supported(n, rows_per_slice), "layer_norm_quantize: unsupported shape");
    if (outer == 0) return LELE_B200_OK;
    ln_quantize_kernel<512, 2><<<lb_ceil_div(outer, 8 * 2), 256, 0, ctx->stream>>>(x, gamma, beta, reinterpret_cast<const float2*>(stats), keys, outer,
                                                                                    rows_per_slice, a_u8, rowsum, row_scale, row_zp);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_layer_norm(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta,
                                    long long outer, int n, float eps, float* out) {
    LB_REQUIRE(ctx && x && out && n > 0 && outer >= 0, "layer_norm: bad arguments");
    LB_ENTER(ctx);
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
        o[j] = j < simd_end ? lb_cephes_expf(d) : lb_libm_expf(d);
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
    LB_ENTER(ctx);
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
    LB_ENTER(ctx);
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
    LB_ENTER(ctx);
    if (outer == 0) return LELE_B200_OK;
    rms_norm_kernel<<<lb_ceil_div(outer, 8), 256, 0, ctx->stream>>>(x, w, outer, n, eps, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
