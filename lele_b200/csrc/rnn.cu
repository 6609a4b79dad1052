// rnn.cu -- ONNX LSTM (gates i,o,f,c) and GRU (gates z,r,h), forward direction, batch 1 per
// sequence (src/kernels/rnn.rs:67-230, 246-357).  B200 shape of the computation:
//   1. W.x_t for ALL steps of ALL sequences hoisted into one GEMM (no dependence on h),
//   2. one persistent CTA per sequence walks the steps: R.h (R pre-transposed so the per-gate
//      dot products read coalesced), then the gate math with lele's x86 polynomial
//      sigmoid/tanh on the first H/8*8 units and libm on the tail (rnn.rs:15-64, 360-432).
// n_seq independent sequences run concurrently (config 3 "batch=16" = 16 sequences).
#include "common.cuh"
#include <stdlib.h>

int lb_sgemm_strided(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                     long long rsb, long long csb, long long bsb, float* C, int batch, int m, int k, int n, float alpha,
                     int pre_mode);

namespace {
__global__ void transpose_small_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    long long total = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), c = (int)(i % cols);
        out[(long long)c * rows + r] = in[i];
    }
}

// gate math of one step for hidden unit k (shared by the two sequence kernels): wc = W.x_t, rc = R.h_{t-1}
template <int G>
__device__ __forceinline__ float rnn_gate_step(int k, int H, int GH, int simd_end, const float* __restrict__ bias, const float* wc, const float* rc,
                                               float* h, float* c) {
    const bool sd = k < simd_end;
    float ht;
    if (G == 4) {   // rnn.rs:156-158 then lstm_gates_avx2
        float g4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int gi = q * H + k;
            float bw = bias ? bias[gi] : 0.0f, br = bias ? bias[GH + gi] : 0.0f;
            g4[q] = __fadd_rn(__fadd_rn(__fadd_rn(wc[gi], rc[gi]), bw), br);
        }
        float ig = sd ? lb_sigmoid_simd(g4[0]) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-g4[0])));
        float og = sd ? lb_sigmoid_simd(g4[1]) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-g4[1])));
        float fg = sd ? lb_sigmoid_simd(g4[2]) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-g4[2])));
        float cg = sd ? lb_tanh_simd(g4[3]) : tanhf(g4[3]);
        float ct = sd ? __fmaf_rn(fg, c[k], __fmul_rn(ig, cg)) : __fadd_rn(__fmul_rn(fg, c[k]), __fmul_rn(ig, cg));
        ht = __fmul_rn(og, sd ? lb_tanh_simd(ct) : tanhf(ct));
        c[k] = ct;
    } else {        // gru_gate_fusion_avx2 rnn.rs:360-432 (linear_before_reset has no effect, :368)
        float bwz = bias ? bias[k] : 0.0f, brz = bias ? bias[GH + k] : 0.0f;
        float bwr = bias ? bias[H + k] : 0.0f, brr = bias ? bias[GH + H + k] : 0.0f;
        float bwh = bias ? bias[2 * H + k] : 0.0f, brh = bias ? bias[GH + 2 * H + k] : 0.0f;
        float zp, rp;
        if (sd) {
            zp = __fadd_rn(__fadd_rn(wc[k], rc[k]), __fadd_rn(bwz, brz));
            rp = __fadd_rn(__fadd_rn(wc[H + k], rc[H + k]), __fadd_rn(bwr, brr));
        } else {
            zp = __fadd_rn(__fadd_rn(__fadd_rn(wc[k], rc[k]), bwz), brz);
            rp = __fadd_rn(__fadd_rn(__fadd_rn(wc[H + k], rc[H + k]), bwr), brr);
        }
        float z = sd ? lb_sigmoid_simd(zp) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-zp)));
        float rg = sd ? lb_sigmoid_simd(rp) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-rp)));
        float hp = __fadd_rn(__fadd_rn(wc[2 * H + k], bwh), __fmul_rn(rg, __fadd_rn(rc[2 * H + k], brh)));
        float hg = sd ? lb_tanh_simd(hp) : tanhf(hp);
        ht = sd ? __fmaf_rn(__fsub_rn(1.0f, z), hg, __fmul_rn(z, h[k]))
                : __fadd_rn(__fmul_rn(__fsub_rn(1.0f, z), hg), __fmul_rn(z, h[k]));
    }
    return ht;
}

// gates: 4 (LSTM) or 3 (GRU)
template <int G>
__global__ void __launch_bounds__(1024)
rnn_seq_kernel(const float* __restrict__ wx /*[n_seq, seq, G*H]*/, const float* __restrict__ rt /*[H, G*H]*/,
               const float* __restrict__ bias /*[2*G*H] or NULL*/, const float* __restrict__ h0, const float* __restrict__ c0,
               int seq, int H, float* __restrict__ y, float* __restrict__ h_out, float* __restrict__ c_out) {
    extern __shared__ float sm[];
    float* h = sm;                 // [H]
    float* c = sm + H;             // [H]   (LSTM only)
    float* wc = sm + 2 * H;        // [G*H] W.x + ... per step
    float* rc = wc + G * H;        // [G*H] R.h
    const int s = blockIdx.x, GH = G * H, simd_end = (H / 8) * 8;
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        h[k] = h0 ? h0[(long long)s * H + k] : 0.0f;
        c[k] = (G == 4 && c0) ? c0[(long long)s * H + k] : 0.0f;
    }
    __syncthreads();
    for (int t = 0; t < seq; ++t) {
        const float* wxt = wx + ((long long)s * seq + t) * GH;
        for (int g = threadIdx.x; g < GH; g += blockDim.x) {
            float acc = 0.0f;
            for (int j = 0; j < H; ++j) acc = fmaf(rt[(long long)j * GH + g], h[j], acc);
            rc[g] = acc;
            wc[g] = wxt[g];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < H; k += blockDim.x) {
            const float ht = rnn_gate_step<G>(k, H, GH, simd_end, bias, wc, rc, h, c);
            // h[k] is only read by thread k in this phase (R.h already done), safe to overwrite
            h[k] = ht;
            y[((long long)s * seq + t) * H + k] = ht;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        h_out[(long long)s * H + k] = h[k];
        if (G == 4) c_out[(long long)s * H + k] = c[k];
    }
}


// Resident-R variant: the recurrence is a chain of seq dependent steps, each a [G*H, H] x [H] product, and with R streamed
// from L2 every step (256 KB for H = 128) the step is bound by one SM's L2 bandwidth (~8 us).  Here thread g keeps its row
// of R for the whole sequence: the first RREG entries in REGISTERS, the rest in shared memory (coalesced, conflict-free), so
// a step is 128 dependent fma + the gate math (~0.5 us).  Same summation order (j ascending) -> identical results.
template <int G, int RREG>
__global__ void __launch_bounds__(512)
rnn_seq_resident_kernel(const float* __restrict__ wx /*[n_seq, seq, G*H]*/, const float* __restrict__ rt /*[H, G*H]*/,
                        const float* __restrict__ bias, const float* __restrict__ h0, const float* __restrict__ c0,
                        int seq, int H, float* __restrict__ y, float* __restrict__ h_out, float* __restrict__ c_out) {
    extern __shared__ __align__(16) float sm[];
    const int GH = G * H, simd_end = (H / 8) * 8;
    float* h = sm;                 // [H]
    float* c = sm + H;             // [H]   (LSTM only)
    float* wc = sm + 2 * H;        // [G*H]
    float* rc = wc + GH;           // [G*H]
    float* rs = rc + GH;           // [(H - RREG), G*H]  rows RREG.. of R^T
    const int s = blockIdx.x, g = threadIdx.x;   // blockDim.x == GH
    float rreg[RREG];
#pragma unroll
    for (int j = 0; j < RREG; ++j) rreg[j] = __ldg(rt + (long long)j * GH + g);
    for (int j = RREG; j < H; ++j) rs[(j - RREG) * GH + g] = __ldg(rt + (long long)j * GH + g);
    for (int k = g; k < H; k += GH) {
        h[k] = h0 ? h0[(long long)s * H + k] : 0.0f;
        c[k] = (G == 4 && c0) ? c0[(long long)s * H + k] : 0.0f;
    }
    __syncthreads();
    for (int t = 0; t < seq; ++t) {
        const float wxv = __ldg(wx + ((long long)s * seq + t) * GH + g);      // in flight under the dot product
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < RREG; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h + j);
            acc = fmaf(rreg[j], hv.x, acc); acc = fmaf(rreg[j + 1], hv.y, acc); acc = fmaf(rreg[j + 2], hv.z, acc); acc = fmaf(rreg[j + 3], hv.w, acc);
        }
        for (int j = RREG; j < H; ++j) acc = fmaf(rs[(j - RREG) * GH + g], h[j], acc);
        rc[g] = acc;
        wc[g] = wxv;
        __syncthreads();
        if (g < H) {
            const float ht = rnn_gate_step<G>(g, H, GH, simd_end, bias, wc, rc, h, c);
            h[g] = ht;
            y[((long long)s * seq + t) * H + g] = ht;
        }
        __syncthreads();
    }
    if (g < H) {
        h_out[(long long)s * H + g] = h[g];
        if (G == 4) c_out[(long long)s * H + g] = c[g];
    }
}

template <int G>
int run_rnn(lele_b200_ctx* ctx, const float* x, const float* w, const float* r, const float* bias, const float* h0,
            const float* c0, int n_seq, int seq, int in_size, int H, float* y, float* h, float* c) {
    const int GH = G * H;
    if (n_seq == 0) return LELE_B200_OK;
    size_t wx_bytes = sizeof(float) * (size_t)n_seq * seq * GH, rt_bytes = sizeof(float) * (size_t)H * GH;
    void* sc;
    int rc = lb_scratch(ctx, wx_bytes + rt_bytes + 512, &sc);
    if (rc) return rc;
    float* wx = (float*)sc;
    float* rt = (float*)((uint8_t*)sc + ((wx_bytes + 255) / 256) * 256);
    if (seq > 0) {
        // wx[n_seq*seq, GH] = X[n_seq*seq, in] x W^T   (W is [GH, in] row-major)
        rc = lb_sgemm_strided(ctx, x, in_size, 1, 0, w, 1, in_size, 0, wx, 1, n_seq * seq, in_size, GH, 1.0f, 0);
        if (rc) return rc;
    }
    transpose_small_kernel<<<lb_ceil_div((long long)GH * H, 256), 256, 0, ctx->stream>>>(r, GH, H, rt);
    LB_LAUNCH_CHECK(ctx);
    {   // resident-R kernel: one thread per gate row, R^T rows 0..63 in registers, the rest in shared memory
        constexpr int RREG = 64;
        const size_t smem_r = sizeof(float) * ((size_t)(2 * H + 2 * GH) + (size_t)(H > RREG ? H - RREG : 0) * GH);
        if (H >= RREG && H % 4 == 0 && GH <= 512 && GH % 32 == 0 && smem_r <= 200 * 1024 && !getenv("LELE_B200_RNN_STREAM_R")) {
            if ((rc = lb_func_smem(ctx, (const void*)rnn_seq_resident_kernel<G, RREG>, smem_r))) return rc;
            rnn_seq_resident_kernel<G, RREG><<<n_seq, GH, smem_r, ctx->stream>>>(wx, rt, bias, h0, c0, seq, H, y, h, c);
            LB_LAUNCH_CHECK(ctx);
            return LELE_B200_OK;
        }
    }
    int threads = GH < 1024 ? ((GH + 31) / 32) * 32 : 1024;
    size_t smem = sizeof(float) * (size_t)(2 * H + 2 * GH);
    LB_REQUIRE(smem <= 200 * 1024, "rnn: hidden size %d too large", H);
    if (smem > 48 * 1024 && (rc = lb_func_smem(ctx, (const void*)rnn_seq_kernel<G>, smem))) return rc;
    rnn_seq_kernel<G><<<n_seq, threads, smem, ctx->stream>>>(wx, rt, bias, h0, c0, seq, H, y, h, c);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
}  // namespace

extern "C" int lele_b200_lstm(lele_b200_ctx* ctx, const float* x, const float* w, const float* r, const float* bias,
                              const float* h0, const float* c0, int n_seq, int seq, int in_size, int hidden, float* y, float* h,
                              float* c) {
    LB_REQUIRE(ctx && x && w && r && y && h && c, "lstm: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(hidden > 0 && in_size > 0 && seq >= 0 && n_seq >= 0, "lstm: bad dims");
    return run_rnn<4>(ctx, x, w, r, bias, h0, c0, n_seq, seq, in_size, hidden, y, h, c);
}
extern "C" int lele_b200_gru(lele_b200_ctx* ctx, const float* x, const float* w, const float* r, const float* bias,
                             const float* h0, int n_seq, int seq, int in_size, int hidden, float* y, float* h) {
    LB_REQUIRE(ctx && x && w && r && y && h, "gru: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(hidden > 0 && in_size > 0 && seq >= 0 && n_seq >= 0, "gru: bad dims");
    return run_rnn<3>(ctx, x, w, r, bias, h0, nullptr, n_seq, seq, in_size, hidden, y, h, nullptr);
}
