// frontend.cu -- lele::features on B200: fused fbank (scale, mean-sub, pre-emphasis, Hann,
// 512-pt real FFT, power, sparse 80-mel, log) + LFR stacking; CMVN; generic rFFT / STFT.
// Reference: src/features/pipeline.rs:67-193, fft.rs, mel.rs, lfr.rs, cmvn.rs,
// src/kernels/fft.rs:51-168, src/kernels/math.rs:2304-2439.
//
// Roofline class: HBM-bound (1.024 MB read + 0.598 MB written per 16 s clip, ~43 MFLOP):
// CUDA cores only, one warp per frame, the FFT lives in shared memory.
#include "common.cuh"
#include <math.h>

#define LB_PI_F 3.14159265358979323846f

// ---------------------------------------------------------------------------
// host-side constant tables (identical arithmetic to the reference: f32 cosf/sinf/log10f/powf)
// ---------------------------------------------------------------------------
static void host_hann(int size, float* out) {  // window.rs:2-12
    if (size == 0) return;
    if (size == 1) { out[0] = 1.0f; return; }
    for (int n = 0; n < size; ++n) out[n] = 0.5f * (1.0f - cosf(2.0f * LB_PI_F * (float)n / (float)(size - 1)));
}
static int host_log2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
static void host_twiddles(int n, float* tw_re, float* tw_im, int* bit_rev) {  // kernels/fft.rs:136-168
    int log2n = host_log2(n);
    for (int i = 0; i < n; ++i) {
        int r = 0, x = i;
        for (int b = 0; b < log2n; ++b) { r = (r << 1) | (x & 1); x >>= 1; }
        bit_rev[i] = r;
    }
    int o = 0;
    for (int size = 2; size <= n; size *= 2) {
        int half = size / 2, step = n / size;
        for (int k = 0; k < half; ++k) {
            float angle = -2.0f * LB_PI_F * (float)(k * step) / (float)n;
            tw_re[o] = cosf(angle);
            tw_im[o] = sinf(angle);
            ++o;
        }
    }
}
static float host_hz_to_mel(float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); }
static float host_mel_to_hz(float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); }
static void host_mel_filterbank(float sr, int n_fft, int n_mels, float f_min, float f_max, float* w) {  // mel.rs:7-45
    int n_freqs = n_fft / 2 + 1, pts = n_mels + 2;
    float mel_min = host_hz_to_mel(f_min), mel_max = host_hz_to_mel(f_max);
    float mel_step = (mel_max - mel_min) / (float)(n_mels + 1);
    std::vector<float> hz(pts);
    for (int i = 0; i < pts; ++i) hz[i] = host_mel_to_hz(mel_min + (float)i * mel_step);
    for (int i = 0; i < n_mels; ++i) {
        float fl = hz[i], fc = hz[i + 1], fr = hz[i + 2];
        for (int j = 0; j < n_freqs; ++j) {
            float f = (float)j * sr / (float)n_fft, val = 0.0f;
            if (f > fl && f < fc) val = (f - fl) / (fc - fl);
            else if (f >= fc && f < fr) val = (fr - f) / (fr - fc);
            w[i * n_freqs + j] = val;
        }
    }
}

extern "C" int lele_b200_hann_window(int size, float* out_host) {
    LB_REQUIRE(size >= 0 && (out_host || size == 0), "hann_window: bad arguments");
    host_hann(size, out_host);
    return LELE_B200_OK;
}
extern "C" int lele_b200_mel_filterbank(float sr, int n_fft, int n_mels, float f_min, float f_max, float* out_host) {
    LB_REQUIRE(out_host && n_fft > 0 && n_mels > 0, "mel_filterbank: bad arguments");
    host_mel_filterbank(sr, n_fft, n_mels, f_min, f_max, out_host);
    return LELE_B200_OK;
}
extern "C" int lele_b200_frontend_num_frames(int n_samples) { return n_samples < 400 ? 0 : (n_samples - 400) / 160 + 1; }
extern "C" int lele_b200_frontend_out_rows(int n_samples) { return (lele_b200_frontend_num_frames(n_samples) + 5) / 6; }

// ---------------------------------------------------------------------------
// radix-2 DIT FFT over shared memory (same butterfly network as kernels/fft.rs:79-134)
// ---------------------------------------------------------------------------
template <bool kWarp>
__device__ __forceinline__ void fft_sync() {
    if (kWarp) __syncwarp(); else __syncthreads();
}
// re/im: n floats each in shared memory, input already in bit-reversed order.
template <bool kWarp>
__device__ __forceinline__ void fft_radix2(float* re, float* im, int n, const float* __restrict__ tw_re,
                                           const float* __restrict__ tw_im, int tid, int nthreads) {
    int tw_off = 0;
    for (int hs = 1; hs < n; hs <<= 1) {
        fft_sync<kWarp>();
        for (int b = tid; b < (n >> 1); b += nthreads) {
            int k = b & (hs - 1);
            int e = ((b - k) << 1) + k, o = e + hs;
            float wr = __ldg(tw_re + tw_off + k), wi = __ldg(tw_im + tw_off + k);
            float ore = re[o], oim = im[o];
            float tr = wr * ore - wi * oim;
            float ti = wr * oim + wi * ore;
            float ere = re[e], eim = im[e];
            re[o] = ere - tr; im[o] = eim - ti;
            re[e] = ere + tr; im[e] = eim + ti;
        }
        tw_off += hs;
    }
    fft_sync<kWarp>();
}

// ---------------------------------------------------------------------------
// fused fbank + LFR: one warp per frame (pipeline.rs:85-190 + lfr.rs:18-54)
// ---------------------------------------------------------------------------
struct FbankTables {
    const float* window;   // [400]
    const float* tw_re;    // [511] stage-major
    const float* tw_im;
    const int* bit_rev;    // [512]
    const int* mel_start;  // [80]
    const int* mel_len;    // [80]
    const int* mel_off;    // [80] offset into mel_w
    const float* mel_w;    // packed non-zero spans
};

constexpr int FB_FL = 400, FB_NF = 512, FB_HOP = 160, FB_NM = 80, FB_WARPS = 8;

__global__ void __launch_bounds__(FB_WARPS * 32)
fbank_lfr_kernel(const float* __restrict__ pcm, long long clip_stride, int n_clips, int frames, int t_lfr,
                 FbankTables tb, float* __restrict__ mel_opt, float* __restrict__ lfr_out) {
    __shared__ float s_re[FB_WARPS][FB_NF];
    __shared__ float s_im[FB_WARPS][FB_NF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gframe = (long long)blockIdx.x * FB_WARPS + warp;
    if (gframe >= (long long)n_clips * frames) return;   // whole warp exits together
    const int clip = (int)(gframe / frames), f = (int)(gframe % frames);
    float* re = s_re[warp];
    float* im = s_im[warp];
    const float* p = pcm + (long long)clip * clip_stride + (long long)f * FB_HOP;

    // 1. scale (x32768) and frame mean
    float x[13];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        int j = lane + 32 * i;
        x[i] = j < FB_FL ? __fmul_rn(__ldg(p + j), 32768.0f) : 0.0f;
        sum += x[i];
    }
    sum = lb_warp_sum(sum);
    const float mean = __fdiv_rn(sum, (float)FB_FL);
    // 2. mean subtraction; stash in im[] so the neighbour for pre-emphasis is readable
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        int j = lane + 32 * i;
        x[i] = __fsub_rn(x[i], mean);
        if (j < FB_FL) im[j] = x[i];
    }
    __syncwarp();
    // 3. pre-emphasis (backward loop in the reference => uses the un-emphasised neighbour),
    // 4. Hann window, 5. scatter into bit-reversed order, zero-pad 400..511
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int j = lane + 32 * i;
        float v = 0.0f;
        if (j < FB_FL) {
            float cur = im[j];
            if (j >= 1) cur = __fsub_rn(cur, __fmul_rn(0.97f, im[j - 1]));
            v = __fmul_rn(cur, __ldg(tb.window + j));
        }
        re[__ldg(tb.bit_rev + j)] = v;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) im[lane + 32 * i] = 0.0f;
    // 6. FFT
    fft_radix2<true>(re, im, FB_NF, tb.tw_re, tb.tw_im, lane, 32);
    // 7. power spectrum for bins 0..256 (Im(0) = Im(256) = 0 as in kernels/fft.rs:124-128)
    for (int k = lane; k <= 256; k += 32) {
        float r = re[k], q = (k == 0 || k == 256) ? 0.0f : im[k];
        re[k] = __fadd_rn(__fmul_rn(r, r), __fmul_rn(q, q));
    }
    __syncwarp();
    // 8. sparse mel (sequential tap order, mel.rs:92-104) + log(max(x, 1e-5))
    for (int mI = lane; mI < FB_NM; mI += 32) {
        int s = __ldg(tb.mel_start + mI), len = __ldg(tb.mel_len + mI), off = __ldg(tb.mel_off + mI);
        float acc = 0.0f;
        for (int t = 0; t < len; ++t) acc = __fadd_rn(acc, __fmul_rn(__ldg(tb.mel_w + off + t), re[s + t]));
        im[mI] = logf(fmaxf(acc, 1e-5f));
    }
    __syncwarp();
    // 9. outputs: raw mel frame (optional) and every LFR slot that clamps onto this frame
    if (mel_opt) {
        float* mo = mel_opt + ((long long)clip * frames + f) * FB_NM;
        for (int c = lane; c < FB_NM; c += 32) mo[c] = im[c];
    }
    int i_lo = (f - 3 + 5) / 6 - 1; if (i_lo < 0) i_lo = 0;          // rows i with 6i-3 <= f
    int i_hi = (f == frames - 1) ? t_lfr - 1 : (f + 3) / 6;           // last frame absorbs the clamp
    if (i_hi > t_lfr - 1) i_hi = t_lfr - 1;
    for (int i = i_lo; i <= i_hi; ++i) {
#pragma unroll
        for (int b = 0; b < 7; ++b) {
            int raw = i * 6 + b - 3;
            int c = raw < 0 ? 0 : (raw > frames - 1 ? frames - 1 : raw);
            if (c == f) {
                float* dst = lfr_out + ((long long)clip * t_lfr + i) * (7 * FB_NM) + b * FB_NM;
                for (int q = lane; q < FB_NM; q += 32) dst[q] = im[q];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// fused fbank + LFR, register-blocked FFT (round 2): one warp per frame, the 512-point transform as three radix-8 passes
// (512 = 8 x 8 x 8, n = 64 n1 + 8 n2 + n3, k = k1 + 8 k2 + 64 k3) with 16 points per lane in registers and two
// conflict-free exchanges through shared memory -- 128 shared-memory accesses per lane instead of the radix-2 network's
// ~600, no bit-reversal scatter, pre-emphasis neighbours by shuffle.  Pass 1 is a real-input butterfly (the frame is real),
// pass 3 only produces bins 0..256.  Twiddles W64^(n2 k1), W512^(n3 (k1 + 8 k2)) come from tables computed in double.
// Same frame arithmetic as fbank_lfr_kernel up to the FFT's summation order (|dX| ~ 1e-7 of the spectrum's maximum;
// the reference's own AVX2 / scalar FFT paths differ from each other at that level, kernels/fft.rs:208).
// ---------------------------------------------------------------------------
constexpr int FX_P = 72;                                  // row pitch of the exchange buffers (8 rows x 64 points + padding)
constexpr float FX_R = 0.70710678118654752440f;           // sqrt(1/2)

// 8-point forward DFT of a real sequence: X[k] = sum_n x[n] W8^(n k)
__device__ __forceinline__ void fx_dft8_real(const float (&x)[8], float (&yr)[8], float (&yi)[8]) {
    const float b0 = x[0] + x[4], b1 = x[1] + x[5], b2 = x[2] + x[6], b3 = x[3] + x[7];
    const float d0 = x[0] - x[4], d1 = x[1] - x[5], d2 = x[2] - x[6], d3 = x[3] - x[7];
    const float s0 = b0 + b2, s1 = b0 - b2, s2 = b1 + b3, t = b1 - b3;
    yr[0] = s0 + s2; yi[0] = 0.0f;
    yr[4] = s0 - s2; yi[4] = 0.0f;
    yr[2] = s1; yi[2] = -t;
    yr[6] = s1; yi[6] = t;
    const float p = (d1 - d3) * FX_R, q = (d1 + d3) * FX_R;
    yr[1] = d0 + p; yi[1] = -d2 - q;
    yr[7] = d0 + p; yi[7] = d2 + q;
    yr[3] = d0 - p; yi[3] = d2 - q;
    yr[5] = d0 - p; yi[5] = q - d2;
}
// 4-point forward DFT (decimation in frequency), outputs in natural order
__device__ __forceinline__ void fx_dft4(const float (&ar)[4], const float (&ai)[4], float (&yr)[4], float (&yi)[4]) {
    const float s0r = ar[0] + ar[2], s0i = ai[0] + ai[2], s1r = ar[0] - ar[2], s1i = ai[0] - ai[2];
    const float s2r = ar[1] + ar[3], s2i = ai[1] + ai[3], ur = ar[1] - ar[3], ui = ai[1] - ai[3];
    yr[0] = s0r + s2r; yi[0] = s0i + s2i;
    yr[2] = s0r - s2r; yi[2] = s0i - s2i;
    yr[1] = s1r + ui;  yi[1] = s1i - ur;                  // s1 + (-i) u
    yr[3] = s1r - ui;  yi[3] = s1i + ur;
}
// 8-point forward DFT of a complex sequence
__device__ __forceinline__ void fx_dft8(const float (&xr)[8], const float (&xi)[8], float (&yr)[8], float (&yi)[8]) {
    float br[4], bi[4], cr[4], ci[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { br[n] = xr[n] + xr[n + 4]; bi[n] = xi[n] + xi[n + 4]; }
    const float d0r = xr[0] - xr[4], d0i = xi[0] - xi[4], d1r = xr[1] - xr[5], d1i = xi[1] - xi[5];
    const float d2r = xr[2] - xr[6], d2i = xi[2] - xi[6], d3r = xr[3] - xr[7], d3i = xi[3] - xi[7];
    cr[0] = d0r;                 ci[0] = d0i;             // c_n = d_n W8^n
    cr[1] = (d1r + d1i) * FX_R;  ci[1] = (d1i - d1r) * FX_R;
    cr[2] = d2i;                 ci[2] = -d2r;
    cr[3] = (d3i - d3r) * FX_R;  ci[3] = (-d3r - d3i) * FX_R;
    float er[4], ei[4], orr[4], oi[4];
    fx_dft4(br, bi, er, ei);
    fx_dft4(cr, ci, orr, oi);
#pragma unroll
    for (int j = 0; j < 4; ++j) { yr[2 * j] = er[j]; yi[2 * j] = ei[j]; yr[2 * j + 1] = orr[j]; yi[2 * j + 1] = oi[j]; }
}

// The 512-point real-input transform of one warp: v[i] = x[lane + 32 i]; re / im = the warp's two 8 x FX_P exchange buffers.
template <class Emit>
__device__ __forceinline__ void fx_fft512(const float (&v)[16], float* re, float* im, const float2* __restrict__ tw64,
                                          const float2* __restrict__ tw512, int lane, Emit emit) {
    // 5a. pass 1: point j = lane + 32 i = 64 n1 + m with m = lane + 32 (i & 1), n1 = i >> 1: two real 8-point transforms over n1,
    //     times W64^(n2 k1) (n2 = m >> 3), stored as A[k1][m]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float in[8], yr[8], yi[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) in[n1] = v[2 * n1 + h];
        fx_dft8_real(in, yr, yi);
        const int m = lane + 32 * h, n2 = m >> 3;
        re[m] = yr[0]; im[m] = yi[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            const float2 w = __ldg(tw64 + ((n2 * k1) & 63));
            re[k1 * FX_P + m] = yr[k1] * w.x - yi[k1] * w.y;
            im[k1 * FX_P + m] = yr[k1] * w.y + yi[k1] * w.x;
        }
    }
    __syncwarp();
    // 5b. pass 2: the lane owns (k1, n3) = (lane / 8 + 4 h, lane % 8): 8-point transforms over n2, times W512^(n3 (k1 + 8 k2)),
    //     stored as B[k1][k2 + 9 n3] (both passes read every input before the buffer is rewritten)
    {
        const int n3 = lane & 7;
        float ar[2][8], ai[2][8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k1 = (lane >> 3) + 4 * h;
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) { ar[h][n2] = re[k1 * FX_P + n2 * 8 + n3]; ai[h][n2] = im[k1 * FX_P + n2 * 8 + n3]; }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k1 = (lane >> 3) + 4 * h;
            float yr[8], yi[8];
            fx_dft8(ar[h], ai[h], yr, yi);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const float2 w = __ldg(tw512 + ((n3 * (k1 + 8 * k2)) & 511));
                re[k1 * FX_P + k2 + 9 * n3] = yr[k2] * w.x - yi[k2] * w.y;
                im[k1 * FX_P + k2 + 9 * n3] = yr[k2] * w.y + yi[k2] * w.x;
            }
        }
    }
    __syncwarp();
    // 5c. pass 3: the lane owns (k1, k2) = (lane / 8 + 4 h, lane % 8): 8-point transforms over n3 give X[k1 + 8 k2 + 64 k3];
    //     bins 0..256 are handed to emit(k, re, im) with Im(0) = Im(256) = 0 (kernels/fft.rs:124-128); emit may overwrite re[] / im[]
    //     (every input of the pass has been read)
    {
        const int k2 = lane & 7;
        float ar[2][8], ai[2][8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k1 = (lane >> 3) + 4 * h;
#pragma unroll
            for (int n3 = 0; n3 < 8; ++n3) { ar[h][n3] = re[k1 * FX_P + k2 + 9 * n3]; ai[h][n3] = im[k1 * FX_P + k2 + 9 * n3]; }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k1 = (lane >> 3) + 4 * h;
            float yr[8], yi[8];
            fx_dft8(ar[h], ai[h], yr, yi);
            const int k0 = k1 + 8 * k2;
#pragma unroll
            for (int k3 = 0; k3 < 4; ++k3) emit(k0 + 64 * k3, yr[k3], (k0 == 0 && k3 == 0) ? 0.0f : yi[k3]);
            if (k0 == 0) emit(256, yr[4], 0.0f);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FB_WARPS * 32)
fbank_lfr_r8_kernel(const float* __restrict__ pcm, long long clip_stride, int n_clips, int frames, int t_lfr,
                    FbankTables tb, const float2* __restrict__ tw64, const float2* __restrict__ tw512,
                    float* __restrict__ mel_opt, float* __restrict__ lfr_out) {
    __shared__ float s_re[FB_WARPS][8 * FX_P];
    __shared__ float s_im[FB_WARPS][8 * FX_P];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gframe = (long long)blockIdx.x * FB_WARPS + warp;
    if (gframe >= (long long)n_clips * frames) return;   // whole warp exits together
    const int clip = (int)(gframe / frames), f = (int)(gframe % frames);
    float* re = s_re[warp];
    float* im = s_im[warp];
    const float* p = pcm + (long long)clip * clip_stride + (long long)f * FB_HOP;

    // 1. scale (x32768) and frame mean; 2. mean subtraction (same operation order as fbank_lfr_kernel)
    float x[13];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int j = lane + 32 * i;
        x[i] = j < FB_FL ? __fmul_rn(__ldg(p + j), 32768.0f) : 0.0f;
        sum += x[i];
    }
    sum = lb_warp_sum(sum);
    const float mean = __fdiv_rn(sum, (float)FB_FL);
#pragma unroll
    for (int i = 0; i < 13; ++i) x[i] = __fsub_rn(x[i], mean);
    // 3. pre-emphasis against the un-emphasised neighbour (sample j - 1 lives in the lane below, or in lane 31 of the
    //    previous register), 4. Hann window; samples 400..511 are the zero padding
    float v[16];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int j = lane + 32 * i;
        const float up = __shfl_up_sync(0xffffffffu, x[i], 1);
        const float wrap = i > 0 ? __shfl_sync(0xffffffffu, x[i - 1], 31) : 0.0f;
        const float prev = lane == 0 ? wrap : up;
        float cur = x[i];
        if (j >= 1) cur = __fsub_rn(cur, __fmul_rn(0.97f, prev));
        v[i] = j < FB_FL ? __fmul_rn(cur, __ldg(tb.window + j)) : 0.0f;
    }
    v[13] = 0.0f; v[14] = 0.0f; v[15] = 0.0f;

    // 5. FFT-512 (three radix-8 passes); 6. power spectrum of bins 0..256 into re[0..256]
    fx_fft512(v, re, im, tw64, tw512, lane, [&](int k, float r, float q) { re[k] = __fadd_rn(__fmul_rn(r, r), __fmul_rn(q, q)); });
    // 7. sparse mel (sequential tap order, mel.rs:92-104) + log(max(x, 1e-5))
    for (int mI = lane; mI < FB_NM; mI += 32) {
        int s = __ldg(tb.mel_start + mI), len = __ldg(tb.mel_len + mI), off = __ldg(tb.mel_off + mI);
        float acc = 0.0f;
        for (int t = 0; t < len; ++t) acc = __fadd_rn(acc, __fmul_rn(__ldg(tb.mel_w + off + t), re[s + t]));
        im[mI] = logf(fmaxf(acc, 1e-5f));
    }
    __syncwarp();
    // 8. outputs: raw mel frame (optional) and every LFR slot that clamps onto this frame
    if (mel_opt) {
        float* mo = mel_opt + ((long long)clip * frames + f) * FB_NM;
        for (int c = lane; c < FB_NM; c += 32) mo[c] = im[c];
    }
    int i_lo = (f - 3 + 5) / 6 - 1; if (i_lo < 0) i_lo = 0;          // rows i with 6i-3 <= f
    int i_hi = (f == frames - 1) ? t_lfr - 1 : (f + 3) / 6;           // last frame absorbs the clamp
    if (i_hi > t_lfr - 1) i_hi = t_lfr - 1;
    for (int i = i_lo; i <= i_hi; ++i) {
#pragma unroll
        for (int b = 0; b < 7; ++b) {
            int raw = i * 6 + b - 3;
            int c = raw < 0 ? 0 : (raw > frames - 1 ? frames - 1 : raw);
            if (c == f) {
                float* dst = lfr_out + ((long long)clip * t_lfr + i) * (7 * FB_NM) + b * FB_NM;
                for (int q = lane; q < FB_NM; q += 32) dst[q] = im[q];
            }
        }
    }
}

// W64^k | W512^k as float2, computed in double (the radix-8 transform's twiddles)
static int get_fft512r8_tables(lele_b200_ctx* ctx, const float2** tw64, const float2** tw512) {
    void* tw = nullptr;
    auto it = ctx->tables.find("fft512r8");
    if (it == ctx->tables.end()) {
        std::vector<float> h(2 * (64 + 512));
        for (int k = 0; k < 64; ++k) { h[2 * k] = (float)cos(-2.0 * M_PI * k / 64.0); h[2 * k + 1] = (float)sin(-2.0 * M_PI * k / 64.0); }
        for (int k = 0; k < 512; ++k) { h[128 + 2 * k] = (float)cos(-2.0 * M_PI * k / 512.0); h[128 + 2 * k + 1] = (float)sin(-2.0 * M_PI * k / 512.0); }
        int rc = lb_table(ctx, "fft512r8", h.data(), h.size() * sizeof(float), &tw);
        if (rc) return rc;
    } else tw = it->second;
    *tw64 = (const float2*)tw; *tw512 = (const float2*)tw + 64;
    return LELE_B200_OK;
}

// rFFT rows / STFT frames of n_fft = 512 on the same register-blocked transform: one warp per frame (the generic kernel below spends a
// CTA and a radix-2 shared-memory network per frame).  mode: 0 re | im planes, 1 interleaved, 2 power.
__global__ void __launch_bounds__(FB_WARPS * 32)
stft512_r8_kernel(const float* __restrict__ sig, int signal_len, int hop, int win, const float* __restrict__ window,
                  const float2* __restrict__ tw64, const float2* __restrict__ tw512, int mode, int frames,
                  float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ float s_re[FB_WARPS][8 * FX_P];
    __shared__ float s_im[FB_WARPS][8 * FX_P];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * FB_WARPS + warp;
    if (row >= frames) return;                           // whole warp exits together
    float* re = s_re[warp];
    float* im = s_im[warp];
    const long long start = (long long)row * hop;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = lane + 32 * i;
        float x = 0.0f;
        if (j < win && start + j < signal_len) {
            x = __ldg(sig + start + j);
            if (window) x = __fmul_rn(x, __ldg(window + j));
        }
        v[i] = x;
    }
    fx_fft512(v, re, im, tw64, tw512, lane, [&](int k, float r, float q) { re[k] = r; im[k] = q; });
    constexpr int NFR = FB_NF / 2 + 1;
    for (int k = lane; k < NFR; k += 32) {
        const float r = re[k], q = im[k];
        if (mode == 0) { out0[(long long)row * NFR + k] = r; out1[(long long)row * NFR + k] = q; }
        else if (mode == 1) { out0[((long long)row * NFR + k) * 2] = r; out0[((long long)row * NFR + k) * 2 + 1] = q; }
        else out0[(long long)row * NFR + k] = __fadd_rn(__fmul_rn(r, r), __fmul_rn(q, q));
    }
}

static int get_fft_tables(lele_b200_ctx* ctx, int n, const float** tw_re, const float** tw_im, const int** br) {
    std::string key = "fft" + std::to_string(n);
    void* p = nullptr;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::vector<float> h(3 * (size_t)n);
        host_twiddles(n, h.data(), h.data() + n, (int*)(h.data() + 2 * n));
        int rc = lb_table(ctx, key, h.data(), h.size() * sizeof(float), &p);
        if (rc) return rc;
    } else p = it->second;
    *tw_re = (const float*)p; *tw_im = (const float*)p + n; *br = (const int*)p + 2 * n;
    return LELE_B200_OK;
}

static int get_fbank_tables(lele_b200_ctx* ctx, FbankTables* tb) {
    int rc = get_fft_tables(ctx, FB_NF, &tb->tw_re, &tb->tw_im, &tb->bit_rev);
    if (rc) return rc;
    void* p = nullptr;
    auto it = ctx->tables.find("fbank");
    if (it == ctx->tables.end()) {
        // layout: window[400] | start[80] | len[80] | off[80] | weights[...]
        std::vector<float> melw((size_t)FB_NM * 257);
        host_mel_filterbank(16000.0f, FB_NF, FB_NM, 20.0f, 8000.0f, melw.data());  // pipeline.rs:46-52
        std::vector<float> blob(400 + 240);
        host_hann(FB_FL, blob.data());
        int* meta = (int*)(blob.data() + 400);
        std::vector<float> packed;
        for (int i = 0; i < FB_NM; ++i) {  // SparseMelBank::new mel.rs:56-90
            const float* row = melw.data() + (size_t)i * 257;
            int s = 0, e = 257;
            while (s < 257 && row[s] == 0.0f) ++s;
            while (e > s && row[e - 1] == 0.0f) --e;
            if (!(s < e)) { s = 0; e = 0; }
            meta[i] = s; meta[80 + i] = e - s; meta[160 + i] = (int)packed.size();
            packed.insert(packed.end(), row + s, row + e);
        }
        size_t head = blob.size();
        blob.insert(blob.end(), packed.begin(), packed.end());
        rc = lb_table(ctx, "fbank", blob.data(), blob.size() * sizeof(float), &p);
        if (rc) return rc;
        (void)head;
    } else p = it->second;
    const float* base = (const float*)p;
    tb->window = base;
    tb->mel_start = (const int*)(base + 400);
    tb->mel_len = (const int*)(base + 480);
    tb->mel_off = (const int*)(base + 560);
    tb->mel_w = base + 640;
    return LELE_B200_OK;
}

extern "C" int lele_b200_frontend_compute(lele_b200_ctx* ctx, const float* pcm, int n_clips, int n_samples,
                                          long long clip_stride, float* mel_opt, float* lfr_out) {
    LB_REQUIRE(ctx && pcm && lfr_out, "frontend_compute: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_clips >= 0 && clip_stride >= n_samples, "frontend_compute: bad clip geometry");
    int frames = lele_b200_frontend_num_frames(n_samples);
    if (frames == 0 || n_clips == 0) return LELE_B200_OK;  // TensorView::empty() (pipeline.rs:70-72)
    int t_lfr = (frames + 5) / 6;
    FbankTables tb;
    int rc = get_fbank_tables(ctx, &tb);
    if (rc) return rc;
    long long total = (long long)n_clips * frames;
    if (lb_env_flag("LELE_B200_FBANK_R8", 1)) {
        const float2 *tw64, *tw512;
        if ((rc = get_fft512r8_tables(ctx, &tw64, &tw512))) return rc;
        fbank_lfr_r8_kernel<<<lb_ceil_div(total, FB_WARPS), FB_WARPS * 32, 0, ctx->stream>>>(
            pcm, clip_stride, n_clips, frames, t_lfr, tb, tw64, tw512, mel_opt, lfr_out);
    } else
        fbank_lfr_kernel<<<lb_ceil_div(total, FB_WARPS), FB_WARPS * 32, 0, ctx->stream>>>(
            pcm, clip_stride, n_clips, frames, t_lfr, tb, mel_opt, lfr_out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// ---------------------------------------------------------------------------
// LFR (lfr.rs:18-54) and CMVN (cmvn.rs:14-66) as stand-alone operators
// ---------------------------------------------------------------------------
__global__ void lfr_kernel(const float* __restrict__ in, int t, int d, int m, int n, int t_lfr, float* __restrict__ out) {
    const int clip = blockIdx.y;
    const long long total = (long long)t_lfr * m * d;
    const int pad = (m - 1) / 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % d);
        int b = (int)((idx / d) % m);
        int i = (int)(idx / ((long long)d * m));
        int raw = i * n + b - pad;
        int src = raw < 0 ? 0 : (raw > t - 1 ? t - 1 : raw);
        out[(long long)clip * total + idx] = in[((long long)clip * t + src) * d + c];
    }
}
extern "C" int lele_b200_lfr(lele_b200_ctx* ctx, const float* in, int n_clips, int t, int d, int m, int n, float* out) {
    LB_REQUIRE(ctx && m > 0 && n > 0 && d > 0, "lfr: bad arguments");
    LB_ENTER(ctx);
    if (t == 0 || n_clips == 0) return LELE_B200_OK;
    int t_lfr = (t + n - 1) / n;
    long long total = (long long)t_lfr * m * d;
    dim3 grid(min(lb_ceil_div(total, 256), 4096), n_clips);
    lfr_kernel<<<grid, 256, 0, ctx->stream>>>(in, t, d, m, n, t_lfr, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// one thread per (clip, dim); sequential accumulation over time = the reference's order,
// coalesced across dims.
__global__ void cmvn_kernel(const float* __restrict__ in, int t, int d, float eps, float* __restrict__ out) {
    const int clip = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d) return;
    const float* x = in + (long long)clip * t * d + k;
    float* y = out + (long long)clip * t * d + k;
    // the sums run over time in the reference's order (one dependent chain per dimension); the loads do not depend on it, so they
    // are issued 16 rows ahead (a thread owns one of only d x clips chains: without the batching every row was a full memory round trip)
    constexpr int U = 16;
    float s = 0.0f, sq = 0.0f;
    for (int i0 = 0; i0 < t; i0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = i0 + u < t ? __ldg(x + (long long)(i0 + u) * d) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u < t) { s = __fadd_rn(s, v[u]); sq = __fadd_rn(sq, __fmul_rn(v[u], v[u])); }
    }
    float tf = (float)t;
    float mean = __fdiv_rn(s, tf);
    float var = fmaxf(__fsub_rn(__fdiv_rn(sq, tf), __fmul_rn(mean, mean)), 0.0f);
    float sd = __fsqrt_rn(__fadd_rn(var, eps));
    for (int i0 = 0; i0 < t; i0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = i0 + u < t ? __ldg(x + (long long)(i0 + u) * d) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u < t) y[(long long)(i0 + u) * d] = __fdiv_rn(__fsub_rn(v[u], mean), sd);
    }
}
extern "C" int lele_b200_cmvn(lele_b200_ctx* ctx, const float* in, int n_clips, int t, int d, float eps, float* out) {
    LB_REQUIRE(ctx && d > 0, "cmvn: bad arguments");
    LB_ENTER(ctx);
    if (t == 0 || n_clips == 0) return LELE_B200_OK;
    dim3 grid(lb_ceil_div(d, 64), n_clips);
    cmvn_kernel<<<grid, 64, 0, ctx->stream>>>(in, t, d, eps, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// ---------------------------------------------------------------------------
// generic rFFT rows and STFT (math.rs:2304-2439): one CTA per row / frame
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rfft_rows_kernel(const float* __restrict__ sig, int signal_len, int n_fft, int hop, int win,
                 const float* __restrict__ window, const float* __restrict__ tw_re, const float* __restrict__ tw_im,
                 const int* __restrict__ bit_rev, int mode /*0 re|im planes, 1 interleaved, 2 power*/,
                 float* __restrict__ out0, float* __restrict__ out1) {
    extern __shared__ float sm[];
    float* re = sm;
    float* im = sm + n_fft;
    const int row = blockIdx.x;
    const long long start = (long long)row * hop;
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
        float v = 0.0f;
        if (i < win && start + i < signal_len) {
            v = sig[start + i];
            if (window) v = __fmul_rn(v, window[i]);
        }
        re[bit_rev[i]] = v;
        im[i] = 0.0f;
    }
    fft_radix2<false>(re, im, n_fft, tw_re, tw_im, threadIdx.x, blockDim.x);
    const int nfr = n_fft / 2 + 1;
    for (int k = threadIdx.x; k < nfr; k += blockDim.x) {
        float r = re[k], q = (k == 0 || k == nfr - 1) ? 0.0f : im[k];
        if (mode == 0) { out0[(long long)row * nfr + k] = r; out1[(long long)row * nfr + k] = q; }
        else if (mode == 1) { out0[((long long)row * nfr + k) * 2] = r; out0[((long long)row * nfr + k) * 2 + 1] = q; }
        else out0[(long long)row * nfr + k] = __fadd_rn(__fmul_rn(r, r), __fmul_rn(q, q));
    }
}

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

extern "C" int lele_b200_rfft(lele_b200_ctx* ctx, const float* x, int n_rows, int n, float* out_re, float* out_im) {
    LB_REQUIRE(ctx && x && out_re && out_im, "rfft: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(is_pow2(n) && n >= 2 && n <= 4096, "rfft: n=%d must be a power of two in [2,4096] (kernels/fft.rs:4)", n);
    if (n_rows == 0) return LELE_B200_OK;
    if (n == FB_NF && lb_env_flag("LELE_B200_FFT_R8", 1)) {
        const float2 *tw64, *tw512;
        int rc8 = get_fft512r8_tables(ctx, &tw64, &tw512);
        if (rc8) return rc8;
        stft512_r8_kernel<<<lb_ceil_div(n_rows, FB_WARPS), FB_WARPS * 32, 0, ctx->stream>>>(x, n_rows * n, n, n, nullptr, tw64, tw512, 0, n_rows, out_re, out_im);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    const float *tr, *ti; const int* br;
    int rc = get_fft_tables(ctx, n, &tr, &ti, &br);
    if (rc) return rc;
    rfft_rows_kernel<<<n_rows, 256, 2 * n * sizeof(float), ctx->stream>>>(x, n_rows * n, n, n, n, nullptr, tr, ti, br, 0, out_re, out_im);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_stft(lele_b200_ctx* ctx, const float* signal, int signal_len, int n_fft, int hop, int win,
                              const float* window, int power, float* out, int* frames_out) {
    LB_REQUIRE(ctx, "stft: NULL ctx");
    LB_ENTER(ctx);
    LB_REQUIRE(is_pow2(n_fft) && n_fft >= 2 && n_fft <= 4096, "stft: n_fft=%d must be a power of two <= 4096", n_fft);
    LB_REQUIRE(hop > 0 && win > 0 && win <= n_fft, "stft: bad hop/win");
    if (signal_len == 0) { if (frames_out) *frames_out = 0; return LELE_B200_OK; }
    int frames = signal_len < win ? 1 : (signal_len - win) / hop + 1;
    if (frames_out) *frames_out = frames;
    const float *tr, *ti; const int* br;
    int rc = get_fft_tables(ctx, n_fft, &tr, &ti, &br);
    if (rc) return rc;
    const float* wdev = window;
    if (!window) {  // default periodic Hann (math.rs:2327-2333)
        std::string key = "phann" + std::to_string(win);
        auto it = ctx->tables.find(key);
        void* p;
        if (it == ctx->tables.end()) {
            std::vector<float> w(win);
            for (int i = 0; i < win; ++i) w[i] = 0.5f * (1.0f - cosf(2.0f * LB_PI_F * (float)i / (float)win));
            rc = lb_table(ctx, key, w.data(), w.size() * sizeof(float), &p);
            if (rc) return rc;
        } else p = it->second;
        wdev = (const float*)p;
    }
    if (n_fft == FB_NF && lb_env_flag("LELE_B200_FFT_R8", 1)) {
        const float2 *tw64, *tw512;
        if ((rc = get_fft512r8_tables(ctx, &tw64, &tw512))) return rc;
        stft512_r8_kernel<<<lb_ceil_div(frames, FB_WARPS), FB_WARPS * 32, 0, ctx->stream>>>(signal, signal_len, hop, win, wdev, tw64, tw512, power ? 2 : 1, frames, out, nullptr);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    rfft_rows_kernel<<<frames, 256, 2 * n_fft * sizeof(float), ctx->stream>>>(signal, signal_len, n_fft, hop, win, wdev, tr, ti, br,
                                                                                power ? 2 : 1, out, nullptr);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
