/*
 * lele_oracle.c -- TEST INFRASTRUCTURE ONLY (see lele_oracle.h).
 * Scalar C restatement of lele's x86_64 CPU operator semantics.
 * Build: gcc -O3 -march=native -ffp-contract=off -fno-fast-math -fPIC -shared
 */
#include "lele_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#if defined(__AVX512F__) && defined(__AVX512BW__) && defined(__AVX512VNNI__)
#include <immintrin.h>
#define LO_HAVE_VNNI512 1
#endif
#if defined(__AVX512F__)
#include <immintrin.h>
#define LO_HAVE_AVX512F 1
#endif

#define LO_PI 3.14159265358979323846f /* std::f32::consts::PI rounds to the same f32 */

/* ------------------------------------------------------------------------- */
/* front-end                                                                  */
/* ------------------------------------------------------------------------- */

/* src/features/window.rs:2-12 : symmetric Hann, cos evaluated in f32 */
void lo_hann_window(int size, float *out) {
    if (size == 0) return;
    if (size == 1) { out[0] = 1.0f; return; }
    for (int n = 0; n < size; ++n)
        out[n] = 0.5f * (1.0f - cosf(2.0f * LO_PI * (float)n / (float)(size - 1)));
}

static int lo_bit_reverse(int n, int log2n) { /* kernels/fft.rs:161 */
    int r = 0, x = n;
    for (int i = 0; i < log2n; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static int lo_log2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

/* src/kernels/fft.rs:136-158 : stage-major twiddle table, angle computed in f32 */
void lo_precompute_twiddles(int n, float *tw_re, float *tw_im, int32_t *bit_rev) {
    int log2n = lo_log2(n);
    for (int i = 0; i < n; ++i) bit_rev[i] = lo_bit_reverse(i, log2n);
    int o = 0;
    for (int size = 2; size <= n; size *= 2) {
        int half = size / 2, step = n / size;
        for (int k = 0; k < half; ++k) {
            float angle = -2.0f * LO_PI * (float)(k * step) / (float)n;
            tw_re[o] = cosf(angle);
            tw_im[o] = sinf(angle);
            ++o;
        }
    }
}

/* src/kernels/fft.rs:79-134 (rfft_forward_f32_precomputed_scalar): radix-2 DIT on a
 * zero-imaginary complex buffer; out has n/2+1 bins, Im(0)=Im(n/2)=0.
 * (The AVX2 arm :172 uses fmsub/fmadd in the butterfly; last-bit differences.) */
static void lo_rfft_pre(const float *input, int n, const float *tw_re, const float *tw_im,
                        const int32_t *bit_rev, float *re, float *im, float *out_re,
                        float *out_im) {
    int half = n / 2 + 1;
    for (int i = 0; i < n; ++i) re[bit_rev[i]] = input[i];
    memset(im, 0, sizeof(float) * (size_t)n);
    int tw_off = 0;
    for (int size = 2; size <= n; size *= 2) {
        int hs = size / 2, nb = n / size;
        for (int b = 0; b < nb; ++b) {
            int base = b * size;
            for (int k = 0; k < hs; ++k) {
                int e = base + k, o = base + hs + k;
                float wr = tw_re[tw_off + k], wi = tw_im[tw_off + k];
                float ore = re[o], oim = im[o];
                float tr = wr * ore - wi * oim;
                float ti = wr * oim + wi * ore;
                re[o] = re[e] - tr;
                im[o] = im[e] - ti;
                re[e] += tr;
                im[e] += ti;
            }
        }
        tw_off += hs;
    }
    out_re[0] = re[0];
    out_im[0] = 0.0f;
    if (half > 1) { out_re[half - 1] = re[n / 2]; out_im[half - 1] = 0.0f; }
    for (int k = 1; k < half - 1; ++k) { out_re[k] = re[k]; out_im[k] = im[k]; }
}

void lo_rfft_forward(const float *input, int n, float *out_re, float *out_im) {
    float *tw_re = malloc(sizeof(float) * n), *tw_im = malloc(sizeof(float) * n);
    int32_t *br = malloc(sizeof(int32_t) * n);
    float *re = malloc(sizeof(float) * n), *im = malloc(sizeof(float) * n);
    lo_precompute_twiddles(n, tw_re, tw_im, br);
    lo_rfft_pre(input, n, tw_re, tw_im, br, re, im, out_re, out_im);
    free(tw_re); free(tw_im); free(br); free(re); free(im);
}

float lo_hz_to_mel_htk(float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); }   /* mel.rs:1 */
float lo_mel_to_hz_htk(float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); } /* mel.rs:4 */

/* src/features/mel.rs:7-45 : HTK triangles on bin-centre frequencies, no area norm */
void lo_mel_filterbank(float sr, int n_fft, int n_mels, float f_min, float f_max, float *w) {
    int n_freqs = n_fft / 2 + 1, pts = n_mels + 2;
    float mel_min = lo_hz_to_mel_htk(f_min), mel_max = lo_hz_to_mel_htk(f_max);
    float mel_step = (mel_max - mel_min) / (float)(n_mels + 1);
    float *hz = malloc(sizeof(float) * pts);
    for (int i = 0; i < pts; ++i) hz[i] = lo_mel_to_hz_htk(mel_min + (float)i * mel_step);
    for (int i = 0; i < n_mels; ++i) {
        float fl = hz[i], fc = hz[i + 1], fr = hz[i + 2];
        for (int j = 0; j < n_freqs; ++j) {
            float f = (float)j * sr / (float)n_fft, val = 0.0f;
            if (f > fl && f < fc) val = (f - fl) / (fc - fl);
            else if (f >= fc && f < fr) val = (fr - f) / (fr - fc);
            w[i * n_freqs + j] = val;
        }
    }
    free(hz);
}

/* src/features/lfr.rs:18-54 */
void lo_lfr(const float *in, int t, int d, int m, int n, float *out) {
    if (t == 0) return;
    int t_lfr = (t + n - 1) / n, pad = (m - 1) / 2, d_out = d * m;
    for (int i = 0; i < t_lfr; ++i)
        for (int b = 0; b < m; ++b) {
            int raw = i * n + b - pad;
            int c = raw < 0 ? 0 : (raw > t - 1 ? t - 1 : raw);
            memcpy(out + (size_t)i * d_out + (size_t)b * d, in + (size_t)c * d, sizeof(float) * d);
        }
}

/* src/features/cmvn.rs:14-66 : per-dim stats over time, sequential f32 accumulation */
void lo_cmvn(const float *in, int t, int d, float eps, float *out) {
    if (t == 0) return;
    float *sums = calloc(d, sizeof(float)), *sq = calloc(d, sizeof(float));
    for (int ti = 0; ti < t; ++ti)
        for (int k = 0; k < d; ++k) {
            float v = in[(size_t)ti * d + k];
            sums[k] += v;
            sq[k] += v * v;
        }
    float tf = (float)t;
    for (int k = 0; k < d; ++k) {
        float mean = sums[k] / tf;
        float var = sq[k] / tf - mean * mean;
        if (!(var > 0.0f)) var = 0.0f; /* f32::max(0.0) */
        float sd = sqrtf(var + eps);
        sums[k] = mean;
        sq[k] = sd;
    }
    for (int ti = 0; ti < t; ++ti)
        for (int k = 0; k < d; ++k)
            out[(size_t)ti * d + k] = (in[(size_t)ti * d + k] - sums[k]) / sq[k];
    free(sums); free(sq);
}

int lo_frontend_num_frames(int n_samples) { /* pipeline.rs:70-73 */
    if (n_samples < 400) return 0;
    return (n_samples - 400) / 160 + 1;
}

/* src/features/pipeline.rs:67-193 with FeatureConfig::default (16 kHz, 80 mel, 25/10 ms,
 * LFR 7/6): frame_len 400, n_fft 512, hop 160, f_min 20 Hz (pipeline.rs:38-66). */
int lo_frontend_compute(const float *pcm, int n_samples, float *mel_out, float *lfr_out) {
    enum { FL = 400, NF = 512, HOP = 160, NM = 80, NB = 257 };
    int frames = lo_frontend_num_frames(n_samples);
    if (frames == 0) return 0;
    float window[FL], tw_re[NF], tw_im[NF], re[NF], im[NF], fre[NB], fim[NB], frame[NF], raw[FL],
        power[NB];
    int32_t br[NF];
    float *melw = malloc(sizeof(float) * NM * NB);
    int start_bin[NM], end_bin[NM];
    lo_hann_window(FL, window);
    lo_precompute_twiddles(NF, tw_re, tw_im, br);
    lo_mel_filterbank(16000.0f, NF, NM, 20.0f, 8000.0f, melw);
    for (int i = 0; i < NM; ++i) { /* SparseMelBank::new mel.rs:56-90 */
        const float *row = melw + i * NB;
        int s = 0, e = NB;
        while (s < NB && row[s] == 0.0f) ++s;
        while (e > s && row[e - 1] == 0.0f) --e;
        if (s < e) { start_bin[i] = s; end_bin[i] = e; } else { start_bin[i] = 0; end_bin[i] = 0; }
    }
    float *mel = mel_out ? mel_out : malloc(sizeof(float) * (size_t)frames * NM);
    for (int i = 0; i < frames; ++i) {
        const float *p = pcm + (size_t)i * HOP;
        for (int j = 0; j < FL; ++j) raw[j] = p[j] * 32768.0f;
        float sum = 0.0f;
        for (int j = 0; j < FL; ++j) sum += raw[j];
        float mean = sum / (float)FL;
        for (int j = 0; j < FL; ++j) raw[j] -= mean;
        for (int j = FL - 1; j >= 1; --j) raw[j] -= 0.97f * raw[j - 1];
        for (int j = 0; j < FL; ++j) frame[j] = raw[j] * window[j];
        for (int j = FL; j < NF; ++j) frame[j] = 0.0f;
        lo_rfft_pre(frame, NF, tw_re, tw_im, br, re, im, fre, fim);
        for (int j = 0; j < NB; ++j) power[j] = fre[j] * fre[j] + fim[j] * fim[j];
        float *mf = mel + (size_t)i * NM;
        for (int k = 0; k < NM; ++k) { /* SparseMelBank::apply mel.rs:92-104 */
            float s = 0.0f;
            for (int j = start_bin[k]; j < end_bin[k]; ++j) s += melw[k * NB + j] * power[j];
            float c = s > 1e-5f ? s : 1e-5f; /* log_compress mel.rs:124 */
            mf[k] = logf(c);
        }
    }
    lo_lfr(mel, frames, NM, 7, 6, lfr_out);
    if (!mel_out) free(mel);
    free(melw);
    return (frames + 5) / 6;
}

/* src/kernels/math.rs:2304-2439 */
int lo_stft(const float *sig, int signal_len, int n_fft, int hop, int win, const float *window,
            int power, float *out) {
    if (signal_len == 0) return 0;
    int frames = signal_len < win ? 1 : (signal_len - win) / hop + 1;
    int nfr = n_fft / 2 + 1;
    float *wd = malloc(sizeof(float) * win);
    if (window) memcpy(wd, window, sizeof(float) * win);
    else
        for (int i = 0; i < win; ++i)
            wd[i] = 0.5f * (1.0f - cosf(2.0f * LO_PI * (float)i / (float)win));
    float *tw_re = malloc(sizeof(float) * n_fft), *tw_im = malloc(sizeof(float) * n_fft);
    int32_t *br = malloc(sizeof(int32_t) * n_fft);
    float *re = malloc(sizeof(float) * n_fft), *im = malloc(sizeof(float) * n_fft);
    float *fd = malloc(sizeof(float) * n_fft), *fre = malloc(sizeof(float) * nfr),
          *fim = malloc(sizeof(float) * nfr);
    lo_precompute_twiddles(n_fft, tw_re, tw_im, br);
    for (int f = 0; f < frames; ++f) {
        int start = f * hop;
        for (int i = 0; i < n_fft; ++i)
            fd[i] = (i < win && start + i < signal_len) ? sig[start + i] * wd[i] : 0.0f;
        lo_rfft_pre(fd, n_fft, tw_re, tw_im, br, re, im, fre, fim);
        for (int q = 0; q < nfr; ++q) {
            if (power) out[(size_t)f * nfr + q] = fre[q] * fre[q] + fim[q] * fim[q];
            else { out[((size_t)f * nfr + q) * 2] = fre[q]; out[((size_t)f * nfr + q) * 2 + 1] = fim[q]; }
        }
    }
    free(wd); free(tw_re); free(tw_im); free(br); free(re); free(im); free(fd); free(fre); free(fim);
    return frames;
}

/* ------------------------------------------------------------------------- */
/* quantised linear                                                           */
/* ------------------------------------------------------------------------- */

static float lo_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* avx/quantization.rs:112-140 (shared by :832): per-tensor params.
 * zp uses f32::round (half away), avx/quantization.rs:139. */
static void lo_dq_params(const float *x, size_t len, float *scale, float *zp) {
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (size_t i = 0; i < len; ++i) { float v = x[i]; if (v < mn) mn = v; if (v > mx) mx = v; }
    float amax = mx > 0.0f ? mx : 0.0f, amin = mn < 0.0f ? mn : 0.0f;
    float range = amax - amin;
    if (!(range > 1e-5f)) range = 1e-5f;
    *scale = range / 255.0f;
    *zp = lo_clampf(roundf(-amin / *scale), 0.0f, 255.0f);
}
/* SIMD body: fma + round-half-even (avx/quantization.rs:158-161); scalar tail:
 * mul, add, round-half-away (:208).  `simd` selects which. */
static float lo_dq_one(float v, float inv_scale, float zp, int simd) {
    float r = simd ? rintf(fmaf(v, inv_scale, zp)) : roundf(v * inv_scale + zp);
    return lo_clampf(r, 0.0f, 255.0f);
}

/* avx/quantization.rs:832-921 : outputs integer-valued f32; SIMD body covers the first
 * len/8*8 elements of the flat tensor. */
void lo_dynamic_quantize_linear(const float *x, size_t len, float *q, float *scale, float *zp) {
    if (len == 0) { *scale = 1.0f; *zp = 0.0f; return; }
    lo_dq_params(x, len, scale, zp);
    float inv = 1.0f / *scale;
    size_t simd_end = (len / 8) * 8;
    for (size_t i = 0; i < len; ++i) q[i] = lo_dq_one(x[i], inv, *zp, i < simd_end);
}

/* quantization.rs:29-72 + scalar core :1137-1236: exact i32 sum of (a-zpa)(b-zpb),
 * then f32 * scale[j] + bias[j] (separate mul and add), optional ReLU.
 * The 2-row AVX2 kernel's VPMADDUBSW i16 saturation (avx/quantization.rs:1597-1600) is
 * NOT reproduced: the documented math (:926-936) is the contract. */
void lo_mat_mul_integer(const float *a, const float *b, int batch, int m, int k, int n,
                        float a_zp, float b_zp, const float *scale, int scale_len,
                        const float *bias, int relu, float *out) {
    int zpa = (int)a_zp, zpb = (int)b_zp;
    int32_t *acc = malloc(sizeof(int32_t) * (size_t)n);
    for (int bi = 0; bi < batch; ++bi)
        for (int i = 0; i < m; ++i) {
            memset(acc, 0, sizeof(int32_t) * (size_t)n);
            const float *ar = a + ((size_t)bi * m + i) * k;
            for (int kk = 0; kk < k; ++kk) {
                int av = (int)(uint8_t)lo_clampf(ar[kk], 0.0f, 255.0f) - zpa;
                const float *br = b + (size_t)kk * n;
                for (int j = 0; j < n; ++j)
                    acc[j] += av * ((int)(uint8_t)lo_clampf(br[j], 0.0f, 255.0f) - zpb);
            }
            float *o = out + ((size_t)bi * m + i) * n;
            for (int j = 0; j < n; ++j) {
                float v = (float)acc[j];
                if (scale) v = v * (scale_len == 1 ? scale[0] : scale[j]);
                if (bias) v = v + bias[j];
                if (relu && v < 0.0f) v = 0.0f;
                o[j] = v;
            }
        }
    free(acc);
}

/* quantization.rs:77-169 -> avx/quantization.rs:225-330 (+ epilogue :1396-1428):
 * per [m,k] slice: min/max -> u8 + row sums (row tail k%8 takes the scalar rounding),
 * exact integer GEMM, y = f32(acc) * (dyn_scale*w_scale[j]) + bias[j], optional ReLU. */
/* prepare_weights (quantization.rs:221-262): [k,n] u8 -> [n,k] + column sums (done once per weight,
 * like the reference's B_WEIGHT_CACHE, avx/quantization.rs:47-95). */
void lo_prepare_weights(const uint8_t *w, int k, int n, uint8_t *wt, int32_t *colsum) {
    for (int j = 0; j < n; ++j) {
        int32_t s = 0;
        /* stored XOR 0x80 (= w - 128 as i8), the reference's VPMADDUBSW/VNNI form (avx/quantization.rs:926-936) */
        for (int kk = 0; kk < k; ++kk) { uint8_t v = w[(size_t)kk * n + j]; wt[(size_t)j * k + kk] = v ^ 0x80; s += v; }
        colsum[j] = s;
    }
}

/* VNNI layout of a prepared weight: [ceil(k/4)][n16][4] i8 (n16 = n rounded up to 16, zero padded), built once per weight
 * like the reference's B_WEIGHT_CACHE.  Returns NULL on hosts without AVX-512 VNNI (the scalar loop is used then). */
int8_t *lo_pack_weights_vnni(const uint8_t *wt, int k, int n) {
#ifdef LO_HAVE_VNNI512
    const int k4 = (k + 3) / 4, n16 = (n + 15) / 16 * 16;
    int8_t *wp = aligned_alloc(64, (size_t)k4 * n16 * 4);
    if (!wp) return NULL;
    memset(wp, 0, (size_t)k4 * n16 * 4);
    const int kq = k / 4;   /* whole quads: a [n][kq] -> [kq][n16] transpose of 4-byte units, 16 x 16 tiles */
    for (int j0 = 0; j0 < n; j0 += 16)
        for (int q0 = 0; q0 < kq; q0 += 16)
            for (int j = j0; j < j0 + 16 && j < n; ++j)
                for (int q = q0; q < q0 + 16 && q < kq; ++q)
                    memcpy(wp + ((size_t)q * n16 + j) * 4, wt + (size_t)j * k + 4 * q, 4);
    for (int j = 0; j < n; ++j)
        for (int kk = kq * 4; kk < k; ++kk) wp[((size_t)(kk >> 2) * n16 + j) * 4 + (kk & 3)] = (int8_t)wt[(size_t)j * k + kk];
    return wp;
#else
    (void)wt; (void)k; (void)n;
    return NULL;
#endif
}
void lo_free_packed(int8_t *wp) { free(wp); }

void lo_fused_quantized_linear_packed(const float *x, int batch, int m, int k, int n, const uint8_t *wt, const int8_t *wp_in,
                                      const int32_t *colsum, const float *w_scale, int w_scale_len,
                                      int w_zp, const float *bias, int relu, float *out) {
    uint8_t *aq = malloc((size_t)m * k);
    int32_t *rsum = malloc(sizeof(int32_t) * (size_t)m);
    float *cs = malloc(sizeof(float) * (size_t)(w_scale_len > 1 ? w_scale_len : 1));
    int k_simd = (k / 8) * 8;
#ifdef LO_HAVE_VNNI512
    int8_t *wp_own = wp_in ? NULL : lo_pack_weights_vnni(wt, k, n);
    const int8_t *wp = wp_in ? wp_in : wp_own;
#else
    (void)wp_in;
#endif
    for (int bi = 0; bi < batch; ++bi) {
        const float *xb = x + (size_t)bi * m * k;
        float scale, zpf;
        lo_dq_params(xb, (size_t)m * k, &scale, &zpf);
        float inv = 1.0f / scale;
        int zpa = (int)zpf;
        if (w_scale_len <= 1) cs[0] = scale * w_scale[0];
        else for (int j = 0; j < w_scale_len; ++j) cs[j] = scale * w_scale[j];
        for (int i = 0; i < m; ++i) {
            int32_t rs = 0;
            for (int kk = 0; kk < k; ++kk) {
                uint8_t qv = (uint8_t)lo_dq_one(xb[(size_t)i * k + kk], inv, zpf, kk < k_simd);
                aq[(size_t)i * k + kk] = qv;
                rs += qv;
            }
            rsum[i] = rs;
        }
        /* The integer core is exact, so any evaluation order gives the same bits.  On AVX-512 VNNI hosts it runs as a
         * register-blocked u8 x i8 GEMM (vpdpbusd = the reference's x86 instruction class, avx/quantization.rs:926-936,
         * 1203-1603): weights repacked to the VNNI layout [k/4][n][4], 4 rows x 64 columns of i32 accumulators per block,
         * activation quads broadcast -- so that the timed CPU baseline is a SIMD micro-kernel like lele's, not a scalar loop.
         * The f32 epilogue is the scalar one lane by lane (separate mul and add, :1417-1423). */
        int j0 = 0;
#ifdef LO_HAVE_VNNI512
        if (wp) {
            const int k4 = (k + 3) / 4, n16 = (n + 15) / 16 * 16, kp = k4 * 4;
            uint8_t *ap = aligned_alloc(64, ((size_t)m * kp + 63) / 64 * 64);
            int32_t *colz = malloc(sizeof(int32_t) * (size_t)n16);
            float *csv = malloc(sizeof(float) * (size_t)n16), *bv = malloc(sizeof(float) * (size_t)n16);
            for (int i = 0; i < m; ++i) {
                memcpy(ap + (size_t)i * kp, aq + (size_t)i * k, (size_t)k);
                memset(ap + (size_t)i * kp + k, 0, (size_t)(kp - k));
            }
            for (int j = 0; j < n16; ++j) {
                colz[j] = j < n ? zpa * colsum[j] : 0;
                csv[j] = j < n ? cs[w_scale_len <= 1 ? 0 : j] : 0.0f;
                bv[j] = (j < n && bias) ? bias[j] : 0.0f;
            }
            const int32_t kzz = k * zpa * w_zp;
            /* k is walked in blocks of KQ quads so that the 64-column weight panel of a block (KQ x 256 B = 32 KB) stays in L1
             * while all row blocks stream past it; partial i32 sums of a column panel wait in `part` between k-blocks */
            enum { KQ = 128 };
            int32_t *part = (k4 > KQ) ? aligned_alloc(64, (((size_t)m + 3) / 4 * 4) * 64 * sizeof(int32_t)) : NULL;
            for (int jb = 0; jb < n16; jb += 64) {
                const int nv = (n16 - jb) >= 64 ? 4 : (n16 - jb) / 16;
                for (int q0 = 0; q0 < k4; q0 += KQ) {
                    const int q1 = (q0 + KQ < k4) ? q0 + KQ : k4;
                    const int first = q0 == 0, last = q1 == k4;
                    for (int i0 = 0; i0 < m; i0 += 4) {
                        const int nr = (m - i0) >= 4 ? 4 : (m - i0);
                        __m512i acc[4][4];
                        if (first) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) acc[r][c] = _mm512_setzero_si512(); }
                        else { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) acc[r][c] = _mm512_load_si512((const void *)(part + ((size_t)(i0 + r) * 4 + c) * 16)); }
                        if (nr == 4 && nv == 4) {       /* full block: constant trip counts keep the 16 accumulators in registers */
                            const uint8_t *a0 = ap + (size_t)i0 * kp;
                            for (int q = q0; q < q1; ++q) {
                                const int8_t *wq = wp + ((size_t)q * n16 + jb) * 4;
                                const __m512i w0 = _mm512_load_si512((const void *)wq), w1 = _mm512_load_si512((const void *)(wq + 64));
                                const __m512i w2 = _mm512_load_si512((const void *)(wq + 128)), w3 = _mm512_load_si512((const void *)(wq + 192));
#define LO_ROW(r)                                                                                              \
                                {                                                                              \
                                    int32_t a4; memcpy(&a4, a0 + (size_t)(r) * kp + 4 * q, 4);                 \
                                    const __m512i av = _mm512_set1_epi32(a4);                                  \
                                    acc[r][0] = _mm512_dpbusd_epi32(acc[r][0], av, w0); acc[r][1] = _mm512_dpbusd_epi32(acc[r][1], av, w1); \
                                    acc[r][2] = _mm512_dpbusd_epi32(acc[r][2], av, w2); acc[r][3] = _mm512_dpbusd_epi32(acc[r][3], av, w3); \
                                }
                                LO_ROW(0) LO_ROW(1) LO_ROW(2) LO_ROW(3)
#undef LO_ROW
                            }
                        } else {
                            for (int q = q0; q < q1; ++q) {
                                __m512i wv[4], av[4];
                                const int8_t *wq = wp + ((size_t)q * n16 + jb) * 4;
                                for (int c = 0; c < nv; ++c) wv[c] = _mm512_load_si512((const void *)(wq + c * 64));
                                for (int r = 0; r < nr; ++r) { int32_t a4; memcpy(&a4, ap + (size_t)(i0 + r) * kp + 4 * q, 4); av[r] = _mm512_set1_epi32(a4); }
                                for (int r = 0; r < nr; ++r) for (int c = 0; c < nv; ++c) acc[r][c] = _mm512_dpbusd_epi32(acc[r][c], av[r], wv[c]);
                            }
                        }
                        if (!last) {
                            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) _mm512_store_si512((void *)(part + ((size_t)(i0 + r) * 4 + c) * 16), acc[r][c]);
                            continue;
                        }
                        for (int r = 0; r < nr; ++r) {
                            const int i = i0 + r;
                            const __m512i rowc = _mm512_set1_epi32((128 - w_zp) * rsum[i] + kzz);
                            for (int c = 0; c < nv; ++c) {
                                const int j = jb + c * 16;
                                const __mmask16 mk = (j + 16 <= n) ? (__mmask16)0xFFFF : (__mmask16)((1u << (n - j)) - 1u);
                                /* dot + (128 - w_zp) * rowsum - zpa * colsum + k * zpa * w_zp  (two's-complement adds: order-free) */
                                __m512i a32 = _mm512_sub_epi32(_mm512_add_epi32(acc[r][c], rowc), _mm512_loadu_si512((const void *)(colz + j)));
                                __m512 v = _mm512_mul_ps(_mm512_cvtepi32_ps(a32), _mm512_loadu_ps(csv + j));
                                if (bias) v = _mm512_add_ps(v, _mm512_loadu_ps(bv + j));
                                if (relu) v = _mm512_mask_blend_ps(_mm512_cmp_ps_mask(v, _mm512_setzero_ps(), _CMP_LT_OQ), v, _mm512_setzero_ps());
                                _mm512_mask_storeu_ps(out + ((size_t)bi * m + i) * n + j, mk, v);
                            }
                        }
                    }
                }
            }
            free(part);
            free(ap); free(colz); free(csv); free(bv);
            j0 = n;
        }
#endif
        for (int j = j0; j < n; ++j) {
            const int8_t *wr = (const int8_t *)(wt + (size_t)j * k);
            const float csj = cs[w_scale_len <= 1 ? 0 : j];
            for (int i = 0; i < m; ++i) {
                const uint8_t *ar = aq + (size_t)i * k;
                int32_t dot = 0;   /* u8 x i8 dot: exact in i32 (auto-vectorises to VNNI vpdpbusd) */
                for (int kk = 0; kk < k; ++kk) dot += (int32_t)ar[kk] * (int32_t)wr[kk];
                /* sum a*w = dot + 128*rowsum;  sum (a-zpa)(w-zpw) = sum a*w - zpw*rowsum - zpa*colsum + k*zpa*zpw */
                int32_t acc = dot + (128 - w_zp) * rsum[i] - zpa * colsum[j] + k * zpa * w_zp;
                float v = (float)acc * csj;
                if (bias) v = v + bias[j];
                if (relu && v < 0.0f) v = 0.0f;
                out[((size_t)bi * m + i) * n + j] = v;
            }
        }
    }
#ifdef LO_HAVE_VNNI512
    free(wp_own);
#endif
    free(aq); free(rsum); free(cs);
}

void lo_fused_quantized_linear_prepared(const float *x, int batch, int m, int k, int n, const uint8_t *wt,
                                        const int32_t *colsum, const float *w_scale, int w_scale_len,
                                        int w_zp, const float *bias, int relu, float *out) {
    lo_fused_quantized_linear_packed(x, batch, m, k, n, wt, NULL, colsum, w_scale, w_scale_len, w_zp, bias, relu, out);
}

void lo_fused_quantized_linear(const float *x, int batch, int m, int k, int n, const uint8_t *w,
                               const float *w_scale, int w_scale_len, int w_zp,
                               const float *bias, int relu, float *out) {
    uint8_t *wt = malloc((size_t)n * k); /* [n,k] so the inner dot is contiguous */
    int32_t *colsum = malloc(sizeof(int32_t) * (size_t)n);
    lo_prepare_weights(w, k, n, wt, colsum);
    lo_fused_quantized_linear_prepared(x, batch, m, k, n, wt, colsum, w_scale, w_scale_len, w_zp, bias, relu, out);
    free(wt); free(colsum);
}

/* ------------------------------------------------------------------------- */
/* f32 GEMM.  Upstream arithmetic lives in faer 0.24 (Cargo.toml:64; not vendored;
 * summation order unknown).  Restated as the textbook sequential-k sum; the
 * reference's own tests pin this boundary to 1e-5..1e-3 abs.        */
/* ------------------------------------------------------------------------- */
static void lo_sgemm_acc(const float *a, long rsa, long csa, const float *b, long rsb, long csb,
                         int m, int k, int n, float alpha, float *c /* [m,n] += */) {
#ifdef LO_HAVE_AVX512F
    if (csb == 1) {
        /* register-blocked form of the loop below (4 rows x 64 columns of accumulators, k outermost inside a block): every
         * output is still  c = fmaf(alpha * a[i,kk], b[kk,j], c)  for kk ascending, i.e. the same bits, at SIMD GEMM speed
         * (the reference's f32 GEMM is faer's AVX micro-kernel) */
        for (int j0 = 0; j0 < n; j0 += 64) {
            const int nv = (n - j0) >= 64 ? 4 : (n - j0 + 15) / 16;
            __mmask16 mk[4];
            for (int cc = 0; cc < 4; ++cc) { const int rem = n - (j0 + 16 * cc); mk[cc] = rem >= 16 ? (__mmask16)0xFFFF : (rem > 0 ? (__mmask16)((1u << rem) - 1u) : 0); }
            for (int i0 = 0; i0 < m; i0 += 4) {
                const int nr = (m - i0) >= 4 ? 4 : (m - i0);
                __m512 acc[4][4];
                for (int r = 0; r < 4; ++r) for (int cc = 0; cc < 4; ++cc)
                    acc[r][cc] = (r < nr && cc < nv) ? _mm512_maskz_loadu_ps(mk[cc], c + (size_t)(i0 + r) * n + j0 + 16 * cc) : _mm512_setzero_ps();
                if (nr == 4 && nv == 4 && mk[3] == (__mmask16)0xFFFF) {
                    for (int kk = 0; kk < k; ++kk) {
                        const float *br = b + kk * rsb + j0;
                        const __m512 b0 = _mm512_loadu_ps(br), b1 = _mm512_loadu_ps(br + 16), b2 = _mm512_loadu_ps(br + 32), b3 = _mm512_loadu_ps(br + 48);
#define LO_ROW(r)                                                                                            \
                        {                                                                                    \
                            const __m512 av = _mm512_set1_ps(alpha * a[(i0 + (r)) * rsa + kk * csa]);        \
                            acc[r][0] = _mm512_fmadd_ps(av, b0, acc[r][0]); acc[r][1] = _mm512_fmadd_ps(av, b1, acc[r][1]); \
                            acc[r][2] = _mm512_fmadd_ps(av, b2, acc[r][2]); acc[r][3] = _mm512_fmadd_ps(av, b3, acc[r][3]); \
                        }
                        LO_ROW(0) LO_ROW(1) LO_ROW(2) LO_ROW(3)
#undef LO_ROW
                    }
                } else {
                    for (int kk = 0; kk < k; ++kk) {
                        const float *br = b + kk * rsb + j0;
                        __m512 bv[4];
                        for (int cc = 0; cc < nv; ++cc) bv[cc] = _mm512_maskz_loadu_ps(mk[cc], br + 16 * cc);
                        for (int r = 0; r < nr; ++r) {
                            const __m512 av = _mm512_set1_ps(alpha * a[(i0 + r) * rsa + kk * csa]);
                            for (int cc = 0; cc < nv; ++cc) acc[r][cc] = _mm512_fmadd_ps(av, bv[cc], acc[r][cc]);
                        }
                    }
                }
                for (int r = 0; r < nr; ++r) for (int cc = 0; cc < nv; ++cc)
                    _mm512_mask_storeu_ps(c + (size_t)(i0 + r) * n + j0 + 16 * cc, mk[cc], acc[r][cc]);
            }
        }
        return;
    }
#endif
    for (int i = 0; i < m; ++i)
        for (int kk = 0; kk < k; ++kk) {
            float av = alpha * a[i * rsa + kk * csa];
            const float *br = b + kk * rsb;
            float *cr = c + (size_t)i * n;
            /* sequential-k FMA chain per output (x86 GEMM micro-kernels are FMA based) */
            if (csb == 1) for (int j = 0; j < n; ++j) cr[j] = fmaf(av, br[j], cr[j]);
            else for (int j = 0; j < n; ++j) cr[j] = fmaf(av, br[j * csb], cr[j]);
        }
}

void lo_matmul(const float *a, const float *b, int batch_a, int batch_b, int m, int k, int n,
               float *out) { /* gemm.rs:112-222 */
    int fb = batch_a > batch_b ? batch_a : batch_b;
    memset(out, 0, sizeof(float) * (size_t)fb * m * n);
    for (int bi = 0; bi < fb; ++bi)
        lo_sgemm_acc(a + (batch_a == 1 ? 0 : (size_t)bi * m * k), k, 1,
                     b + (batch_b == 1 ? 0 : (size_t)bi * k * n), n, 1, m, k, n, 1.0f,
                     out + (size_t)bi * m * n);
}

void lo_matmul_fused_add(const float *a, const float *b, const float *bias, int bias_len,
                         int batch_a, int batch_b, int m, int k, int n, float *out) {
    /* gemm.rs:223-432: bias.len()==n -> pre-fill rows then accumulate; otherwise
     * matmul followed by a length-modulo broadcast add. */
    int fb = batch_a > batch_b ? batch_a : batch_b;
    if (bias_len == n) {
        for (size_t r = 0; r < (size_t)fb * m; ++r) memcpy(out + r * n, bias, sizeof(float) * n);
        for (int bi = 0; bi < fb; ++bi)
            lo_sgemm_acc(a + (batch_a == 1 ? 0 : (size_t)bi * m * k), k, 1,
                         b + (batch_b == 1 ? 0 : (size_t)bi * k * n), n, 1, m, k, n, 1.0f,
                         out + (size_t)bi * m * n);
    } else {
        lo_matmul(a, b, batch_a, batch_b, m, k, n, out);
        size_t tot = (size_t)fb * m * n;
        for (size_t i = 0; i < tot; ++i) out[i] += bias[i % (size_t)bias_len];
    }
}

void lo_gemm(const float *a, const float *b, const float *c, int c_len, float alpha, float beta,
             int trans_a, int trans_b, int m, int k, int n, float *out) { /* gemm.rs:433-535 */
    size_t tot = (size_t)m * n;
    if (c && beta != 0.0f) {
        if ((size_t)c_len == tot) for (size_t i = 0; i < tot; ++i) out[i] = c[i] * beta;
        else if (c_len == n) for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = c[j] * beta;
        else if (c_len == m) for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = c[i] * beta;
        else if (c_len == 1) { float v = c[0] * beta; for (size_t i = 0; i < tot; ++i) out[i] = v; }
        else for (size_t i = 0; i < tot; ++i) out[i] = c[i % (size_t)c_len] * beta;
    } else memset(out, 0, sizeof(float) * tot);
    lo_sgemm_acc(a, trans_a ? 1 : k, trans_a ? m : 1, b, trans_b ? 1 : n, trans_b ? k : 1, m, k, n,
                 alpha, out);
}

/* ------------------------------------------------------------------------- */
/* activations: x86 SIMD body (first len/8*8 elements) + scalar libm tail      */
/* ------------------------------------------------------------------------- */

/* avx/math.rs:11-66 (Cephes-style expf, FMA Horner, 2^n by exponent bits) */
float lo_cephes_expf(float x) {
    if (x < -87.33654f) x = -87.33654f;
    if (x > 88.72284f) x = 88.72284f;
    float fx = rintf(x * 1.44269504088896341f);
    x = fmaf(-fx, 0.693359375f, x);
    x = fmaf(-fx, -2.12194440e-4f, x);
    float y = fmaf(0.000198712018891638893f, x, 0.00139712726883569741f);
    y = fmaf(y, x, 0.00833345670066840443f);
    y = fmaf(y, x, 0.0416657844442129135f);
    y = fmaf(y, x, 0.166666671633720398f);
    y = fmaf(y, x, 0.5f);
    y = fmaf(y, x, 1.0f);
    y = fmaf(y, x, 1.0f);
    int32_t e = ((int32_t)fx + 127) << 23;
    float p;
    memcpy(&p, &e, 4);
    return y * p;
}
static float lo_sigmoid_simd(float x) { return 1.0f / (1.0f + lo_cephes_expf(-x)); } /* avx/math.rs:69 */
static float lo_tanh_simd(float x) { /* avx/math.rs:81-97 */
    float e = lo_cephes_expf(-x * 2.0f);
    float r = fabsf((1.0f - e) / (1.0f + e));
    return copysignf(r, x);
}
static float lo_erf_simd(float x) { /* avx/math.rs:113-150 (A&S 7.1.26) */
    float ax = fabsf(x);
    float t = 1.0f / fmaf(0.3275911f, ax, 1.0f);
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    float ev = lo_cephes_expf(-(ax * ax));
    float r = fmaf(-(poly * t), ev, 1.0f);
    uint32_t ri, xi;
    memcpy(&ri, &r, 4); memcpy(&xi, &x, 4);
    ri |= (xi & 0x80000000u); /* OR the sign bit in, as _mm256_or_ps does */
    memcpy(&r, &ri, 4);
    return r;
}
static float lo_sigmoid_scalar(float x) { return 1.0f / (1.0f + expf(-x)); } /* activations.rs */

void lo_unary(int op, const float *x, size_t len, float *out) {
    size_t simd_end = (len / 8) * 8;
    for (size_t i = 0; i < len; ++i) {
        float v = x[i];
        int simd = i < simd_end;
        switch (op) {
        case LO_RELU: out[i] = v > 0.0f ? v : 0.0f; break;
        case LO_SIGMOID: out[i] = simd ? lo_sigmoid_simd(v) : lo_sigmoid_scalar(v); break;
        case LO_TANH: out[i] = simd ? lo_tanh_simd(v) : tanhf(v); break;
        case LO_SILU: out[i] = simd ? v * lo_sigmoid_simd(v) : v / (1.0f + expf(-v)); break;
        case LO_ERF: out[i] = simd ? lo_erf_simd(v) : erff(v); break;
        case LO_GELU: out[i] = 0.5f * v * (1.0f + (simd ? lo_erf_simd(v * 0.70710678f) : erff(v * 0.70710678f))); break;
        case LO_EXP: out[i] = simd ? lo_cephes_expf(v) : expf(v); break;
        case LO_SOFTPLUS: out[i] = v > 20.0f ? v : logf(1.0f + expf(v)); break; /* math.rs:1046 */
        default: out[i] = v;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* norms                                                                      */
/* ------------------------------------------------------------------------- */

/* norm.rs:226-312 -> avx/norm.rs:10-133: mean = sum*(1/n), var = sumsq*(1/n) - mean^2
 * (not clamped), inv = 1/sqrt(var+eps), y = fma((x-mean)*inv, gamma, beta) in the SIMD
 * body (first n/8*8 of each row), plain mul+add in the scalar tail.  The SIMD lanes'
 * partial-sum order is not reproduced (sequential sum here). */
/* Row reduction in the exact order of the AVX2 kernels (avx/norm.rs:28-82, :169-205): four
 * 8-lane accumulators over 32-element blocks, merged (s0+s1)+(s2+s3); remaining 8-blocks added
 * to the merged vector; horizontal ((v0+v4)+(v2+v6))+((v1+v5)+(v3+v7)); scalar tail.
 * sq != 0: accumulate x*x (FMA in the vector part, mul+add in the tail). */
static float lo_avx_row_sum(const float *x, int n, int sq) {
    float p[32], v[8];
    int j = 0;
    for (int i = 0; i < 32; ++i) p[i] = 0.0f;
    for (; j + 32 <= n; j += 32)
        for (int i = 0; i < 32; ++i) p[i] = sq ? fmaf(x[j + i], x[j + i], p[i]) : p[i] + x[j + i];
    for (int l = 0; l < 8; ++l) v[l] = (p[l] + p[8 + l]) + (p[16 + l] + p[24 + l]);
    for (; j + 8 <= n; j += 8)
        for (int l = 0; l < 8; ++l) v[l] = sq ? fmaf(x[j + l], x[j + l], v[l]) : v[l] + x[j + l];
    float s = ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]));
    for (; j < n; ++j) s = sq ? s + x[j] * x[j] : s + x[j];
    return s;
}

void lo_layer_norm(const float *x, const float *gamma, const float *beta, int outer, int n,
                   float eps, float *out) {
    float inv_n = 1.0f / (float)n;
    int simd_end = (n / 8) * 8;
    for (int r = 0; r < outer; ++r) {
        const float *xr = x + (size_t)r * n;
        float *o = out + (size_t)r * n;
        float sum = lo_avx_row_sum(xr, n, 0), sq = lo_avx_row_sum(xr, n, 1);
        float mean = sum * inv_n;
        float var = sq * inv_n - mean * mean;
        float inv = 1.0f / sqrtf(var + eps);
        for (int j = 0; j < n; ++j) {
            float g = gamma ? gamma[j] : 1.0f, b = beta ? beta[j] : 0.0f;
            float s = (xr[j] - mean) * inv;
            o[j] = j < simd_end ? fmaf(s, g, b) : s * g + b;
        }
    }
}

/* norm.rs:8-224 (inner_size==1) -> avx/norm.rs:139-229: max, exp(x-max) with the
 * polynomial exp on the SIMD body and libm expf on the tail, multiply by 1/sum. */
void lo_softmax(const float *x, int outer, int n, float *out) {
    int simd_end = (n / 8) * 8;
    for (int r = 0; r < outer; ++r) {
        const float *xr = x + (size_t)r * n;
        float *o = out + (size_t)r * n;
        float mx = -3.402823466e+38f;
        for (int j = 0; j < n; ++j) if (xr[j] > mx) mx = xr[j];
        for (int j = 0; j < n; ++j) o[j] = j < simd_end ? lo_cephes_expf(xr[j] - mx) : expf(xr[j] - mx);
        float sum = lo_avx_row_sum(o, n, 0); /* 4x8-lane accumulators, hsum_ps, scalar tail */
        float inv = 1.0f / sum;
        for (int j = 0; j < n; ++j) o[j] *= inv;
    }
}

/* norm.rs:313-419 : y = scale*(x-mean)/sqrt(var+eps)+bias per channel, NC[inner] */
void lo_batch_norm(const float *x, const float *scale, const float *bias, const float *mean,
                   const float *var, int nb, int c, int inner, float eps, float *out) {
    for (int b = 0; b < nb; ++b)
        for (int ch = 0; ch < c; ++ch) {
            float inv = 1.0f / sqrtf(var[ch] + eps);
            float s = scale[ch] * inv, sh = bias[ch] - mean[ch] * s;
            size_t base = ((size_t)b * c + ch) * inner;
            for (int i = 0; i < inner; ++i) out[base + i] = x[base + i] * s + sh;
        }
}

/* norm.rs:420-506 : x * w / sqrt(mean(x^2)+eps) */
void lo_rms_norm(const float *x, const float *w, int outer, int n, float eps, float *out) {
    for (int r = 0; r < outer; ++r) {
        const float *xr = x + (size_t)r * n;
        float sq = 0.0f;
        for (int j = 0; j < n; ++j) sq += xr[j] * xr[j];
        float inv = 1.0f / sqrtf(sq / (float)n + eps);
        for (int j = 0; j < n; ++j) out[(size_t)r * n + j] = xr[j] * inv * (w ? w[j] : 1.0f);
    }
}

/* ------------------------------------------------------------------------- */
/* convolutions                                                               */
/* ------------------------------------------------------------------------- */
int lo_conv1d_out_len(int l, int k, int pad_l, int pad_r, int stride, int dil) {
    return (l + pad_l + pad_r - dil * (k - 1) - 1) / stride + 1; /* conv1d.rs:889 */
}

/* conv1d.rs:853-1342 : NCL, weight [OC, IC/g, K]; all dispatch arms compute the same
 * cross-correlation; bias then optional ReLU. */
void lo_conv1d(const float *x, const float *w, const float *bias, int nb, int ic, int l, int oc,
               int k, int group, int pad_l, int pad_r, int stride, int dil, int relu, float *out) {
    int ol = lo_conv1d_out_len(l, k, pad_l, pad_r, stride, dil);
    int icg = ic / group, ocg = oc / group;
    for (int b = 0; b < nb; ++b)
        for (int o = 0; o < oc; ++o) {
            int g = o / ocg;
            for (int t = 0; t < ol; ++t) {
                float s = 0.0f;
                for (int c = 0; c < icg; ++c) {
                    const float *xr = x + ((size_t)b * ic + (size_t)g * icg + c) * l;
                    const float *wr = w + ((size_t)o * icg + c) * k;
                    for (int kk = 0; kk < k; ++kk) {
                        int pos = t * stride + kk * dil - pad_l;
                        if (pos >= 0 && pos < l) s += wr[kk] * xr[pos];
                    }
                }
                if (bias) s += bias[o];
                if (relu && s < 0.0f) s = 0.0f;
                out[((size_t)b * oc + o) * ol + t] = s;
            }
        }
}

/* conv2d.rs:176-880 (semantics as tests/regression_kernels.rs:23-69 ref_conv2d):
 * NCHW, weight [OC, IC/g, kh, kw], pads [t,l,b,r]; SiLU = v/(1+exp(-v)). */
void lo_conv2d(const float *x, const float *w, const float *bias, int nb, int ic, int h, int wd,
               int oc, int kh, int kw, int group, const int *pads, const int *strides,
               const int *dils, int act, float *out, int *oh_out, int *ow_out) {
    int pt = pads[0], pl = pads[1], pb = pads[2], pr = pads[3];
    int sh = strides[0], sw = strides[1], dh = dils[0], dw = dils[1];
    int oh = (h + pt + pb - dh * (kh - 1) - 1) / sh + 1;
    int ow = (wd + pl + pr - dw * (kw - 1) - 1) / sw + 1;
    if (oh_out) *oh_out = oh;
    if (ow_out) *ow_out = ow;
    if (!out) return;
    int icg = ic / group, ocg = oc / group;
    for (int b = 0; b < nb; ++b)
        for (int o = 0; o < oc; ++o) {
            int g = o / ocg;
            for (int y = 0; y < oh; ++y)
                for (int xx = 0; xx < ow; ++xx) {
                    float s = 0.0f;
                    for (int c = 0; c < icg; ++c)
                        for (int ky = 0; ky < kh; ++ky) {
                            int iy = y * sh + ky * dh - pt;
                            if (iy < 0 || iy >= h) continue;
                            for (int kx = 0; kx < kw; ++kx) {
                                int ix = xx * sw + kx * dw - pl;
                                if (ix < 0 || ix >= wd) continue;
                                s += x[(((size_t)b * ic + (size_t)g * icg + c) * h + iy) * wd + ix] *
                                     w[(((size_t)o * icg + c) * kh + ky) * kw + kx];
                            }
                        }
                    if (bias) s += bias[o];
                    if (act == 1) s = s > 0.0f ? s : 0.0f;
                    else if (act == 2) s = s / (1.0f + expf(-s));
                    out[(((size_t)b * oc + o) * oh + y) * ow + xx] = s;
                }
        }
}

/* conv2d.rs:2976-3128 : rank-4, group 1, weight [IC, OC, kh, kw]; GEMM col = W^T X with a
 * sequential ic sum (:3069-3087), scatter-add in (kh,kw,ih,iw) order, bias last. */
void lo_conv_transpose(const float *x, const float *w, const float *bias, int nb, int ic, int h,
                       int wd, int oc, int kh, int kw, const int *pads, const int *strides,
                       const int *dils, float *out, int *oh_out, int *ow_out) {
    int pt = pads[0], pl = pads[1], pb = pads[2], pr = pads[3];
    int sh = strides[0], sw = strides[1], dh = dils[0], dw = dils[1];
    int oh = (h - 1) * sh - (pt + pb) + dh * (kh - 1) + 1;
    int ow = (wd - 1) * sw - (pl + pr) + dw * (kw - 1) + 1;
    if (oh_out) *oh_out = oh;
    if (ow_out) *ow_out = ow;
    if (!out) return;
    memset(out, 0, sizeof(float) * (size_t)nb * oc * oh * ow);
    int hw = h * wd, col_rows = oc * kh * kw;
    for (int n = 0; n < nb; ++n) {
        for (int o = 0; o < oc; ++o)
            for (int ky = 0; ky < kh; ++ky)
                for (int kx = 0; kx < kw; ++kx) {
                    int r = (o * kh + ky) * kw + kx;
                    for (int iy = 0; iy < h; ++iy) {
                        int oy = iy * sh + ky * dh;
                        if (oy < pt || oy >= oh + pt) continue;
                        for (int ix = 0; ix < wd; ++ix) {
                            int ox = ix * sw + kx * dw;
                            if (ox < pl || ox >= ow + pl) continue;
                            float s = 0.0f;
                            for (int c = 0; c < ic; ++c)
                                s += w[(size_t)c * col_rows + r] * x[((size_t)n * ic + c) * hw + iy * wd + ix];
                            out[(((size_t)n * oc + o) * oh + (oy - pt)) * ow + (ox - pl)] += s;
                        }
                    }
                }
        if (bias)
            for (int o = 0; o < oc; ++o)
                for (int i = 0; i < oh * ow; ++i) out[((size_t)n * oc + o) * oh * ow + i] += bias[o];
    }
}

/* conv2d.rs:1051-1260 : padded cells ignored (max starts at -inf) */
void lo_max_pool2d(const float *x, int nb, int c, int h, int w, int kh, int kw, const int *pads,
                   const int *strides, const int *dils, int ceil_mode, float *out, int *oh_out,
                   int *ow_out) {
    int pt = pads[0], pl = pads[1], pb = pads[2], pr = pads[3];
    int sh = strides[0], sw = strides[1], dh = dils[0], dw = dils[1];
    int nh = h + pt + pb - dh * (kh - 1) - 1, nw = w + pl + pr - dw * (kw - 1) - 1;
    int oh = (ceil_mode ? (nh + sh - 1) / sh : nh / sh) + 1;
    int ow = (ceil_mode ? (nw + sw - 1) / sw : nw / sw) + 1;
    if (oh_out) *oh_out = oh;
    if (ow_out) *ow_out = ow;
    if (!out) return;
    for (int b = 0; b < nb * c; ++b)
        for (int y = 0; y < oh; ++y)
            for (int xx = 0; xx < ow; ++xx) {
                float m = -INFINITY;
                for (int ky = 0; ky < kh; ++ky) {
                    int iy = y * sh + ky * dh - pt;
                    if (iy < 0 || iy >= h) continue;
                    for (int kx = 0; kx < kw; ++kx) {
                        int ix = xx * sw + kx * dw - pl;
                        if (ix < 0 || ix >= w) continue;
                        float v = x[((size_t)b * h + iy) * w + ix];
                        if (v > m) m = v;
                    }
                }
                out[((size_t)b * oh + y) * ow + xx] = m;
            }
}

/* ------------------------------------------------------------------------- */
/* recurrent                                                                  */
/* ------------------------------------------------------------------------- */
static void lo_gemv(const float *a, const float *x, int m, int k, float *y) {
    for (int i = 0; i < m; ++i) {
        float s = 0.0f;
        for (int j = 0; j < k; ++j) s += a[(size_t)i * k + j] * x[j];
        y[i] = s;
    }
}

/* rnn.rs:67-230 : gates i,o,f,c ; lstm_gates_avx2 rnn.rs:15-64 uses the polynomial
 * sigmoid/tanh on the first H/8*8 units and scalar libm on the tail. */
void lo_lstm(const float *x, const float *w, const float *r, const float *bias, const float *h0,
             const float *c0, int seq, int in_size, int hidden, float *y, float *h, float *c) {
    int m = 4 * hidden, simd_end = (hidden / 8) * 8;
    float *wc = malloc(sizeof(float) * m), *rc = malloc(sizeof(float) * m), *g = malloc(sizeof(float) * m);
    for (int k = 0; k < hidden; ++k) { h[k] = h0 ? h0[k] : 0.0f; c[k] = c0 ? c0[k] : 0.0f; }
    for (int t = 0; t < seq; ++t) {
        lo_gemv(w, x + (size_t)t * in_size, m, in_size, wc);
        lo_gemv(r, h, m, hidden, rc);
        for (int q = 0; q < m; ++q)
            g[q] = wc[q] + rc[q] + (bias ? bias[q] : 0.0f) + (bias ? bias[m + q] : 0.0f);
        for (int k = 0; k < hidden; ++k) {
            int sd = k < simd_end;
            float ig = sd ? lo_sigmoid_simd(g[k]) : lo_sigmoid_scalar(g[k]);
            float og = sd ? lo_sigmoid_simd(g[hidden + k]) : lo_sigmoid_scalar(g[hidden + k]);
            float fg = sd ? lo_sigmoid_simd(g[2 * hidden + k]) : lo_sigmoid_scalar(g[2 * hidden + k]);
            float cg = sd ? lo_tanh_simd(g[3 * hidden + k]) : tanhf(g[3 * hidden + k]);
            float ct = sd ? fmaf(fg, c[k], ig * cg) : fg * c[k] + ig * cg;
            float ht = og * (sd ? lo_tanh_simd(ct) : tanhf(ct));
            c[k] = ct;
            h[k] = ht;
            y[(size_t)t * hidden + k] = ht;
        }
    }
    free(wc); free(rc); free(g);
}

/* rnn.rs:246-357 + gru_gate_fusion_avx2 :360-432 : gates z,r,h;
 * h~ = tanh(Wh x + bWh + r*(Rh h + bRh)) regardless of linear_before_reset (:368). */
void lo_gru(const float *x, const float *w, const float *r, const float *bias, const float *h0,
            int seq, int in_size, int hidden, float *y, float *h) {
    int m = 3 * hidden, simd_end = (hidden / 8) * 8;
    float *wc = malloc(sizeof(float) * m), *rc = malloc(sizeof(float) * m);
    float *hn = malloc(sizeof(float) * hidden);
    for (int k = 0; k < hidden; ++k) h[k] = h0 ? h0[k] : 0.0f;
    for (int t = 0; t < seq; ++t) {
        lo_gemv(w, x + (size_t)t * in_size, m, in_size, wc);
        lo_gemv(r, h, m, hidden, rc);
        for (int k = 0; k < hidden; ++k) {
            int sd = k < simd_end;
            float bwz = bias ? bias[k] : 0.0f, brz = bias ? bias[m + k] : 0.0f;
            float bwr = bias ? bias[hidden + k] : 0.0f, brr = bias ? bias[m + hidden + k] : 0.0f;
            float bwh = bias ? bias[2 * hidden + k] : 0.0f, brh = bias ? bias[m + 2 * hidden + k] : 0.0f;
            float zp, rp;
            if (sd) { zp = (wc[k] + rc[k]) + (bwz + brz); rp = (wc[hidden + k] + rc[hidden + k]) + (bwr + brr); }
            else { zp = wc[k] + rc[k] + bwz + brz; rp = wc[hidden + k] + rc[hidden + k] + bwr + brr; }
            float z = sd ? lo_sigmoid_simd(zp) : lo_sigmoid_scalar(zp);
            float rg = sd ? lo_sigmoid_simd(rp) : lo_sigmoid_scalar(rp);
            float hp = (wc[2 * hidden + k] + bwh) + rg * (rc[2 * hidden + k] + brh);
            float hg = sd ? lo_tanh_simd(hp) : tanhf(hp);
            float ht = sd ? fmaf(1.0f - z, hg, z * h[k]) : (1.0f - z) * hg + z * h[k];
            hn[k] = ht;
        }
        memcpy(h, hn, sizeof(float) * hidden);
        memcpy(y + (size_t)t * hidden, hn, sizeof(float) * hidden);
    }
    free(wc); free(rc); free(hn);
}
