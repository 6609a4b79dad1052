/*
 * lele_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar C restatement of the CPU semantics of miuda-ai/lele's AOT operator hot
 * path (x86_64 arm of src/kernels + src/features).  It is the parity checker and
 * the timed CPU baseline; it is never linked into, imported by, or called from
 * the product path (lele_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Pinning status: pinned against every known-answer vector the reference's own
 * tests hold for this path (tests/test_oracle_kats.py lists them with the
 * reference file:line of each).  NOT pinned at model level: the reference cannot
 * be compiled here (no Rust toolchain) and ships no model files, so
 * SenseVoice-level parity is operator-level parity of a synthetic-weight network
 * ("model-level parity unpinned").
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Compile with -ffp-contract=off: Rust never contracts a*b+c;
 * where the reference's x86 path uses an explicit FMA this file calls fmaf().
 */
#ifndef LELE_ORACLE_H
#define LELE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- front-end (src/features) ---- */
void lo_hann_window(int size, float *out);                                  /* window.rs:2 */
void lo_precompute_twiddles(int n, float *tw_re, float *tw_im, int32_t *bit_rev); /* kernels/fft.rs:136 */
void lo_rfft_forward(const float *input, int n, float *out_re, float *out_im);     /* kernels/fft.rs:79 */
float lo_hz_to_mel_htk(float hz);                                           /* mel.rs:1 */
float lo_mel_to_hz_htk(float mel);                                          /* mel.rs:4 */
void lo_mel_filterbank(float sr, int n_fft, int n_mels, float f_min, float f_max, float *w); /* mel.rs:7 */
/* SenseVoiceFrontend::compute pipeline.rs:67 (25ms/10ms/80 mel/LFR 7,6 at 16 kHz).
 * mel_out may be NULL; returns T_lfr (0 when pcm shorter than one frame). */
int lo_frontend_num_frames(int n_samples);
int lo_frontend_compute(const float *pcm, int n_samples, float *mel_out /*[frames,80]*/,
                        float *lfr_out /*[T_lfr,560]*/);
void lo_lfr(const float *in, int t, int d, int m, int n, float *out);       /* lfr.rs:18 */
void lo_cmvn(const float *in, int t, int d, float eps, float *out);         /* cmvn.rs:14 */
/* math.rs:2304 / :2372 ; window may be NULL (periodic Hann default) */
int lo_stft(const float *sig, int signal_len, int n_fft, int hop, int win, const float *window,
            int power, float *out);

/* ---- quantised linear (src/kernels/quantization.rs + avx/quantization.rs) ---- */
void lo_dynamic_quantize_linear(const float *x, size_t len, float *q, float *scale, float *zp); /* avx/quantization.rs:832 */
/* quantization.rs:29 (mat_mul_integer_with_scale_bias_activation): a,b hold integer
 * values as f32; scale NULL|len 1|len n; bias NULL|len n. */
void lo_mat_mul_integer(const float *a, const float *b, int batch, int m, int k, int n,
                        float a_zp, float b_zp, const float *scale, int scale_len,
                        const float *bias, int relu, float *out);
/* quantization.rs:77 -> avx/quantization.rs:225.  w is the u8 weight [k,n];
 * dynamic quantisation is per [m,k] slice of the batch. */
void lo_fused_quantized_linear(const float *x, int batch, int m, int k, int n, const uint8_t *w,
                               const float *w_scale, int w_scale_len, int w_zp,
                               const float *bias, int relu, float *out);

/* prepared form (prepare_weights, quantization.rs:221): wt [n,k], colsum [n] */
void lo_prepare_weights(const uint8_t *w, int k, int n, uint8_t *wt, int32_t *colsum);
/* optional VNNI repack of a prepared weight (NULL on hosts without AVX-512 VNNI) + the entry that takes it */
int8_t *lo_pack_weights_vnni(const uint8_t *wt, int k, int n);
void lo_free_packed(int8_t *wp);
void lo_fused_quantized_linear_packed(const float *x, int batch, int m, int k, int n, const uint8_t *wt, const int8_t *wp,
                                      const int32_t *colsum, const float *w_scale, int w_scale_len,
                                      int w_zp, const float *bias, int relu, float *out);
void lo_fused_quantized_linear_prepared(const float *x, int batch, int m, int k, int n, const uint8_t *wt,
                                        const int32_t *colsum, const float *w_scale, int w_scale_len,
                                        int w_zp, const float *bias, int relu, float *out);

/* ---- f32 GEMM (src/kernels/gemm.rs; arithmetic delegated to faer 0.24 upstream) ---- */
void lo_matmul(const float *a, const float *b, int batch_a, int batch_b, int m, int k, int n, float *out); /* gemm.rs:112 */
void lo_matmul_fused_add(const float *a, const float *b, const float *bias, int bias_len,
                         int batch_a, int batch_b, int m, int k, int n, float *out);    /* gemm.rs:223 */
void lo_gemm(const float *a, const float *b, const float *c, int c_len, float alpha, float beta,
             int trans_a, int trans_b, int m, int k, int n, float *out);                /* gemm.rs:433 */

/* ---- norms (src/kernels/norm.rs + avx/norm.rs) ---- */
void lo_layer_norm(const float *x, const float *gamma, const float *beta, int outer, int n,
                   float eps, float *out);                                               /* norm.rs:226 */
void lo_softmax(const float *x, int outer, int n, float *out);                           /* norm.rs:8 (last axis) */
void lo_batch_norm(const float *x, const float *scale, const float *bias, const float *mean,
                   const float *var, int nb, int c, int inner, float eps, float *out);   /* norm.rs:313 */
void lo_rms_norm(const float *x, const float *w, int outer, int n, float eps, float *out); /* norm.rs:420 */

/* ---- activations with the x86 SIMD-body/scalar-tail split (avx/math.rs) ---- */
enum { LO_RELU = 0, LO_SIGMOID, LO_TANH, LO_SILU, LO_ERF, LO_GELU, LO_EXP, LO_SOFTPLUS };
void lo_unary(int op, const float *x, size_t len, float *out);
float lo_cephes_expf(float x);                                                           /* avx/math.rs:11 */

/* ---- convolutions ---- */
int lo_conv1d_out_len(int l, int k, int pad_l, int pad_r, int stride, int dil);
void lo_conv1d(const float *x, const float *w, const float *bias, int nb, int ic, int l, int oc,
               int k, int group, int pad_l, int pad_r, int stride, int dil, int relu, float *out); /* conv1d.rs:853 */
/* act: 0 none, 1 relu, 2 silu  (conv2d.rs:176) ; pads t,l,b,r */
void lo_conv2d(const float *x, const float *w, const float *bias, int nb, int ic, int h, int wd,
               int oc, int kh, int kw, int group, const int *pads, const int *strides,
               const int *dils, int act, float *out, int *oh_out, int *ow_out);
void lo_conv_transpose(const float *x, const float *w, const float *bias, int nb, int ic, int h,
                       int wd, int oc, int kh, int kw, const int *pads, const int *strides,
                       const int *dils, float *out, int *oh_out, int *ow_out);           /* conv2d.rs:2976 */
void lo_max_pool2d(const float *x, int nb, int c, int h, int w, int kh, int kw, const int *pads,
                   const int *strides, const int *dils, int ceil_mode, float *out, int *oh_out,
                   int *ow_out);                                                         /* conv2d.rs:1051 */

/* ---- recurrent (src/kernels/rnn.rs) ---- */
void lo_lstm(const float *x, const float *w, const float *r, const float *bias, const float *h0,
             const float *c0, int seq, int in_size, int hidden, float *y, float *h, float *c); /* rnn.rs:67 */
void lo_gru(const float *x, const float *w, const float *r, const float *bias, const float *h0,
            int seq, int in_size, int hidden, float *y, float *h);                       /* rnn.rs:246 */

/* ---- SenseVoice-shaped synthetic network: the op-by-op call sequence a lele_gen
 *      model.rs would make (examples/sensevoice/src/main.rs:140, wasm_bench.rs:888-1113) ---- */
typedef struct lo_sv_model lo_sv_model;
/* blob layout is produced by lele_b200/sensevoice_weights.py (shared with the CUDA runner). */
lo_sv_model *lo_sv_create(const uint8_t *blob, size_t nbytes);
void lo_sv_destroy(lo_sv_model *m);
int lo_sv_vocab(const lo_sv_model *m);
/* feats: CMVN'd [t,560]; logits out [t+4, vocab]; n_layers_limit<0 = all. Returns rows. */
int lo_sv_forward(const lo_sv_model *m, const float *feats, int t, int lang, int textnorm,
                  int n_layers_limit, float *logits);
/* Full path PCM -> ids (front-end + CMVN + encoder + argmax). ids [t_lfr+4]. returns rows */
int lo_sv_pcm_to_ids(const lo_sv_model *m, const float *pcm, int n_samples, int lang,
                     int textnorm, int32_t *ids, float *logits_opt);

#ifdef __cplusplus
}
#endif
#endif
