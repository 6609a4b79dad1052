/*
 * sensevoice_ref.c -- TEST INFRASTRUCTURE ONLY (see lele_oracle.h).
 *
 * The SenseVoiceSmall-shaped network executed the way a lele_gen-compiled model.rs
 * executes it: one clip at a time, a straight-line sequence of lele::kernels::* calls
 * (examples/sensevoice/src/main.rs:140; layer composition documented by the reference
 * in src/bin/wasm_bench.rs:888-1113: layer_norm -> int8 QKV linear -> QK^T -> softmax
 * -> attn*V -> int8 out-proj -> layer_norm -> int8 FFN1+ReLU -> int8 FFN2 + ~9
 * element-wise ops).  The SANM details the reference does not show (FSMN memory block =
 * depthwise conv1d k=11 over time on V, 4 prompt rows, sqrt(d) scaling + sinusoidal
 * positions, after_norm / tp_norm, CTC head) follow the public FunASR SenseVoiceSmall
 * architecture; weights are synthetic (no model file exists).
 *
 * Blob layout: lele_b200/sensevoice_weights.py (shared with the CUDA runner).
 */
#include "lele_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { SV_G_EMBED = 0, SV_G_POS, SV_G_AFTER_G, SV_G_AFTER_B, SV_G_TP_G, SV_G_TP_B, SV_G_CTC_W,
       SV_G_CTC_SCALE, SV_G_CTC_BIAS, SV_G_CTC_ZP, SV_NUM_GLOBAL };
enum { SV_L_LN1_G = 0, SV_L_LN1_B, SV_L_QKV_W, SV_L_QKV_SCALE, SV_L_QKV_BIAS, SV_L_QKV_ZP,
       SV_L_FSMN_W, SV_L_OUT_W, SV_L_OUT_SCALE, SV_L_OUT_BIAS, SV_L_OUT_ZP, SV_L_LN2_G,
       SV_L_LN2_B, SV_L_FFN1_W, SV_L_FFN1_SCALE, SV_L_FFN1_BIAS, SV_L_FFN1_ZP, SV_L_FFN2_W,
       SV_L_FFN2_SCALE, SV_L_FFN2_BIAS, SV_L_FFN2_ZP, SV_NUM_LAYER };

struct lo_sv_model {
    const uint8_t *blob;
    int n_layers, d, d_in, ffn, heads, fsmn_k, vocab, n_embed, max_t, n_stage1, n_tensors;
    const uint64_t *table;
    /* prepared weights (one-time transpose + column sums, the reference's B_WEIGHT_CACHE) */
    uint8_t **wt;      /* [n_layers*4 + 1] */
    int32_t **colsum;
    int8_t **wp;       /* VNNI repack of wt (NULL entries on hosts without AVX-512 VNNI) */
};

static const void *sv_t(const lo_sv_model *m, int idx) { return m->blob + m->table[2 * idx]; }
static const void *sv_l(const lo_sv_model *m, int layer, int which) {
    return sv_t(m, SV_NUM_GLOBAL + layer * SV_NUM_LAYER + which);
}

lo_sv_model *lo_sv_create(const uint8_t *blob, size_t nbytes) {
    if (nbytes < 256) return NULL;
    const int32_t *h = (const int32_t *)blob;
    if (h[0] != 0x454C454C || h[1] != 1) return NULL;
    lo_sv_model *m = calloc(1, sizeof(*m));
    m->blob = blob;
    m->n_layers = h[2]; m->d = h[3]; m->d_in = h[4]; m->ffn = h[5]; m->heads = h[6];
    m->fsmn_k = h[7]; m->vocab = h[8]; m->n_embed = h[9]; m->max_t = h[10]; m->n_stage1 = h[11];
    m->n_tensors = h[12];
    m->table = (const uint64_t *)(blob + 256);
    int nl = m->n_layers * 4 + 1;
    m->wt = calloc(nl, sizeof(uint8_t *));
    m->colsum = calloc(nl, sizeof(int32_t *));
    m->wp = calloc(nl, sizeof(int8_t *));
    for (int l = 0; l < m->n_layers; ++l) {
        int cur = l == 0 ? m->d_in : m->d;
        int ks[4] = { cur, m->d, m->d, m->ffn }, ns[4] = { 3 * m->d, m->d, m->ffn, m->d };
        int which[4] = { SV_L_QKV_W, SV_L_OUT_W, SV_L_FFN1_W, SV_L_FFN2_W };
        for (int q = 0; q < 4; ++q) {
            m->wt[l * 4 + q] = malloc((size_t)ks[q] * ns[q]);
            m->colsum[l * 4 + q] = malloc(sizeof(int32_t) * ns[q]);
            lo_prepare_weights(sv_l(m, l, which[q]), ks[q], ns[q], m->wt[l * 4 + q], m->colsum[l * 4 + q]);
            m->wp[l * 4 + q] = lo_pack_weights_vnni(m->wt[l * 4 + q], ks[q], ns[q]);
        }
    }
    m->wt[nl - 1] = malloc((size_t)m->d * m->vocab);
    m->colsum[nl - 1] = malloc(sizeof(int32_t) * m->vocab);
    lo_prepare_weights(sv_t(m, SV_G_CTC_W), m->d, m->vocab, m->wt[nl - 1], m->colsum[nl - 1]);
    m->wp[nl - 1] = lo_pack_weights_vnni(m->wt[nl - 1], m->d, m->vocab);
    return m;
}
void lo_sv_destroy(lo_sv_model *m) {
    if (!m) return;
    int nl = m->n_layers * 4 + 1;
    for (int i = 0; i < nl; ++i) { free(m->wt[i]); free(m->colsum[i]); lo_free_packed(m->wp[i]); }
    free(m->wt); free(m->colsum); free(m->wp); free(m);
}
int lo_sv_vocab(const lo_sv_model *m) { return m->vocab; }

static void sv_add(const float *a, const float *b, size_t n, float *o) { for (size_t i = 0; i < n; ++i) o[i] = a[i] + b[i]; }

int lo_sv_forward(const lo_sv_model *m, const float *feats, int t, int lang, int textnorm,
                  int n_layers_limit, float *logits) {
    const int d = m->d, din = m->d_in, H = m->heads, dk = d / H, T = t + 4, ffn = m->ffn;
    const int n_layers = (n_layers_limit >= 0 && n_layers_limit < m->n_layers) ? n_layers_limit : m->n_layers;
    const float *embed = sv_t(m, SV_G_EMBED), *pos = sv_t(m, SV_G_POS);
    size_t wide = (size_t)T * (din > d ? din : d);
    float *x = malloc(sizeof(float) * wide), *h = malloc(sizeof(float) * wide);
    float *qkv = malloc(sizeof(float) * (size_t)T * 3 * d);
    float *vt = malloc(sizeof(float) * (size_t)T * d), *ft = malloc(sizeof(float) * (size_t)T * d);
    float *fsmn = malloc(sizeof(float) * (size_t)T * d);
    float *qh = malloc(sizeof(float) * (size_t)T * d), *kh = malloc(sizeof(float) * (size_t)T * d),
          *vh = malloc(sizeof(float) * (size_t)T * d), *oh = malloc(sizeof(float) * (size_t)T * d),
          *om = malloc(sizeof(float) * (size_t)T * d), *att = malloc(sizeof(float) * (size_t)T * d);
    float *sc = malloc(sizeof(float) * (size_t)H * T * T), *pr = malloc(sizeof(float) * (size_t)H * T * T);
    float *f1 = malloc(sizeof(float) * (size_t)T * ffn), *f2 = malloc(sizeof(float) * (size_t)T * d);

    /* gather(embed,[lang,1,2,textnorm]) ++ concat(axis 0) with the features */
    int ids[4] = { lang, 1, 2, textnorm };
    for (int r = 0; r < 4; ++r) memcpy(x + (size_t)r * din, embed + (size_t)ids[r] * din, sizeof(float) * din);
    memcpy(x + (size_t)4 * din, feats, sizeof(float) * (size_t)t * din);
    /* mul by sqrt(d) (scalar operand), add positional table rows */
    float sq = sqrtf((float)d);
    for (size_t i = 0; i < (size_t)T * din; ++i) x[i] = x[i] * sq;
    for (size_t i = 0; i < (size_t)T * din; ++i) x[i] = x[i] + pos[i];

    float qscale = 1.0f / sqrtf((float)dk);
    int cur = din;
    for (int l = 0; l < n_layers; ++l) {
        lo_layer_norm(x, sv_l(m, l, SV_L_LN1_G), sv_l(m, l, SV_L_LN1_B), T, cur, 1e-5f, h);
        lo_fused_quantized_linear_packed(h, 1, T, cur, 3 * d, m->wt[l * 4 + 0], m->wp[l * 4 + 0], m->colsum[l * 4 + 0], sv_l(m, l, SV_L_QKV_SCALE),
                                  3 * d, *(const uint8_t *)sv_l(m, l, SV_L_QKV_ZP),
                                  sv_l(m, l, SV_L_QKV_BIAS), 0, qkv);
        /* split + head transposes: q [H,T,dk] (scaled), k^T [H,dk,T], v [H,T,dk]; v^T [d,T] */
        for (int i = 0; i < T; ++i)
            for (int c = 0; c < d; ++c) {
                int hd = c / dk, e = c % dk;
                float qv = qkv[(size_t)i * 3 * d + c], kv = qkv[(size_t)i * 3 * d + d + c],
                      vv = qkv[(size_t)i * 3 * d + 2 * d + c];
                qh[((size_t)hd * T + i) * dk + e] = qv * qscale;
                kh[((size_t)hd * dk + e) * T + i] = kv;
                vh[((size_t)hd * T + i) * dk + e] = vv;
                vt[(size_t)c * T + i] = vv;
            }
        /* FSMN memory: depthwise conv1d over time (k=11, pad 5/5, no bias) + v */
        int pad = (m->fsmn_k - 1) / 2;
        lo_conv1d(vt, sv_l(m, l, SV_L_FSMN_W), NULL, 1, d, T, d, m->fsmn_k, d, pad, m->fsmn_k - 1 - pad, 1, 1, 0, ft);
        for (int i = 0; i < T; ++i)
            for (int c = 0; c < d; ++c)
                fsmn[(size_t)i * d + c] = ft[(size_t)c * T + i] + qkv[(size_t)i * 3 * d + 2 * d + c];
        /* attention: matmul, softmax(last axis), matmul */
        lo_matmul(qh, kh, H, H, T, dk, T, sc);
        lo_softmax(sc, H * T, T, pr);
        lo_matmul(pr, vh, H, H, T, T, dk, oh);
        for (int hd = 0; hd < H; ++hd)
            for (int i = 0; i < T; ++i)
                memcpy(om + (size_t)i * d + (size_t)hd * dk, oh + ((size_t)hd * T + i) * dk, sizeof(float) * dk);
        lo_fused_quantized_linear_packed(om, 1, T, d, d, m->wt[l * 4 + 1], m->wp[l * 4 + 1], m->colsum[l * 4 + 1], sv_l(m, l, SV_L_OUT_SCALE), d,
                                  *(const uint8_t *)sv_l(m, l, SV_L_OUT_ZP), sv_l(m, l, SV_L_OUT_BIAS), 0, att);
        sv_add(att, fsmn, (size_t)T * d, att);
        if (cur == d) sv_add(x, att, (size_t)T * d, x);
        else memcpy(x, att, sizeof(float) * (size_t)T * d);
        cur = d;
        lo_layer_norm(x, sv_l(m, l, SV_L_LN2_G), sv_l(m, l, SV_L_LN2_B), T, d, 1e-5f, h);
        lo_fused_quantized_linear_packed(h, 1, T, d, ffn, m->wt[l * 4 + 2], m->wp[l * 4 + 2], m->colsum[l * 4 + 2], sv_l(m, l, SV_L_FFN1_SCALE), ffn,
                                  *(const uint8_t *)sv_l(m, l, SV_L_FFN1_ZP), sv_l(m, l, SV_L_FFN1_BIAS), 1, f1);
        lo_fused_quantized_linear_packed(f1, 1, T, ffn, d, m->wt[l * 4 + 3], m->wp[l * 4 + 3], m->colsum[l * 4 + 3], sv_l(m, l, SV_L_FFN2_SCALE), d,
                                  *(const uint8_t *)sv_l(m, l, SV_L_FFN2_ZP), sv_l(m, l, SV_L_FFN2_BIAS), 0, f2);
        sv_add(x, f2, (size_t)T * d, x);
        if (l == m->n_stage1 - 1) {
            lo_layer_norm(x, sv_t(m, SV_G_AFTER_G), sv_t(m, SV_G_AFTER_B), T, d, 1e-5f, h);
            memcpy(x, h, sizeof(float) * (size_t)T * d);
        }
    }
    if (n_layers == m->n_layers) {
        lo_layer_norm(x, sv_t(m, SV_G_TP_G), sv_t(m, SV_G_TP_B), T, cur, 1e-5f, h);
        lo_fused_quantized_linear_packed(h, 1, T, cur, m->vocab, m->wt[m->n_layers * 4], m->wp[m->n_layers * 4], m->colsum[m->n_layers * 4], sv_t(m, SV_G_CTC_SCALE), m->vocab,
                                  *(const uint8_t *)sv_t(m, SV_G_CTC_ZP), sv_t(m, SV_G_CTC_BIAS), 0, logits);
    } else {
        /* truncated run (tests): expose the hidden state instead of logits */
        memcpy(logits, x, sizeof(float) * (size_t)T * cur);
    }
    free(x); free(h); free(qkv); free(vt); free(ft); free(fsmn); free(qh); free(kh); free(vh);
    free(oh); free(om); free(att); free(sc); free(pr); free(f1); free(f2);
    return T;
}

int lo_sv_pcm_to_ids(const lo_sv_model *m, const float *pcm, int n_samples, int lang,
                     int textnorm, int32_t *ids, float *logits_opt) {
    int frames = lo_frontend_num_frames(n_samples);
    if (frames == 0) return 0;
    int t = (frames + 5) / 6, T = t + 4;
    float *lfr = malloc(sizeof(float) * (size_t)t * 560), *cm = malloc(sizeof(float) * (size_t)t * 560);
    lo_frontend_compute(pcm, n_samples, NULL, lfr);
    lo_cmvn(lfr, t, 560, 1e-5f, cm);
    float *logits = logits_opt ? logits_opt : malloc(sizeof(float) * (size_t)T * m->vocab);
    lo_sv_forward(m, cm, t, lang, textnorm, -1, logits);
    for (int i = 0; i < T; ++i) { /* greedy argmax; Iterator::max_by keeps the LAST max (tokenizer.rs:55-59) */
        const float *r = logits + (size_t)i * m->vocab;
        int best = 0;
        for (int j = 1; j < m->vocab; ++j) if (r[j] >= r[best]) best = j;
        ids[i] = best;
    }
    if (!logits_opt) free(logits);
    free(lfr); free(cm);
    return T;
}
