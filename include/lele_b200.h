/*
 * lele_b200.h -- C ABI of the Blackwell (sm_100a) back-end for lele's AOT operator path.
 *
 * Drop-in boundary: lele's generated `model.rs` calls free functions in
 * `lele::kernels::*` / `lele::features::*`; a `cfg(feature = "cuda")` arm inside each of those
 * functions binds the entry point listed here (the same pattern the reference already uses
 * for Apple Accelerate: `unsafe extern "C" { fn cblas_sgemm(..) }`, src/kernels/gemm.rs:30-50,
 * linked from build.rs:10-12).  INTEGRATION.md shows the Rust stubs.
 *
 * Conventions
 *  - Plain C: pointers + sizes, no torch / C++ types.  Every tensor pointer is a DEVICE
 *    pointer unless the parameter is documented "host".  Tensors are dense row-major f32
 *    unless stated (lele's TensorView<f32>, src/tensor.rs:5).
 *  - Every call is asynchronous on the context's stream; lele_b200_sync() joins it.
 *  - Return value: 0 = ok, non-zero = error (lele panics on precondition failures --
 *    conv2d.rs:196-205, rnn.rs:85-90 -- so the Rust shim turns non-zero into
 *    panic!("{}", lele_b200_last_error())).  No CPU fallback exists: without a CUDA device
 *    every compute entry point fails with LELE_B200_ERR_CUDA.
 *  - "n_clips"/"n_slices" arguments are the batch the B200 back-end adds below the
 *    boundary: the reference processes one clip per call (batch is baked to 1 in generated
 *    code), so per-tensor semantics (dynamic quantisation, CMVN) are applied per
 *    slice.
 */
#ifndef LELE_B200_H
#define LELE_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lele_b200_ctx lele_b200_ctx;
typedef struct lele_b200_qweights lele_b200_qweights;
typedef struct lele_b200_sensevoice lele_b200_sensevoice;
typedef struct lele_b200_comm lele_b200_comm;
typedef struct lele_b200_graph lele_b200_graph;

enum { LELE_B200_OK = 0, LELE_B200_ERR_ARG = 1, LELE_B200_ERR_CUDA = 2, LELE_B200_ERR_UNSUPPORTED = 3 };

/* ---- context / memory: src/tensor.rs (TensorView, weight views) + kernels/utils.rs:10 ensure_capacity,
 *      generated <Model>Workspace (src/compiler/mod.rs:1057-1092) ---- */
const char* lele_b200_last_error(void);
int lele_b200_device_count(void);
/* stream: a cudaStream_t (may be NULL = create a private non-blocking stream). */
int lele_b200_ctx_create(int device, void* stream, lele_b200_ctx** out);
int lele_b200_ctx_destroy(lele_b200_ctx* ctx);
int lele_b200_sync(lele_b200_ctx* ctx);
unsigned long long lele_b200_launch_count(const lele_b200_ctx* ctx);
int lele_b200_malloc(lele_b200_ctx* ctx, size_t nbytes, void** dptr);
int lele_b200_free(lele_b200_ctx* ctx, void* dptr);
/* page-locked host buffers for asynchronous, full-rate h2d / d2h */
int lele_b200_malloc_host(lele_b200_ctx* ctx, size_t nbytes, void** hptr);
int lele_b200_free_host(lele_b200_ctx* ctx, void* hptr);
int lele_b200_memset(lele_b200_ctx* ctx, void* dptr, int value, size_t nbytes);
int lele_b200_h2d(lele_b200_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);
int lele_b200_d2h(lele_b200_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes);
int lele_b200_d2d(lele_b200_ctx* ctx, void* dst_dev, const void* src_dev, size_t nbytes);
/* Arena: device mirror of a host `Vec<f32>` workspace buffer (ws.buf_N) or of the weights
 * blob, keyed by the host base pointer; grow-only like ensure_capacity. */
int lele_b200_arena_bind(lele_b200_ctx* ctx, const void* host_base, size_t nbytes, void** dptr);
int lele_b200_arena_release(lele_b200_ctx* ctx, const void* host_base);

/* Streams and graphs.  stream_fork: `lane`'s stream waits for everything enqueued on ctx so far; stream_join: ctx's stream waits
 * for everything enqueued on `lane` (same device).  capture_begin .. capture_end records every call made on ctx (and on lanes
 * forked from it and joined back) into a CUDA graph instead of executing it -- valid once the arena has its steady-state sizes,
 * i.e. after one ordinary forward; nothing on the captured path may synchronise (lele_b200_sync, d2h of results, arena growth).
 * lane_launches = kernel launches the caller counted on the lane contexts during the capture (launch_count deltas), so that
 * graph_launch keeps lele_b200_launch_count(ctx) truthful. */
int lele_b200_stream_fork(lele_b200_ctx* ctx, lele_b200_ctx* lane);
int lele_b200_stream_join(lele_b200_ctx* ctx, lele_b200_ctx* lane);
int lele_b200_capture_begin(lele_b200_ctx* ctx);
int lele_b200_capture_end(lele_b200_ctx* ctx, unsigned long long lane_launches, lele_b200_graph** out);
int lele_b200_graph_launch(lele_b200_ctx* ctx, lele_b200_graph* graph);
int lele_b200_graph_destroy(lele_b200_ctx* ctx, lele_b200_graph* graph);

/* ---- lele::features (src/features/ *.rs) ---- */
int lele_b200_hann_window(int size, float* out_host);                          /* window.rs:2 */
int lele_b200_mel_filterbank(float sample_rate, int n_fft, int n_mels, float f_min, float f_max,
                             float* out_host /*[n_mels, n_fft/2+1]*/);           /* mel.rs:7 */
int lele_b200_frontend_num_frames(int n_samples);                               /* pipeline.rs:73 */
int lele_b200_frontend_out_rows(int n_samples);                                 /* lfr.rs:34 */
/* SenseVoiceFrontend::compute (pipeline.rs:67) for n_clips equal-length clips.
 * pcm [n_clips, clip_stride>=n_samples]; mel_opt NULL or [n_clips, frames, 80];
 * lfr_out [n_clips, T_lfr, 560]. */
int lele_b200_frontend_compute(lele_b200_ctx* ctx, const float* pcm, int n_clips, int n_samples,
                               long long clip_stride, float* mel_opt, float* lfr_out);
int lele_b200_rfft(lele_b200_ctx* ctx, const float* x, int n_rows, int n, float* out_re,
                   float* out_im);                                               /* features/fft.rs:18, kernels/fft.rs:51 */
int lele_b200_lfr(lele_b200_ctx* ctx, const float* in, int n_clips, int t, int d, int m, int n,
                  float* out);                                                  /* lfr.rs:18 */
int lele_b200_cmvn(lele_b200_ctx* ctx, const float* in, int n_clips, int t, int d, float eps,
                   float* out);                                                 /* cmvn.rs:14 */
/* math.rs:2304 (power=0: [frames, n_fft/2+1, 2]) and :2372 (power=1: [frames, n_fft/2+1]);
 * window NULL = periodic Hann.  n_fft power of two <= 4096. Returns frames via *frames_out. */
int lele_b200_stft(lele_b200_ctx* ctx, const float* signal, int signal_len, int n_fft, int hop,
                   int win, const float* window, int power, float* out, int* frames_out);

/* ---- norm.rs ---- */
int lele_b200_layer_norm(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta,
                         long long outer, int n, float eps, float* out);        /* norm.rs:226 */
int lele_b200_softmax(lele_b200_ctx* ctx, const float* x, long long outer, int n, float* out); /* norm.rs:8 */
int lele_b200_batch_norm(lele_b200_ctx* ctx, const float* x, const float* scale, const float* bias,
                         const float* mean, const float* var, int nb, int c, long long inner,
                         float eps, float* out);                                /* norm.rs:313 */
int lele_b200_rms_norm(lele_b200_ctx* ctx, const float* x, const float* w, long long outer, int n,
                       float eps, float* out);                                  /* norm.rs:420 */

/* ---- quantization.rs ---- */
/* dynamic_quantize_linear (quantization.rs:1628): per slice; q holds integer values as f32;
 * scale/zp [n_slices]. */
int lele_b200_dynamic_quantize_linear(lele_b200_ctx* ctx, const float* x, int n_slices,
                                      long long slice_len, float* q, float* scale, float* zp);
/* mat_mul_integer[_with_scale_bias[_relu]] (quantization.rs:8-72): a [batch,m,k], b [k,n] hold
 * integer values as f32; scale NULL | [1] | [n]; bias NULL | [n]. */
int lele_b200_mat_mul_integer(lele_b200_ctx* ctx, const float* a, const float* b, int batch, int m,
                              int k, int n, float a_zp, float b_zp, const float* scale,
                              int scale_len, const float* bias, int relu, float* out);
/* same with a batched b [batch_b, k, n] (quantization.rs:1157-1173): batch_a == batch_b, or the side with batch 1 is broadcast;
 * out [max(batch_a, batch_b), m, n]. */
int lele_b200_mat_mul_integer_batched(lele_b200_ctx* ctx, const float* a, const float* b, int batch_a, int batch_b,
                                      int m, int k, int n, float a_zp, float b_zp, const float* scale, int scale_len,
                                      const float* bias, int relu, float* out);
/* prepare_weights (quantization.rs:221) / B_WEIGHT_CACHE (avx/quantization.rs:47-95): one-time
 * transpose of the u8 weight [k,n] to the K-major layout the tensor cores read + column sums.
 * w_scale_len 1 or n; bias NULL or [n]. */
int lele_b200_prepare_weights(lele_b200_ctx* ctx, const uint8_t* w, int k, int n,
                              const float* w_scale, int w_scale_len, int w_zp, const float* bias,
                              lele_b200_qweights** out);
int lele_b200_qweights_destroy(lele_b200_ctx* ctx, lele_b200_qweights* w);
/* fused_quantized_linear (quantization.rs:77): x [n_slices, m, k] -> out [n_slices, m, n];
 * min/max/scale/zero-point per slice, exact integer GEMM on tcgen05 (kind::i8), fused
 * zero-point corrections + scale + bias (+ReLU) epilogue. */
int lele_b200_fused_quantized_linear(lele_b200_ctx* ctx, const float* x, int n_slices, int m,
                                     const lele_b200_qweights* w, int relu, float* out);

/* ---- gemm.rs ---- */
int lele_b200_matmul(lele_b200_ctx* ctx, const float* a, const float* b, int batch_a, int batch_b,
                     int m, int k, int n, float* out);                          /* gemm.rs:112 */
int lele_b200_matmul_fused_add(lele_b200_ctx* ctx, const float* a, const float* b, const float* bias,
                               int bias_len, int batch_a, int batch_b, int m, int k, int n,
                               float* out);                                     /* gemm.rs:223 */
int lele_b200_gemm(lele_b200_ctx* ctx, const float* a, const float* b, const float* c, int c_len,
                   float alpha, float beta, int trans_a, int trans_b, int m, int k, int n,
                   float* out);                                                 /* gemm.rs:433 */

/* ---- conv1d.rs / conv2d.rs ---- */
int lele_b200_conv1d(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias, int nb,
                     int ic, int l, int oc, int k, int group, int pad_l, int pad_r, int stride,
                     int dilation, int relu, float* out);                       /* conv1d.rs:837,853 */
/* act: 0 none (conv2d.rs:107), 1 ReLU (conv2d_fused :155), 2 SiLU (conv2d_silu :124); pads t,l,b,r */
int lele_b200_conv2d(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias, int nb,
                     int ic, int h, int wd, int oc, int kh, int kw, int group, const int* pads_host,
                     const int* strides_host, const int* dilations_host, int act, float* out);
/* conv_integer (conv2d.rs:2216, emitted by ops/nn.rs:328): (x - x_zp) (*) (w - w_zp); padded positions hold RAW zeros, i.e. contribute
 * (0 - x_zp) (conv2d.rs:2025).  x, w hold integer values as f32; x_zp / w_zp = first element of the zero-point tensor or 0. */
int lele_b200_conv_integer(lele_b200_ctx* ctx, const float* x, const float* w, float x_zp, float w_zp, int nb, int ic,
                           int h, int wd, int oc, int kh, int kw, int group, const int* pads_host,
                           const int* strides_host, const int* dilations_host, float* out);
int lele_b200_conv_transpose(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias,
                             int nb, int ic, int h, int wd, int oc, int kh, int kw,
                             const int* pads_host, const int* strides_host,
                             const int* dilations_host, float* out);            /* conv2d.rs:2952 */
int lele_b200_max_pool2d(lele_b200_ctx* ctx, const float* x, int nb, int c, int h, int w, int kh,
                         int kw, const int* pads_host, const int* strides_host,
                         const int* dilations_host, int ceil_mode, float* out); /* conv2d.rs:1051 */
/* mode 0 = asymmetric, 1 = half-pixel + round (conv2d.rs:1261) */
int lele_b200_resize_nearest(lele_b200_ctx* ctx, const float* x, int nb, int c, int h, int w, int oh,
                             int ow, int mode, float* out);

/* ---- rnn.rs (batch_size 1 per sequence; n_seq independent sequences run concurrently) ---- */
int lele_b200_lstm(lele_b200_ctx* ctx, const float* x, const float* w, const float* r,
                   const float* bias, const float* h0, const float* c0, int n_seq, int seq,
                   int in_size, int hidden, float* y, float* h, float* c);      /* rnn.rs:67 */
int lele_b200_gru(lele_b200_ctx* ctx, const float* x, const float* w, const float* r,
                  const float* bias, const float* h0, int n_seq, int seq, int in_size, int hidden,
                  float* y, float* h);                                          /* rnn.rs:246 */

/* ---- math.rs element-wise (NumPy broadcasting, utils.rs:107) ---- */
enum { LELE_B200_ADD = 0, LELE_B200_SUB, LELE_B200_MUL, LELE_B200_DIV, LELE_B200_MAX, LELE_B200_POW,
       LELE_B200_MOD, LELE_B200_PRELU, LELE_B200_EQUAL, LELE_B200_LESS };
/* shapes are right-aligned and broadcast; rank <= 8; out has the broadcast shape. */
int lele_b200_binary(lele_b200_ctx* ctx, int op, const float* a, const long long* a_shape_host,
                     int a_rank, const float* b, const long long* b_shape_host, int b_rank,
                     float* out);
enum { LELE_B200_RELU = 0, LELE_B200_SIGMOID, LELE_B200_TANH, LELE_B200_SILU, LELE_B200_ERF,
       LELE_B200_GELU, LELE_B200_EXP, LELE_B200_SOFTPLUS, LELE_B200_LOG, LELE_B200_SQRT,
       LELE_B200_NEG, LELE_B200_RECIPROCAL, LELE_B200_SIN, LELE_B200_COS, LELE_B200_NOT,
       LELE_B200_FAST_GELU };
int lele_b200_unary(lele_b200_ctx* ctx, int op, const float* x, long long len, float* out);
int lele_b200_clip(lele_b200_ctx* ctx, const float* x, long long len, float lo, float hi, float* out);
/* kind: 0 sum, 1 mean, 2 max, 3 l2 (math.rs:1527-1921); reduces `axis_len` with `inner` trailing */
int lele_b200_reduce(lele_b200_ctx* ctx, int kind, const float* x, long long outer, int axis_len,
                     long long inner, float* out);
int lele_b200_where(lele_b200_ctx* ctx, const float* cond, const long long* c_shape_host, int c_rank,
                    const float* x, const long long* x_shape_host, int x_rank, const float* y,
                    const long long* y_shape_host, int y_rank, float* out);     /* manipulation.rs:1215 */

/* ---- manipulation.rs / shape.rs (bit-exact copies) ---- */
/* generic strided gather: out[i] = in[offset + sum_d coord_d(i) * in_stride_d]; covers
 * transpose (manipulation.rs:644), slice (:209), expand (math.rs:2168), split (:1091). */
int lele_b200_strided_copy(lele_b200_ctx* ctx, const float* in, long long in_offset,
                           const long long* out_shape_host, const long long* in_strides_host,
                           int rank, float* out);
int lele_b200_concat(lele_b200_ctx* ctx, const float* const* inputs_host_array,
                     const long long* axis_lens_host, int n_inputs, long long outer,
                     long long inner, float* out);                              /* manipulation.rs:108 */
/* mode 0 constant, 1 edge, 2 reflect; pads_host [begin_0..begin_r-1, end_0..end_r-1] */
int lele_b200_pad(lele_b200_ctx* ctx, const float* in, const long long* shape_host, int rank,
                  const long long* pads_host, int mode, float value, float* out); /* manipulation.rs:382 */
/* indices are f32 or i64-as-f32 values (AsI64); negative wrap (manipulation.rs:589) */
int lele_b200_gather(lele_b200_ctx* ctx, const float* data, long long outer, int axis_dim,
                     long long inner, const float* indices, long long n_indices, float* out);
int lele_b200_gather_elements(lele_b200_ctx* ctx, const float* data, const float* indices,
                              long long outer, int axis_dim, int idx_dim, long long inner,
                              float* out);                                      /* conv2d.rs:1438 */
int lele_b200_tile(lele_b200_ctx* ctx, const float* in, const long long* shape_host,
                   const long long* repeats_host, int rank, float* out);        /* math.rs:2249 */
/* last axis, stable, indices returned as f32 (conv2d.rs:1385) */
int lele_b200_topk(lele_b200_ctx* ctx, const float* x, long long outer, int n, int k, float* values,
                   float* indices);
/* argmax over the last axis; ties -> LAST index (Iterator::max_by, sensevoice tokenizer.rs:55) */
int lele_b200_argmax_last(lele_b200_ctx* ctx, const float* x, long long outer, int n, int32_t* out);
/* greedy decode filter (examples/sensevoice/src/tokenizer.rs:37-75): per clip keep, in frame order,
 * the ids that are neither blank (0) nor flagged in skip_mask[vocab] (the "<|...|>" special tokens; NULL = none);
 * out_ids [n_clips, t] is padded with -1, out_len [n_clips] holds the kept counts.  All pointers are device pointers. */
int lele_b200_greedy_filter(lele_b200_ctx* ctx, const int32_t* ids, int n_clips, int t, const uint8_t* skip_mask,
                            int vocab, int32_t* out_ids, int32_t* out_len);

/* ---- SenseVoice-shaped graph runner: the batched replay of the model.rs call sequence
 *      (examples/sensevoice/src/main.rs:73-140) with the weights blob resident in HBM ---- */
int lele_b200_sensevoice_create(lele_b200_ctx* ctx, const uint8_t* blob_dev, size_t nbytes,
                                const uint8_t* blob_header_host /*first 256 B + table*/,
                                size_t header_bytes, int max_clips, int max_samples,
                                lele_b200_sensevoice** out);
int lele_b200_sensevoice_destroy(lele_b200_ctx* ctx, lele_b200_sensevoice* m);
int lele_b200_sensevoice_rows(const lele_b200_sensevoice* m, int n_samples);    /* T' = T_lfr + 4 */
int lele_b200_sensevoice_vocab(const lele_b200_sensevoice* m);
/* device-resident inputs: pcm_dev [n_clips, n_samples] -> ids_dev [n_clips, T'] (i32);
 * logits_dev_opt NULL or [n_clips, T', vocab]; n_layers_limit < 0 = all (tests use a prefix,
 * in which case logits_dev_opt receives the hidden state [n_clips, T', d]). */
int lele_b200_sensevoice_forward(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_dev,
                                 int n_clips, int n_samples, int lang, int textnorm,
                                 int n_layers_limit, int32_t* ids_dev, float* logits_dev_opt);
/* same from CMVN'd features [n_clips, t, 560] (model.forward(speech,..), main.rs:140) */
int lele_b200_sensevoice_forward_features(lele_b200_ctx* ctx, lele_b200_sensevoice* m,
                                          const float* feats_dev, int n_clips, int t, int lang,
                                          int textnorm, int n_layers_limit, int32_t* ids_dev,
                                          float* logits_dev_opt);
/* host-buffer entry (what the reference-facing plugin call looks like): pinned or pageable
 * host PCM in, host ids out; H2D + forward + D2H on the context stream, then sync. */
int lele_b200_sensevoice_transcribe_host(lele_b200_ctx* ctx, lele_b200_sensevoice* m,
                                         const float* pcm_host, int n_clips, int n_samples,
                                         int lang, int textnorm, int32_t* ids_host);
/* Pipelined form of the above (serving loop): submit batch i into slot i % 2, collect it with transcribe_wait(slot).
 * The H2D copy of the next batch and the D2H of the previous ids run on their own streams, so copies overlap the
 * forward; pcm_host / ids_host (pinned) must stay valid until the matching wait.  A slot cannot be re-submitted
 * before it was waited for. */
int lele_b200_sensevoice_transcribe_host_async(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_host,
                                               int n_clips, int n_samples, int lang, int textnorm,
                                               int32_t* ids_host, int slot);
int lele_b200_sensevoice_transcribe_wait(lele_b200_ctx* ctx, lele_b200_sensevoice* m, int slot);
/* Multi-GPU serving (one process per GPU, clips sharded): with a communicator attached, transcribe_host_async gathers the ids of all
 * ranks on the device (lele_b200_comm_gather out of the slot's id buffer) and only `root` copies them to ITS ids_host
 * [world][n_clips][T'] (rank order) -- one device-to-host copy per batch for the whole job; other ranks pass ids_host = NULL.
 * Every rank must submit the same n_clips / n_samples.  comm = NULL detaches. */
int lele_b200_sensevoice_set_comm(lele_b200_sensevoice* m, lele_b200_comm* comm, int root);
/* per-kernel-class device time of the last forward, ms (the analogue of kernels/timing.rs
 * print()): names_host receives up to `cap` const char*, ms_host the summed device time of
 * the class and calls_host its launch-group count; *n_out = classes written.  Only filled when
 * lele_b200_sensevoice_set_profiling(m, 1) was called before the forward (events are recorded
 * around every launch; not for timed runs). */
int lele_b200_sensevoice_set_profiling(lele_b200_sensevoice* m, int enable);
/* Borrow of a workspace buffer after a forward (the analogue of forward_with_workspace returning
 * borrows of ws.buf_N, src/compiler/mod.rs:1269-1351).  name: "lfr","feats","x0","x","h","qkv",
 * "fsmn","att","f1","keys".  *dptr is a device pointer valid until the next forward. */
int lele_b200_sensevoice_workspace(lele_b200_sensevoice* m, const char* name, void** dptr,
                                   size_t* nbytes);
int lele_b200_sensevoice_last_profile(lele_b200_sensevoice* m, const char** names_host,
                                      float* ms_host, int* calls_host, int cap, int* n_out);

/* ---- multi-GPU: clips shard across one process per GPU; the only collectives of the path are one broadcast of the weights
 *      blob and one gather of the greedy ids per batch, over NCCL / NVLink.  NCCL is bound at run time (dlopen; LELE_B200_NCCL_LIB
 *      overrides the library path); a single-GPU host never needs it.  Rendez-vous: rank 0 creates the 128-byte id and the host hands it
 *      to the other ranks by its own means.  Every call runs on the context stream. ---- */
int lele_b200_comm_nccl_version(void);                                        /* 0 when NCCL cannot be loaded */
int lele_b200_comm_unique_id(void* id128_host);
int lele_b200_comm_create(lele_b200_ctx* ctx, const void* id128_host, int world, int rank, lele_b200_comm** out);
int lele_b200_comm_destroy(lele_b200_comm* comm);
int lele_b200_comm_rank(const lele_b200_comm* comm);
int lele_b200_comm_world(const lele_b200_comm* comm);
/* in place: dptr [nbytes] on every rank receives root's bytes (ncclBroadcast) */
int lele_b200_comm_broadcast(lele_b200_ctx* ctx, lele_b200_comm* comm, void* dptr, size_t nbytes, int root);
/* root's recv_dev [world][nbytes] receives every rank's send_dev [nbytes] in rank order (grouped ncclSend / ncclRecv) */
int lele_b200_comm_gather(lele_b200_ctx* ctx, lele_b200_comm* comm, const void* send_dev, void* recv_dev, size_t nbytes, int root);

#ifdef __cplusplus
}
#endif
#endif /* LELE_B200_H */
