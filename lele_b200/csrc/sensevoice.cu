// sensevoice.cu -- the graph runner for the SenseVoiceSmall-shaped network: the batched replay
// of what a lele_gen-compiled model.rs does per clip (examples/sensevoice/src/main.rs:73-140;
// layer composition src/bin/wasm_bench.rs:888-1113), with the weights blob resident in HBM
// (src/compiler/mod.rs:1082 keeps a borrowed &[u8]; here blob_dev) and the workspace arena
// (src/compiler/mod.rs:1057-1070) pre-sized for max_clips.
//
// The batch lives below the operator boundary: B clips are stacked along the row (M) dimension
// of every GEMM, while everything the reference computes "per tensor" stays per clip
// (dynamic-quantisation min/max/scale/zero-point, CMVN) -- SURVEY.md 7.2.
//
// Host-side C++ (the analogue of the generated straight-line run_chunk_k body); all math is in
// the CUDA kernels of this directory.
#include "gemm_i8_tc.cuh"
#include <math.h>
#include <stdlib.h>

// ---- internal entry points from the other translation units ----
bool lb_layer_norm_quantize_supported(int n, int rows_per_slice);
bool lb_layer_norm_quantize_cluster_supported(int n, int T);
bool lb_layer_norm_quantize_stream_supported(lele_b200_ctx* ctx, int n, int T);
int lb_layer_norm_quantize_stream(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, int clips, int T, float eps,
                                  uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp, void* records, int rec_stride);
int lb_layer_norm_quantize_stream_records(int T);
int lb_layer_norm_quantize_cluster(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, int clips, int T, float eps,
                                   uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp, unsigned* keys_out);
int lb_layer_norm_stats(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer, int n, float eps,
                        float* stats, unsigned* minmax_keys, int rows_per_slice);
int lb_layer_norm_quantize(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer, int n,
                           const float* stats, const unsigned* keys, int rows_per_slice, uint8_t* a_u8, int32_t* rowsum,
                           float* row_scale, int32_t* row_zp);
int lb_layer_norm_minmax(lele_b200_ctx* ctx, const float* x, const float* gamma, const float* beta, long long outer, int n,
                         float eps, float* out, unsigned* minmax_keys, int rows_per_slice);
int lb_sgemm_strided_ldc(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                         long long rsb, long long csb, long long bsb, float* C, long long bsc, long long ldc, int batch, int m,
                         int k, int n, float alpha);
int lb_argmax_keys_to_ids(lele_b200_ctx* ctx, const unsigned long long* keys, long long n, int32_t* out);
bool lb_attention_tc_supported(int T, int d, int H);
size_t lb_attention_tc_scratch_bytes(int B, int T, int d, int H);
int lb_attention_tc(lele_b200_ctx* ctx, const float* qkv, int B, int T, int d, int H, float qscale, void* scratch, float* att,
                    unsigned* minmax_keys, int operands_ready);
void lb_attention_tc_operands(void* scratch, int B, int T, int d, int H, float** vt, int* tp);

namespace {
enum { SV_G_EMBED = 0, SV_G_POS, SV_G_AFTER_G, SV_G_AFTER_B, SV_G_TP_G, SV_G_TP_B, SV_G_CTC_W, SV_G_CTC_SCALE, SV_G_CTC_BIAS,
       SV_G_CTC_ZP, SV_NUM_GLOBAL };
enum { SV_L_LN1_G = 0, SV_L_LN1_B, SV_L_QKV_W, SV_L_QKV_SCALE, SV_L_QKV_BIAS, SV_L_QKV_ZP, SV_L_FSMN_W, SV_L_OUT_W, SV_L_OUT_SCALE,
       SV_L_OUT_BIAS, SV_L_OUT_ZP, SV_L_LN2_G, SV_L_LN2_B, SV_L_FFN1_W, SV_L_FFN1_SCALE, SV_L_FFN1_BIAS, SV_L_FFN1_ZP, SV_L_FFN2_W,
       SV_L_FFN2_SCALE, SV_L_FFN2_BIAS, SV_L_FFN2_ZP, SV_NUM_LAYER };

enum ProfClass { P_FRONTEND = 0, P_CMVN, P_PREP, P_LAYERNORM, P_QUANTIZE, P_GEMM_I8, P_FSMN, P_ATTN_QK, P_SOFTMAX, P_ATTN_PV,
                 P_MINMAX, P_ARGMAX, P_MISC, P_ATTN_TC,
                 // per-site split of P_GEMM_I8 (also accumulated into it)
                 P_G_QKV, P_G_OUT, P_G_FFN1_MAX, P_G_FFN1, P_G_FFN2, P_G_CTC, P_NUM };
const char* kProfNames[P_NUM] = {"frontend_fbank_lfr", "cmvn", "prompt_scale_pos", "layer_norm", "quantize_rows", "gemm_i8_tcgen05",
                                 "fsmn_dwconv", "attn_qk_sgemm", "softmax", "attn_pv_sgemm", "slice_minmax", "argmax", "misc",
                                 "attn_tcgen05_tf32x3",
                                 "gemm_i8:qkv", "gemm_i8:out_proj", "gemm_i8:ffn1_max_pass", "gemm_i8:ffn1", "gemm_i8:ffn2", "gemm_i8:ctc"};

// x0[b, r, :] = (r < 4 ? embed[id_r] : feats[b, r-4]) * sqrt(d) + pos[r]
__global__ void __launch_bounds__(128)
prompt_scale_pos_kernel(const float* __restrict__ feats, const float* __restrict__ embed, const float* __restrict__ pos,
                        int4 ids, int n_clips, int t, int din, float sq, float* __restrict__ x0) {
    // one CTA per output row: the row's source (a prompt embedding or a feature row) is decided once, no per-element index divisions
    const int T = t + 4;
    const int row = blockIdx.x, b = row / T, r = row - b * T;
    const float* src;
    if (r < 4) { const int id = r == 0 ? ids.x : (r == 1 ? ids.y : (r == 2 ? ids.z : ids.w)); src = embed + (long long)id * din; }
    else src = feats + ((long long)b * t + (r - 4)) * din;
    const float* pr = pos + (long long)r * din;
    float* dst = x0 + (long long)row * din;
    for (int c = threadIdx.x; c < din; c += 128) dst[c] = __fadd_rn(__fmul_rn(__ldg(src + c), sq), __ldg(pr + c));
}

// FSMN memory block: depthwise conv1d over time (k taps, zero pad, no bias) on V + V.
// v lives inside qkv [M, 3d] at column offset 2d.  Same tap order / unfused mul+add as the
// reference conv1d then `add`.
__global__ void __launch_bounds__(256)
fsmn_kernel(const float* __restrict__ qkv, const float* __restrict__ w /*[d,1,k]*/, int n_clips, int T, int d, int k,
            float* __restrict__ out /*[M,d]*/) {
    const int pad = (k - 1) / 2;
    const long long total = (long long)n_clips * T * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % d), tt = (int)((i / d) % T), b = (int)(i / ((long long)d * T));
        const float* vb = qkv + ((long long)b * T) * 3 * d + 2 * d + c;
        float s = 0.0f;
        for (int kk = 0; kk < k; ++kk) {
            int pos = tt + kk - pad;
            if (pos >= 0 && pos < T) s = __fadd_rn(s, __fmul_rn(__ldg(w + (long long)c * k + kk), vb[(long long)pos * 3 * d]));
        }
        out[i] = __fadd_rn(s, vb[(long long)tt * 3 * d]);
    }
}

// Fast path: one thread per channel walks FS_TCH consecutive time steps with the K-tap window in
// registers (one strided-but-coalesced load per output instead of K), no index divisions.
// grid (d/128, ceil(T/FS_TCH), B).  Same tap order and unfused mul+add as fsmn_kernel.
constexpr int FS_TCH = 16;
template <int K>
__global__ void __launch_bounds__(128)
fsmn_window_kernel(const float* __restrict__ qkv, const float* __restrict__ w, int T, int d, float* __restrict__ out) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= d) return;
    constexpr int PAD = (K - 1) / 2;
    const int t0 = blockIdx.y * FS_TCH, b = blockIdx.z;
    const float* vb = qkv + ((long long)b * T) * 3 * d + 2 * d + c;
    float wk[K], win[FS_TCH + K - 1];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) wk[kk] = __ldg(w + (long long)c * K + kk);
#pragma unroll
    for (int i = 0; i < FS_TCH + K - 1; ++i) {
        int pos = t0 + i - PAD;
        win[i] = (pos >= 0 && pos < T) ? vb[(long long)pos * 3 * d] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < FS_TCH; ++i) {
        const int tt = t0 + i;
        if (tt < T) {
            float s = 0.0f;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                int pos = tt + kk - PAD;
                if (pos >= 0 && pos < T) s = __fadd_rn(s, __fmul_rn(wk[kk], win[i + kk]));
            }
            out[((long long)b * T + tt) * d + c] = __fadd_rn(s, win[i + PAD]);
        }
    }
}

// The same block fed from V^T [B][H][128][Tp] (keys contiguous), the only copy of v the product path keeps: the QKV projection's
// epilogue writes v transposed for the attention kernel and skips the row-major v columns (36 MB per layer less HBM traffic at
// B = 64).  One CTA = (clip, 32 channels): the channel rows are staged in shared memory (coalesced along time), each warp then
// walks time steps with lane = channel, so the [M, d] output rows are written 128 B at a time.  Tap order and the unfused
// mul + add are those of fsmn_kernel (bit-identical results).
template <int K>
__global__ void __launch_bounds__(256)
fsmn_vt_kernel(const float* __restrict__ vt, const float* __restrict__ w, int T, int Tp, int d, float* __restrict__ out) {
    extern __shared__ float fs_tile[];                 // [32][Tp + 1]
    constexpr int PAD = (K - 1) / 2;
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = Tp + 1;
    const float* src = vt + ((size_t)b * d + c0) * (size_t)Tp;     // channel c = h * 128 + dk -> row (b * d + c) of V^T
    for (int i = threadIdx.x; i < 32 * Tp; i += 256) { const int ch = i / Tp, t = i - ch * Tp; fs_tile[ch * pitch + t] = src[i]; }
    float wk[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) wk[kk] = __ldg(w + (size_t)(c0 + lane) * K + kk);
    __syncthreads();
    const float* row = fs_tile + lane * pitch;
    for (int tt = warp; tt < T; tt += 8) {
        float s = 0.0f;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const int pos = tt + kk - PAD;
            if (pos >= 0 && pos < T) s = __fadd_rn(s, __fmul_rn(wk[kk], row[pos]));
        }
        out[((size_t)b * T + tt) * d + c0 + lane] = __fadd_rn(s, row[tt]);
    }
}

// v2 of the block above: division-free staging (a warp copies whole channel rows), and each warp walks ONE contiguous time segment with
// the K-tap window in registers (one shared-memory read per output instead of K; the taps that fall outside [0, T) read 0.0f: adding
// w * 0 = +-0 to a sum that started at +0 never changes it, so the result equals the tap-skipping loop bit for bit).
template <int K>
__global__ void __launch_bounds__(256)
fsmn_vt_win_kernel(const float* __restrict__ vt, const float* __restrict__ w, int T, int Tp, int d, float* __restrict__ out) {
    extern __shared__ float fs_tile[];                 // [32][Tp + 1]
    constexpr int PAD = (K - 1) / 2;
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = Tp + 1;
    const float* src = vt + ((size_t)b * d + c0) * (size_t)Tp;
    for (int ch = warp; ch < 32; ch += 8)
        for (int t = lane; t < Tp; t += 32) fs_tile[ch * pitch + t] = __ldg(src + (size_t)ch * Tp + t);
    float wk[K];
#pragma unroll
    for (int kk = 0; kk < K; ++kk) wk[kk] = __ldg(w + (size_t)(c0 + lane) * K + kk);
    __syncthreads();
    const float* row = fs_tile + lane * pitch;
    const int seg = (T + 7) >> 3;
    const int t0 = warp * seg, t1 = min(t0 + seg, T);
    if (t0 >= t1) return;
    float win[K];                                      // win[(tt + kk) % K] = v[tt + kk - PAD] while step tt is computed
#pragma unroll
    for (int kk = 0; kk < K; ++kk) { const int pos = t0 + kk - PAD; win[kk] = (pos >= 0 && pos < T) ? row[pos] : 0.0f; }
    float* op = out + ((size_t)b * T + t0) * d + c0 + lane;
    for (int base = t0; base < t1; base += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {                  // compile-time rotation of the window registers
            const int tt = base + j;
            if (tt < t1) {
                float s = 0.0f;
#pragma unroll
                for (int kk = 0; kk < K; ++kk) s = __fadd_rn(s, __fmul_rn(wk[kk], win[(j + kk) % K]));
                *op = __fadd_rn(s, win[(j + PAD) % K]);
                op += d;
                const int nx = tt + K - PAD;           // the element that enters the window for step tt + 1
                win[j % K] = nx < T ? row[nx] : 0.0f;
            }
        }
    }
}

__global__ void scale_copy_q_kernel(const float* __restrict__ qkv, long long M, int d, float qscale, float* __restrict__ q) {
    const long long total = M * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / d; int c = (int)(i % d);
        q[i] = __fmul_rn(qkv[r * 3 * d + c], qscale);
    }
}

// softmax over score rows (same arithmetic as norm.cu softmax_kernel), in place
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ s, long long rows, int n) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* xr = s + row * n;
    const int simd_end = (n / 8) * 8;
    float mx = -3.402823466e+38f;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, xr[j]);
    mx = lb_warp_max(mx);
    for (int j = lane; j < n; j += 32) {
        float dlt = __fsub_rn(xr[j], mx);
        xr[j] = j < simd_end ? lb_cephes_expf(dlt) : lb_libm_expf(dlt);
    }
    __syncwarp();
    const float sum = lb_avx_order_reduce(n, lane, [&](float acc, int j) { return __fadd_rn(acc, xr[j]); },
                                          [&](float acc, int j) { return __fadd_rn(acc, xr[j]); });
    const float inv = __fdiv_rn(1.0f, sum);
    for (int j = lane; j < n; j += 32) xr[j] = __fmul_rn(xr[j], inv);
}

__global__ void init_u64_kernel(unsigned long long* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0ull;
}
int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }
}  // namespace

struct lele_b200_sensevoice {
    int n_layers, d, d_in, ffn, heads, fsmn_k, vocab, n_embed, max_t, n_stage1, n_tensors;
    const uint8_t* blob = nullptr;
    std::vector<unsigned long long> table;       // (offset, nbytes) pairs
    std::vector<lele_b200_qweights*> lin;         // 4 per layer + ctc
    int max_clips = 0, max_samples = 0, max_T = 0;
    // workspace arena
    float *lfr = nullptr, *feats = nullptr, *x0 = nullptr, *x = nullptr, *h = nullptr, *qkv = nullptr, *qs = nullptr, *fsmn = nullptr,
          *att = nullptr, *f1 = nullptr, *scores = nullptr, *pcm_stage = nullptr;
    int32_t* ids_stage = nullptr;
    unsigned* keys = nullptr;
    unsigned long long* amax_keys = nullptr;
    int* fq_counters = nullptr;       // [sites][clips] arrival counters of the kernels whose CTAs meet through global memory (one-pass FFN1,
                                      // streaming LayerNorm + quantiser), zeroed per forward; same site index as `keys`
    void* ln_records = nullptr;       // [sites][clips][rec_stride] 16-byte {min key, 1, max key, 1} records of the streaming LayerNorm + quantiser, zeroed per forward
    int rec_stride = 0; size_t rec_bytes = 0;
    int lane_split = 0;               // clip lanes size their persistent grids to #SMs / lanes (each lane owns a share of the device): LELE_B200_LANE_SPLIT=0 -> full grids
    int lane_sms = 0;                 // (views) the SM share of this lane, 0 = the whole device
    int ln_stream = 0;                // LayerNorm + quantiser as the streaming persistent kernel (LELE_B200_LNQ_STREAM=0 -> cluster kernel)
    void* qscratch = nullptr;
    void* qscratch2 = nullptr;        // second quantised-operand set: FFN1's fused output quantiser writes it while reading the first
    void* attn_scratch = nullptr;
    int fsmn_v2 = 1;                  // FSMN block: register-window kernel (LELE_B200_FSMN_V2=0 -> the tap-gathering kernel)
    int fsmn_fork = 1;                // FSMN block on the side stream (LELE_B200_FSMN_FORK=0: in line, before attention)
    int ffn_onepass = 1;              // FFN1 in ONE pass (dequantised tile parked in TMEM until the clip's max is known); LELE_B200_FFN_FUSED=0 -> two passes
    int ffn_twopass = 1;              // FFN1 as max-only pass + quantising pass (no f32 [M, ffn] round trip); LELE_B200_FFN_TWOPASS=0 disables
    int fuse_lnq = 1;                 // LayerNorm + quantiser fused for the encoder width (LELE_B200_FUSE_LNQ=0 disables)
    int attn_simt = 0;   // LELE_B200_ATTN_SIMT=1: CUDA-core attention (cross-check of the tcgen05 path)
    int qkv_full = 0;    // LELE_B200_QKV_FULL=1: the QKV projection also writes the row-major v columns (default: v exists only as V^T)
    // profiling
    int profiling = 0;
    // CUDA-graph replay of the whole forward (the ~1000 launches of one step are CPU-bound otherwise)
    struct GraphKey { const float* pcm; int B, n_samples, lang, textnorm, n_layers; int32_t* ids; float* logits; };
    int use_graph = 1;              // LELE_B200_GRAPH=0 disables
    // capture policy: library-owned staging buffers (pcm_stage, pcm_slot[]) are captured on first use; a caller-owned buffer set is
    // captured the SECOND time the same (pointers, geometry) key is seen, so a host that passes fresh device buffers on every call
    // (the numpy-form wrappers) never pays capture + instantiate and never evicts the graphs the serving entries replay
    GraphKey seen_key[4] = {};
    int seen_next = 0;
    bool warmed = false;            // one eager forward has run (lazy tables / attributes / scratch are in place)
    // small cache of captured forwards (the pipelined host entry alternates between two staging buffers)
    static constexpr int N_GRAPHS = 4;
    cudaGraphExec_t graph_exec[N_GRAPHS] = {nullptr, nullptr, nullptr, nullptr};
    GraphKey graph_key[N_GRAPHS] = {};
    unsigned long long graph_launches[N_GRAPHS] = {0, 0, 0, 0};
    unsigned long long graph_used[N_GRAPHS] = {0, 0, 0, 0}, graph_clock = 0;   // least-recently-used replacement
    // pipelined host entry (transcribe_host_async): 2 slots of staging, H2D / D2H on their own streams so the copy of
    // batch i+1 overlaps the forward of batch i
    static constexpr int N_SLOTS = 2;
    float* pcm_slot[N_SLOTS] = {nullptr, nullptr};
    int32_t* ids_slot[N_SLOTS] = {nullptr, nullptr};
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    // multi-GPU: with a communicator attached the pipelined host entry gathers every rank's ids on the device
    // (lele_b200_comm_gather straight out of ids_slot) and only the root copies them to its host buffer -- one D2H per batch
    lele_b200_comm* comm = nullptr;
    int comm_root = 0;
    int32_t* ids_gather[2] = {nullptr, nullptr};   // root: [world][max_clips * max_T]
    cudaEvent_t ev_h2d[N_SLOTS] = {nullptr, nullptr}, ev_fwd[N_SLOTS] = {nullptr, nullptr}, ev_d2h[N_SLOTS] = {nullptr, nullptr};
    bool slot_busy[N_SLOTS] = {false, false};
    // side stream: the HBM-bound FSMN block runs concurrently with the latency-bound attention (both only read qkv)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // clip lanes: the batch is split into n_lanes groups of clips that run the encoder on their own streams (own
    // workspace slices, own min/max keys).  Every kernel of the layer body is a short (20-120 us) launch separated
    // from its successor by a drain / launch / pipeline-ramp gap; with two independent chains in flight one lane's
    // CTAs fill the SMs while the other lane's kernel drains.  Clips are independent end to end, so this is the same
    // computation (bit-identical per clip).  Measured: +1.5 % with 2 lanes before the attention kernel became persistent, -0.6 % after
    // (both kernels of a pair now hold every SM for their whole life), so the default is 1; LELE_B200_LANES=2..4 enables.
    static constexpr int MAX_LANES = 4;
    int n_lanes = 1;
    lele_b200_sensevoice* lane[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};   // shallow views (pointers offset per call)
    lele_b200_ctx* lane_ctx[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};      // [0] unused (lane 0 runs on the caller's ctx)
    cudaEvent_t ev_lane_fork = nullptr, ev_lane_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    bool is_view = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    float prof_ms[P_NUM] = {0};
    int prof_calls[P_NUM] = {0};

    const void* tensor(int idx) const { return blob + table[2 * idx]; }
    const void* lt(int layer, int which) const { return tensor(SV_NUM_GLOBAL + layer * SV_NUM_LAYER + which); }
};

namespace {
struct ProfScope {
    lele_b200_sensevoice* m; lele_b200_ctx* ctx; int cls; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(lele_b200_sensevoice* m_, lele_b200_ctx* c_, int cls_) : m(m_), ctx(c_), cls(cls_) {
        if (!m->profiling) return;
        auto next = [&]() { if (m->ev_used == m->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); m->ev_pool.push_back(e); } return m->ev_pool[m->ev_used++]; };
        a = next(); b = next();
        cudaEventRecord(a, ctx->stream);
    }
    ~ProfScope() { if (m->profiling) { cudaEventRecord(b, ctx->stream); m->spans.push_back({cls, a, b}); } }
};
#define SV_RUN(cls, expr) do { ProfScope _ps(m, ctx, cls); int _rc = (expr); if (_rc) return _rc; } while (0)

#define SV_LINEAR(...) do { int _rc = sv_linear(__VA_ARGS__); if (_rc) return _rc; } while (0)

// the quantised GEMM; an epilogue that carries fq_keys asks for the one-pass linear -> ReLU -> quantiser kernel
static int sv_gemm(lele_b200_ctx* ctx, const LbQuantScratch& qs, const lele_b200_qweights* w, long long M, const LbI8Epilogue& ep) {
    if (ep.fq_keys) return lb_gemm_i8_tc_fused_q(ctx, qs.a_u8, w->wt, (int)M, w->n, w->k, ep);
    return lb_gemm_i8(ctx, qs.a_u8, w->wt, (int)M, w->n, w->k, ep);
}

// quantise the activation rows (per-clip keys already hold min/max) then the tcgen05 GEMM;
// profiled as two classes so the GEMM's own duration feeds the roofline.
int sv_linear(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* x, const unsigned* keys, long long M, int T,
              const lele_b200_qweights* w, const LbQuantScratch& qs, LbI8Epilogue ep, int gcls) {
    SV_RUN(P_QUANTIZE, lb_quantize_rows(ctx, x, keys, M, T, w->k, qs.a_u8, qs.rowsum, qs.row_scale, qs.row_zp));
    lb_fill_weight_fields(ep, w, qs);
    SV_RUN(gcls, sv_gemm(ctx, qs, w, M, ep));
    return LELE_B200_OK;
}

// LayerNorm -> dynamic quantiser -> tcgen05 GEMM.  For the encoder width (512) the normalised f32 rows are never
// materialised (norm.cu: statistics + min/max pass, then a re-deriving quantise pass); m->h holds the row statistics.
int sv_ln_linear(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* x, const float* gamma, const float* beta, int n, unsigned* keys,
                 long long M, int T, const lele_b200_qweights* w, const LbQuantScratch& qs, LbI8Epilogue ep, int gcls) {
    if (m->fuse_lnq == 1 && m->ln_stream && !m->is_view && m->ln_records && M % T == 0 && lb_layer_norm_quantize_stream_supported(ctx, n, T)) {
        // streaming persistent kernel: x read once, rows quantised one round after their clip's min / max met in global memory
        const size_t site_idx = (size_t)(keys - m->keys) / ((size_t)2 * LB_MM_SLOTS * (size_t)(M / T));
        SV_RUN(P_LAYERNORM, lb_layer_norm_quantize_stream(ctx, x, gamma, beta, (int)(M / T), T, 1e-5f, qs.a_u8, w->w_signed ? nullptr : qs.rowsum, qs.row_scale, qs.row_zp,
                                                          (uint8_t*)m->ln_records + (size_t)16 * m->rec_stride * (size_t)(M / T) * site_idx, m->rec_stride));
        lb_fill_weight_fields(ep, w, qs);
        SV_RUN(gcls, sv_gemm(ctx, qs, w, M, ep));
        return LELE_B200_OK;
    }
    if (m->fuse_lnq == 1 && lb_layer_norm_quantize_cluster_supported(n, T) && M % T == 0) {
        // one cluster per clip: x read once, normalised rows live in shared memory until the clip's min/max is known
        SV_RUN(P_LAYERNORM, lb_layer_norm_quantize_cluster(ctx, x, gamma, beta, (int)(M / T), T, 1e-5f, qs.a_u8, qs.rowsum, qs.row_scale, qs.row_zp, keys));
        lb_fill_weight_fields(ep, w, qs);
        SV_RUN(gcls, sv_gemm(ctx, qs, w, M, ep));
        return LELE_B200_OK;
    }
    if (m->fuse_lnq && lb_layer_norm_quantize_supported(n, T)) {
        SV_RUN(P_LAYERNORM, lb_layer_norm_stats(ctx, x, gamma, beta, M, n, 1e-5f, m->h, keys, T));
        SV_RUN(P_QUANTIZE, lb_layer_norm_quantize(ctx, x, gamma, beta, M, n, m->h, keys, T, qs.a_u8, qs.rowsum, qs.row_scale, qs.row_zp));
        lb_fill_weight_fields(ep, w, qs);
        SV_RUN(gcls, sv_gemm(ctx, qs, w, M, ep));
        return LELE_B200_OK;
    }
    SV_RUN(P_LAYERNORM, lb_layer_norm_minmax(ctx, x, gamma, beta, M, n, 1e-5f, m->h, keys, T));
    return sv_linear(ctx, m, m->h, keys, M, T, w, qs, ep, gcls);
}

int sv_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) { lb_set_error("sensevoice: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return LELE_B200_ERR_CUDA; }
    return LELE_B200_OK;
}
}  // namespace

extern "C" int lele_b200_sensevoice_create(lele_b200_ctx* ctx, const uint8_t* blob_dev, size_t nbytes, const uint8_t* hdr_host,
                                           size_t header_bytes, int max_clips, int max_samples, lele_b200_sensevoice** out) {
    LB_REQUIRE(ctx && blob_dev && hdr_host && out, "sensevoice_create: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(header_bytes >= 256, "sensevoice_create: header too short");
    const int32_t* hd = (const int32_t*)hdr_host;
    LB_REQUIRE(hd[0] == 0x454C454C && hd[1] == 1, "sensevoice_create: bad blob magic/version");
    lele_b200_sensevoice* m = new lele_b200_sensevoice();
    struct Guard { lele_b200_sensevoice* p; lele_b200_ctx* c; ~Guard() { if (p) lele_b200_sensevoice_destroy(c, p); } } guard{m, ctx};   // LB_REQUIRE returns below must not leak m
    m->n_layers = hd[2]; m->d = hd[3]; m->d_in = hd[4]; m->ffn = hd[5]; m->heads = hd[6]; m->fsmn_k = hd[7]; m->vocab = hd[8];
    m->n_embed = hd[9]; m->max_t = hd[10]; m->n_stage1 = hd[11]; m->n_tensors = hd[12];
    LB_REQUIRE(header_bytes >= 256 + 16 * (size_t)m->n_tensors, "sensevoice_create: header does not contain the tensor table");
    LB_REQUIRE(m->n_tensors == SV_NUM_GLOBAL + m->n_layers * SV_NUM_LAYER, "sensevoice_create: tensor count mismatch");
    LB_REQUIRE(m->d % m->heads == 0 && m->d % 4 == 0 && m->d_in % 4 == 0, "sensevoice_create: unsupported dims");
    m->blob = blob_dev;
    m->table.resize(2 * (size_t)m->n_tensors);
    memcpy(m->table.data(), hdr_host + 256, 16 * (size_t)m->n_tensors);
    for (int i = 0; i < m->n_tensors; ++i)
        LB_REQUIRE(m->table[2 * i] + m->table[2 * i + 1] <= nbytes, "sensevoice_create: tensor %d out of blob bounds", i);
    m->max_clips = max_clips; m->max_samples = max_samples;
    int frames = lele_b200_frontend_num_frames(max_samples);
    int t = (frames + 5) / 6;
    m->max_T = t + 4;
    LB_REQUIRE(m->max_T <= m->max_t, "sensevoice_create: %d rows exceed the positional table (%d)", m->max_T, m->max_t);

    // one-time weight preparation (the analogue of B_WEIGHT_CACHE / prepare_weights)
    auto prep = [&](const void* w, int k, int n, const void* sc, const void* zp_dev, const void* bias, lele_b200_qweights** q) -> int {
        uint8_t zp_h = 0;
        cudaError_t e = cudaMemcpyAsync(&zp_h, zp_dev, 1, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { lb_set_error("sensevoice_create: reading zero point: %s", cudaGetErrorString(e)); return LELE_B200_ERR_CUDA; }
        return lele_b200_prepare_weights(ctx, (const uint8_t*)w, k, n, (const float*)sc, n, zp_h, (const float*)bias, q);
    };
    m->lin.assign((size_t)m->n_layers * 4 + 1, nullptr);
    int rc = 0;
    for (int l = 0; l < m->n_layers && !rc; ++l) {
        int cur = l == 0 ? m->d_in : m->d;
        rc = prep(m->lt(l, SV_L_QKV_W), cur, 3 * m->d, m->lt(l, SV_L_QKV_SCALE), m->lt(l, SV_L_QKV_ZP), m->lt(l, SV_L_QKV_BIAS), &m->lin[l * 4 + 0]);
        if (!rc) rc = prep(m->lt(l, SV_L_OUT_W), m->d, m->d, m->lt(l, SV_L_OUT_SCALE), m->lt(l, SV_L_OUT_ZP), m->lt(l, SV_L_OUT_BIAS), &m->lin[l * 4 + 1]);
        if (!rc) rc = prep(m->lt(l, SV_L_FFN1_W), m->d, m->ffn, m->lt(l, SV_L_FFN1_SCALE), m->lt(l, SV_L_FFN1_ZP), m->lt(l, SV_L_FFN1_BIAS), &m->lin[l * 4 + 2]);
        if (!rc) rc = prep(m->lt(l, SV_L_FFN2_W), m->ffn, m->d, m->lt(l, SV_L_FFN2_SCALE), m->lt(l, SV_L_FFN2_ZP), m->lt(l, SV_L_FFN2_BIAS), &m->lin[l * 4 + 3]);
    }
    if (!rc) rc = prep(m->tensor(SV_G_CTC_W), m->d, m->vocab, m->tensor(SV_G_CTC_SCALE), m->tensor(SV_G_CTC_ZP), m->tensor(SV_G_CTC_BIAS), &m->lin[(size_t)m->n_layers * 4]);
    if (rc) return rc;   // (guard frees m)

    const size_t B = max_clips, M = B * m->max_T;
    const int wide = m->d_in > m->d ? m->d_in : m->d;
    const int kmax = m->ffn > wide ? m->ffn : wide;
    rc = sv_alloc((void**)&m->lfr, sizeof(float) * B * t * m->d_in);
    if (!rc) rc = sv_alloc((void**)&m->feats, sizeof(float) * B * t * m->d_in);
    if (!rc) rc = sv_alloc((void**)&m->x0, sizeof(float) * M * m->d_in);
    if (!rc) rc = sv_alloc((void**)&m->x, sizeof(float) * M * m->d);
    if (!rc) rc = sv_alloc((void**)&m->h, sizeof(float) * M * wide);
    if (!rc) rc = sv_alloc((void**)&m->qkv, sizeof(float) * M * 3 * m->d);
    if (!rc) rc = sv_alloc((void**)&m->qs, sizeof(float) * M * m->d);
    if (!rc) rc = sv_alloc((void**)&m->fsmn, sizeof(float) * M * m->d);
    if (!rc) rc = sv_alloc((void**)&m->att, sizeof(float) * M * m->d);
    if (!rc) rc = sv_alloc((void**)&m->f1, sizeof(float) * M * m->ffn);
    if (!rc) rc = sv_alloc((void**)&m->scores, sizeof(float) * B * m->heads * m->max_T * m->max_T);
    if (!rc) rc = sv_alloc((void**)&m->keys, sizeof(unsigned) * 2 * LB_MM_SLOTS * B * ((size_t)m->n_layers * 4 + 1));
    if (!rc) rc = sv_alloc((void**)&m->amax_keys, sizeof(unsigned long long) * M);
    if (!rc) rc = sv_alloc((void**)&m->fq_counters, sizeof(int) * ((size_t)m->n_layers * 4 + 1) * B);
    m->rec_stride = lb_layer_norm_quantize_stream_records(m->max_T);
    m->rec_bytes = (size_t)16 * m->rec_stride * B * ((size_t)m->n_layers * 4 + 1);
    if (!rc) rc = sv_alloc(&m->ln_records, m->rec_bytes);
    if (!rc) rc = sv_alloc(&m->qscratch, lb_quant_scratch_bytes((long long)M, kmax) + 4096 * lele_b200_sensevoice::MAX_LANES);   // + per-lane carve alignment
    if (!rc) rc = sv_alloc(&m->qscratch2, lb_quant_scratch_bytes((long long)M, kmax) + 4096 * lele_b200_sensevoice::MAX_LANES);
    { const char* e = getenv("LELE_B200_FFN_TWOPASS"); m->ffn_twopass = (e && e[0] == '0') ? 0 : 1; }
    m->ffn_onepass = lb_env_flag("LELE_B200_FFN_FUSED", 1) ? 1 : 0;
    m->fsmn_fork = lb_env_flag("LELE_B200_FSMN_FORK", 1) ? 1 : 0;
    m->fsmn_v2 = lb_env_flag("LELE_B200_FSMN_V2", 1) ? 1 : 0;
    m->ln_stream = lb_env_flag("LELE_B200_LNQ_STREAM", 0) ? 1 : 0;   // measured: 12 us in-kernel but 26 us in the replayed step (a 192 KB persistent CTA cannot start under its predecessor's tail; the 70 KB cluster CTAs can) -> opt-in
    m->lane_split = lb_env_flag("LELE_B200_LANE_SPLIT", 0) ? 1 : 0;     // measured (2 lanes x 74 SMs): 3.54 vs 3.22 ms on the 8-layer stack -> opt-in
    if (!rc) rc = sv_alloc(&m->attn_scratch, lb_attention_tc_scratch_bytes(max_clips, m->max_T, m->d, m->heads));
    { const char* e = getenv("LELE_B200_ATTN_SIMT"); m->attn_simt = (e && e[0] == '1') ? 1 : 0; }
    { const char* e = getenv("LELE_B200_GRAPH"); m->use_graph = (e && e[0] == '0') ? 0 : 1; }
    m->qkv_full = lb_env_flag("LELE_B200_QKV_FULL", 0) ? 1 : 0;
    { const char* e = getenv("LELE_B200_FUSE_LNQ"); m->fuse_lnq = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }   // 0 unfused, 1 cluster, 2 two-pass
    if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); m->side = nullptr; }
    if (!rc) rc = sv_alloc((void**)&m->pcm_stage, sizeof(float) * B * (size_t)max_samples);
    if (!rc) rc = sv_alloc((void**)&m->ids_stage, sizeof(int32_t) * M);
    if (rc) return rc;   // (guard frees m)
    {   // clip lanes (see the struct): views are shallow copies whose workspace pointers are re-based per call
        const char* e = getenv("LELE_B200_LANES");
        m->n_lanes = e ? atoi(e) : 1;
        if (m->n_lanes < 1) m->n_lanes = 1;
        if (m->n_lanes > lele_b200_sensevoice::MAX_LANES) m->n_lanes = lele_b200_sensevoice::MAX_LANES;
        if (m->n_lanes > 1) {
            bool ok = cudaEventCreateWithFlags(&m->ev_lane_fork, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; i < m->n_lanes && ok; ++i) {
                lele_b200_sensevoice* v = new lele_b200_sensevoice(*m);
                v->is_view = true; for (int g = 0; g < lele_b200_sensevoice::N_GRAPHS; ++g) v->graph_exec[g] = nullptr; v->n_lanes = 1;
                for (int j = 0; j < lele_b200_sensevoice::MAX_LANES; ++j) { v->lane[j] = nullptr; v->lane_ctx[j] = nullptr; v->ev_lane_join[j] = nullptr; }
                m->lane[i] = v;
                if (i == 0) continue;
                v->side = nullptr; v->ev_fork = nullptr; v->ev_join = nullptr;
                ok = lele_b200_ctx_create(ctx->device, nullptr, &m->lane_ctx[i]) == LELE_B200_OK &&
                     cudaEventCreateWithFlags(&m->ev_lane_join[i], cudaEventDisableTiming) == cudaSuccess;
                if (ok && (cudaStreamCreateWithFlags(&v->side, cudaStreamNonBlocking) != cudaSuccess ||
                           cudaEventCreateWithFlags(&v->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                           cudaEventCreateWithFlags(&v->ev_join, cudaEventDisableTiming) != cudaSuccess)) { cudaGetLastError(); v->side = nullptr; }
            }
            if (!ok) { cudaGetLastError(); m->n_lanes = 1; }
        }
    }
    guard.p = nullptr;
    *out = m;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_destroy(lele_b200_ctx* ctx, lele_b200_sensevoice* m) {
    if (!m) return LELE_B200_OK;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (auto* q : m->lin) lele_b200_qweights_destroy(nullptr, q);
    void* bufs[] = {m->lfr, m->feats, m->x0, m->x, m->h, m->qkv, m->qs, m->fsmn, m->att, m->f1, m->scores, m->keys, m->amax_keys, m->fq_counters, m->ln_records,
                    m->qscratch, m->qscratch2, m->pcm_stage, m->ids_stage, m->attn_scratch};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto e : m->ev_pool) cudaEventDestroy(e);
    for (int g = 0; g < lele_b200_sensevoice::N_GRAPHS; ++g) if (m->graph_exec[g]) cudaGraphExecDestroy(m->graph_exec[g]);
    for (int sl = 0; sl < lele_b200_sensevoice::N_SLOTS; ++sl) {
        if (m->pcm_slot[sl]) cudaFree(m->pcm_slot[sl]);
        if (m->ids_slot[sl]) cudaFree(m->ids_slot[sl]);
        if (m->ev_h2d[sl]) cudaEventDestroy(m->ev_h2d[sl]);
        if (m->ev_fwd[sl]) cudaEventDestroy(m->ev_fwd[sl]);
        if (m->ev_d2h[sl]) cudaEventDestroy(m->ev_d2h[sl]);
    }
    for (int sl = 0; sl < 2; ++sl) if (m->ids_gather[sl]) cudaFree(m->ids_gather[sl]);
    if (m->h2d_stream) cudaStreamDestroy(m->h2d_stream);
    if (m->d2h_stream) cudaStreamDestroy(m->d2h_stream);
    if (m->side) { cudaStreamSynchronize(m->side); cudaStreamDestroy(m->side); }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    for (int i = 0; i < lele_b200_sensevoice::MAX_LANES; ++i) {
        lele_b200_sensevoice* v = m->lane[i];
        if (v && i > 0) {
            if (v->side) { cudaStreamSynchronize(v->side); cudaStreamDestroy(v->side); }
            if (v->ev_fork) cudaEventDestroy(v->ev_fork);
            if (v->ev_join) cudaEventDestroy(v->ev_join);
        }
        delete v;                                        // shallow view: owns nothing else
        if (m->lane_ctx[i]) lele_b200_ctx_destroy(m->lane_ctx[i]);
        if (m->ev_lane_join[i]) cudaEventDestroy(m->ev_lane_join[i]);
    }
    if (m->ev_lane_fork) cudaEventDestroy(m->ev_lane_fork);
    delete m;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_rows(const lele_b200_sensevoice* m, int n_samples) {
    (void)m;
    int frames = lele_b200_frontend_num_frames(n_samples);
    return frames == 0 ? 0 : (frames + 5) / 6 + 4;
}
extern "C" int lele_b200_sensevoice_vocab(const lele_b200_sensevoice* m) { return m ? m->vocab : 0; }

static int sv_encoder(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* feats, int B, int t, int lang, int textnorm,
                      int n_layers_limit, int32_t* ids_dev, float* logits_opt) {
    const int d = m->d, din = m->d_in, H = m->heads, dk = d / H, T = t + 4, ffn = m->ffn;
    const long long M = (long long)B * T;
    const int n_layers = (n_layers_limit >= 0 && n_layers_limit < m->n_layers) ? n_layers_limit : m->n_layers;
    LB_REQUIRE(lang >= 0 && lang < m->n_embed && textnorm >= 0 && textnorm < m->n_embed, "sensevoice: prompt id out of range");
    const int n_sites = m->n_layers * 4 + 1;
    SV_RUN(P_MISC, lb_minmax_init(ctx, m->keys, n_sites * B));
    // one-pass FFN1 (accumulators wait in TMEM for the clip's max): its CTAs spin on each other, so it runs only when this forward owns the
    // device's SMs -- not inside a clip lane (lanes are concurrent forwards)
    // kernels whose CTAs wait for each other need every CTA resident: the forward owns the device, or it is a clip lane whose grids are
    // sized to its share of the SMs (then the concurrent lanes' persistent grids fit side by side)
    const bool meet_ok = (!m->is_view || m->lane_sms > 0) && m->fq_counters;
    const bool fq_allowed = meet_ok && m->ffn_onepass;
    if (meet_ok) LB_CHECK_CUDA(cudaMemsetAsync(m->fq_counters, 0, sizeof(int) * (size_t)n_sites * B, ctx->stream));
    if (meet_ok && m->ln_stream && m->ln_records) LB_CHECK_CUDA(cudaMemsetAsync(m->ln_records, 0, (size_t)16 * m->rec_stride * B * n_sites, ctx->stream));
    auto site = [&](int s) { return m->keys + (size_t)2 * LB_MM_SLOTS * B * s; };
    const LbQuantScratch qs = lb_quant_scratch_carve(m->qscratch, M, ffn > din ? ffn : din);
    const LbQuantScratch qs2 = lb_quant_scratch_carve(m->qscratch2, M, ffn > din ? ffn : din);

    {   // gather(embed, prompt ids) ++ concat ++ mul sqrt(d) ++ add pos
        ProfScope ps(m, ctx, P_PREP);
        prompt_scale_pos_kernel<<<(unsigned)M, 128, 0, ctx->stream>>>(feats, (const float*)m->tensor(SV_G_EMBED),
            (const float*)m->tensor(SV_G_POS), make_int4(lang, 1, 2, textnorm), B, t, din, sqrtf((float)d), m->x0);
        LB_LAUNCH_CHECK(ctx);
    }
    const float qscale = 1.0f / sqrtf((float)dk);
    const float* xin = m->x0;
    int cur = din;
    const bool attn_tc = !m->attn_simt && lb_attention_tc_supported(T, d, H);
    int attn_ops_ready = 0;
    bool fsmn_from_vt = false; const float* vt_ptr = nullptr; int vt_tp = 0;
    for (int l = 0; l < n_layers; ++l) {
        // ---- self-attention block ----
        {
            LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
            ep.out = m->qkv; ep.rows_per_slice = T;
            // the projection's epilogue also emits the V^T operand copy of the fused attention kernel
            if (attn_tc && m->lin[l * 4 + 0]->k % 16 == 0 && !getenv("LELE_B200_FORCE_SIMT") && !getenv("LELE_B200_GEMM_NO_TMA_STORE")) {
                lb_attention_tc_operands(m->attn_scratch, B, T, d, H, &ep.vt, &ep.vt_tp);
                attn_ops_ready = 1;
                ep.skip_v_out = (!m->qkv_full && m->fsmn_k == 11 && d % 32 == 0) ? 1 : 0;   // attention and the FSMN block both read V^T
            } else attn_ops_ready = 0;
            fsmn_from_vt = attn_ops_ready && ep.skip_v_out;
            vt_ptr = ep.vt; vt_tp = ep.vt_tp;
            int rc_ = sv_ln_linear(ctx, m, xin, (const float*)m->lt(l, SV_L_LN1_G), (const float*)m->lt(l, SV_L_LN1_B), cur, site(l * 4 + 0), M, T,
                                   m->lin[l * 4 + 0], qs, ep, P_G_QKV);
            if (rc_) return rc_;
        }
        const bool fork = !m->profiling && m->side != nullptr && m->fsmn_fork;
        cudaStream_t fs = fork ? m->side : ctx->stream;
        if (fork) {
            LB_CHECK_CUDA(cudaEventRecord(m->ev_fork, ctx->stream));
            LB_CHECK_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
        }
        {
            ProfScope ps(m, ctx, P_FSMN);
            if (fsmn_from_vt) {
                const size_t sm = sizeof(float) * 32 * (size_t)(vt_tp + 1);
                if (m->fsmn_v2) fsmn_vt_win_kernel<11><<<dim3(d / 32, B), 256, sm, fs>>>(vt_ptr, (const float*)m->lt(l, SV_L_FSMN_W), T, vt_tp, d, m->fsmn);
                else fsmn_vt_kernel<11><<<dim3(d / 32, B), 256, sm, fs>>>(vt_ptr, (const float*)m->lt(l, SV_L_FSMN_W), T, vt_tp, d, m->fsmn);
            } else if (m->fsmn_k == 11)
                fsmn_window_kernel<11><<<dim3(lb_ceil_div(d, 128), lb_ceil_div(T, FS_TCH), B), 128, 0, fs>>>(
                    m->qkv, (const float*)m->lt(l, SV_L_FSMN_W), T, d, m->fsmn);
            else
                fsmn_kernel<<<grid_for(M * d), 256, 0, fs>>>(m->qkv, (const float*)m->lt(l, SV_L_FSMN_W), B, T, d, m->fsmn_k, m->fsmn);
            LB_LAUNCH_CHECK(ctx);
        }
        if (fork) LB_CHECK_CUDA(cudaEventRecord(m->ev_join, m->side));
        if (attn_tc) {
            // fused tcgen05 attention (3xTF32), per-clip min/max of the output fused in its epilogue
            SV_RUN(P_ATTN_TC, lb_attention_tc(ctx, m->qkv, B, T, d, H, qscale, m->attn_scratch, m->att, site(l * 4 + 1), attn_ops_ready));
        } else {
            {
                ProfScope ps(m, ctx, P_FSMN);
                scale_copy_q_kernel<<<grid_for(M * d), 256, 0, ctx->stream>>>(m->qkv, M, d, qscale, m->qs);
                LB_LAUNCH_CHECK(ctx);
            }
            for (int hd = 0; hd < H; ++hd) {   // scores[b,hd] = (q*scale) k^T
                SV_RUN(P_ATTN_QK, lb_sgemm_strided_ldc(ctx, m->qs + hd * dk, d, 1, (long long)T * d,
                                                       m->qkv + d + hd * dk, 1, 3 * d, (long long)T * 3 * d,
                                                       m->scores + (long long)hd * T * T, (long long)H * T * T, T, B, T, dk, T, 1.0f));
            }
            {
                ProfScope ps(m, ctx, P_SOFTMAX);
                softmax_rows_kernel<<<lb_ceil_div((long long)B * H * T, 8), 256, 0, ctx->stream>>>(m->scores, (long long)B * H * T, T);
                LB_LAUNCH_CHECK(ctx);
            }
            for (int hd = 0; hd < H; ++hd) {   // att[b, :, hd*dk:(hd+1)*dk] = P v
                SV_RUN(P_ATTN_PV, lb_sgemm_strided_ldc(ctx, m->scores + (long long)hd * T * T, T, 1, (long long)H * T * T,
                                                       m->qkv + 2 * d + hd * dk, 3 * d, 1, (long long)T * 3 * d,
                                                       m->att + hd * dk, (long long)T * d, d, B, T, T, dk, 1.0f));
            }
            SV_RUN(P_MINMAX, lb_slice_minmax(ctx, m->att, B, (long long)T * d, site(l * 4 + 1)));
        }
        if (fork) LB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_join, 0));   // out-proj epilogue adds the FSMN memory
        {
            LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
            ep.out = m->x; ep.rows_per_slice = T; ep.add1 = m->fsmn; ep.add2 = (cur == d) ? xin : nullptr;   // x = x + (lin + fsmn)
            const lele_b200_qweights* wo = m->lin[l * 4 + 1];
            LbI8Epilogue epf = ep;
            epf.a_f32 = m->att; epf.a_keys = site(l * 4 + 1);
            lb_fill_weight_fields(epf, wo, qs);
            if (lb_gemm_i8_afuse_supported(ctx, M, wo->n, wo->k, epf)) {
                // the attention output is quantised INSIDE the projection (its per-clip min / max came out of the attention epilogue):
                // no quantiser launch, no u8 copy of the tensor
                SV_RUN(P_G_OUT, lb_gemm_i8(ctx, nullptr, wo->wt, (int)M, wo->n, wo->k, epf));
            } else
                SV_LINEAR(ctx, m,m->att, site(l * 4 + 1), M, T, wo, qs, ep, P_G_OUT);
        }
        xin = m->x; cur = d;
        // ---- feed-forward block ----
        const lele_b200_qweights* w1 = m->lin[l * 4 + 2];
        const bool twopass = m->ffn_twopass && T >= 32 && w1->n % 32 == 0 && w1->k % 16 == 0 && !getenv("LELE_B200_FORCE_SIMT");
        const lele_b200_qweights* w2f = m->lin[l * 4 + 3];
        const bool onepass = twopass && fq_allowed && w2f->w_signed && lb_gemm_i8_fused_q_supported(ctx, M, w1->n, w1->k, T, w1->w_signed);
        if (onepass) {
            // FFN1 once: LayerNorm + quantiser, then the GEMM whose epilogue reduces the per-clip max, parks the dequantised tile in TMEM
            // until every CTA has contributed, and quantises it -> FFN2 consumes the u8 operand; no max-only pass, no f32 [M, ffn] tensor
            LbI8Epilogue e2; memset(&e2, 0, sizeof(e2));
            e2.rows_per_slice = T; e2.relu = 1; e2.q_out = qs2.a_u8; e2.q_row_scale = qs2.row_scale; e2.q_row_zp = qs2.row_zp;
            e2.fq_keys = site(l * 4 + 3); e2.fq_counters = m->fq_counters + (size_t)(l * 4 + 3) * B;
            int rc_ = sv_ln_linear(ctx, m, m->x, (const float*)m->lt(l, SV_L_LN2_G), (const float*)m->lt(l, SV_L_LN2_B), d, site(l * 4 + 2), M, T, w1, qs, e2, P_G_FFN1);
            if (rc_) return rc_;
            LbI8Epilogue e3; memset(&e3, 0, sizeof(e3));
            e3.out = m->x; e3.rows_per_slice = T; e3.add2 = m->x;                                          // x = x + ffn
            lb_fill_weight_fields(e3, w2f, qs2);                                                           // s8 weights: FFN2 needs no row sums
            SV_RUN(P_G_FFN2, lb_gemm_i8(ctx, qs2.a_u8, w2f->wt, (int)M, w2f->n, w2f->k, e3));
        } else if (twopass) {
            // FFN1 twice over the same quantised LN output: pass 1 reduces only the per-clip max of the ReLU output (nothing is
            // written), pass 2 recomputes the tile and quantises it in the epilogue -> the [M, ffn] f32 tensor of the reference
            // (and its quantiser pass) never exists; FFN2 consumes the u8 operand directly.
            LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
            ep.rows_per_slice = T; ep.relu = 1; ep.minmax_keys = site(l * 4 + 3); ep.q_rowsum = qs2.rowsum;
            int rc_ = sv_ln_linear(ctx, m, m->x, (const float*)m->lt(l, SV_L_LN2_G), (const float*)m->lt(l, SV_L_LN2_B), d, site(l * 4 + 2), M, T, w1, qs, ep, P_G_FFN1_MAX);
            if (rc_) return rc_;
            LbI8Epilogue e2; memset(&e2, 0, sizeof(e2));
            e2.rows_per_slice = T; e2.relu = 1; e2.q_out = qs2.a_u8; e2.q_rowsum = qs2.rowsum; e2.q_row_scale = qs2.row_scale; e2.q_row_zp = qs2.row_zp;
            e2.q_keys = site(l * 4 + 3);
            lb_fill_weight_fields(e2, w1, qs);
            SV_RUN(P_G_FFN1, lb_gemm_i8(ctx, qs.a_u8, w1->wt, (int)M, w1->n, w1->k, e2));
            LbI8Epilogue e3; memset(&e3, 0, sizeof(e3));
            e3.out = m->x; e3.rows_per_slice = T; e3.add2 = m->x;                                          // x = x + ffn
            const lele_b200_qweights* w2 = m->lin[l * 4 + 3];
            lb_fill_weight_fields(e3, w2, qs2);
            SV_RUN(P_G_FFN2, lb_gemm_i8(ctx, qs2.a_u8, w2->wt, (int)M, w2->n, w2->k, e3));
        } else {
            {
                LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
                ep.out = m->f1; ep.rows_per_slice = T; ep.relu = 1;
                ep.minmax_keys = T >= 32 ? site(l * 4 + 3) : nullptr;     // fused per-clip min/max of the ReLU output
                int rc_ = sv_ln_linear(ctx, m, m->x, (const float*)m->lt(l, SV_L_LN2_G), (const float*)m->lt(l, SV_L_LN2_B), d, site(l * 4 + 2), M, T,
                                       m->lin[l * 4 + 2], qs, ep, P_G_FFN1);
                if (rc_) return rc_;
                if (T < 32) SV_RUN(P_MINMAX, lb_slice_minmax(ctx, m->f1, B, (long long)T * ffn, site(l * 4 + 3)));   // very short clips
            }
            {
                LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
                ep.out = m->x; ep.rows_per_slice = T; ep.add2 = m->x;                                          // x = x + ffn
                SV_LINEAR(ctx, m,m->f1, site(l * 4 + 3), M, T, m->lin[l * 4 + 3], qs, ep, P_G_FFN2);
            }
        }
        if (l == m->n_stage1 - 1) {   // after_norm (output replaces the residual stream)
            // in place for the register-resident row kernels (a warp holds its whole row before writing it back)
            float* ln_out = (d == 512 || d == 560) ? m->x : m->h;
            SV_RUN(P_LAYERNORM, lb_layer_norm_minmax(ctx, m->x, (const float*)m->tensor(SV_G_AFTER_G), (const float*)m->tensor(SV_G_AFTER_B), M, d,
                                                     1e-5f, ln_out, nullptr, T));
            if (ln_out != m->x) LB_CHECK_CUDA(cudaMemcpyAsync(m->x, m->h, sizeof(float) * M * d, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (n_layers < m->n_layers) {   // truncated run (tests): expose the hidden state
        if (logits_opt) LB_CHECK_CUDA(cudaMemcpyAsync(logits_opt, xin, sizeof(float) * M * cur, cudaMemcpyDeviceToDevice, ctx->stream));
        return LELE_B200_OK;
    }
    const int ctc_site = m->n_layers * 4;
    if (ids_dev) {
        ProfScope ps(m, ctx, P_ARGMAX);
        init_u64_kernel<<<lb_ceil_div(M, 256), 256, 0, ctx->stream>>>(m->amax_keys, M);
        LB_LAUNCH_CHECK(ctx);
    }
    {
        LbI8Epilogue ep; memset(&ep, 0, sizeof(ep));
        ep.out = logits_opt; ep.rows_per_slice = T; ep.argmax_keys = ids_dev ? m->amax_keys : nullptr;
        int rc_ = sv_ln_linear(ctx, m, xin, (const float*)m->tensor(SV_G_TP_G), (const float*)m->tensor(SV_G_TP_B), cur, site(ctc_site), M, T,
                               m->lin[(size_t)m->n_layers * 4], qs, ep, P_G_CTC);
        if (rc_) return rc_;
    }
    if (ids_dev) SV_RUN(P_ARGMAX, lb_argmax_keys_to_ids(ctx, m->amax_keys, M, ids_dev));
    return LELE_B200_OK;
}

// The batch split into clip lanes (lele_b200_sensevoice::n_lanes): lane i runs clips [c_i, c_i + n_i) through
// sv_encoder on its own stream with re-based workspace pointers; lane 0 stays on the caller's stream, which forks the
// others with an event and joins them at the end (works eagerly and under stream capture alike).
static int sv_encoder_lanes(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* feats, int B, int t, int lang, int textnorm,
                            int n_layers_limit, int32_t* ids_dev, float* logits_opt) {
    const int nl = (m->profiling || m->n_lanes < 2 || B < 2 * m->n_lanes) ? 1 : m->n_lanes;
    if (nl == 1) return sv_encoder(ctx, m, feats, B, t, lang, textnorm, n_layers_limit, ids_dev, logits_opt);
    const int T = t + 4, d = m->d, din = m->d_in, H = m->heads, ffn = m->ffn, dk = d / H;
    const int wide = din > d ? din : d, kmax = ffn > din ? ffn : din;
    const size_t n_sites = (size_t)m->n_layers * 4 + 1;
    const int n_layers = (n_layers_limit >= 0 && n_layers_limit < m->n_layers) ? n_layers_limit : m->n_layers;
    const long long out_w = n_layers < m->n_layers ? (n_layers == 0 ? din : d) : m->vocab;
    const size_t Tp = (size_t)((T + 3) / 4 * 4);
    LB_CHECK_CUDA(cudaEventRecord(m->ev_lane_fork, ctx->stream));
    const int lane_sms = ctx->num_sms / nl;   // each lane's persistent kernels take their share of the SMs, so the lanes run side by side
    int c0 = 0, rc = 0;
    size_t q_off = 0;
    for (int i = 0; i < nl; ++i) {
        const int nb = B / nl + (i < B % nl ? 1 : 0);
        lele_b200_sensevoice* v = m->lane[i];
        lele_b200_ctx* lctx = i == 0 ? ctx : m->lane_ctx[i];
        const long long r0 = (long long)c0 * T;
        v->x0 = m->x0 + r0 * din; v->x = m->x + r0 * d; v->h = m->h + r0 * wide; v->qkv = m->qkv + r0 * 3 * d; v->qs = m->qs + r0 * d;
        v->fsmn = m->fsmn + r0 * d; v->att = m->att + r0 * d; v->f1 = m->f1 + r0 * ffn;
        v->scores = m->scores + (long long)c0 * H * T * T;
        v->keys = m->keys + 2 * LB_MM_SLOTS * n_sites * (size_t)c0;
        v->fq_counters = m->fq_counters + n_sites * (size_t)c0;
        v->lane_sms = m->lane_split ? lane_sms : 0;
        const int sms_saved = lctx->num_sms;
        if (m->lane_split) lctx->num_sms = lane_sms;
        v->amax_keys = m->amax_keys + r0;
        v->qscratch = (uint8_t*)m->qscratch + q_off; v->qscratch2 = (uint8_t*)m->qscratch2 + q_off;
        v->attn_scratch = (float*)m->attn_scratch + (size_t)c0 * H * dk * Tp;
        v->profiling = 0;
        if (i > 0) LB_CHECK_CUDA(cudaStreamWaitEvent(lctx->stream, m->ev_lane_fork, 0));
        if (!rc) rc = sv_encoder(lctx, v, feats + (long long)c0 * t * din, nb, t, lang, textnorm, n_layers_limit, ids_dev ? ids_dev + r0 : nullptr,
                                 logits_opt ? logits_opt + r0 * out_w : nullptr);
        lctx->num_sms = sms_saved;
        if (i > 0) {   // join even after an error so that an enclosing stream capture stays well formed
            LB_CHECK_CUDA(cudaEventRecord(m->ev_lane_join[i], lctx->stream));
            LB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_lane_join[i], 0));
            ctx->launches += lctx->launches; lctx->launches = 0;
        }
        c0 += nb;
        q_off += (lb_quant_scratch_bytes((long long)nb * T, kmax) + 4095) / 4096 * 4096;
    }
    return rc;
}

static int sv_finish_profile(lele_b200_ctx* ctx, lele_b200_sensevoice* m) {
    if (!m->profiling) return LELE_B200_OK;
    LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < P_NUM; ++i) { m->prof_ms[i] = 0; m->prof_calls[i] = 0; }
    for (auto& s : m->spans) {
        float ms = 0; cudaEventElapsedTime(&ms, s.a, s.b); m->prof_ms[s.cls] += ms; m->prof_calls[s.cls]++;
        if (s.cls >= P_G_QKV) { m->prof_ms[P_GEMM_I8] += ms; m->prof_calls[P_GEMM_I8]++; }
    }
    m->spans.clear(); m->ev_used = 0;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_forward_features(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* feats_dev, int n_clips,
                                                     int t, int lang, int textnorm, int n_layers_limit, int32_t* ids_dev,
                                                     float* logits_dev_opt) {
    LB_REQUIRE(ctx && m && feats_dev, "sensevoice_forward_features: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_clips >= 1 && n_clips <= m->max_clips && t >= 1 && t + 4 <= m->max_T, "sensevoice_forward_features: batch/length exceeds workspace");
    int rc = sv_encoder_lanes(ctx, m, feats_dev, n_clips, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
    if (rc) return rc;
    return sv_finish_profile(ctx, m);
}

static int sv_forward_eager(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_dev, int n_clips, int n_samples, int t,
                            int lang, int textnorm, int n_layers_limit, int32_t* ids_dev, float* logits_dev_opt) {
    SV_RUN(P_FRONTEND, lele_b200_frontend_compute(ctx, pcm_dev, n_clips, n_samples, n_samples, nullptr, m->lfr));
    SV_RUN(P_CMVN, lele_b200_cmvn(ctx, m->lfr, n_clips, t, m->d_in, 1e-5f, m->feats));
    return sv_encoder_lanes(ctx, m, m->feats, n_clips, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
}

extern "C" int lele_b200_sensevoice_forward(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_dev, int n_clips, int n_samples,
                                            int lang, int textnorm, int n_layers_limit, int32_t* ids_dev, float* logits_dev_opt) {
    LB_REQUIRE(ctx && m && pcm_dev, "sensevoice_forward: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_clips >= 1 && n_clips <= m->max_clips && n_samples <= m->max_samples, "sensevoice_forward: batch/length exceeds workspace");
    int frames = lele_b200_frontend_num_frames(n_samples);
    LB_REQUIRE(frames > 0, "sensevoice_forward: clip shorter than one frame (400 samples)");
    int t = (frames + 5) / 6;
    if (m->profiling || !m->use_graph) {
        int rc = sv_forward_eager(ctx, m, pcm_dev, n_clips, n_samples, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
        if (rc) return rc;
        m->warmed = true;
        return sv_finish_profile(ctx, m);
    }
    const lele_b200_sensevoice::GraphKey key = {pcm_dev, n_clips, n_samples, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt};
    auto same = [](const lele_b200_sensevoice::GraphKey& a, const lele_b200_sensevoice::GraphKey& b) {   // (not memcmp: the struct has padding)
        return a.pcm == b.pcm && a.B == b.B && a.n_samples == b.n_samples && a.lang == b.lang && a.textnorm == b.textnorm &&
               a.n_layers == b.n_layers && a.ids == b.ids && a.logits == b.logits;
    };
    for (int g = 0; g < lele_b200_sensevoice::N_GRAPHS; ++g)
        if (m->graph_exec[g] && same(key, m->graph_key[g])) {
            LB_CHECK_CUDA(cudaGraphLaunch(m->graph_exec[g], ctx->stream));
            ctx->launches += m->graph_launches[g];
            m->graph_used[g] = ++m->graph_clock;
            return LELE_B200_OK;
        }
    // Caller-owned buffers (the numpy-form entry points allocate fresh ones per call) would miss on every call and evict the
    // graphs of the library-owned staging slots the serving entries replay: those calls run eagerly and capture nothing.
    bool capture_now = pcm_dev == m->pcm_stage || pcm_dev == m->pcm_slot[0] || pcm_dev == m->pcm_slot[1];
    for (int i = 0; i < 4 && !capture_now; ++i) capture_now = m->seen_key[i].pcm != nullptr && same(key, m->seen_key[i]);
    if (!capture_now) {
        m->seen_key[m->seen_next] = key; m->seen_next = (m->seen_next + 1) % 4;
        int rc = sv_forward_eager(ctx, m, pcm_dev, n_clips, n_samples, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
        if (rc) return rc;
        m->warmed = true;
        return LELE_B200_OK;
    }
    // new shape / pointers: one eager pass (also the warm-up that performs every lazy allocation), then capture
    int rc = sv_forward_eager(ctx, m, pcm_dev, n_clips, n_samples, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
    if (rc) return rc;
    m->warmed = true;
    int gi = 0;
    for (int g = 1; g < lele_b200_sensevoice::N_GRAPHS; ++g)
        if (!m->graph_exec[g] ? m->graph_exec[gi] != nullptr : (m->graph_exec[gi] && m->graph_used[g] < m->graph_used[gi])) gi = g;
    if (m->graph_exec[gi]) { cudaGraphExecDestroy(m->graph_exec[gi]); m->graph_exec[gi] = nullptr; }
    const unsigned long long l0 = ctx->launches;
    LB_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    rc = sv_forward_eager(ctx, m, pcm_dev, n_clips, n_samples, t, lang, textnorm, n_layers_limit, ids_dev, logits_dev_opt);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    m->graph_launches[gi] = ctx->launches - l0;
    ctx->launches = l0;            // captured, not executed
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { lb_set_error("sensevoice_forward: graph capture failed: %s", cudaGetErrorString(ce)); return LELE_B200_ERR_CUDA; }
    ce = cudaGraphInstantiate(&m->graph_exec[gi], graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { m->graph_exec[gi] = nullptr; lb_set_error("sensevoice_forward: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return LELE_B200_ERR_CUDA; }
    m->graph_key[gi] = key;
    m->graph_used[gi] = ++m->graph_clock;
    cudaGraphUpload(m->graph_exec[gi], ctx->stream);   // pre-stage the executable graph on the device: the first replay then costs what every replay costs
    cudaGetLastError();
    return LELE_B200_OK;           // this call's result was produced by the eager pass above
}

extern "C" int lele_b200_sensevoice_transcribe_host(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_host, int n_clips,
                                                    int n_samples, int lang, int textnorm, int32_t* ids_host) {
    LB_REQUIRE(ctx && m && pcm_host && ids_host, "sensevoice_transcribe_host: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_clips >= 1 && n_clips <= m->max_clips && n_samples <= m->max_samples, "sensevoice_transcribe_host: batch/length exceeds workspace");
    int T = lele_b200_sensevoice_rows(m, n_samples);
    LB_REQUIRE(T > 0, "sensevoice_transcribe_host: clip shorter than one frame");
    LB_CHECK_CUDA(cudaMemcpyAsync(m->pcm_stage, pcm_host, sizeof(float) * (size_t)n_clips * n_samples, cudaMemcpyHostToDevice, ctx->stream));
    int rc = lele_b200_sensevoice_forward(ctx, m, m->pcm_stage, n_clips, n_samples, lang, textnorm, -1, m->ids_stage, nullptr);
    if (rc) return rc;
    LB_CHECK_CUDA(cudaMemcpyAsync(ids_host, m->ids_stage, sizeof(int32_t) * (size_t)n_clips * T, cudaMemcpyDeviceToHost, ctx->stream));
    LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return LELE_B200_OK;
}

// Pipelined host entry: submit batch i into slot i % 2 and collect it later; the H2D copy of the next batch (its own
// stream) overlaps the forward of the current one, the D2H of the ids (a third stream) overlaps the next forward.
// pcm_host / ids_host should be pinned; both must stay valid until the matching wait.
static int sv_pipeline_init(lele_b200_sensevoice* m) {
    if (m->h2d_stream) return LELE_B200_OK;
    LB_CHECK_CUDA(cudaStreamCreateWithFlags(&m->h2d_stream, cudaStreamNonBlocking));
    LB_CHECK_CUDA(cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
    const size_t B = m->max_clips, M = B * m->max_T;
    for (int sl = 0; sl < lele_b200_sensevoice::N_SLOTS; ++sl) {
        int rc = sv_alloc((void**)&m->pcm_slot[sl], sizeof(float) * B * (size_t)m->max_samples);
        if (!rc) rc = sv_alloc((void**)&m->ids_slot[sl], sizeof(int32_t) * M);
        if (rc) return rc;
        LB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_h2d[sl], cudaEventDisableTiming));
        LB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fwd[sl], cudaEventDisableTiming));
        LB_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_d2h[sl], cudaEventDisableTiming));
    }
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_transcribe_host_async(lele_b200_ctx* ctx, lele_b200_sensevoice* m, const float* pcm_host, int n_clips,
                                                          int n_samples, int lang, int textnorm, int32_t* ids_host, int slot) {
    LB_REQUIRE(ctx && m && pcm_host, "sensevoice_transcribe_host_async: NULL argument");
    LB_REQUIRE(ids_host || (m->comm && lele_b200_comm_rank(m->comm) != m->comm_root), "sensevoice_transcribe_host_async: ids_host may be NULL only on a non-root rank of an attached communicator");
    LB_ENTER(ctx);
    LB_REQUIRE(slot >= 0 && slot < lele_b200_sensevoice::N_SLOTS, "sensevoice_transcribe_host_async: slot %d out of range", slot);
    LB_REQUIRE(n_clips >= 1 && n_clips <= m->max_clips && n_samples <= m->max_samples, "sensevoice_transcribe_host_async: batch/length exceeds workspace");
    LB_REQUIRE(!m->slot_busy[slot], "sensevoice_transcribe_host_async: slot %d still in flight (call transcribe_wait first)", slot);
    int T = lele_b200_sensevoice_rows(m, n_samples);
    LB_REQUIRE(T > 0, "sensevoice_transcribe_host_async: clip shorter than one frame");
    int rc = sv_pipeline_init(m);
    if (rc) return rc;
    // the slot's staging buffers were last read by the forward / D2H of its previous batch, which transcribe_wait(slot) joined
    LB_CHECK_CUDA(cudaMemcpyAsync(m->pcm_slot[slot], pcm_host, sizeof(float) * (size_t)n_clips * n_samples, cudaMemcpyHostToDevice, m->h2d_stream));
    LB_CHECK_CUDA(cudaEventRecord(m->ev_h2d[slot], m->h2d_stream));
    LB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_h2d[slot], 0));
    rc = lele_b200_sensevoice_forward(ctx, m, m->pcm_slot[slot], n_clips, n_samples, lang, textnorm, -1, m->ids_slot[slot], nullptr);
    if (rc) return rc;
    const size_t n_ids = (size_t)n_clips * T;
    const int world = m->comm ? lele_b200_comm_world(m->comm) : 1;
    if (world > 1) {
        const bool root = lele_b200_comm_rank(m->comm) == m->comm_root;
        if (root && !m->ids_gather[slot]) {
            rc = sv_alloc((void**)&m->ids_gather[slot], sizeof(int32_t) * (size_t)world * m->max_clips * m->max_T);
            if (rc) return rc;
        }
        rc = lele_b200_comm_gather(ctx, m->comm, m->ids_slot[slot], root ? m->ids_gather[slot] : nullptr, sizeof(int32_t) * n_ids, m->comm_root);
        if (rc) return rc;
        LB_CHECK_CUDA(cudaEventRecord(m->ev_fwd[slot], ctx->stream));
        LB_CHECK_CUDA(cudaStreamWaitEvent(m->d2h_stream, m->ev_fwd[slot], 0));
        if (root) LB_CHECK_CUDA(cudaMemcpyAsync(ids_host, m->ids_gather[slot], sizeof(int32_t) * n_ids * world, cudaMemcpyDeviceToHost, m->d2h_stream));
    } else {
        LB_CHECK_CUDA(cudaEventRecord(m->ev_fwd[slot], ctx->stream));
        LB_CHECK_CUDA(cudaStreamWaitEvent(m->d2h_stream, m->ev_fwd[slot], 0));
        LB_CHECK_CUDA(cudaMemcpyAsync(ids_host, m->ids_slot[slot], sizeof(int32_t) * n_ids, cudaMemcpyDeviceToHost, m->d2h_stream));
    }
    LB_CHECK_CUDA(cudaEventRecord(m->ev_d2h[slot], m->d2h_stream));
    m->slot_busy[slot] = true;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_transcribe_wait(lele_b200_ctx* ctx, lele_b200_sensevoice* m, int slot) {
    LB_REQUIRE(ctx && m, "sensevoice_transcribe_wait: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(slot >= 0 && slot < lele_b200_sensevoice::N_SLOTS, "sensevoice_transcribe_wait: slot %d out of range", slot);
    if (!m->slot_busy[slot]) return LELE_B200_OK;
    LB_CHECK_CUDA(cudaEventSynchronize(m->ev_d2h[slot]));
    m->slot_busy[slot] = false;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_set_comm(lele_b200_sensevoice* m, lele_b200_comm* comm, int root) {
    LB_REQUIRE(m, "sensevoice_set_comm: NULL model");
    LB_REQUIRE(!comm || (root >= 0 && root < lele_b200_comm_world(comm)), "sensevoice_set_comm: root %d outside the communicator", root);
    for (int sl = 0; sl < lele_b200_sensevoice::N_SLOTS; ++sl)
        LB_REQUIRE(!m->slot_busy[sl], "sensevoice_set_comm: slot %d is in flight", sl);
    m->comm = comm; m->comm_root = root;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sensevoice_workspace(lele_b200_sensevoice* m, const char* name, void** dptr, size_t* nbytes) {
    LB_REQUIRE(m && name && dptr && nbytes, "sensevoice_workspace: NULL argument");
    const size_t B = m->max_clips, M = B * m->max_T;
    const int wide = m->d_in > m->d ? m->d_in : m->d;
    const int t = m->max_T - 4;
    struct { const char* n; void* p; size_t b; } tab[] = {
        {"lfr", m->lfr, sizeof(float) * B * t * m->d_in}, {"feats", m->feats, sizeof(float) * B * t * m->d_in},
        {"x0", m->x0, sizeof(float) * M * m->d_in}, {"x", m->x, sizeof(float) * M * m->d}, {"h", m->h, sizeof(float) * M * wide},
        {"qkv", m->qkv, sizeof(float) * M * 3 * m->d}, {"fsmn", m->fsmn, sizeof(float) * M * m->d}, {"att", m->att, sizeof(float) * M * m->d},
        {"f1", m->f1, sizeof(float) * M * m->ffn}, {"vt", m->attn_scratch, lb_attention_tc_scratch_bytes(m->max_clips, m->max_T, m->d, m->heads)}, {"keys", m->keys, sizeof(unsigned) * 2 * LB_MM_SLOTS * B * ((size_t)m->n_layers * 4 + 1)}};
    for (auto& e : tab)
        if (strcmp(e.n, name) == 0) { *dptr = e.p; *nbytes = e.b; return LELE_B200_OK; }
    lb_set_error("sensevoice_workspace: unknown buffer '%s'", name);
    return LELE_B200_ERR_ARG;
}

extern "C" int lele_b200_sensevoice_set_profiling(lele_b200_sensevoice* m, int enable) {
    LB_REQUIRE(m, "sensevoice_set_profiling: NULL model");
    m->profiling = enable ? 1 : 0;
    return LELE_B200_OK;
}
extern "C" int lele_b200_sensevoice_last_profile(lele_b200_sensevoice* m, const char** names, float* ms, int* calls, int cap,
                                                 int* n_out) {
    LB_REQUIRE(m && names && ms && calls && n_out, "sensevoice_last_profile: NULL argument");
    int n = 0;
    for (int i = 0; i < P_NUM && n < cap; ++i) { names[n] = kProfNames[i]; ms[n] = m->prof_ms[i]; calls[n] = m->prof_calls[i]; ++n; }
    *n_out = n;
    return LELE_B200_OK;
}
