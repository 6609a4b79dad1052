// gemm_i8_tc.cuh -- interface of the tcgen05 (kind::i8) u8 x u8 -> s32 GEMM with the fused
// lele quantised-linear epilogue.  See gemm_i8_tc.cu.
#pragma once
#include "common.cuh"

struct LbI8Epilogue {
    // per output row (filled by the activation quantiser, quant.cu)
    const int32_t* rowsum;     // [M] sum_k a_q[row,k]
    const float* row_scale;    // [M] dynamic scale of the row's slice
    const int32_t* row_zp;     // [M] activation zero point of the row's slice
    // per output column (filled once by lele_b200_prepare_weights), padded to a multiple of 256
    const int32_t* colsum;     // sum_k w[k,col]
    const float* w_scale;      // per-channel weight scale (scalar scales are expanded)
    const float* bias;         // zeros when absent
    int w_zp;
    int w_signed;              // Wt holds (w - 128) as s8 (prepare_weights does this when w_zp == 128): the tensor core multiplies u8 x s8,
                               // colsum = sum_k (w - 128), and the row term w_zp * rowsum of the zero-point correction vanishes
    int has_bias;
    int relu;
    // optional fusions used by the SenseVoice runner (all NULL for the plain operator)
    const float* add1;         // out = (v + add1)            e.g. + fsmn memory
    const float* add2;         // out = add2 + (v [+ add1])   e.g. residual stream
    unsigned* minmax_keys;     // per-slice min/max keys of `out` ([n_slices][2])
    int rows_per_slice;
    unsigned long long* argmax_keys;  // [M] max over columns of (fkey(out) << 32 | col)
    float* out;                // [M, N]; may be NULL when only argmax_keys is wanted
    // fused operand preparation for attn_tc.cu (QKV projection only, N = 3 * heads * 128, rows_per_slice = T):
    // besides out = [q | k | v] the epilogue writes V^T to vt [clips][heads][128][vt_tp] (keys contiguous, the K-major
    // B operand of P.V).  q / k are read straight from `out`; the tf32 lo residuals are computed on chip by attn_tc.cu.
    float* vt;
    int vt_tp;
    int skip_v_out;            // the v third of `out` is not written (nobody reads it: attention and the FSMN block both take V^T)
    // fused output quantiser (two-pass linear -> dynamic quantiser, no f32 round trip): pass 1 = this GEMM with out == NULL,
    // minmax_keys set and q_rowsum set (max-only, zeroes q_rowsum); pass 2 = the same GEMM with q_out set: the epilogue
    // quantises with the per-slice (scale, zp) derived from q_keys (the keys pass 1 reduced) and emits the next GEMM's
    // operand directly: q_out u8 [M, N], q_rowsum [M] (atomic adds), q_row_scale / q_row_zp [M].  N % 32 == 0.
    uint8_t* q_out;
    int32_t* q_rowsum;
    float* q_row_scale;
    int32_t* q_row_zp;
    const unsigned* q_keys;
    // fused INPUT quantiser (lb_gemm_i8_tc with a_f32 set, A == NULL): the A operand is quantised in the kernel from the f32 activation
    // a_f32 [M, K] with the per-slice (scale, zp) derived from a_keys; rowsum / row_scale / row_zp are not read.  K % 128 == 0, K <= 512,
    // M <= #SMs * 128, s8 weights, add1-only (+ in-place add2) epilogue.
    const float* a_f32;
    const unsigned* a_keys;
    // single-pass variant (lb_gemm_i8_tc_fused_q): the accumulators wait in TMEM for the clip's min / max instead of being
    // recomputed; fq_keys = the output's per-clip key slots (initialised), fq_counters = [n_clips] zeroed arrival counters
    unsigned* fq_keys;
    int* fq_counters;
};

// A: u8 [M, K] row-major (K-major); Wt: u8 [N, K] row-major (K-major).  K % 16 == 0.
int lb_gemm_i8_tc(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K,
                  const LbI8Epilogue& ep);

// one-pass linear -> (ReLU) -> dynamic quantiser: q_out / q_row_scale / q_row_zp + fq_keys / fq_counters; s8 weights only
// the fused input quantiser above is available for this problem (the caller then skips its quantiser launch)
bool lb_gemm_i8_afuse_supported(lele_b200_ctx* ctx, long long M, int N, int K, const LbI8Epilogue& ep);
bool lb_gemm_i8_fused_q_supported(lele_b200_ctx* ctx, long long M, int N, int K, int T, int w_signed);
int lb_gemm_i8_tc_fused_q(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K, const LbI8Epilogue& ep);

// ---- shared between quant.cu and the graph runner (sensevoice.cu) ----
struct lele_b200_qweights {
    uint8_t* wt = nullptr;      // [n, k]  K-major copy of the u8 weight (lele's b_t, quantization.rs:206); w ^ 0x80 (= s8 w - 128) when w_signed
    int32_t* colsum = nullptr;  // [n_pad] sum_k w, or sum_k (w - 128) when w_signed
    float* w_scale = nullptr;   // [n_pad] (scalar scales expanded)
    float* bias = nullptr;      // [n_pad] zeros when absent
    int k = 0, n = 0, n_pad = 0, w_zp = 0, has_bias = 0, w_signed = 0;
};
struct LbQuantScratch { uint8_t* a_u8; int32_t* rowsum; float* row_scale; int32_t* row_zp; };
size_t lb_quant_scratch_bytes(long long M, int K);
LbQuantScratch lb_quant_scratch_carve(void* base, long long M, int K);
int lb_minmax_init(lele_b200_ctx* ctx, unsigned* keys, int n_slices);
int lb_slice_minmax(lele_b200_ctx* ctx, const float* x, int n_slices, long long slice_len, unsigned* keys);
int lb_quantize_rows(lele_b200_ctx* ctx, const float* x, const unsigned* keys, long long M, int rows_per_slice, int K,
                     uint8_t* a_u8, int32_t* rowsum, float* row_scale, int32_t* row_zp);
// dispatches to the tcgen05 kernel (K % 16 == 0) or the CUDA-core kernel
int lb_gemm_i8(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K, const LbI8Epilogue& ep);
void lb_fill_weight_fields(LbI8Epilogue& ep, const lele_b200_qweights* w, const LbQuantScratch& s);
