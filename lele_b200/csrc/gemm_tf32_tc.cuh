// gemm_tf32_tc.cuh -- interface of the tcgen05 3xTF32 f32 GEMM (gemm_tf32_tc.cu)
#pragma once
#include "common.cuh"

struct LbGemmTcEpilogue {
    float alpha = 1.0f;
    int pre_mode = 0;                 // 1: C already holds the pre-fill (bias / beta*C): C = C + alpha*acc
    const float* bias_row = nullptr;  // [m]: v = v + bias_row[row] (conv: per output channel), applied before act
    int act = 0;                      // 0 none, 1 ReLU, 2 SiLU (SIMD body for columns < simd_end, libm tail; avx/math.rs:344-470)
    int simd_end = 0;
};
// B operand produced by the kernel's converter warps instead of TMA
struct LbGatherB {
    int mode;            // 1: N-major matrix, element (n, k) = ptr[b*bs + k*ldk + n];  2: im2col of an NCHW image (implicit GEMM)
    const float* ptr;
    long long ldk, bs;   // bs = batch stride in floats (0 = shared)
    int h, w, kh, kw, pt, pl, sh, sw, dh, dw, ow;   // mode 2: n = oy*ow + ox, k = (c*kh + ky)*kw + kx
    int k_valid;         // mode 2: true IC*kh*kw when the GEMM's K was padded up to a multiple of 4 (0 = K is the true size)
};
// TMA-addressable operands: 16-byte aligned bases, pitches and batch strides multiples of 4 floats, k >= 4
bool lb_gemm_tc_supported(const float* A, long long lda, long long bsa, const float* B, long long ldb, long long bsb, int m, int n, int k);
// C[b][m,n] (row pitch ldc, batch stride bsc) = epilogue(A[b][m,k] . B[b][n,k]^T); both operands K-major;
// bsa / bsb == 0 broadcast one matrix over the batch
int lb_gemm_tf32x3_nt(lele_b200_ctx* ctx, const float* A, long long lda, long long bsa, const float* B, long long ldb, long long bsb, float* C,
                      long long ldc, long long bsc, int batch, int m, int n, int k, const LbGemmTcEpilogue& ep);
int lb_gemm_tf32x3_gather(lele_b200_ctx* ctx, const float* A, long long lda, long long bsa, const LbGatherB& gb, float* C, long long ldc, long long bsc,
                          int batch, int m, int n, int k, const LbGemmTcEpilogue& ep);
// out[b][c][r] = in[b][r][c]
int lb_transpose_f32(lele_b200_ctx* ctx, const float* in, long long ld_in, long long bs_in, float* out, long long ld_out, long long bs_out, int batch,
                     int rows, int cols);
