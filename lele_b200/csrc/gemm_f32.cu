// gemm_f32.cu -- f32 matmul / matmul_fused_add / gemm (src/kernels/gemm.rs:112,223,433).
// Upstream arithmetic is faer 0.24 (un-vendored); the reference's own tests pin this boundary
// to 1e-5..1e-3 abs, i.e. any correctly-rounded-ish f32 GEMM.  CUDA cores, fp32 FMA,
// 64x64x16 shared-memory tiles, 4x4 register blocking; strides make transposes free.
#include "gemm_tf32_tc.cuh"

namespace {
constexpr int TM = 64, TN = 64, TK = 16;

// C[b] (m x n, row-major) = pre[b] + alpha * op(A[b]) op(B[b]);  pre_mode: 0 none, 1 C already
// holds the pre-fill (read-modify-write)
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long rsa, long long csa, long long bsa,
             const float* __restrict__ B, long long rsb, long long csb, long long bsb,
             float* __restrict__ C, long long bsc, long long ldc, int m, int k, int n, float alpha, int pre_mode) {
    __shared__ float sa[TK][TM + 4];
    __shared__ float sb[TK][TN + 4];
    const int b = blockIdx.z;
    A += b * bsa; B += b * bsb; C += b * bsc;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < k; k0 += TK) {
        // A tile: TM x TK, B tile: TK x TN; 256 threads load 4 elements each
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int idx = threadIdx.x + i * 256;
            int am, ak;
            if (csa == 1) { ak = idx & 15; am = idx >> 4; } else { am = idx & 63; ak = idx >> 6; }   // coalesce along the unit stride
            int gm = m0 + am, gk = k0 + ak;
            sa[ak][am] = (gm < m && gk < k) ? A[gm * rsa + gk * csa] : 0.0f;
            int bk, bn;
            if (csb == 1) { bn = idx & 63; bk = idx >> 6; } else { bk = idx & 15; bn = idx >> 4; }
            int gk2 = k0 + bk, gn = n0 + bn;
            sb[bk][bn] = (gk2 < k && gn < n) ? B[gk2 * rsb + gn * csb] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= n) continue;
            float v = alpha == 1.0f ? acc[i][j] : alpha * acc[i][j];
            long long o = (long long)gm * ldc + gn;
            C[o] = pre_mode ? C[o] + v : v;
        }
    }
}

int launch_sgemm(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                 long long rsb, long long csb, long long bsb, float* C, int batch, int m, int k, int n, float alpha,
                 int pre_mode, long long bsc = -1, long long ldc = -1) {
    if (bsc < 0) bsc = (long long)m * n;
    if (ldc < 0) ldc = n;
    if (batch == 0 || m == 0 || n == 0) return LELE_B200_OK;
    dim3 grid(lb_ceil_div(n, TN), lb_ceil_div(m, TM), batch);
    LB_REQUIRE(batch <= 65535 && grid.y <= 65535, "matmul: batch/m too large for one launch");
    sgemm_kernel<<<grid, 256, 0, ctx->stream>>>(A, rsa, csa, bsa, B, rsb, csb, bsb, C, bsc, ldc, m, k, n, alpha, pre_mode);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// pre-fill helpers
__global__ void fill_rows_kernel(float* __restrict__ out, long long rows, int n, const float* __restrict__ bias) {
    long long total = rows * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = bias[i % n];
}
__global__ void add_mod_kernel(float* __restrict__ out, long long total, const float* __restrict__ bias, int len) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(out[i], bias[i % len]);
}
// gemm.rs:484-511: C pre-scaled by beta with length-based broadcast
__global__ void gemm_prefill_kernel(float* __restrict__ out, int m, int n, const float* __restrict__ c, int c_len, float beta) {
    long long total = (long long)m * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (c && beta != 0.0f) {
            long long idx;
            if ((long long)c_len == total) idx = i;
            else if (c_len == n) idx = i % n;
            else if (c_len == m) idx = i / n;
            else if (c_len == 1) idx = 0;
            else idx = i % c_len;
            v = __fmul_rn(c[idx], beta);
        }
        out[i] = v;
    }
}
int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); }
}  // namespace

int lb_sgemm_strided(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                     long long rsb, long long csb, long long bsb, float* C, int batch, int m, int k, int n, float alpha,
                     int pre_mode) {
    return launch_sgemm(ctx, A, rsa, csa, bsa, B, rsb, csb, bsb, C, batch, m, k, n, alpha, pre_mode);
}
// general form: explicit batch stride and row pitch of C (lets attention write heads in place)
int lb_sgemm_strided_ldc(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                         long long rsb, long long csb, long long bsb, float* C, long long bsc, long long ldc, int batch, int m,
                         int k, int n, float alpha) {
    return launch_sgemm(ctx, A, rsa, csa, bsa, B, rsb, csb, bsb, C, batch, m, k, n, alpha, 0, bsc, ldc);
}


// Tensor-core dispatch (gemm_tf32_tc.cu): op(A) [m,k] and op(B)^T [n,k] must be K-major; operands stored the other way
// round are transposed into the context scratch first (one extra pass over that operand).  Small problems and
// operands TMA cannot address (k % 4 != 0, unaligned) stay on the CUDA-core kernel.
//   a_kmajor: A stored [m,k] (else [k,m]);  b_nmajor: B stored [k,n] (else [n,k])
static bool tc_worthwhile(int batch, int m, int k, int n) {
    return (long long)batch * m * n * k >= (1ll << 22) && k >= 16 && n >= 16 && m >= 16;
}
static int gemm_tc_or_simt(lele_b200_ctx* ctx, const float* a, bool a_kmajor, long long bsa, const float* b, bool b_nmajor, long long bsb,
                           float* c, int batch, int m, int k, int n, float alpha, int pre_mode) {
    const bool tc = tc_worthwhile(batch, m, k, n) && k % 4 == 0 && ((long long)m * k) % 4 == 0 && lb_gemm_tc_supported(a, k, 0, a, k, 0, m, n, k) &&
                    (b_nmajor || (((long long)n * k) % 4 == 0 && (((uintptr_t)b) & 15) == 0));
    if (tc) {
        const float* A = a;
        long long bsA = bsa;
        if (!a_kmajor) {   // stored [k, m] -> [m, k] (one transposing pass into the context scratch; rare: gemm with transA)
            const int nba = bsa ? batch : 1;
            void* sc; int rc = lb_scratch(ctx, sizeof(float) * (size_t)nba * m * k, &sc);
            if (rc) return rc;
            if ((rc = lb_transpose_f32(ctx, a, m, bsa, (float*)sc, k, (long long)m * k, nba, k, m))) return rc;
            A = (const float*)sc; bsA = bsa ? (long long)m * k : 0;
        }
        LbGemmTcEpilogue ep; ep.alpha = alpha; ep.pre_mode = pre_mode;
        if (b_nmajor) {    // [k, n]: gathered by the kernel itself (no transpose pass)
            LbGatherB gb; memset(&gb, 0, sizeof(gb));
            gb.mode = 1; gb.ptr = b; gb.ldk = n; gb.bs = bsb;
            return lb_gemm_tf32x3_gather(ctx, A, k, bsA, gb, c, n, (long long)m * n, batch, m, n, k, ep);
        }
        return lb_gemm_tf32x3_nt(ctx, A, k, bsA, b, k, bsb, c, n, (long long)m * n, batch, m, n, k, ep);
    }
    return launch_sgemm(ctx, a, a_kmajor ? k : 1, a_kmajor ? 1 : m, bsa, b, b_nmajor ? n : 1, b_nmajor ? 1 : k, bsb, c, batch, m, k, n, alpha, pre_mode);
}

extern "C" int lele_b200_matmul(lele_b200_ctx* ctx, const float* a, const float* b, int batch_a, int batch_b, int m, int k,
                                int n, float* out) {
    LB_REQUIRE(ctx && a && b && out, "matmul: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(batch_a >= 1 && batch_b >= 1 && (batch_a == batch_b || batch_a == 1 || batch_b == 1),
               "matmul: batch mismatch %d vs %d (gemm.rs:134)", batch_a, batch_b);
    int fb = batch_a > batch_b ? batch_a : batch_b;
    return gemm_tc_or_simt(ctx, a, true, batch_a == 1 ? 0 : (long long)m * k, b, true, batch_b == 1 ? 0 : (long long)k * n, out, fb, m, k, n, 1.0f, 0);
}

extern "C" int lele_b200_matmul_fused_add(lele_b200_ctx* ctx, const float* a, const float* b, const float* bias, int bias_len,
                                          int batch_a, int batch_b, int m, int k, int n, float* out) {
    LB_REQUIRE(ctx && a && b && bias && out && bias_len > 0, "matmul_fused_add: bad arguments");
    LB_ENTER(ctx);
    int fb = batch_a > batch_b ? batch_a : batch_b;
    long long total = (long long)fb * m * n;
    if (total == 0) return LELE_B200_OK;
    if (bias_len == n) {  // pre-fill rows with bias, then accumulate (gemm.rs:247-330)
        fill_rows_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(out, (long long)fb * m, n, bias);
        LB_LAUNCH_CHECK(ctx);
        return gemm_tc_or_simt(ctx, a, true, batch_a == 1 ? 0 : (long long)m * k, b, true, batch_b == 1 ? 0 : (long long)k * n, out, fb, m, k, n, 1.0f, 1);
    }
    int rc = gemm_tc_or_simt(ctx, a, true, batch_a == 1 ? 0 : (long long)m * k, b, true, batch_b == 1 ? 0 : (long long)k * n, out, fb, m, k, n, 1.0f, 0);
    if (rc) return rc;
    add_mod_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(out, total, bias, bias_len);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_gemm(lele_b200_ctx* ctx, const float* a, const float* b, const float* c, int c_len, float alpha,
                              float beta, int trans_a, int trans_b, int m, int k, int n, float* out) {
    LB_REQUIRE(ctx && a && b && out, "gemm: NULL argument");
    LB_ENTER(ctx);
    if ((long long)m * n == 0) return LELE_B200_OK;
    gemm_prefill_kernel<<<grid_for((long long)m * n), 256, 0, ctx->stream>>>(out, m, n, c, c_len, beta);
    LB_LAUNCH_CHECK(ctx);
    return gemm_tc_or_simt(ctx, a, !trans_a, 0, b, !trans_b, 0, out, 1, m, k, n, alpha, 1);
}
