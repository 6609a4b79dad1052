// elementwise.cu -- lele::kernels::math element-wise family (src/kernels/math.rs:414-2300):
// NumPy-broadcast binary ops, activations (x86 SIMD-body polynomial / scalar-tail libm split,
// avx/math.rs:232-600), clip, reductions, where.  HBM-bound: float4 grid-stride loops.
#include "common.cuh"

namespace {
constexpr int MAXR = 8;
struct Bcast {
    int rank;
    long long out_shape[MAXR];
    long long sa[MAXR], sb[MAXR], sc[MAXR];   // element strides (0 = broadcast)
    long long total;
};

// right-align shapes, compute broadcast shape + strides (utils.rs:107-132 broadcast_shapes)
int make_bcast(Bcast& bc, const long long* a_shape, int ar, const long long* b_shape, int br,
               const long long* c_shape = nullptr, int cr = 0) {
    int r = ar > br ? ar : br; if (cr > r) r = cr;
    if (r > MAXR) { lb_set_error("broadcast: rank %d > %d", r, MAXR); return LELE_B200_ERR_ARG; }
    bc.rank = r; bc.total = 1;
    long long as[MAXR], bs[MAXR], cs[MAXR];
    for (int i = 0; i < r; ++i) {
        int ia = i - (r - ar), ib = i - (r - br), ic = i - (r - cr);
        as[i] = ia >= 0 ? a_shape[ia] : 1; bs[i] = ib >= 0 ? b_shape[ib] : 1; cs[i] = (c_shape && ic >= 0) ? c_shape[ic] : 1;
        long long d = as[i] > bs[i] ? as[i] : bs[i]; if (cs[i] > d) d = cs[i];
        if ((as[i] != d && as[i] != 1) || (bs[i] != d && bs[i] != 1) || (cs[i] != d && cs[i] != 1)) {
            lb_set_error("broadcast: incompatible shapes at dim %d", i); return LELE_B200_ERR_ARG;
        }
        bc.out_shape[i] = d; bc.total *= d;
    }
    long long ca = 1, cb = 1, cc = 1;
    for (int i = r - 1; i >= 0; --i) {
        bc.sa[i] = as[i] == 1 ? 0 : ca; ca *= as[i];
        bc.sb[i] = bs[i] == 1 ? 0 : cb; cb *= bs[i];
        bc.sc[i] = cs[i] == 1 ? 0 : cc; cc *= cs[i];
    }
    return LELE_B200_OK;
}

__device__ __forceinline__ float bin_op(int op, float a, float b) {
    switch (op) {
        case LELE_B200_ADD: return __fadd_rn(a, b);
        case LELE_B200_SUB: return __fsub_rn(a, b);
        case LELE_B200_MUL: return __fmul_rn(a, b);
        case LELE_B200_DIV: return __fdiv_rn(a, b);
        case LELE_B200_MAX: return fmaxf(a, b);
        case LELE_B200_POW: return powf(a, b);
        case LELE_B200_MOD: return b == 0.0f ? 0.0f : __fsub_rn(a, __fmul_rn(b, floorf(__fdiv_rn(a, b))));   // math.rs:1163
        case LELE_B200_PRELU: return a < 0.0f ? __fmul_rn(a, b) : a;                                            // math.rs:2012
        case LELE_B200_EQUAL: return a == b ? 1.0f : 0.0f;
        case LELE_B200_LESS: return a < b ? 1.0f : 0.0f;
    }
    return 0.0f;
}

__global__ void binary_flat_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, long long total,
                                   int a_scalar, int b_scalar, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = !a_scalar && !b_scalar && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0);
    if (vec) {
        const long long nv = total >> 2;
        for (long long q = i; q < nv; q += stride) {
            float4 x = __ldg(reinterpret_cast<const float4*>(a) + q), y = __ldg(reinterpret_cast<const float4*>(b) + q);
            reinterpret_cast<float4*>(out)[q] = make_float4(bin_op(op, x.x, y.x), bin_op(op, x.y, y.y), bin_op(op, x.z, y.z), bin_op(op, x.w, y.w));
        }
        for (long long q = (nv << 2) + i; q < total; q += stride) out[q] = bin_op(op, a[q], b[q]);
    } else {
        const float as = a_scalar ? a[0] : 0.0f, bs = b_scalar ? b[0] : 0.0f;
        for (long long q = i; q < total; q += stride) out[q] = bin_op(op, a_scalar ? as : a[q], b_scalar ? bs : b[q]);
    }
}
__global__ void binary_bcast_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, Bcast bc, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bc.total; i += (long long)gridDim.x * blockDim.x) {
        long long rem = i, oa = 0, ob = 0;
        for (int d = bc.rank - 1; d >= 0; --d) {
            long long c = rem % bc.out_shape[d]; rem /= bc.out_shape[d];
            oa += c * bc.sa[d]; ob += c * bc.sb[d];
        }
        out[i] = bin_op(op, a[oa], b[ob]);
    }
}
// Periodic broadcast: one operand has the output's shape, the other is broadcast with index (i / inner) % period
// (trailing-dims operand such as a bias row: inner = 1; per-channel NC[HW] operand: inner = H*W).  float4 over the big
// operand; the small operand is read per element (L1-resident) or as float4 when inner == 1.
__global__ void binary_periodic_kernel(int op, const float* __restrict__ big, const float* __restrict__ small, long long total, long long inner,
                                       long long period, int small_is_a, float* __restrict__ out) {
    const long long nv = total >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += stride) {
        const long long i = q << 2;
        const float4 x = __ldg(reinterpret_cast<const float4*>(big) + q);
        float s0, s1, s2, s3;
        if (inner == 1) {                                   // period % 4 == 0 (checked by the caller): 4 consecutive small elements
            const float4 y = __ldg(reinterpret_cast<const float4*>(small + (i % period)));
            s0 = y.x; s1 = y.y; s2 = y.z; s3 = y.w;
        } else {                                            // inner % 4 == 0: the 4 elements share one small element
            s0 = s1 = s2 = s3 = __ldg(small + (i / inner) % period);
        }
        float4 r;
        if (small_is_a) r = make_float4(bin_op(op, s0, x.x), bin_op(op, s1, x.y), bin_op(op, s2, x.z), bin_op(op, s3, x.w));
        else r = make_float4(bin_op(op, x.x, s0), bin_op(op, x.y, s1), bin_op(op, x.z, s2), bin_op(op, x.w, s3));
        reinterpret_cast<float4*>(out)[q] = r;
    }
}
__global__ void where_kernel(const float* __restrict__ c, const float* __restrict__ x, const float* __restrict__ y, Bcast bc,
                             float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bc.total; i += (long long)gridDim.x * blockDim.x) {
        long long rem = i, oc = 0, ox = 0, oy = 0;
        for (int d = bc.rank - 1; d >= 0; --d) {
            long long q = rem % bc.out_shape[d]; rem /= bc.out_shape[d];
            oc += q * bc.sa[d]; ox += q * bc.sb[d]; oy += q * bc.sc[d];
        }
        out[i] = c[oc] != 0.0f ? x[ox] : y[oy];
    }
}

__device__ __forceinline__ float un_op(int op, float v, bool simd) {
    switch (op) {
        case LELE_B200_RELU: return fmaxf(v, 0.0f);
        case LELE_B200_SIGMOID: return simd ? lb_sigmoid_simd(v) : __fdiv_rn(1.0f, __fadd_rn(1.0f, lb_libm_expf(-v)));
        case LELE_B200_TANH: return simd ? lb_tanh_simd(v) : tanhf(v);
        case LELE_B200_SILU: return simd ? __fmul_rn(v, lb_sigmoid_simd(v)) : __fdiv_rn(v, __fadd_rn(1.0f, lb_libm_expf(-v)));
        case LELE_B200_ERF: return simd ? lb_erf_simd(v) : erff(v);
        case LELE_B200_GELU: {   // 0.5 x (1 + erf(x / sqrt2))  math.rs:906
            float t = __fmul_rn(v, 0.70710678f);
            float e = simd ? lb_erf_simd(t) : erff(t);
            return __fmul_rn(__fmul_rn(0.5f, v), __fadd_rn(1.0f, e));
        }
        case LELE_B200_FAST_GELU: {  // tanh form with 0.044715  math.rs:950
            float u = __fmul_rn(0.7978845608f, __fadd_rn(v, __fmul_rn(0.044715f, __fmul_rn(v, __fmul_rn(v, v)))));
            return __fmul_rn(__fmul_rn(0.5f, v), __fadd_rn(1.0f, tanhf(u)));
        }
        case LELE_B200_EXP: return simd ? lb_cephes_expf(v) : expf(v);
        case LELE_B200_SOFTPLUS: return v > 20.0f ? v : logf(__fadd_rn(1.0f, lb_libm_expf(v)));   // math.rs:1046
        case LELE_B200_LOG: return logf(v);
        case LELE_B200_SQRT: return __fsqrt_rn(v);
        case LELE_B200_NEG: return -v;
        case LELE_B200_RECIPROCAL: return __fdiv_rn(1.0f, v);
        case LELE_B200_SIN: return sinf(v);
        case LELE_B200_COS: return cosf(v);
        case LELE_B200_NOT: return v == 0.0f ? 1.0f : 0.0f;
    }
    return v;
}
__global__ void unary_kernel(int op, const float* __restrict__ x, long long len, float* __restrict__ out) {
    const long long simd_end = (len / 8) * 8, stride = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0) {      // float4 body (a float4 never straddles the 8-element SIMD boundary)
        const long long nv = len >> 2;
        for (long long q = t; q < nv; q += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
            const bool simd = (q << 2) < simd_end;
            reinterpret_cast<float4*>(out)[q] = make_float4(un_op(op, v.x, simd), un_op(op, v.y, simd), un_op(op, v.z, simd), un_op(op, v.w, simd));
        }
        for (long long i = (nv << 2) + t; i < len; i += stride) out[i] = un_op(op, x[i], i < simd_end);
        return;
    }
    for (long long i = t; i < len; i += stride) out[i] = un_op(op, x[i], i < simd_end);
}
__global__ void clip_kernel(const float* __restrict__ x, long long len, float lo, float hi, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(x[i], lo), hi);
}
// reduce over the middle axis of [outer, axis_len, inner]; sequential over the axis in input
// order per output element (math.rs:1527-1921 accumulates in linear input order).
__global__ void reduce_kernel(int kind, const float* __restrict__ x, long long outer, int axis_len, long long inner,
                              float* __restrict__ out) {
    const long long total = outer * inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long o = i / inner, in = i % inner;
        const float* p = x + o * axis_len * inner + in;
        float acc = kind == 2 ? -INFINITY : 0.0f;
        for (int a = 0; a < axis_len; ++a) {
            float v = p[(long long)a * inner];
            if (kind == 2) acc = fmaxf(acc, v);
            else if (kind == 3) acc = __fadd_rn(acc, __fmul_rn(v, v));
            else acc = __fadd_rn(acc, v);
        }
        if (kind == 1) acc = __fmul_rn(acc, __fdiv_rn(1.0f, (float)axis_len));
        if (kind == 3) acc = __fsqrt_rn(acc);
        out[i] = acc;
    }
}
int grid_for(long long total, int per = 1) { long long g = (total / per + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); }
}  // namespace

extern "C" int lele_b200_binary(lele_b200_ctx* ctx, int op, const float* a, const long long* a_shape, int a_rank, const float* b,
                                const long long* b_shape, int b_rank, float* out) {
    LB_REQUIRE(ctx && a && b && out, "binary: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(op >= LELE_B200_ADD && op <= LELE_B200_LESS, "binary: unknown op %d", op);
    Bcast bc;
    int rc = make_bcast(bc, a_shape, a_rank, b_shape, b_rank);
    if (rc) return rc;
    if (bc.total == 0) return LELE_B200_OK;
    long long na = 1, nb = 1;
    for (int i = 0; i < a_rank; ++i) na *= a_shape[i];
    for (int i = 0; i < b_rank; ++i) nb *= b_shape[i];
    if ((na == bc.total || na == 1) && (nb == bc.total || nb == 1)) {   // same-shape / scalar fast paths (math.rs:414-470)
        binary_flat_kernel<<<grid_for(bc.total, 4), 256, 0, ctx->stream>>>(op, a, b, bc.total, na == 1 && bc.total != 1, nb == 1 && bc.total != 1, out);
    } else {
        // periodic fast path: one operand full-size, the other's non-broadcast dims form one contiguous run of output dims
        bool done = false;
        if ((na == bc.total) != (nb == bc.total) && bc.total % 4 == 0 && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0) {
            const bool small_is_a = nb == bc.total;
            const long long* ss = small_is_a ? bc.sa : bc.sb;
            int lo = -1, hi = -1; bool ok = true;
            for (int d = 0; d < bc.rank; ++d) if (bc.out_shape[d] != 1 && ss[d] != 0) { if (lo < 0) lo = d; hi = d; }
            long long period = 1, inner = 1, expect = 1;
            if (lo < 0) ok = false;
            for (int d = bc.rank - 1; d >= 0 && ok; --d) {
                if (d > hi) inner *= bc.out_shape[d];
                else if (d >= lo) { if (bc.out_shape[d] != 1 && ss[d] != expect) ok = false; expect *= bc.out_shape[d]; period *= bc.out_shape[d]; }
            }
            if (ok && ((inner == 1 && period % 4 == 0) || (inner > 1 && inner % 4 == 0))) {
                binary_periodic_kernel<<<grid_for(bc.total, 4), 256, 0, ctx->stream>>>(op, small_is_a ? b : a, small_is_a ? a : b, bc.total, inner, period, small_is_a ? 1 : 0, out);
                done = true;
            }
        }
        if (!done) binary_bcast_kernel<<<grid_for(bc.total), 256, 0, ctx->stream>>>(op, a, b, bc, out);
    }
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_where(lele_b200_ctx* ctx, const float* cond, const long long* c_shape, int c_rank, const float* x,
                               const long long* x_shape, int x_rank, const float* y, const long long* y_shape, int y_rank,
                               float* out) {
    LB_REQUIRE(ctx && cond && x && y && out, "where: NULL argument");
    LB_ENTER(ctx);
    Bcast bc;
    int rc = make_bcast(bc, c_shape, c_rank, x_shape, x_rank, y_shape, y_rank);
    if (rc) return rc;
    if (bc.total == 0) return LELE_B200_OK;
    where_kernel<<<grid_for(bc.total), 256, 0, ctx->stream>>>(cond, x, y, bc, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_unary(lele_b200_ctx* ctx, int op, const float* x, long long len, float* out) {
    LB_REQUIRE(ctx && (len == 0 || (x && out)), "unary: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(op >= LELE_B200_RELU && op <= LELE_B200_FAST_GELU, "unary: unknown op %d", op);
    if (len == 0) return LELE_B200_OK;
    unary_kernel<<<grid_for(len), 256, 0, ctx->stream>>>(op, x, len, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_clip(lele_b200_ctx* ctx, const float* x, long long len, float lo, float hi, float* out) {
    LB_REQUIRE(ctx && (len == 0 || (x && out)), "clip: NULL argument");
    LB_ENTER(ctx);
    if (len == 0) return LELE_B200_OK;
    clip_kernel<<<grid_for(len), 256, 0, ctx->stream>>>(x, len, lo, hi, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_reduce(lele_b200_ctx* ctx, int kind, const float* x, long long outer, int axis_len, long long inner,
                                float* out) {
    LB_REQUIRE(ctx && x && out && kind >= 0 && kind <= 3 && axis_len > 0, "reduce: bad arguments");
    LB_ENTER(ctx);
    if (outer * inner == 0) return LELE_B200_OK;
    reduce_kernel<<<grid_for(outer * inner), 256, 0, ctx->stream>>>(kind, x, outer, axis_len, inner, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
