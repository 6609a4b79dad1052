// conv.cu -- Conv1d / Conv2d / ConvTranspose (src/kernels/conv1d.rs:837-1342,
// conv2d.rs:107-880, conv2d.rs:2952-3128).
//  * dense conv2d (group 1): 1x1/s1/p0 -> GEMM W[OC,IC] x X[IC,HW]; otherwise im2col into
//    scratch + GEMM per image (the reference's own decomposition, conv2d.rs:600-690), then a
//    fused bias + {none, ReLU, SiLU} pass (avx/math.rs:344-470 body/tail split).
//  * depthwise / grouped / conv1d: direct CUDA-core kernels (HBM-bound).
//  * conv_transpose: gather form (each output pixel sums its contributing taps) - no
//    zero-fill + scatter-add round trip.
#include "gemm_tf32_tc.cuh"

int lb_sgemm_strided(lele_b200_ctx* ctx, const float* A, long long rsa, long long csa, long long bsa, const float* B,
                     long long rsb, long long csb, long long bsb, float* C, int batch, int m, int k, int n, float alpha,
                     int pre_mode);

namespace {
int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

__global__ void conv1d_direct_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                     int nb, int ic, int l, int oc, int k, int group, int pad_l, int stride, int dil, int relu,
                                     int ol, float* __restrict__ out) {
    const int icg = ic / group, ocg = oc / group;
    const long long total = (long long)nb * oc * ol;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int t = (int)(i % ol), o = (int)((i / ol) % oc), b = (int)(i / ((long long)ol * oc));
        int g = o / ocg;
        float s = 0.0f;
        for (int c = 0; c < icg; ++c) {
            const float* xr = x + ((long long)b * ic + (long long)g * icg + c) * l;
            const float* wr = w + ((long long)o * icg + c) * k;
            for (int kk = 0; kk < k; ++kk) {
                int pos = t * stride + kk * dil - pad_l;
                if (pos >= 0 && pos < l) s = fmaf(wr[kk], xr[pos], s);
            }
        }
        if (bias) s = __fadd_rn(s, bias[o]);
        if (relu) s = fmaxf(s, 0.0f);
        out[i] = s;
    }
}

struct Conv2dGeom { int ic, h, w, oc, kh, kw, group, pt, pl, sh, sw, dh, dw, oh, ow; };

__global__ void conv2d_direct_kernel(const float* __restrict__ x, const float* __restrict__ w, int nb, Conv2dGeom g,
                                     float* __restrict__ out) {
    const int icg = g.ic / g.group, ocg = g.oc / g.group;
    const long long total = (long long)nb * g.oc * g.oh * g.ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % g.ow), oy = (int)((i / g.ow) % g.oh);
        int o = (int)((i / ((long long)g.ow * g.oh)) % g.oc), b = (int)(i / ((long long)g.ow * g.oh * g.oc));
        int grp = o / ocg;
        float s = 0.0f;
        for (int c = 0; c < icg; ++c)
            for (int ky = 0; ky < g.kh; ++ky) {
                int iy = oy * g.sh + ky * g.dh - g.pt;
                if (iy < 0 || iy >= g.h) continue;
                for (int kx = 0; kx < g.kw; ++kx) {
                    int ix = ox * g.sw + kx * g.dw - g.pl;
                    if (ix < 0 || ix >= g.w) continue;
                    s = fmaf(x[(((long long)b * g.ic + (long long)grp * icg + c) * g.h + iy) * g.w + ix],
                             w[(((long long)o * icg + c) * g.kh + ky) * g.kw + kx], s);
                }
            }
        out[i] = s;
    }
}
// im2col for one image: col[(c*kh+ky)*kw+kx][oy*ow+ox]   (conv2d.rs:892)
__global__ void im2col_kernel(const float* __restrict__ x, Conv2dGeom g, float* __restrict__ col) {
    const long long hw = (long long)g.oh * g.ow, total = (long long)g.ic * g.kh * g.kw * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i % hw); long long r = i / hw;
        int kx = (int)(r % g.kw), ky = (int)((r / g.kw) % g.kh), c = (int)(r / ((long long)g.kw * g.kh));
        int oy = p / g.ow, ox = p % g.ow;
        int iy = oy * g.sh + ky * g.dh - g.pt, ix = ox * g.sw + kx * g.dw - g.pl;
        col[i] = (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) ? x[((long long)c * g.h + iy) * g.w + ix] : 0.0f;
    }
}
// conv_transpose, step 2: out[n,o,oy,ox] = sum over the taps (ky, kx ascending, as the reference's scatter order) of
// colm[n][(o*kh+ky)*kw+kx][iy*wd+ix] + bias[o]; colm = W^T X came from the tensor-core GEMM
__global__ void col2im_gather_kernel(const float* __restrict__ colm, const float* __restrict__ bias, int nb, int h, int wd, int oc, int kh, int kw,
                                     int pt, int pl, int sh, int sw, int dh, int dw, int oh, int ow, float* __restrict__ out) {
    const long long total = (long long)nb * oc * oh * ow, hw = (long long)h * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
        const int o = (int)((i / ((long long)ow * oh)) % oc), n = (int)(i / ((long long)ow * oh * oc));
        float s = 0.0f;
        for (int ky = 0; ky < kh; ++ky) {
            const int ty = oy + pt - ky * dh;
            if (ty < 0 || ty % sh) continue;
            const int iy = ty / sh;
            if (iy >= h) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int tx = ox + pl - kx * dw;
                if (tx < 0 || tx % sw) continue;
                const int ix = tx / sw;
                if (ix >= wd) continue;
                s = __fadd_rn(s, colm[(((long long)n * oc + o) * kh * kw + (long long)ky * kw + kx) * hw + (long long)iy * wd + ix]);
            }
        }
        if (bias) s = __fadd_rn(s, bias[o]);
        out[i] = s;
    }
}
// bias + activation over [planes, hw]; SIMD body = first hw/8*8 of each plane (avx/math.rs:344-470)
__global__ void bias_act_kernel(float* __restrict__ data, long long planes, int oc, long long hw, const float* __restrict__ bias,
                                int act) {
    const long long total = planes * hw, simd_end = (hw / 8) * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long j = i % hw; int ch = (int)((i / hw) % oc);
        float v = data[i];
        if (bias) v = __fadd_rn(v, bias[ch]);
        if (act == 1) v = fmaxf(v, 0.0f);
        else if (act == 2) v = j < simd_end ? __fmul_rn(v, lb_sigmoid_simd(v)) : __fdiv_rn(v, __fadd_rn(1.0f, expf(-v)));
        data[i] = v;
    }
}

__global__ void conv_transpose_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                      int nb, int ic, int h, int wd, int oc, int kh, int kw, int pt, int pl, int sh, int sw,
                                      int dh, int dw, int oh, int ow, float* __restrict__ out) {
    const long long total = (long long)nb * oc * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
        int o = (int)((i / ((long long)ow * oh)) % oc), n = (int)(i / ((long long)ow * oh * oc));
        float s = 0.0f;
        for (int ky = 0; ky < kh; ++ky) {
            int ty = oy + pt - ky * dh;
            if (ty < 0 || ty % sh) continue;
            int iy = ty / sh;
            if (iy >= h) continue;
            for (int kx = 0; kx < kw; ++kx) {
                int tx = ox + pl - kx * dw;
                if (tx < 0 || tx % sw) continue;
                int ix = tx / sw;
                if (ix >= wd) continue;
                float t = 0.0f;   // sequential ic sum per tap, as the reference's col = W^T X (conv2d.rs:3069-3087)
                for (int c = 0; c < ic; ++c)
                    t = fmaf(w[(((long long)c * oc + o) * kh + ky) * kw + kx], x[(((long long)n * ic + c) * h + iy) * wd + ix], t);
                s = __fadd_rn(s, t);
            }
        }
        if (bias) s = __fadd_rn(s, bias[o]);
        out[i] = s;
    }
}
}  // namespace

bool lb_conv2d_tc_pixel_supported(const float* w, long long w_pitch, int oc, long long kdim);
int lb_conv_transpose_tc_scatter(lele_b200_ctx* ctx, const float* x, const float* wt, const float* bias, int nb, int ic, int h, int wd, int oc,
                                 int kh, int kw, int sh, int sw, float* out);
int lb_conv2d_tc_pixel(lele_b200_ctx* ctx, const float* x, const float* w, long long w_pitch, int k_valid, const float* bias, int nb, int ic, int h,
                       int wd, int oc, int kh, int kw, int pt, int pl, int sh, int sw, int dh, int dw, int oh, int ow, int act, float* out);
namespace {
__global__ void pad_rows_kernel(const float* __restrict__ w, int rows, int k, int kpad, float* __restrict__ out) {
    const long long total = (long long)rows * kpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % kpad); const long long r = i / kpad;
        out[i] = c < k ? w[r * k + c] : 0.0f;
    }
}
}  // namespace

extern "C" int lele_b200_conv1d(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias, int nb, int ic, int l,
                                int oc, int k, int group, int pad_l, int pad_r, int stride, int dilation, int relu, float* out) {
    LB_REQUIRE(ctx && x && w && out, "conv1d: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(group >= 1 && ic % group == 0 && oc % group == 0 && stride >= 1 && dilation >= 1 && k >= 1, "conv1d: bad attributes");
    int ol = (l + pad_l + pad_r - dilation * (k - 1) - 1) / stride + 1;
    LB_REQUIRE(ol >= 0, "conv1d: negative output length");
    long long total = (long long)nb * oc * ol;
    if (total == 0) return LELE_B200_OK;
    conv1d_direct_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(x, w, bias, nb, ic, l, oc, k, group, pad_l, stride, dilation, relu, ol, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_conv2d(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias, int nb, int ic, int h, int wd,
                                int oc, int kh, int kw, int group, const int* pads, const int* strides, const int* dils, int act,
                                float* out) {
    LB_REQUIRE(ctx && x && w && out && pads && strides && dils, "conv2d: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(group >= 1 && ic % group == 0 && oc % group == 0, "Conv2d: channels not divisible by group (conv2d.rs:196-205)");
    Conv2dGeom g;
    g.ic = ic; g.h = h; g.w = wd; g.oc = oc; g.kh = kh; g.kw = kw; g.group = group; g.pt = pads[0]; g.pl = pads[1];
    g.sh = strides[0]; g.sw = strides[1]; g.dh = dils[0]; g.dw = dils[1];
    g.oh = (h + pads[0] + pads[2] - g.dh * (kh - 1) - 1) / g.sh + 1;
    g.ow = (wd + pads[1] + pads[3] - g.dw * (kw - 1) - 1) / g.sw + 1;
    LB_REQUIRE(g.oh > 0 && g.ow > 0, "conv2d: non-positive output size");
    const long long hw = (long long)g.oh * g.ow;
    if (nb == 0) return LELE_B200_OK;
    int rc;
    bool fused_epilogue = false;
    LbGemmTcEpilogue tep; tep.bias_row = bias; tep.act = act; tep.simd_end = (int)((hw / 8) * 8);
    const long long kdim = (long long)ic * kh * kw;
    const bool tc_ok = group == 1 && kdim % 4 == 0 && oc >= 16 && hw >= 64 && kdim >= 16 && hw * kdim * oc >= (1ll << 22) &&
                       lb_gemm_tc_supported(w, kdim, 0, w, kdim, 0, oc, (int)hw, (int)kdim);
    // Pixel-major implicit GEMM (conv_tc.cu): pixels = the tensor core's M (full 128-row tiles whatever OC is), OC = its N; persistent,
    // epilogue overlapped with the next tile.  Takes every group-1 layer with 8 <= OC <= 256 and enough work, 1x1 included.
    if (group == 1 && oc >= 8 && oc <= 256 && kdim >= 8 && (long long)nb * hw >= 128 && (long long)nb * hw * kdim * oc >= (1ll << 22) &&
        g.dh >= 1 && g.dw >= 1) {
        const float* wp = w; long long pitch = kdim;
        if (kdim % 4 != 0) {          // K = IC*kh*kw not a multiple of 4 (a 3-channel stem: 27): rows re-pitched (zeros) so TMA can address them
            pitch = (kdim + 3) / 4 * 4;
            void* sc;
            if ((rc = lb_scratch(ctx, sizeof(float) * (size_t)oc * pitch + 256, &sc))) return rc;
            pad_rows_kernel<<<grid_for((long long)oc * pitch), 256, 0, ctx->stream>>>(w, oc, (int)kdim, (int)pitch, (float*)sc);
            LB_LAUNCH_CHECK(ctx);
            wp = (const float*)sc;
        }
        if (lb_conv2d_tc_pixel_supported(wp, pitch, oc, kdim))
            return lb_conv2d_tc_pixel(ctx, x, wp, pitch, (int)kdim, bias, nb, ic, h, wd, oc, kh, kw, g.pt, g.pl, g.sh, g.sw, g.dh, g.dw, g.oh, g.ow, act, out);
    }
    if (group == 1 && kh == 1 && kw == 1 && g.sh == 1 && g.sw == 1 && pads[0] == 0 && pads[1] == 0 && pads[2] == 0 && pads[3] == 0) {
        // 1x1: out[b] = W[OC,IC] x X[b][IC,HW]
        if (tc_ok) {   // tensor cores: the kernel gathers X[b] [IC, HW] as its N-major B operand; bias + activation fused in the epilogue
            LbGatherB gb; memset(&gb, 0, sizeof(gb));
            gb.mode = 1; gb.ptr = x; gb.ldk = hw; gb.bs = (long long)ic * hw;
            if ((rc = lb_gemm_tf32x3_gather(ctx, w, ic, 0, gb, out, hw, (long long)oc * hw, nb, oc, (int)hw, ic, tep))) return rc;
            fused_epilogue = true;
        } else {
            rc = lb_sgemm_strided(ctx, w, ic, 1, 0, x, hw, 1, (long long)ic * hw, out, nb, oc, ic, (int)hw, 1.0f, 0);
            if (rc) return rc;
        }
    } else if (!tc_ok && group == 1 && kdim % 4 != 0 && oc >= 16 && hw >= 64 && kdim >= 16 && hw * kdim * oc >= (1ll << 22)) {
        // K = IC*kh*kw not a multiple of 4 (a 3-channel stem: 27): the weight rows are re-pitched to the next multiple of 4 (zeros) so TMA can
        // address them; the im2col producers treat k >= K as padding (k_valid) -- the layer then runs on the tensor cores like the others
        const int kpad = (int)((kdim + 3) / 4 * 4);
        void* sc;
        if ((rc = lb_scratch(ctx, sizeof(float) * (size_t)oc * kpad + 256, &sc))) return rc;   // (scratch2 may hold conv_integer's staged operands)
        float* wp = (float*)sc;
        pad_rows_kernel<<<grid_for((long long)oc * kpad), 256, 0, ctx->stream>>>(w, oc, (int)kdim, kpad, wp);
        LB_LAUNCH_CHECK(ctx);
        if (lb_gemm_tc_supported(wp, kpad, 0, wp, kpad, 0, oc, (int)hw, kpad)) {
            LbGatherB gb; memset(&gb, 0, sizeof(gb));
            gb.mode = 2; gb.ptr = x; gb.bs = (long long)ic * h * wd; gb.h = h; gb.w = wd; gb.kh = kh; gb.kw = kw; gb.pt = g.pt; gb.pl = g.pl;
            gb.sh = g.sh; gb.sw = g.sw; gb.dh = g.dh; gb.dw = g.dw; gb.ow = g.ow; gb.k_valid = (int)kdim;
            if ((rc = lb_gemm_tf32x3_gather(ctx, wp, kpad, 0, gb, out, hw, (long long)oc * hw, nb, oc, (int)hw, kpad, tep))) return rc;
            fused_epilogue = true;
        } else {
            conv2d_direct_kernel<<<grid_for((long long)nb * oc * hw), 256, 0, ctx->stream>>>(x, w, nb, g, out);
            LB_LAUNCH_CHECK(ctx);
        }
    } else if (tc_ok) {
        // implicit GEMM on the tensor cores: out[b] = W[OC,K] . im2col(x[b])[HW,K]^T with the im2col element computed on the
        // fly by the kernel's operand producers (no col buffer), the whole batch in one launch, bias + activation fused
        LbGatherB gb; memset(&gb, 0, sizeof(gb));
        gb.mode = 2; gb.ptr = x; gb.bs = (long long)ic * h * wd; gb.h = h; gb.w = wd; gb.kh = kh; gb.kw = kw; gb.pt = g.pt; gb.pl = g.pl;
        gb.sh = g.sh; gb.sw = g.sw; gb.dh = g.dh; gb.dw = g.dw; gb.ow = g.ow;
        if ((rc = lb_gemm_tf32x3_gather(ctx, w, kdim, 0, gb, out, hw, (long long)oc * hw, nb, oc, (int)hw, (int)kdim, tep))) return rc;
        fused_epilogue = true;
    } else if (group == 1 && kdim >= 32) {
        void* col;
        if ((rc = lb_scratch(ctx, sizeof(float) * (size_t)kdim * hw, &col))) return rc;
        for (int b = 0; b < nb; ++b) {
            im2col_kernel<<<grid_for(kdim * hw), 256, 0, ctx->stream>>>(x + (long long)b * ic * h * wd, g, (float*)col);
            LB_LAUNCH_CHECK(ctx);
            rc = lb_sgemm_strided(ctx, w, kdim, 1, 0, (const float*)col, hw, 1, 0, out + (long long)b * oc * hw, 1, oc, (int)kdim, (int)hw, 1.0f, 0);
            if (rc) return rc;
        }
    } else {
        conv2d_direct_kernel<<<grid_for((long long)nb * oc * hw), 256, 0, ctx->stream>>>(x, w, nb, g, out);
        LB_LAUNCH_CHECK(ctx);
    }
    if (fused_epilogue) return LELE_B200_OK;
    if (bias || act) {
        bias_act_kernel<<<grid_for((long long)nb * oc * hw), 256, 0, ctx->stream>>>(out, (long long)nb * oc, oc, hw, bias, act);
        LB_LAUNCH_CHECK(ctx);
    }
    return LELE_B200_OK;
}

// ConvInteger (conv2d.rs:2216-2420 -> conv2d_with_zero_points :1507-2000): an f32 convolution of (x - x_zp) with (w - w_zp) whose im2col
// pads with RAW zeros, so a padded position contributes (0 - x_zp) (conv2d.rs:2025).  Here: one staging kernel writes the zero-padded,
// shifted input and the shifted weight, then the convolution runs un-padded on the same paths as conv2d (tcgen05 implicit GEMM when
// the shape qualifies).
namespace {
__global__ void conv_integer_stage_kernel(const float* __restrict__ x, int nbc, int h, int w, int pt, int pl, int hp, int wp, float x_zp,
                                          float* __restrict__ xs, const float* __restrict__ wgt, long long w_len, float w_zp, float* __restrict__ ws) {
    const long long total = (long long)nbc * hp * wp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total + w_len; i += (long long)gridDim.x * blockDim.x) {
        if (i >= total) { ws[i - total] = __fsub_rn(wgt[i - total], w_zp); continue; }
        const int c = (int)(i % wp), r = (int)((i / wp) % hp); const long long p = i / ((long long)wp * hp);
        const int sr = r - pt, sc = c - pl;
        const float v = (sr >= 0 && sr < h && sc >= 0 && sc < w) ? x[(p * h + sr) * w + sc] : 0.0f;
        xs[i] = __fsub_rn(v, x_zp);
    }
}
}  // namespace

extern "C" int lele_b200_conv_integer(lele_b200_ctx* ctx, const float* x, const float* w, float x_zp, float w_zp, int nb, int ic, int h,
                                      int wd, int oc, int kh, int kw, int group, const int* pads, const int* strides, const int* dils,
                                      float* out) {
    LB_REQUIRE(ctx && x && w && out && pads && strides && dils, "conv_integer: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(group >= 1 && ic % group == 0 && oc % group == 0, "ConvInteger: channels not divisible by group (conv2d.rs:196-205)");
    LB_REQUIRE(pads[0] >= 0 && pads[1] >= 0 && pads[2] >= 0 && pads[3] >= 0, "conv_integer: negative pads");
    if (nb == 0) return LELE_B200_OK;
    const int hp = h + pads[0] + pads[2], wp = wd + pads[1] + pads[3];
    const long long n_x = (long long)nb * ic * hp * wp, n_w = (long long)oc * (ic / group) * kh * kw;
    void* sc; int rc;
    if ((rc = lb_scratch2(ctx, sizeof(float) * (size_t)(n_x + n_w) + 512, &sc))) return rc;
    float* xs = (float*)sc; float* ws = xs + (n_x + 63) / 64 * 64;
    conv_integer_stage_kernel<<<grid_for(n_x + n_w), 256, 0, ctx->stream>>>(x, nb * ic, h, wd, pads[0], pads[1], hp, wp, x_zp, xs, w, n_w, w_zp, ws);
    LB_LAUNCH_CHECK(ctx);
    const int zero_pads[4] = {0, 0, 0, 0};
    return lele_b200_conv2d(ctx, xs, ws, nullptr, nb, ic, hp, wp, oc, kh, kw, group, zero_pads, strides, dils, 0, out);
}

extern "C" int lele_b200_conv_transpose(lele_b200_ctx* ctx, const float* x, const float* w, const float* bias, int nb, int ic, int h,
                                        int wd, int oc, int kh, int kw, const int* pads, const int* strides, const int* dils,
                                        float* out) {
    LB_REQUIRE(ctx && x && w && out && pads && strides && dils, "conv_transpose: NULL argument");
    LB_ENTER(ctx);
    int oh = (h - 1) * strides[0] - (pads[0] + pads[2]) + dils[0] * (kh - 1) + 1;
    int ow = (wd - 1) * strides[1] - (pads[1] + pads[3]) + dils[1] * (kw - 1) + 1;
    LB_REQUIRE(oh > 0 && ow > 0, "conv_transpose: output dimensions must be positive, got out_h=%d out_w=%d (conv2d.rs:3025)", oh, ow);
    long long total = (long long)nb * oc * oh * ow;
    if (total == 0) return LELE_B200_OK;
    const long long hw = (long long)h * wd; const int mk = oc * kh * kw;
    if (kh == strides[0] && kw == strides[1] && pads[0] == 0 && pads[1] == 0 && pads[2] == 0 && pads[3] == 0 && dils[0] == 1 && dils[1] == 1 &&
        ic % 4 == 0 && ic >= 8 && mk >= 8 && mk <= 256 && (long long)nb * hw >= 128 && (long long)nb * hw * ic * mk >= (1ll << 22)) {
        // kernel == stride: non-overlapping taps, every output written once -> one pixel-major GEMM with the scatter in its epilogue
        void* sc; int rc;
        if ((rc = lb_scratch(ctx, sizeof(float) * (size_t)mk * ic + 256, &sc))) return rc;
        float* wt = (float*)sc;
        if ((rc = lb_transpose_f32(ctx, w, mk, 0, wt, ic, 0, 1, ic, mk))) return rc;       // W [ic, oc*kh*kw] -> [oc*kh*kw, ic]
        if (lb_conv2d_tc_pixel_supported(wt, ic, mk, ic))
            return lb_conv_transpose_tc_scatter(ctx, x, wt, bias, nb, ic, h, wd, oc, kh, kw, strides[0], strides[1], out);
    }
    if (ic % 4 == 0 && ic >= 16 && mk >= 16 && hw >= 64 && hw * ic * mk >= (1ll << 22) && lb_gemm_tc_supported(x, ic, 0, x, ic, 0, mk, (int)hw, ic)) {
        // the reference's own structure (conv2d.rs:3069-3128): col = W^T X as a GEMM -- here on the tensor cores -- then the
        // scatter-add, done as a gather per output element in the same tap order
        void* sc; int rc;
        const size_t n_wt = (size_t)mk * ic, n_col = (size_t)nb * mk * hw;
        if ((rc = lb_scratch(ctx, sizeof(float) * (n_wt + n_col) + 512, &sc))) return rc;
        float* wt = (float*)sc; float* colm = wt + (n_wt + 63) / 64 * 64;
        if ((rc = lb_transpose_f32(ctx, w, mk, 0, wt, ic, 0, 1, ic, mk))) return rc;       // W [ic, oc*kh*kw] -> [oc*kh*kw, ic] (K-major A operand)
        LbGatherB gb; memset(&gb, 0, sizeof(gb));
        gb.mode = 1; gb.ptr = x; gb.ldk = hw; gb.bs = (long long)ic * hw;                    // X[b] [ic, hw] gathered as the N-major B operand
        LbGemmTcEpilogue tep;
        if ((rc = lb_gemm_tf32x3_gather(ctx, wt, ic, 0, gb, colm, hw, (long long)mk * hw, nb, mk, (int)hw, ic, tep))) return rc;
        col2im_gather_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(colm, bias, nb, h, wd, oc, kh, kw, pads[0], pads[1], strides[0], strides[1], dils[0], dils[1], oh, ow, out);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    conv_transpose_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(x, w, bias, nb, ic, h, wd, oc, kh, kw, pads[0], pads[1], strides[0],
                                                                   strides[1], dils[0], dils[1], oh, ow, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
