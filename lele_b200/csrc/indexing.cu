// indexing.cu -- bit-exact copy / index operators (src/kernels/manipulation.rs, shape.rs,
// conv2d.rs:1051-1506): strided gather (transpose / slice / expand / split), concat, pad,
// gather, gather_elements, tile, topk, argmax, resize_nearest, max_pool2d.
// reshape / flatten / squeeze / unsqueeze stay zero-copy views on the host side (shape.rs).
#include "common.cuh"

namespace {
constexpr int MAXR = 8;
struct StridedArgs { int rank; long long shape[MAXR]; long long stride[MAXR]; long long total; long long offset; };
int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

__global__ void strided_copy_kernel(const float* __restrict__ in, StridedArgs a, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        long long rem = i, off = a.offset;
        for (int d = a.rank - 1; d >= 0; --d) { long long c = rem % a.shape[d]; rem /= a.shape[d]; off += c * a.stride[d]; }
        out[i] = in[off];
    }
}
// Row copy: the innermost output dim is contiguous in the input (stride 1) and a multiple of 4 floats, both sides 16-byte
// aligned (transpose [0,2,1,3], slice / split along an outer axis, expand of outer dims): the row index is decomposed once
// per row and the row moves as float4 -- HBM-rate instead of one div/mod chain per element.
__global__ void __launch_bounds__(256)
strided_rows_kernel(const float* __restrict__ in, StridedArgs a, long long n_rows, int row_v4, float* __restrict__ out) {
    const int lanes_per_row = row_v4 >= 32 ? 32 : (row_v4 >= 16 ? 16 : (row_v4 >= 8 ? 8 : 4));
    const int rows_per_block = 256 / lanes_per_row;
    const int sub = threadIdx.x % lanes_per_row;
    for (long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / lanes_per_row; row < n_rows; row += (long long)gridDim.x * rows_per_block) {
        long long rem = row, off = a.offset;
        for (int d = a.rank - 2; d >= 0; --d) { const long long c = rem % a.shape[d]; rem /= a.shape[d]; off += c * a.stride[d]; }
        const float4* src = reinterpret_cast<const float4*>(in + off);
        float4* dst = reinterpret_cast<float4*>(out + row * (long long)row_v4 * 4);
        for (int q = sub; q < row_v4; q += lanes_per_row) dst[q] = __ldg(src + q);
    }
}
// gather along axis 0-like layouts with a long contiguous inner run: one sub-warp per gathered row, float4 copies
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ data, long long outer, int axis_dim, int inner_v4, const float* __restrict__ indices, long long n_idx,
                   float* __restrict__ out, int* __restrict__ err) {
    const int lanes_per_row = inner_v4 >= 32 ? 32 : (inner_v4 >= 16 ? 16 : (inner_v4 >= 8 ? 8 : 4));
    const int rows_per_block = 256 / lanes_per_row;
    const int sub = threadIdx.x % lanes_per_row;
    const long long n_rows = outer * n_idx;
    for (long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / lanes_per_row; row < n_rows; row += (long long)gridDim.x * rows_per_block) {
        const long long o = row / n_idx, k = row - o * n_idx;
        long long idx = (long long)indices[k];
        if (idx < 0) idx += axis_dim;                             // negative index wrap (manipulation.rs:610)
        if (idx < 0 || idx >= axis_dim) { if (sub == 0) *err = 1; idx = 0; }   // the reference panics (slice bounds check); reported at sync
        const float4* src = reinterpret_cast<const float4*>(data + (o * axis_dim + idx) * (long long)inner_v4 * 4);
        float4* dst = reinterpret_cast<float4*>(out + row * (long long)inner_v4 * 4);
        for (int q = sub; q < inner_v4; q += lanes_per_row) dst[q] = __ldg(src + q);
    }
}
// 32x32 tiled transpose of the two innermost output dims when the innermost INPUT stride is not 1
// but some other output dim has input stride 1: covers [0,2,1,3]/[0,2,3,1]/2-D transposes
// (manipulation.rs:644-1080 fast paths) with coalesced reads and writes.
__global__ void __launch_bounds__(256)
transpose_tiled_kernel(const float* __restrict__ in, long long in_off, long long batch_in_stride, long long rows, long long cols,
                       long long in_row_stride /*stride of out-row index in input*/, long long in_col_stride /*stride of out-col index*/,
                       float* __restrict__ out) {
    // here in_row_stride == 1 (input contiguous along the output ROW index)
    __shared__ float tile[32][33];
    const long long b = blockIdx.z;
    const float* src = in + in_off + b * batch_in_stride;
    float* dst = out + b * rows * cols;
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {   // read: fast index = row (unit stride in input)
        long long r = r0 + tx, c = c0 + j;
        if (r < rows && c < cols) tile[j][tx] = src[r * in_row_stride + c * in_col_stride];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {   // write: fast index = col
        long long r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) dst[r * cols + c] = tile[tx][j];
    }
}

struct ConcatArgs { const float* in[16]; long long axis_len[16]; long long axis_off[16]; int n; };
// row variant: every input is an [outer, axis_len_s * inner] block of the [outer, total_axis * inner] output; when every
// block width and offset is a multiple of 4 floats the blocks move as float4 rows, one warp per (outer index, input) row
__global__ void __launch_bounds__(256)
concat_rows_kernel(ConcatArgs a, long long outer, long long inner, long long total_axis, long long chunk_v4, int chunks_per_row,
                   float* __restrict__ out) {
    // work item = (row = (outer index, input), chunk of the row): a row of a feature-map concat is megabytes long (outer = batch,
    // inner = H * W), so rows are cut into chunks of chunk_v4 float4 and every warp of the grid gets work
    const int lane = threadIdx.x & 31;
    const long long n_items = outer * a.n * chunks_per_row;
    for (long long item = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); item < n_items; item += (long long)gridDim.x * 8) {
        const long long row = item / chunks_per_row; const int c = (int)(item - row * chunks_per_row);
        const long long o = row / a.n; const int s = (int)(row - o * a.n);
        const long long w4 = a.axis_len[s] * inner / 4;
        const long long q0 = (long long)c * chunk_v4, q1 = q0 + chunk_v4 < w4 ? q0 + chunk_v4 : w4;
        const float4* src = reinterpret_cast<const float4*>(a.in[s] + o * a.axis_len[s] * inner);
        float4* dst = reinterpret_cast<float4*>(out + (o * total_axis + a.axis_off[s]) * inner);
        for (long long q = q0 + lane; q < q1; q += 32) dst[q] = __ldg(src + q);
    }
}
__global__ void concat_kernel(ConcatArgs a, long long outer, long long inner, long long total_axis, float* __restrict__ out) {
    const long long total = outer * total_axis * inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long in_i = i % inner, ax = (i / inner) % total_axis, o = i / (inner * total_axis);
        int s = 0;
        while (s + 1 < a.n && ax >= a.axis_off[s + 1]) ++s;
        out[i] = a.in[s][(o * a.axis_len[s] + (ax - a.axis_off[s])) * inner + in_i];
    }
}

struct PadArgs { int rank; long long in_shape[MAXR]; long long out_shape[MAXR]; long long begin[MAXR]; long long total; int mode; float value; };
__global__ void pad_kernel(const float* __restrict__ in, PadArgs a, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        long long rem = i, off = 0, mul = 1;
        bool inside = true;
        for (int d = a.rank - 1; d >= 0; --d) {
            long long c = rem % a.out_shape[d] - a.begin[d]; rem /= a.out_shape[d];
            long long n = a.in_shape[d];
            if (c < 0 || c >= n) {
                if (a.mode == 0) inside = false;
                else if (a.mode == 1) c = c < 0 ? 0 : n - 1;                       // edge  (manipulation.rs:484)
                else { if (n == 1) c = 0; else { long long p = 2 * (n - 1); c = ((c % p) + p) % p; if (c >= n) c = p - c; } }  // reflect (:486)
            }
            off += c * mul; mul *= n;
        }
        out[i] = inside ? in[off] : a.value;
    }
}

__global__ void gather_kernel(const float* __restrict__ data, long long outer, int axis_dim, long long inner,
                              const float* __restrict__ indices, long long n_idx, float* __restrict__ out, int* __restrict__ err) {
    const long long total = outer * n_idx * inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long k = i % inner, q = (i / inner) % n_idx, o = i / (inner * n_idx);
        long long idx = (long long)indices[q];
        if (idx < 0) idx += axis_dim;
        if (idx < 0 || idx >= axis_dim) { *err = 1; idx = 0; }
        out[i] = data[(o * axis_dim + idx) * inner + k];
    }
}
__global__ void gather_elements_kernel(const float* __restrict__ data, const float* __restrict__ indices, long long outer,
                                       int axis_dim, int idx_dim, long long inner, float* __restrict__ out, int* __restrict__ err) {
    const long long total = outer * idx_dim * inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long k = i % inner, o = i / (inner * idx_dim);
        long long idx = (long long)indices[i];
        if (idx < 0) idx += axis_dim;
        if (idx < 0 || idx >= axis_dim) { *err = 2; idx = 0; }
        out[i] = data[(o * axis_dim + idx) * inner + k];
    }
}
struct TileArgs { int rank; long long in_shape[MAXR]; long long out_shape[MAXR]; long long total; };
__global__ void tile_kernel(const float* __restrict__ in, TileArgs a, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        long long rem = i, off = 0, mul = 1;
        for (int d = a.rank - 1; d >= 0; --d) {
            long long c = (rem % a.out_shape[d]) % a.in_shape[d]; rem /= a.out_shape[d];
            off += c * mul; mul *= a.in_shape[d];
        }
        out[i] = in[off];
    }
}

// top-k of each row: one warp per row, k rounds of (max value, lowest index) selection.
// Stable sort_by(partial_cmp) descending => ties keep the lower index first (conv2d.rs:1385-1437).
__global__ void __launch_bounds__(256)
topk_kernel(const float* __restrict__ x, long long outer, int n, int k, float* __restrict__ values, float* __restrict__ indices) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* xr = x + row * n;
    float last_v = INFINITY; int last_i = -1;
    for (int t = 0; t < k; ++t) {
        // candidate ordering: (value desc, index asc); must come strictly after (last_v, last_i)
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            float v = xr[j];
            bool after = (v < last_v) || (v == last_v && j > last_i);
            if (after && (v > bv || (v == bv && j < bi))) { bv = v; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { values[row * k + t] = bv; indices[row * k + t] = (float)bi; }
        last_v = bv; last_i = bi;
    }
}
// top-k of long rows (the detection heads of a vision graph: k = 300 of n = 8400 / 24000): one CTA per row.
//   1. radix select on order-preserving keys (4 x 8-bit passes, shared-memory histogram) finds the k-th largest key `thr` and how many
//      of the elements equal to it belong to the answer;
//   2. everything above thr is collected (any order), the ties at thr in index order (block-wide ballot scan), lowest indices first --
//      exactly what a stable descending sort keeps (conv2d.rs:1385-1437: sort_by(partial_cmp), ties keep the lower index);
//   3. the k candidates are bitonic-sorted in shared memory on (key, ~index) and written out with the ORIGINAL values (a -0.0 stays -0.0).
// O(n) instead of the O(n k / 32) of the warp-per-row kernel above: 24000 -> 300 took 13.5 ms there, ~30 us here.
constexpr int TOPK_NT = 1024, TOPK_MAXK = 2048;
__device__ __forceinline__ unsigned topk_key(float v) { return v != v ? 0u : lb_fkey(v == 0.0f ? 0.0f : v); }   // NaN sorts last, both zeros tie
__global__ void __launch_bounds__(TOPK_NT)
topk_select_kernel(const float* __restrict__ x, int n, int k, float* __restrict__ values, float* __restrict__ indices) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need, s_cnt, s_running;
    __shared__ unsigned warp_cnt[TOPK_NT / 32];
    __shared__ unsigned long long cand[TOPK_MAXK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xr = x + (long long)blockIdx.x * n;
    unsigned prefix = 0u, mask = 0u, need = (unsigned)k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        for (int j = tid; j < n; j += TOPK_NT) {
            const unsigned key = topk_key(xr[j]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned acc = 0u; int b = 255;
            for (; b > 0; --b) { if (acc + hist[b] >= need) break; acc += hist[b]; }
            s_prefix = prefix | ((unsigned)b << shift); s_need = need - acc;
        }
        __syncthreads();
        prefix = s_prefix; need = s_need; mask |= 0xffu << shift;
        __syncthreads();
    }
    const unsigned thr = prefix;                       // `need` of the elements whose key == thr are in the answer
    if (tid == 0) { s_cnt = 0u; s_running = 0u; }
    __syncthreads();
    for (int j = tid; j < n; j += TOPK_NT) {
        const unsigned key = topk_key(xr[j]);
        if (key > thr) { const unsigned slot = atomicAdd(&s_cnt, 1u); cand[slot] = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)j); }
    }
    __syncthreads();
    const unsigned c_gt = s_cnt;                       // == k - need
    for (int base = 0; base < n; base += TOPK_NT) {
        if (s_running >= need) break;                  // (block-uniform: read after the barrier below / above)
        const int j = base + tid;
        const bool eq = j < n && topk_key(xr[j]) == thr;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        unsigned before = s_running;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        const unsigned pos = before + __popc(m & ((1u << lane) - 1u));
        if (eq && pos < need) cand[c_gt + pos] = ((unsigned long long)thr << 32) | (0xffffffffu - (unsigned)j);
        __syncthreads();
        if (tid == 0) { unsigned t = 0u; for (int w = 0; w < TOPK_NT / 32; ++w) t += warp_cnt[w]; s_running += t; }
        __syncthreads();
    }
    int P = 1; while (P < k) P <<= 1;
    for (int i = k + tid; i < P; i += TOPK_NT) cand[i] = 0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)          // bitonic sort, descending
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P; i += TOPK_NT) {
                const int partner = i ^ stride;
                if (partner > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = cand[i], b = cand[partner];
                    if (desc ? a < b : a > b) { cand[i] = b; cand[partner] = a; }
                }
            }
            __syncthreads();
        }
    for (int t = tid; t < k; t += TOPK_NT) {
        const unsigned idx = 0xffffffffu - (unsigned)(cand[t] & 0xffffffffull);
        values[(long long)blockIdx.x * k + t] = xr[idx];
        indices[(long long)blockIdx.x * k + t] = (float)idx;
    }
}
// argmax with LAST-max tie rule (Iterator::max_by)
__global__ void __launch_bounds__(256)
argmax_last_kernel(const float* __restrict__ x, long long outer, int n, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* xr = x + row * n;
    unsigned long long best = 0ull;
    for (int j = lane; j < n; j += 32) {
        unsigned long long key = ((unsigned long long)lb_fkey_argmax(xr[j]) << 32) | (unsigned)j;
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o); best = ok > best ? ok : best; }
    if (lane == 0) out[row] = (int32_t)(best & 0xffffffffu);
}
__global__ void argmax_keys_to_ids_kernel(const unsigned long long* __restrict__ keys, long long n, int32_t* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(keys[i] & 0xffffffffu);
}

__global__ void resize_nearest_kernel(const float* __restrict__ x, long long nc, int h, int w, int oh, int ow, int mode,
                                      float* __restrict__ out) {
    const float hs = __fdiv_rn((float)h, (float)oh), ws = __fdiv_rn((float)w, (float)ow);
    const long long total = nc * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % ow), oy = (int)((i / ow) % oh); long long c = i / ((long long)ow * oh);
        int iy, ix;
        if (mode == 0) {
            iy = (int)fminf(floorf(__fmul_rn((float)oy, hs)), (float)(h - 1));
            ix = (int)fminf(floorf(__fmul_rn((float)ox, ws)), (float)(w - 1));
        } else {
            iy = (int)fminf(fmaxf(roundf(__fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), hs), 0.5f)), 0.0f), (float)(h - 1));
            ix = (int)fminf(fmaxf(roundf(__fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), ws), 0.5f)), 0.0f), (float)(w - 1));
        }
        out[i] = x[(c * h + iy) * w + ix];
    }
}
struct PoolArgs { int h, w, oh, ow, kh, kw, pt, pl, sh, sw, dh, dw; };
// plane-tiled variant: blockIdx.z = (n, c) plane, 32 x 8 outputs per block, 32-bit index math, window rows walked with
// the bounds hoisted (the flat kernel pays three 64-bit divisions per output); max is order-independent -> bit-exact
__global__ void __launch_bounds__(256)
max_pool2d_plane_kernel(const float* __restrict__ x, PoolArgs a, float* __restrict__ out) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= a.ow || oy >= a.oh) return;
    const float* xp = x + (long long)blockIdx.z * a.h * a.w;
    float m = -INFINITY;
    const int ix0 = ox * a.sw - a.pl, iy0 = oy * a.sh - a.pt;
    for (int ky = 0; ky < a.kh; ++ky) {
        const int iy = iy0 + ky * a.dh;
        if (iy < 0 || iy >= a.h) continue;
        const float* row = xp + iy * a.w;
        for (int kx = 0; kx < a.kw; ++kx) {
            const int ix = ix0 + kx * a.dw;
            if (ix >= 0 && ix < a.w) m = fmaxf(m, __ldg(row + ix));
        }
    }
    out[((long long)blockIdx.z * a.oh + oy) * a.ow + ox] = m;
}
// stride 1, dilation 1 (the SPPF pools of the detection models: 5 x 5, pad 2): separable in shared memory.  A block owns a 64 x 16 output
// tile of one plane: the input tile (+ halo, -inf outside the image: padded cells never win, conv2d.rs:1230) is staged once, a row pass
// takes the kw-wide maxima, a column pass the kh-high ones -- kw + kh shared-memory reads per output instead of kw * kh cached global
// loads.  max is order-independent, so the result is bit-identical to the window walk.
constexpr int MP_TW = 64, MP_TH = 16, MP_KMAX = 7;
__global__ void __launch_bounds__(256)
max_pool2d_sep_kernel(const float* __restrict__ x, PoolArgs a, float* __restrict__ out) {
    __shared__ float s_in[(MP_TH + MP_KMAX - 1) * (MP_TW + MP_KMAX - 1)];
    __shared__ float s_h[(MP_TH + MP_KMAX - 1) * MP_TW];
    const int ox0 = blockIdx.x * MP_TW, oy0 = blockIdx.y * MP_TH;
    const int iw_t = MP_TW + a.kw - 1, ih_t = MP_TH + a.kh - 1;
    const float* xp = x + (long long)blockIdx.z * a.h * a.w;
    const int tx = threadIdx.x & (MP_TW - 1), ty0 = threadIdx.x >> 6;        // 64 columns x 4 rows of threads: no index divisions
    for (int ty = ty0; ty < ih_t; ty += 4) {
        const int iy = oy0 - a.pt + ty;
        const bool row_in = iy >= 0 && iy < a.h;
        for (int cx = tx; cx < iw_t; cx += MP_TW) {
            const int ix = ox0 - a.pl + cx;
            s_in[ty * iw_t + cx] = (row_in && ix >= 0 && ix < a.w) ? __ldg(xp + iy * a.w + ix) : -INFINITY;
        }
    }
    __syncthreads();
    for (int ty = ty0; ty < ih_t; ty += 4) {
        const float* r = s_in + ty * iw_t + tx;
        float m = r[0];
        for (int kx = 1; kx < a.kw; ++kx) m = fmaxf(m, r[kx]);
        s_h[ty * MP_TW + tx] = m;
    }
    __syncthreads();
    const int ox = ox0 + tx;
    if (ox < a.ow) {
        for (int ty = ty0; ty < MP_TH; ty += 4) {
            const int oy = oy0 + ty;
            if (oy >= a.oh) break;
            float m = s_h[ty * MP_TW + tx];
            for (int ky = 1; ky < a.kh; ++ky) m = fmaxf(m, s_h[(ty + ky) * MP_TW + tx]);
            out[((long long)blockIdx.z * a.oh + oy) * a.ow + ox] = m;
        }
    }
}
__global__ void max_pool2d_kernel(const float* __restrict__ x, long long nc, PoolArgs a, float* __restrict__ out) {
    const long long total = nc * a.oh * a.ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % a.ow), oy = (int)((i / a.ow) % a.oh); long long c = i / ((long long)a.ow * a.oh);
        float m = -INFINITY;
        for (int ky = 0; ky < a.kh; ++ky) {
            int iy = oy * a.sh + ky * a.dh - a.pt;
            if (iy < 0 || iy >= a.h) continue;
            for (int kx = 0; kx < a.kw; ++kx) {
                int ix = ox * a.sw + kx * a.dw - a.pl;
                if (ix < 0 || ix >= a.w) continue;
                m = fmaxf(m, x[(c * a.h + iy) * a.w + ix]);
            }
        }
        out[i] = m;
    }
}
}  // namespace

int lb_argmax_keys_to_ids(lele_b200_ctx* ctx, const unsigned long long* keys, long long n, int32_t* out) {
    argmax_keys_to_ids_kernel<<<lb_ceil_div(n, 256), 256, 0, ctx->stream>>>(keys, n, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_strided_copy(lele_b200_ctx* ctx, const float* in, long long in_offset, const long long* out_shape,
                                      const long long* in_strides, int rank, float* out) {
    LB_REQUIRE(ctx && in && out && rank >= 0 && rank <= MAXR, "strided_copy: bad arguments");
    LB_ENTER(ctx);
    StridedArgs a; a.rank = rank; a.total = 1; a.offset = in_offset;
    for (int i = 0; i < rank; ++i) { a.shape[i] = out_shape[i]; a.stride[i] = in_strides[i]; a.total *= out_shape[i]; }
    if (a.total == 0) return LELE_B200_OK;
    // tiled path: innermost two output dims (rows, cols) where the input is contiguous along rows
    if (rank >= 2 && in_strides[rank - 2] == 1 && in_strides[rank - 1] != 1 && out_shape[rank - 1] >= 8 && out_shape[rank - 2] >= 8) {
        long long batch = 1; bool regular = true; long long bstride = 0;
        // leading dims must collapse to a single batch stride
        if (rank > 2) {
            for (int i = 0; i < rank - 2; ++i) batch *= out_shape[i];
            bstride = in_strides[rank - 3];
            long long expect = bstride;
            for (int i = rank - 3; i >= 0; --i) { if (out_shape[i] != 1 && in_strides[i] != expect) regular = false; expect *= out_shape[i]; }
        }
        if (regular && batch <= 65535 && lb_ceil_div(out_shape[rank - 2], 32) <= 65535) {
            dim3 grid(lb_ceil_div(out_shape[rank - 1], 32), lb_ceil_div(out_shape[rank - 2], 32), (unsigned)batch);
            transpose_tiled_kernel<<<grid, 256, 0, ctx->stream>>>(in, in_offset, bstride, out_shape[rank - 2], out_shape[rank - 1], 1,
                                                                 in_strides[rank - 1], out);
            LB_LAUNCH_CHECK(ctx);
            return LELE_B200_OK;
        }
    }
    // row-copy path: contiguous innermost run of >= 16 bytes, everything 16-byte aligned
    if (rank >= 1 && in_strides[rank - 1] == 1 && out_shape[rank - 1] % 4 == 0 && out_shape[rank - 1] >= 4 && out_shape[rank - 1] <= (1ll << 30) &&
        ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0 && in_offset % 4 == 0) {
        bool aligned = true;
        for (int i = 0; i < rank - 1; ++i) if (out_shape[i] != 1 && in_strides[i] % 4 != 0) aligned = false;
        if (aligned) {
            const long long n_rows = a.total / out_shape[rank - 1];
            const int row_v4 = (int)(out_shape[rank - 1] / 4);
            const int lanes = row_v4 >= 32 ? 32 : (row_v4 >= 16 ? 16 : (row_v4 >= 8 ? 8 : 4));
            strided_rows_kernel<<<grid_for(n_rows * lanes), 256, 0, ctx->stream>>>(in, a, n_rows, row_v4, out);
            LB_LAUNCH_CHECK(ctx);
            return LELE_B200_OK;
        }
    }
    strided_copy_kernel<<<grid_for(a.total), 256, 0, ctx->stream>>>(in, a, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_concat(lele_b200_ctx* ctx, const float* const* inputs, const long long* axis_lens, int n_inputs,
                                long long outer, long long inner, float* out) {
    LB_REQUIRE(ctx && inputs && axis_lens && out, "concat: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(n_inputs >= 1 && n_inputs <= 16, "concat: %d inputs (supported: 1..16 per call)", n_inputs);
    ConcatArgs a; a.n = 0; long long off = 0;
    for (int i = 0; i < n_inputs; ++i) {
        if (axis_lens[i] == 0) continue;   // empty inputs skipped (manipulation.rs:120)
        a.in[a.n] = inputs[i]; a.axis_len[a.n] = axis_lens[i]; a.axis_off[a.n] = off; off += axis_lens[i]; ++a.n;
    }
    if (a.n == 0 || outer * off * inner == 0) return LELE_B200_OK;
    bool rows_ok = ((((uintptr_t)out) & 15) == 0) && (off * inner) % 4 == 0;
    for (int i = 0; i < a.n; ++i) rows_ok = rows_ok && (a.axis_len[i] * inner) % 4 == 0 && (a.axis_off[i] * inner) % 4 == 0 && ((((uintptr_t)a.in[i]) & 15) == 0) && a.axis_len[i] * inner >= 32;
    if (rows_ok) {
        long long max_w4 = 0;
        for (int i = 0; i < a.n; ++i) max_w4 = a.axis_len[i] * inner / 4 > max_w4 ? a.axis_len[i] * inner / 4 : max_w4;
        const long long chunk_v4 = 1024;                                     // 16 KB per warp visit
        const int chunks_per_row = (int)((max_w4 + chunk_v4 - 1) / chunk_v4);
        concat_rows_kernel<<<grid_for(outer * a.n * chunks_per_row * 32), 256, 0, ctx->stream>>>(a, outer, inner, off, chunk_v4, chunks_per_row, out);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    concat_kernel<<<grid_for(outer * off * inner), 256, 0, ctx->stream>>>(a, outer, inner, off, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_pad(lele_b200_ctx* ctx, const float* in, const long long* shape, int rank, const long long* pads, int mode,
                             float value, float* out) {
    LB_REQUIRE(ctx && in && out && rank >= 1 && rank <= MAXR && mode >= 0 && mode <= 2, "pad: bad arguments");
    LB_ENTER(ctx);
    PadArgs a; a.rank = rank; a.total = 1; a.mode = mode; a.value = value;
    for (int i = 0; i < rank; ++i) {
        long long b = pads[i] < 0 ? 0 : pads[i], e = pads[i + rank] < 0 ? 0 : pads[i + rank];   // negatives clamp to 0 (manipulation.rs:390)
        a.in_shape[i] = shape[i]; a.begin[i] = b; a.out_shape[i] = shape[i] + b + e; a.total *= a.out_shape[i];
    }
    if (a.total == 0) return LELE_B200_OK;
    pad_kernel<<<grid_for(a.total), 256, 0, ctx->stream>>>(in, a, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_gather(lele_b200_ctx* ctx, const float* data, long long outer, int axis_dim, long long inner,
                                const float* indices, long long n_indices, float* out) {
    LB_REQUIRE(ctx && data && indices && out, "gather: NULL argument");
    LB_ENTER(ctx);
    if (outer * n_indices * inner == 0) return LELE_B200_OK;
    if (inner % 4 == 0 && inner >= 16 && inner <= (1ll << 30) && ((((uintptr_t)data) | ((uintptr_t)out)) & 15) == 0) {
        const int inner_v4 = (int)(inner / 4);
        const int lanes = inner_v4 >= 32 ? 32 : (inner_v4 >= 16 ? 16 : (inner_v4 >= 8 ? 8 : 4));
        gather_rows_kernel<<<grid_for(outer * n_indices * lanes), 256, 0, ctx->stream>>>(data, outer, axis_dim, inner_v4, indices, n_indices, out, ctx->dev_err);
        ctx->dev_err_armed = true;
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    gather_kernel<<<grid_for(outer * n_indices * inner), 256, 0, ctx->stream>>>(data, outer, axis_dim, inner, indices, n_indices, out, ctx->dev_err);
    ctx->dev_err_armed = true;
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_gather_elements(lele_b200_ctx* ctx, const float* data, const float* indices, long long outer, int axis_dim,
                                         int idx_dim, long long inner, float* out) {
    LB_REQUIRE(ctx && data && indices && out, "gather_elements: NULL argument");
    LB_ENTER(ctx);
    if (outer * idx_dim * inner == 0) return LELE_B200_OK;
    gather_elements_kernel<<<grid_for(outer * idx_dim * inner), 256, 0, ctx->stream>>>(data, indices, outer, axis_dim, idx_dim, inner, out, ctx->dev_err);
    ctx->dev_err_armed = true;
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_tile(lele_b200_ctx* ctx, const float* in, const long long* shape, const long long* repeats, int rank,
                              float* out) {
    LB_REQUIRE(ctx && in && out && rank >= 1 && rank <= MAXR, "tile: bad arguments");
    LB_ENTER(ctx);
    TileArgs a; a.rank = rank; a.total = 1;
    for (int i = 0; i < rank; ++i) { a.in_shape[i] = shape[i]; a.out_shape[i] = shape[i] * repeats[i]; a.total *= a.out_shape[i]; }
    if (a.total == 0) return LELE_B200_OK;
    tile_kernel<<<grid_for(a.total), 256, 0, ctx->stream>>>(in, a, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_topk(lele_b200_ctx* ctx, const float* x, long long outer, int n, int k, float* values, float* indices) {
    LB_REQUIRE(ctx && x && values && indices && n > 0, "topk: bad arguments");
    LB_ENTER(ctx);
    LB_REQUIRE(k >= 0 && k <= n, "topk: k=%d must be in [0, n=%d] (caller applies k=min(k,last), conv2d.rs:1396)", k, n);
    if (outer == 0 || k == 0) return LELE_B200_OK;
    if (k <= TOPK_MAXK && (long long)n * k >= (1ll << 16) && outer < (1ll << 31)) {      // long rows: select + sort (one CTA per row)
        topk_select_kernel<<<(unsigned)outer, TOPK_NT, 0, ctx->stream>>>(x, n, k, values, indices);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    topk_kernel<<<lb_ceil_div(outer, 8), 256, 0, ctx->stream>>>(x, outer, n, k, values, indices);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
// Greedy decode filter (examples/sensevoice/src/tokenizer.rs:37-75): per clip, keep the frame ids that are neither blank
// (id 0) nor flagged in skip_mask (the "<|...|>" special tokens), in frame order.  One warp per clip, ballot compaction.
__global__ void __launch_bounds__(32)
greedy_filter_kernel(const int32_t* __restrict__ ids, int t, const uint8_t* __restrict__ skip_mask, int vocab, int32_t* __restrict__ out_ids,
                     int32_t* __restrict__ out_len) {
    const int clip = blockIdx.x, lane = threadIdx.x;
    const int32_t* in = ids + (long long)clip * t;
    int32_t* out = out_ids + (long long)clip * t;
    int n = 0;
    for (int t0 = 0; t0 < t; t0 += 32) {
        const int i = t0 + lane;
        const int id = i < t ? in[i] : 0;
        const bool keep = i < t && id != 0 && id > 0 && id < vocab && !(skip_mask && skip_mask[id]);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) out[n + __popc(m & ((1u << lane) - 1u))] = id;
        n += __popc(m);
    }
    for (int i = n + lane; i < t; i += 32) out[i] = -1;      // padding
    if (lane == 0) out_len[clip] = n;
}

extern "C" int lele_b200_greedy_filter(lele_b200_ctx* ctx, const int32_t* ids, int n_clips, int t, const uint8_t* skip_mask, int vocab,
                                       int32_t* out_ids, int32_t* out_len) {
    LB_REQUIRE(ctx && ids && out_ids && out_len && n_clips >= 0 && t >= 0 && vocab > 0, "greedy_filter: bad arguments");
    LB_ENTER(ctx);
    if (n_clips == 0) return LELE_B200_OK;
    greedy_filter_kernel<<<n_clips, 32, 0, ctx->stream>>>(ids, t, skip_mask, vocab, out_ids, out_len);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

extern "C" int lele_b200_argmax_last(lele_b200_ctx* ctx, const float* x, long long outer, int n, int32_t* out) {
    LB_REQUIRE(ctx && x && out && n > 0, "argmax_last: bad arguments");
    LB_ENTER(ctx);
    if (outer == 0) return LELE_B200_OK;
    argmax_last_kernel<<<lb_ceil_div(outer, 8), 256, 0, ctx->stream>>>(x, outer, n, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_resize_nearest(lele_b200_ctx* ctx, const float* x, int nb, int c, int h, int w, int oh, int ow, int mode,
                                        float* out) {
    LB_REQUIRE(ctx && x && out, "resize_nearest: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(oh > 0 && ow > 0, "Resize: output dimensions must be positive, got out_h=%d out_w=%d (conv2d.rs:1321)", oh, ow);
    long long total = (long long)nb * c * oh * ow;
    if (total == 0) return LELE_B200_OK;
    resize_nearest_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(x, (long long)nb * c, h, w, oh, ow, mode, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
extern "C" int lele_b200_max_pool2d(lele_b200_ctx* ctx, const float* x, int nb, int c, int h, int w, int kh, int kw, const int* pads,
                                    const int* strides, const int* dils, int ceil_mode, float* out) {
    LB_REQUIRE(ctx && x && out && pads && strides && dils, "max_pool2d: NULL argument");
    LB_ENTER(ctx);
    PoolArgs a;
    a.h = h; a.w = w; a.kh = kh; a.kw = kw; a.pt = pads[0]; a.pl = pads[1]; a.sh = strides[0]; a.sw = strides[1]; a.dh = dils[0]; a.dw = dils[1];
    int nh = h + pads[0] + pads[2] - a.dh * (kh - 1) - 1, nw = w + pads[1] + pads[3] - a.dw * (kw - 1) - 1;
    a.oh = (ceil_mode ? (nh + a.sh - 1) / a.sh : nh / a.sh) + 1;
    a.ow = (ceil_mode ? (nw + a.sw - 1) / a.sw : nw / a.sw) + 1;
    long long total = (long long)nb * c * a.oh * a.ow;
    if (total <= 0) return LELE_B200_OK;
    if (a.sh == 1 && a.sw == 1 && a.dh == 1 && a.dw == 1 && kh <= MP_KMAX && kw <= MP_KMAX && a.ow >= 32 && a.oh >= 8 &&
        (long long)nb * c <= 65535 && lb_ceil_div(a.oh, MP_TH) <= 65535 && lb_env_flag("LELE_B200_POOL_SEP", 1)) {
        max_pool2d_sep_kernel<<<dim3(lb_ceil_div(a.ow, MP_TW), lb_ceil_div(a.oh, MP_TH), nb * c), 256, 0, ctx->stream>>>(x, a, out);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    if ((long long)nb * c <= 65535 && lb_ceil_div(a.oh, 8) <= 65535) {
        max_pool2d_plane_kernel<<<dim3(lb_ceil_div(a.ow, 32), lb_ceil_div(a.oh, 8), nb * c), 256, 0, ctx->stream>>>(x, a, out);
        LB_LAUNCH_CHECK(ctx);
        return LELE_B200_OK;
    }
    max_pool2d_kernel<<<grid_for(total), 256, 0, ctx->stream>>>(x, (long long)nb * c, a, out);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
