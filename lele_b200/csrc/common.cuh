// common.cuh -- shared device/host helpers for the lele_b200 CUDA back-end (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_map>

#include "../../include/lele_b200.h"

// ----------------------------------------------------------------------------
// context: one per (device, stream).  Mirrors what the Rust shim would hold per model
// instance (SURVEY 8b "Threading": re-entrant per (device, stream) handle).
// ----------------------------------------------------------------------------
struct lele_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    // grow-only scratch (the analogue of lele's thread_local SCRATCH_A/RS/CS, avx/quantization.rs:90-95)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // second grow-only buffer for entries that stage operands and then call another entry which uses `scratch` itself (conv_integer)
    void* scratch2 = nullptr;
    size_t scratch2_bytes = 0;
    // arena: host Vec base pointer -> device mirror (src/tensor.rs static buffer arena mapped to HBM)
    struct Mirror { void* dptr; size_t bytes; };
    std::unordered_map<const void*, Mirror> arena;
    unsigned long long launches = 0;   // kernels launched through this ctx (bench "gpu_launches")
    // cached device constants (FFT twiddles per n, Hann window, sparse mel bank ...)
    std::unordered_map<std::string, void*> tables;
    // encoded TMA descriptors keyed by (kind, pointer, geometry): workspace / weight pointers are stable
    // across forwards, so cuTensorMapEncodeTiled runs once per tensor instead of once per launch.  Each
    // entry stores its full key tuple (a 64-bit hash collision must not hand back another tensor's
    // descriptor); the cache is bounded (LB_TMAP_CACHE_MAX, oldest-half eviction) because the operator
    // API sees fresh pointers on every call, and entries die with the allocation they describe
    // (lele_b200_free / arena_release / scratch growth call lb_tmap_forget_range).
    struct TmapEntry { unsigned long long key[10]; unsigned long long stamp; unsigned char blob[128]; };
    std::unordered_map<unsigned long long, TmapEntry> tmaps;
    unsigned long long tmap_clock = 0;
    // device-side precondition failures (the reference panics, e.g. a gather index out of range, manipulation.rs:589): kernels
    // clamp the access and raise this word; lele_b200_sync() reads it back (only when a checking kernel ran since the last sync)
    int* dev_err = nullptr;
    bool dev_err_armed = false;
    // fork / join event of this context's stream (lele_b200_stream_fork / _join) and the launch counter at capture begin
    cudaEvent_t ev = nullptr;
    unsigned long long capture_l0 = 0;
    bool capturing = false;
    // allocations released while the stream was being captured (a free must join the stream first, which a capture forbids):
    // handed to cudaFree at the next point where joining is legal
    std::vector<void*> deferred_free;
};
#define LB_TMAP_CACHE_MAX 4096
// key = {kind tag, pointer, dim/stride/box words...}; returns true and fills `blob128` on a verified hit
bool lb_tmap_lookup(lele_b200_ctx* ctx, const unsigned long long (&key)[10], void* blob128);
void lb_tmap_store(lele_b200_ctx* ctx, const unsigned long long (&key)[10], const void* blob128);
void lb_tmap_forget_range(lele_b200_ctx* ctx, const void* base, size_t bytes);   // bytes == 0: every entry whose pointer == base
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), raised (never lowered) per (device, kernel)
int lb_func_smem(lele_b200_ctx* ctx, const void* func, size_t bytes);
// makes the context's device current for the calling thread (every entry point starts with it)
int lb_enter(lele_b200_ctx* ctx);
bool lb_stream_capturing(lele_b200_ctx* ctx);
void lb_flush_deferred_free(lele_b200_ctx* ctx);   // call only where the stream was just joined
#define LB_ENTER(ctx) do { int _rc_enter = lb_enter(ctx); if (_rc_enter) return _rc_enter; } while (0)
static inline unsigned long long lb_hash_mix(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
}
// returns the cached device copy of a host table, uploading it on first use
int lb_table(lele_b200_ctx* ctx, const std::string& key, const void* host, size_t bytes, void** out);

void lb_set_error(const char* fmt, ...);
int lb_scratch(lele_b200_ctx* ctx, size_t bytes, void** out);
int lb_scratch2(lele_b200_ctx* ctx, size_t bytes, void** out);

#define LB_CHECK_CUDA(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            lb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cudaGetLastError(); /* reported here: must not resurface at the next launch check */ \
            return LELE_B200_ERR_CUDA;                                                       \
        }                                                                                    \
    } while (0)

#define LB_REQUIRE(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            lb_set_error(__VA_ARGS__);                                                       \
            return LELE_B200_ERR_ARG;                                                        \
        }                                                                                    \
    } while (0)

#define LB_LAUNCH_CHECK(ctx)                                                                 \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            lb_set_error("%s:%d: launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LELE_B200_ERR_CUDA;                                                       \
        }                                                                                    \
    } while (0)

static inline int lb_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (PDL): the hot per-layer kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, call lb_pdl_launch_dependents() on entry and
// lb_pdl_wait() after their data-independent prologue (barrier init, TMEM allocation, descriptor
// prefetch, weight-side loads).  The next kernel's CTAs then become resident as this kernel's CTAs retire
// and run their prologue under its tail instead of after a full drain + launch gap.  Every such kernel
// executes the wait before touching activations, so completion order stays transitive along the stream.
// LELE_B200_PDL=0 launches them plainly (the device-side instructions are no-ops then).
bool lb_pdl_enabled();
bool lb_env_flag(const char* name, int dflt);   // "0" -> false, anything else -> true, unset -> dflt
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t lb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                        Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster_x > 1) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
    if (lb_pdl_enabled()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void lb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void lb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ----------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------
#ifdef __CUDACC__
// clamp(rint(y), 0, 255) with round-half-even, NaN -> 0 -- the quantisers' inner operation -- without a float -> integer conversion (F2I
// issues on the quarter-rate special-function pipe): clamp first (min / max on the ALU pipe; clamping before or after the rounding is the
// same, 0 and 255 are integers), then add 1.5 * 2^23: at that magnitude one ulp is 1, so the addition itself rounds to the nearest-even
// integer and the sum's low mantissa byte is the result.  Bit-identical to the conversion (parity tests); measured on the LayerNorm +
// quantiser kernel's quantising phase: 4.3 k -> 3.8 k cycles (the rest of that phase is the burst of u8 stores).  LB_Q8_CVT: the conversion.
__device__ __forceinline__ unsigned lb_q8(float y) {
#ifdef LB_Q8_CVT
    return min(__float2uint_rn(y), 255u);
#else
    return __float_as_uint(__fadd_rn(fminf(fmaxf(y, 0.0f), 255.0f), 12582912.0f)) & 0xffu;
#endif
}
// four of them packed little-endian (the byte selects take the low byte of each sum: no mask needed)
__device__ __forceinline__ unsigned lb_q8x4(float y0, float y1, float y2, float y3, int& sum) {
#ifdef LB_Q8_CVT
    const unsigned q0 = lb_q8(y0), q1 = lb_q8(y1), q2 = lb_q8(y2), q3 = lb_q8(y3);
    sum += (int)(q0 + q1 + q2 + q3);
    return q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
#else
    const unsigned b0 = __float_as_uint(__fadd_rn(fminf(fmaxf(y0, 0.0f), 255.0f), 12582912.0f)), b1 = __float_as_uint(__fadd_rn(fminf(fmaxf(y1, 0.0f), 255.0f), 12582912.0f));
    const unsigned b2 = __float_as_uint(__fadd_rn(fminf(fmaxf(y2, 0.0f), 255.0f), 12582912.0f)), b3 = __float_as_uint(__fadd_rn(fminf(fmaxf(y3, 0.0f), 255.0f), 12582912.0f));
    const unsigned pk = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
    sum = (int)__dp4a(pk, 0x01010101u, (unsigned)sum);
    return pk;
#endif
}
__device__ __forceinline__ float lb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float lb_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float lb_warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int lb_warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Row reduction in the summation order of lele's AVX2 row kernels (avx/norm.rs:28-82 LayerNorm,
// :169-205 softmax): four 8-lane accumulators over 32-element blocks (= one partial per warp
// lane, element j -> lane j%32), merged (s0+s1)+(s2+s3), then the remaining 8-blocks into the
// merged vector, the horizontal add ((v0+v4)+(v2+v6))+((v1+v5)+(v3+v7)), then the n%8 scalar
// tail.  A warp reproduces that order exactly, so LayerNorm statistics / softmax sums -- and the
// dynamic-quantisation min/max derived from them -- are bit-identical to the x86 reference order.
// step_vec(acc, j) / step_tail(acc, j) fold element j into acc.  Result broadcast to all lanes.
// second half of the reduction: p holds this lane's partial over the full 32-element blocks
template <class SV, class ST>
__device__ __forceinline__ float lb_avx_order_finish(float p, int n, int lane, SV step_vec, ST step_tail) {
    const int n32 = n & ~31, n8 = n & ~7;
    p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 8));
    p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 16));
    for (int j = n32 + (lane & 7); j < n8; j += 8) p = step_vec(p, j);
    p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 4));
    p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 2));
    p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 1));
    p = __shfl_sync(0xffffffffu, p, 0);
    for (int j = n8; j < n; ++j) p = step_tail(p, j);
    return p;
}
template <class SV, class ST>
__device__ __forceinline__ float lb_avx_order_reduce(int n, int lane, SV step_vec, ST step_tail) {
    float p = 0.0f;
    const int n32 = n & ~31;
    for (int j = lane; j < n32; j += 32) p = step_vec(p, j);
    return lb_avx_order_finish(p, n, lane, step_vec, step_tail);
}

// order-preserving float <-> uint key, so per-clip min/max can use integer atomics
__device__ __forceinline__ unsigned lb_fkey(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// arg-max keys: the reference compares with partial_cmp, for which -0.0 == +0.0 (ties -> last index), so both zeros share a key
__device__ __forceinline__ unsigned lb_fkey_argmax(float f) { return lb_fkey(f == 0.0f ? 0.0f : f); }
__device__ __forceinline__ float lb_fkey_inv(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
#define LB_KEY_MIN_INIT 0xffffffffu /* identity for atomicMin over keys */
#define LB_KEY_MAX_INIT 0x00000000u /* identity for atomicMax over keys */
// Per-slice min/max keys are sharded over LB_MM_SLOTS address pairs (slot = blockIdx.x % slots) so the
// ~10^4 atomics a launch issues do not serialise on 2 addresses per clip; consumers reduce the slots.
// Layout: keys[slice][LB_MM_SLOTS][2] (min key, max key).
#define LB_MM_SLOTS 8
__device__ __forceinline__ void lb_mm_update(unsigned* keys, long long slice, float mn, float mx) {
    unsigned* k = keys + ((size_t)slice * LB_MM_SLOTS + (blockIdx.x & (LB_MM_SLOTS - 1))) * 2;
    atomicMin(k, lb_fkey(mn));
    atomicMax(k + 1, lb_fkey(mx));
}
__device__ __forceinline__ void lb_mm_update_keys(unsigned* keys, long long slice, unsigned kmin, unsigned kmax) {
    unsigned* k = keys + ((size_t)slice * LB_MM_SLOTS + (blockIdx.x & (LB_MM_SLOTS - 1))) * 2;
    atomicMin(k, kmin);
    atomicMax(k + 1, kmax);
}
__device__ __forceinline__ void lb_mm_read(const unsigned* keys, long long slice, float& mn, float& mx) {
    const unsigned* k = keys + (size_t)slice * LB_MM_SLOTS * 2;
    unsigned a = LB_KEY_MIN_INIT, b = LB_KEY_MAX_INIT;
#pragma unroll
    for (int s = 0; s < LB_MM_SLOTS; ++s) { a = min(a, k[2 * s]); b = max(b, k[2 * s + 1]); }
    mn = lb_fkey_inv(a); mx = lb_fkey_inv(b);
}

// Polynomial expf used by lele's x86 SIMD bodies (avx/math.rs:11-66): clamp, rint(x*log2e),
// two-step ln2 reduction, degree-7 FMA Horner, 2^n through the exponent bits.
__device__ __forceinline__ float lb_cephes_expf(float x) {
    x = fmaxf(x, -87.33654f);
    x = fminf(x, 88.72284f);
    float fx = rintf(__fmul_rn(x, 1.44269504088896341f));
    x = __fmaf_rn(-fx, 0.693359375f, x);
    x = __fmaf_rn(-fx, -2.12194440e-4f, x);
    float y = __fmaf_rn(0.000198712018891638893f, x, 0.00139712726883569741f);
    y = __fmaf_rn(y, x, 0.00833345670066840443f);
    y = __fmaf_rn(y, x, 0.0416657844442129135f);
    y = __fmaf_rn(y, x, 0.166666671633720398f);
    y = __fmaf_rn(y, x, 0.5f);
    y = __fmaf_rn(y, x, 1.0f);
    y = __fmaf_rn(y, x, 1.0f);
    int e = ((int)fx + 127) << 23;
    return __fmul_rn(y, __int_as_float(e));
}
// The scalar tails of lele's SIMD row kernels call libm's expf (avx/norm.rs:196, glibc: correctly rounded in all but astronomically
// rare cases); CUDA's expf carries up to 2 ulp.  exp in double, rounded once to float, reproduces the libm result, so the row
// tails -- 7 of the 271 columns of a SenseVoice score row -- do not seed 1-ulp differences that a later quantiser turns into
// flipped codes: with it the CUDA-core-attention path is bit-identical to the CPU restatement through all 70 layers at T' = 271
// (tests/test_gpu_sensevoice.py::test_full_size_simt_attention_is_bit_identical).
__device__ __forceinline__ float lb_libm_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float lb_sigmoid_simd(float x) {  // avx/math.rs:69
    // 1 / d with a correctly rounded reciprocal: by IEEE-754 the same bits as the division _mm256_div_ps(1, d) performs, at a
    // third of the instructions of the general __fdiv_rn sequence (this sits in the conv / LSTM epilogues, once per element)
    return __frcp_rn(__fadd_rn(1.0f, lb_cephes_expf(-x)));
}
__device__ __forceinline__ float lb_tanh_simd(float x) {  // avx/math.rs:81-97
    float e = lb_cephes_expf(__fmul_rn(-x, 2.0f));
    float r = fabsf(__fdiv_rn(__fsub_rn(1.0f, e), __fadd_rn(1.0f, e)));
    return copysignf(r, x);
}
__device__ __forceinline__ float lb_erf_simd(float x) {  // avx/math.rs:113-150
    float ax = fabsf(x);
    float t = __fdiv_rn(1.0f, __fmaf_rn(0.3275911f, ax, 1.0f));
    float poly = __fmaf_rn(1.061405429f, t, -1.453152027f);
    poly = __fmaf_rn(poly, t, 1.421413741f);
    poly = __fmaf_rn(poly, t, -0.284496736f);
    poly = __fmaf_rn(poly, t, 0.254829592f);
    float ev = lb_cephes_expf(-__fmul_rn(ax, ax));
    float r = __fmaf_rn(-__fmul_rn(poly, t), ev, 1.0f);
    return __uint_as_float(__float_as_uint(r) | (__float_as_uint(x) & 0x80000000u));
}
#endif  // __CUDACC__
