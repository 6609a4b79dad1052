// gemm_tf32_tc.cu -- f32 GEMM on the 5th-gen tensor cores at f32 accuracy (3xTF32).
//
// The dense f32 products of the operator set -- matmul / matmul_fused_add / gemm (src/kernels/gemm.rs:112,223,433),
// conv2d as GEMM over im2col (src/kernels/conv2d.rs:311-880), conv_transpose's col = W^T X (conv2d.rs:3069) -- go
// through faer's f32 micro-kernels in the reference.  Here:
//
//     C[b][M,N] (+)= alpha * A[b][M,K] . B[b][N,K]^T          ("NT": both operands K-major, f32 in HBM)
//
// on tcgen05.mma kind::tf32 with the 3xTF32 split  a.b ~= hi(a)hi(b) + hi(a)lo(b) + lo(a)hi(b)  (~2^-21 relative, the
// f32 class; plain TF32 would be 2^-11 and miss the 1e-4 bar).  The tensor core reads the top 19 bits of a 32-bit
// operand, so hi(x) is the raw f32 tile TMA lands in shared memory; lo(x) = x - trunc(x) is computed ON CHIP by the
// converter warps into a second tile with the same swizzled layout (attn_tc.cu uses the same scheme), so each operand
// crosses HBM / L2 exactly once.
//
// CTA: 128 x 128 output tile, K in 32-float chunks (one 128-byte swizzle row), 3-stage ring of
// {A raw, A lo, B raw, B lo} (64 KB per stage); warp 0 TMA producer, warp 1 MMA issuer (hi.hi right after the TMA
// lands, the cross terms after the converters), warp 2 TMEM allocator, warps 4-11 converters; warps 4-7 then run
// the epilogue (TMEM -> registers -> swizzled staging -> coalesced 128-byte row stores with the optional
// read-modify-write / per-row bias / activation).  Operand tails in M, N, K are zero-filled by TMA.
//
// B-operand producers (template GATHER): 0 = TMA over a K-major matrix; 1 = the converter warps GATHER an N-major
// matrix (element (n,k) = ptr[k*ldk + n]: matmul's [K,N] operand, a 1x1 convolution's input, conv_transpose's input)
// straight from global memory into the swizzled tile and write hi and lo in one go -- no transposing pre-pass;
// 2 = implicit-GEMM convolution: element (p, r) of the im2col matrix is computed from the NCHW input on the fly
// (lanes = consecutive output positions -> coalesced along x), so the im2col matrix of the reference
// (conv2d.rs:892, thread-local COL_BUF) never exists and a whole batch is one launch.
#include "gemm_tf32_tc.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 128, KC = 32;          // KC floats = 128 B
constexpr int TILE_A = BM * 128, TILE_B = BN * 128; // 16 KB each
constexpr int STAGE = 2 * TILE_A + 2 * TILE_B;      // 64 KB
constexpr int NSTAGE = 3;
constexpr int EPI_STAGING = 4 * 4096;               // one 32x32 f32 tile per epilogue warp
constexpr int SMEM_MAIN = NSTAGE * STAGE + EPI_STAGING;
constexpr int SMEM_BYTES = SMEM_MAIN + 1024 + 256;
constexpr int NCONV_WARPS = 8;
constexpr int NUM_THREADS = (4 + NCONV_WARPS) * 32; // 384
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
        if (clock64() - t0 > 4000000000ll) { printf("lele_b200 gemm_tf32_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
// K-major, SWIZZLE_128B operand descriptor (see gemm_i8_tc.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10), K-major, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts_v4f(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void lo_convert_16B(uint32_t src, uint32_t dst) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
    sts_v4f(dst, __fsub_rn(v.x, tf32_hi(v.x)), __fsub_rn(v.y, tf32_hi(v.y)), __fsub_rn(v.z, tf32_hi(v.z)), __fsub_rn(v.w, tf32_hi(v.w)));
}

struct GemmArgs {
    LbGatherB gb;
    int M, N, K, n_kchunks;
    int a_bcast, b_bcast;      // operand shared by every batch slice (its tensor map has one slice)
    float* C; long long ldc, bsc;
    float alpha;
    int pre_mode;              // 1: C already holds the pre-fill (bias / beta*C): C = C + alpha*acc
    const float* bias_row;     // [M] or NULL: v = v + bias_row[row]          (conv: per output channel)
    int act;                   // 0 none, 1 ReLU, 2 SiLU (SIMD body / scalar tail split at simd_end, avx/math.rs:344-470)
    int simd_end;
};

template <int GATHER>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32x3_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + SMEM_MAIN);
    uint64_t* full = bars;                  // [NSTAGE] TMA -> converters, MMA (hi.hi)
    uint64_t* conv = bars + NSTAGE;         // [NSTAGE] converters -> MMA (cross terms)
    uint64_t* empty = bars + 2 * NSTAGE;    // [NSTAGE] MMA -> TMA
    uint64_t* acc_full = bars + 3 * NSTAGE; // MMA -> epilogue
    uint32_t* tmem_base_smem = (uint32_t*)(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y, b = blockIdx.z;
    const int NK = args.n_kchunks;

    if (warp == 0 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], NCONV_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            for (int kc = 0; kc < NK; ++kc) {
                const int s = kc % NSTAGE; const uint32_t ph = (uint32_t)(kc / NSTAGE) & 1u;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;
                mbar_expect_tx(&full[s], GATHER ? TILE_A : TILE_A + TILE_B);
                tma_load_3d(st, &map_a, &full[s], kc * KC, m_blk * BM, args.a_bcast ? 0 : b);
                if (!GATHER) tma_load_3d(st + 2 * TILE_A, &map_b, &full[s], kc * KC, n_blk * BN, args.b_bcast ? 0 : b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kc = 0; kc < NK; ++kc) {
                const int s = kc % NSTAGE; const uint32_t ph = (uint32_t)(kc / NSTAGE) & 1u;
                const uint32_t base = smem_u32(smem + s * STAGE);
                const uint64_t ah = make_smem_desc(base), al = make_smem_desc(base + TILE_A);
                const uint64_t bh = make_smem_desc(base + 2 * TILE_A), bl = make_smem_desc(base + 2 * TILE_A + TILE_B);
                if (!GATHER) {
                    mbar_wait(&full[s], ph);                       // raw tiles landed: hi.hi can start
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, ah + (uint64_t)(k * 2), bh + (uint64_t)(k * 2), (kc == 0 && k == 0) ? 0u : 1u);
                }
                mbar_wait(&conv[s], ph);                           // lo tiles (GATHER: the whole B tile) written
                tc_fence_after();
                if (GATHER) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, ah + (uint64_t)(k * 2), bh + (uint64_t)(k * 2), (kc == 0 && k == 0) ? 0u : 1u);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_tf32(tmem_base, ah + (uint64_t)(k * 2), bl + (uint64_t)(k * 2), 1u);
                    umma_tf32(tmem_base, al + (uint64_t)(k * 2), bh + (uint64_t)(k * 2), 1u);
                }
                umma_commit(&empty[s]);
                if (kc == NK - 1) umma_commit(acc_full);
            }
        }
    } else if (warp >= 4) {
        // ---- converters: lo tiles of A and B, 2048 float4 per k-chunk over 256 threads ----
        const int t256 = threadIdx.x - 128;
        // GATHER: this thread owns B-tile row (output position) p_l and 16 of the chunk's 32 k's
        const int p_l = t256 & 127, kg = t256 >> 7;
        const int p = n_blk * BN + p_l;
        const bool p_ok = p < args.N;
        const LbGatherB& gb = args.gb;
        const float* gsrc = gb.ptr + (long long)b * gb.bs;
        int iy0 = 0, ix0 = 0;
        if (GATHER == 2) { const int oy = p / gb.ow, ox = p - oy * gb.ow; iy0 = oy * gb.sh - gb.pt; ix0 = ox * gb.sw - gb.pl; }
        // The gathered operand is software-pipelined one k-chunk ahead: the 16 global loads of chunk kc+1 are issued before chunk kc
        // is converted and handed to the MMA warp, so a load round trip (the im2col reads miss L1 on every new input row) is hidden
        // behind one chunk of conversion + barrier traffic instead of being paid per chunk by the 8 converter warps.
        auto gather = [&](int kc, float (&g)[16]) {
            // the loads do not depend on the TMA: (the stage itself is written only after full[s], which the producer arms only
            // once the MMAs that read the stage's previous contents have retired)
            const int k0 = kc * KC + kg * 16;
            if (GATHER == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) g[j] = (p_ok && k0 + j < args.K) ? __ldg(gsrc + (long long)(k0 + j) * gb.ldk + p) : 0.0f;
            } else {
                const int khw = gb.kh * gb.kw;
                const int kmax = gb.k_valid > 0 ? gb.k_valid : args.K;      // (K padded to a multiple of 4 for the weight's TMA pitch: the pad reads nothing)
                int c = k0 / khw; int rem = k0 - c * khw; int ky = rem / gb.kw; int kx = rem - ky * gb.kw;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int iy = iy0 + ky * gb.dh, ix = ix0 + kx * gb.dw;
                    const bool ok = p_ok && k0 + j < kmax && iy >= 0 && iy < gb.h && ix >= 0 && ix < gb.w;
                    g[j] = ok ? __ldg(gsrc + ((long long)c * gb.h + iy) * gb.w + ix) : 0.0f;
                    if (++kx == gb.kw) { kx = 0; if (++ky == gb.kh) { ky = 0; ++c; } }
                }
            }
        };
        float g[16], gn[16];
        if (GATHER) gather(0, g);
        for (int kc = 0; kc < NK; ++kc) {
            const int s = kc % NSTAGE; const uint32_t ph = (uint32_t)(kc / NSTAGE) & 1u;
            const uint32_t st = smem_u32(smem + s * STAGE);
            if (GATHER && kc + 1 < NK) gather(kc + 1, gn);
            mbar_wait(&full[s], ph);
#pragma unroll
            for (int i = t256; i < TILE_A / 16; i += 256) lo_convert_16B(st + (uint32_t)i * 16u, st + (uint32_t)TILE_A + (uint32_t)i * 16u);
            if (!GATHER) {
#pragma unroll
                for (int i = t256; i < TILE_B / 16; i += 256)
                    lo_convert_16B(st + 2u * TILE_A + (uint32_t)i * 16u, st + 2u * TILE_A + (uint32_t)TILE_B + (uint32_t)i * 16u);
            } else {
                const uint32_t rowb = st + 2u * TILE_A + (uint32_t)p_l * 128u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = (uint32_t)(((kg * 4 + q) ^ (p_l & 7)) << 4);
                    const float a0 = g[q * 4 + 0], a1 = g[q * 4 + 1], a2 = g[q * 4 + 2], a3 = g[q * 4 + 3];
                    sts_v4f(rowb + off, a0, a1, a2, a3);
                    sts_v4f(rowb + (uint32_t)TILE_B + off, __fsub_rn(a0, tf32_hi(a0)), __fsub_rn(a1, tf32_hi(a1)), __fsub_rn(a2, tf32_hi(a2)), __fsub_rn(a3, tf32_hi(a3)));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv[s]);
            if (GATHER) {
#pragma unroll
                for (int j = 0; j < 16; ++j) g[j] = gn[j];
            }
        }
        if (warp < 8) {
            // ---- epilogue: warp = TMEM lane quadrant; thread = row, 4 chunks of 32 columns ----
            const int quad = warp & 3;
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
            const uint32_t stg = smem_u32(smem + NSTAGE * STAGE) + (uint32_t)quad * 4096u;
            const int row0 = m_blk * BM + quad * 32;
            const int nrows = min(32, args.M - row0);
            float* Cb = args.C + (long long)b * args.bsc;
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const float alpha = args.alpha;
            const float brow = (args.bias_row && row0 + lane < args.M) ? __ldg(args.bias_row + row0 + lane) : 0.0f;
            for (int chunk = 0; chunk < BN / 32; ++chunk) {
                const int col0 = n_blk * BN + chunk * 32;
                uint32_t v[32];
                tmem_ld32(trow + (uint32_t)(chunk * 32), v);
                if (col0 >= args.N || nrows <= 0) continue;          // warp-uniform
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = __uint_as_float(v[q * 4 + e]);
                        if (alpha != 1.0f) t = __fmul_rn(alpha, t);
                        if (args.bias_row) t = __fadd_rn(t, brow);
                        o[e] = t;
                    }
                    sts_v4f(stg + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
                }
                __syncwarp();
                // transposed read: lane = column -> 128-byte coalesced row stores
                const int col = col0 + lane;
                if (col < args.N) {
                    const uint32_t rd = stg + (uint32_t)(lane & 3) * 4u;
                    const uint32_t ch = (uint32_t)(lane >> 2);
                    float* outp = Cb + (long long)row0 * args.ldc + col;
#pragma unroll 4
                    for (int rr = 0; rr < nrows; ++rr) {
                        float t = lds_f32(rd + (uint32_t)rr * 128u + ((ch ^ (uint32_t)(rr & 7)) << 4));
                        const long long o = (long long)rr * args.ldc;
                        if (args.pre_mode) t = __fadd_rn(outp[o], t);
                        if (args.act == 1) t = fmaxf(t, 0.0f);
                        else if (args.act == 2) t = col < args.simd_end ? __fmul_rn(t, lb_sigmoid_simd(t)) : __fdiv_rn(t, __fadd_rn(1.0f, expf(-t)));
                        outp[o] = t;
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// [rows, cols] (ld_in) -> [cols, rows] (ld_out), batched; 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ in, long long ld_in, long long bs_in, float* __restrict__ out, long long ld_out, long long bs_out,
                     int rows, int cols) {
    __shared__ float tile[32][33];
    in += (long long)blockIdx.z * bs_in; out += (long long)blockIdx.z * bs_out;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (c < cols && r < rows) out[(long long)c * ld_out + r] = tile[tx][j];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// f32 operand [batch][rows][K] (row pitch ld, batch stride bs; bs == 0 broadcasts one matrix) -> dims (K, rows, batch)
int make_operand_map(lele_b200_ctx* ctx, CUtensorMap* map, const float* ptr, long long rows, long long K, long long ld, long long bs, int batch, int box_rows) {
    const unsigned long long key[10] = {0x74663332ull, (unsigned long long)(uintptr_t)ptr, (unsigned long long)rows, (unsigned long long)K, (unsigned long long)ld,
                                        (unsigned long long)bs, (unsigned long long)batch, (unsigned long long)box_rows};
    if (lb_tmap_lookup(ctx, key, map)) return LELE_B200_OK;
    EncodeTiledFn fn = encode_fn();
    if (!fn) { lb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return LELE_B200_ERR_CUDA; }
    const bool bcast = (bs == 0 || batch == 1);
    cuuint64_t d[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(bcast ? 1 : batch)};
    cuuint64_t st[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(bcast ? (cuuint64_t)rows * ld * 4 : (cuuint64_t)bs * 4)};
    cuuint32_t bx[3] = {(cuuint32_t)KC, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lb_set_error("cuTensorMapEncodeTiled(f32 operand) failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return LELE_B200_ERR_CUDA; }
    lb_tmap_store(ctx, key, map);
    return LELE_B200_OK;
}
}  // namespace

// Whether the tensor-core path can take operands with these pitches / pointers (TMA: 16-byte aligned base and pitch).
bool lb_gemm_tc_supported(const float* A, long long lda, long long bsa, const float* B, long long ldb, long long bsb, int m, int n, int k) {
    if (getenv("LELE_B200_FORCE_SIMT") || getenv("LELE_B200_SGEMM_SIMT")) return false;
    if (m < 1 || n < 1 || k < 4) return false;
    if ((((uintptr_t)A | (uintptr_t)B) & 15) != 0) return false;
    if (lda % 4 || ldb % 4 || bsa % 4 || bsb % 4) return false;
    return true;
}

static int launch_tc(lele_b200_ctx* ctx, const float* A, long long lda, long long bsa, const float* B, long long ldb, long long bsb, const LbGatherB* gb,
                     float* C, long long ldc, long long bsc, int batch, int m, int n, int k, const LbGemmTcEpilogue& ep) {
    LB_REQUIRE(batch >= 1 && batch <= 65535 && lb_ceil_div(m, BM) <= 65535, "gemm_tf32x3_nt: batch / m too large for one launch");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_operand_map(ctx, &ma, A, m, k, lda, bsa, batch, BM))) return rc;
    if (gb) mb = ma;
    else if ((rc = make_operand_map(ctx, &mb, B, n, k, ldb, bsb, batch, BN))) return rc;
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    if (gb) a.gb = *gb;
    a.M = m; a.N = n; a.K = k; a.n_kchunks = lb_ceil_div(k, KC);
    a.a_bcast = (bsa == 0 || batch == 1) ? 1 : 0; a.b_bcast = (bsb == 0 || batch == 1) ? 1 : 0;
    a.C = C; a.ldc = ldc; a.bsc = bsc; a.alpha = ep.alpha; a.pre_mode = ep.pre_mode; a.bias_row = ep.bias_row; a.act = ep.act; a.simd_end = ep.simd_end;
    {
        int rc_a = lb_func_smem(ctx, (const void*)gemm_tf32x3_nt_kernel<0>, SMEM_BYTES);
        if (!rc_a) rc_a = lb_func_smem(ctx, (const void*)gemm_tf32x3_nt_kernel<1>, SMEM_BYTES);
        if (!rc_a) rc_a = lb_func_smem(ctx, (const void*)gemm_tf32x3_nt_kernel<2>, SMEM_BYTES);
        if (rc_a) return rc_a;
    }
    dim3 grid(lb_ceil_div(n, BN), lb_ceil_div(m, BM), batch);
    const int mode = gb ? gb->mode : 0;
    if (mode == 0) gemm_tf32x3_nt_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, a);
    else if (mode == 1) gemm_tf32x3_nt_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, a);
    else gemm_tf32x3_nt_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, a);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

// C[b][m,n] (row pitch ldc, batch stride bsc) = epilogue(A[b][m,k] . B[b][n,k]^T); bsa / bsb == 0 broadcast one matrix
int lb_gemm_tf32x3_nt(lele_b200_ctx* ctx, const float* A, long long lda, long long bsa, const float* B, long long ldb, long long bsb, float* C,
                      long long ldc, long long bsc, int batch, int m, int n, int k, const LbGemmTcEpilogue& ep) {
    LB_REQUIRE(lb_gemm_tc_supported(A, lda, bsa, B, ldb, bsb, m, n, k), "gemm_tf32x3_nt: operands not TMA-addressable");
    return launch_tc(ctx, A, lda, bsa, B, ldb, bsb, nullptr, C, ldc, bsc, batch, m, n, k, ep);
}

// Same product with the B operand gathered by the kernel (LbGatherB): N-major matrix or implicit im2col
int lb_gemm_tf32x3_gather(lele_b200_ctx* ctx, const float* A, long long lda, long long bsa, const LbGatherB& gb, float* C, long long ldc, long long bsc,
                          int batch, int m, int n, int k, const LbGemmTcEpilogue& ep) {
    LB_REQUIRE(lb_gemm_tc_supported(A, lda, bsa, A, lda, bsa, m, n, k) && (gb.mode == 1 || gb.mode == 2) && gb.ptr, "gemm_tf32x3_gather: bad operands");
    return launch_tc(ctx, A, lda, bsa, nullptr, 0, gb.bs, &gb, C, ldc, bsc, batch, m, n, k, ep);
}

int lb_transpose_f32(lele_b200_ctx* ctx, const float* in, long long ld_in, long long bs_in, float* out, long long ld_out, long long bs_out, int batch,
                     int rows, int cols) {
    if (batch == 0 || rows == 0 || cols == 0) return LELE_B200_OK;
    LB_REQUIRE(batch <= 65535 && lb_ceil_div(rows, 32) <= 65535, "transpose: batch / rows too large for one launch");
    transpose_f32_kernel<<<dim3(lb_ceil_div(cols, 32), lb_ceil_div(rows, 32), batch), 256, 0, ctx->stream>>>(in, ld_in, bs_in, out, ld_out, bs_out, rows, cols);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
