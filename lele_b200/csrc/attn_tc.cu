// attn_tc.cu -- self-attention of the SANM layers on the 5th-gen tensor cores, f32-accurate.
//
// Replaces the reference's  mul(q, d_k^-1/2) -> matmul(q, k^T) -> softmax -> matmul(p, v)
// (src/kernels/gemm.rs:112 via faer, src/kernels/norm.rs:8; [4,271,128]x[4,128,271] per clip,
// src/bin/wasm_bench.rs:938-1023) with ONE fused kernel per (clip, head, 128-query tile):
//
//   S = Q K^T          tcgen05.mma kind::tf32, 3xTF32 split (hi*hi + hi*lo + lo*hi: ~2^-21 relative,
//                      i.e. f32-grade; plain TF32 would be 2^-11 and miss the 1e-4 bar)
//   P = exp(S - max)   8 softmax warps read S from TMEM (two threads per query row, even / odd 32-key
//                      chunks), same polynomial exp / libm tail split as the CPU reference
//   O = P V            P is written to shared memory as the tf32 hi/lo A-operand, 32 keys at a time,
//                      while the MMA warp consumes the previous chunk; O accumulates in TMEM
//   out = O / sum      + fused per-clip min/max (feeds the next dynamic quantiser)
//
// The [h,T,T] score / probability tensors never touch HBM (the reference materialises both).
// A small pre-pass (attn_split_kernel) writes the tf32 hi/lo operand copies: q*scale, k ([M,d]) and
// V^T ([B,H,dk,Tp], keys contiguous so it is a K-major B operand).
//
// Geometry: dk = 128, T <= 288 keys (SenseVoice: 271).  Other shapes use the CUDA-core path.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int AQ = 128;               // query rows per CTA (UMMA M)
constexpr int NH = 144;               // keys per S half (UMMA N), two halves = 288
constexpr int KC = 32;                // floats per 128-byte swizzled row chunk
constexpr int DK = 128;
constexpr int TILE_Q = AQ * 128;      // 16 KB  [128 rows][128 B]
constexpr int TILE_K = 2 * NH * 128;  // 36 KB  [288 rows][128 B]
constexpr int STAGE_A = 2 * TILE_Q + 2 * TILE_K;   // 104 KB
constexpr int NSTAGE_A = 2;
constexpr int TILE_P = AQ * 128;      // 16 KB
constexpr int TILE_V = DK * 128;      // 16 KB  [128 dims][32 keys]
constexpr int STAGE_C = 2 * TILE_P + 2 * TILE_V;   // 64 KB
constexpr int NSTAGE_C = 3;
constexpr int SMEM_MAIN = NSTAGE_A * STAGE_A;      // 208 KB (>= NSTAGE_C * STAGE_C = 192 KB)
constexpr int SMEM_BYTES = SMEM_MAIN + 1024 + 256 + 2 * 128 * 4;   // + barriers + row-sum exchange
constexpr int NUM_THREADS = 384;   // TMA, MMA, TMEM-alloc, spare + 8 softmax/epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int O_COL = 2 * NH;         // O accumulator starts at TMEM column 288

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);   // yield the issue slot: the single-thread TMA / MMA roles share SM sub-partitions with softmax warps
        if (clock64() - t0 > 4000000000ll) { printf("lele_b200 attn_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {   // K-major, SWIZZLE_128B (see gemm_i8_tc.cu)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10), K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct AttnArgs {
    int B, T, H, n_qtiles, n_kchunks;
    int rows_per_slice;           // = T (one clip per slice)
    float* out;                   // att [B*T, H*DK]
    unsigned* minmax_keys;        // [B][2] or NULL
};

// ------------------------------------------------------------------------------------------
// pre-pass: tf32 hi/lo operand copies
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_split_qk_kernel(const float* __restrict__ qkv, long long M, int d, float qscale, float* __restrict__ q_hi, float* __restrict__ q_lo,
                     float* __restrict__ k_hi, float* __restrict__ k_lo) {
    const long long total = M * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / d; int c = (int)(i - r * d);
        float q = __fmul_rn(qkv[r * 3 * d + c], qscale);          // mul(q, d_k^-1/2) exactly as the reference
        float k = qkv[r * 3 * d + d + c];
        float qh = tf32_hi(q), kh = tf32_hi(k);
        q_hi[i] = qh; q_lo[i] = __fsub_rn(q, qh);
        k_hi[i] = kh; k_lo[i] = __fsub_rn(k, kh);
    }
}
// V [T, dk] per (b,h) -> V^T [dk, Tp] (keys contiguous, zero padded), hi/lo
__global__ void __launch_bounds__(256)
attn_split_vt_kernel(const float* __restrict__ qkv, int T, int Tp, int d, int H, float* __restrict__ vt_hi, float* __restrict__ vt_lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z / H, h = blockIdx.z % H;
    const int t0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int t = t0 + j;
        tile[j][tx] = t < T ? qkv[((long long)b * T + t) * 3 * d + 2 * d + h * DK + e0 + tx] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int t = t0 + tx;
        if (t < Tp) {
            float v = tile[tx][j], vh = tf32_hi(v);
            long long o = (((long long)b * H + h) * DK + e0 + j) * Tp + t;
            vt_hi[o] = vh; vt_lo[o] = __fsub_rn(v, vh);
        }
    }
}

// ------------------------------------------------------------------------------------------
// fused attention
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
               const __grid_constant__ CUtensorMap map_kh, const __grid_constant__ CUtensorMap map_kl,
               const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl, const AttnArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + SMEM_MAIN);
    uint64_t* full_a = bars;                  // [2] TMA -> MMA (phase A)
    uint64_t* empty_a = bars + 2;             // [2] MMA -> TMA
    uint64_t* s_full = bars + 4;              // S complete (also: phase-A smem is free)
    uint64_t* v_full = bars + 5;              // [3] V chunk landed
    uint64_t* p_full = bars + 8;              // [3] P chunk written by the softmax warps
    uint64_t* pv_empty = bars + 11;           // [3] MMA consumed the stage
    uint64_t* o_full = bars + 14;             // O complete
    uint32_t* tmem_base_smem = (uint32_t*)(bars + 16);
    float* xsum = (float*)(smem + SMEM_MAIN + 256);   // [2][128] partial row sums of the two softmax groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % args.n_qtiles;
    const int bh = blockIdx.x / args.n_qtiles;
    const int h = bh % args.H, b = bh / args.H;
    const int T = args.T, NKC = args.n_kchunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_qh); prefetch_tmap(&map_ql); prefetch_tmap(&map_kh); prefetch_tmap(&map_kl); prefetch_tmap(&map_vh); prefetch_tmap(&map_vl);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        mbar_init(s_full, 1); mbar_init(o_full, 1);
        for (int s = 0; s < NSTAGE_C; ++s) { mbar_init(&v_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&pv_empty[s], 1); }
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            // phase A: 4 k-chunks of 32 dims through a 2-stage ring
            for (int kc = 0; kc < DK / KC; ++kc) {
                const int s = kc & 1; const uint32_t ph = (kc >> 1) & 1;
                mbar_wait(&empty_a[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_A;
                mbar_expect_tx(&full_a[s], STAGE_A);
                tma_load_4d(st, &map_qh, &full_a[s], kc * KC, h, qt * AQ, b);
                tma_load_4d(st + TILE_Q, &map_ql, &full_a[s], kc * KC, h, qt * AQ, b);
                tma_load_4d(st + 2 * TILE_Q, &map_kh, &full_a[s], kc * KC, h, 0, b);
                tma_load_4d(st + 2 * TILE_Q + NH * 128, &map_kh, &full_a[s], kc * KC, h, NH, b);
                tma_load_4d(st + 2 * TILE_Q + TILE_K, &map_kl, &full_a[s], kc * KC, h, 0, b);
                tma_load_4d(st + 2 * TILE_Q + TILE_K + NH * 128, &map_kl, &full_a[s], kc * KC, h, NH, b);
            }
            // phase C reuses the same shared memory: wait until every phase-A MMA has retired
            mbar_wait(s_full, 0);
            for (int c = 0; c < NKC; ++c) {
                const int s = c % NSTAGE_C; const uint32_t ph = (c / NSTAGE_C) & 1;
                mbar_wait(&pv_empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_C;
                mbar_expect_tx(&v_full[s], 2 * TILE_V);
                tma_load_4d(st + 2 * TILE_P, &map_vh, &v_full[s], c * KC, 0, h, b);
                tma_load_4d(st + 2 * TILE_P + TILE_V, &map_vl, &v_full[s], c * KC, 0, h, b);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t ID_S = idesc_tf32(NH), ID_O = idesc_tf32(DK);
            for (int kc = 0; kc < DK / KC; ++kc) {
                const int s = kc & 1; const uint32_t ph = (kc >> 1) & 1;
                mbar_wait(&full_a[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * STAGE_A);
                const uint64_t qh = make_smem_desc(base), ql = make_smem_desc(base + TILE_Q);
                const uint64_t kh = make_smem_desc(base + 2 * TILE_Q), kl = make_smem_desc(base + 2 * TILE_Q + TILE_K);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t dS = tmem_base + (uint32_t)(half * NH);
                    const uint64_t hoff = (uint64_t)((half * NH * 128) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {       // 8 floats (32 B) per MMA: +2 in the >>4 address field
                        const uint64_t ko = (uint64_t)(k * 2);
                        const uint32_t first = (kc == 0 && k == 0) ? 0u : 1u;
                        umma_tf32(dS, qh + ko, kh + hoff + ko, ID_S, first);
                        umma_tf32(dS, qh + ko, kl + hoff + ko, ID_S, 1u);
                        umma_tf32(dS, ql + ko, kh + hoff + ko, ID_S, 1u);
                    }
                }
                umma_commit(&empty_a[s]);
                if (kc == DK / KC - 1) umma_commit(s_full);
            }
            for (int c = 0; c < NKC; ++c) {
                const int s = c % NSTAGE_C; const uint32_t ph = (c / NSTAGE_C) & 1;
                mbar_wait(&v_full[s], ph);
                mbar_wait(&p_full[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * STAGE_C);
                const uint64_t phd = make_smem_desc(base), pld = make_smem_desc(base + TILE_P);
                const uint64_t vhd = make_smem_desc(base + 2 * TILE_P), vld = make_smem_desc(base + 2 * TILE_P + TILE_V);
                const uint32_t dO = tmem_base + (uint32_t)O_COL;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ko = (uint64_t)(k * 2);
                    umma_tf32(dO, phd + ko, vhd + ko, ID_O, (c == 0 && k == 0) ? 0u : 1u);
                    umma_tf32(dO, phd + ko, vld + ko, ID_O, 1u);
                    umma_tf32(dO, pld + ko, vhd + ko, ID_O, 1u);
                }
                umma_commit(&pv_empty[s]);
                if (c == NKC - 1) umma_commit(o_full);
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue: 8 warps, two threads per query row =====================
        // group 0 (warps 4-7) takes the even 32-key chunks, group 1 (warps 8-11) the odd ones; both read the
        // whole row for the max (cheap), each accumulates its share of the row sum, exchanged through smem.
        const int quad = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int r = quad * 32 + lane;                 // row inside the tile == TMEM lane
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int n8 = T & ~7;
        mbar_wait(s_full, 0);
        tc_fence_after();
        // ---- pass 1: row max over the T valid keys ----
        float mx = -3.402823466e+38f;
        for (int c = 0; c < NKC; ++c) {
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)(c * KC), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c * KC + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        // ---- pass 2: e = exp(s - max) (polynomial exp on the x86 SIMD body j < T/8*8, libm on the tail, as the
        //      reference softmax), tf32 hi/lo split into the swizzled K-major P tiles ----
        float psum = 0.0f;
        for (int c = grp; c < NKC; c += 2) {
            const int s = c % NSTAGE_C; const uint32_t ph = (c / NSTAGE_C) & 1;
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)(c * KC), v);
            float e[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int j = c * KC + i;
                const float dlt = __fsub_rn(__uint_as_float(v[i]), mx);
                e[i] = j < n8 ? lb_cephes_expf(dlt) : (j < T ? expf(dlt) : 0.0f);
                psum += e[i];
            }
            mbar_wait(&pv_empty[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE_C;
            float* ph_row = (float*)(st + r * 128);
            float* pl_row = (float*)(st + TILE_P + r * 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int phys = (q ^ (r & 7)) * 4;      // 16-byte chunk XOR (row % 8)
                float4 hi, lo;
                hi.x = tf32_hi(e[q * 4 + 0]); lo.x = __fsub_rn(e[q * 4 + 0], hi.x);
                hi.y = tf32_hi(e[q * 4 + 1]); lo.y = __fsub_rn(e[q * 4 + 1], hi.y);
                hi.z = tf32_hi(e[q * 4 + 2]); lo.z = __fsub_rn(e[q * 4 + 2], hi.z);
                hi.w = tf32_hi(e[q * 4 + 3]); lo.w = __fsub_rn(e[q * 4 + 3], hi.w);
                *reinterpret_cast<float4*>(ph_row + phys) = hi;
                *reinterpret_cast<float4*>(pl_row + phys) = lo;
            }
            fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[s]);
        }
        // exchange the two partial row sums
        xsum[grp * AQ + r] = psum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = __fdiv_rn(1.0f, __fadd_rn(xsum[r], xsum[AQ + r]));
        // ---- epilogue: O / sum -> att, per-clip min/max; transposed through smem for coalesced stores ----
        mbar_wait(o_full, 0);
        tc_fence_after();
        float* stg = (float*)smem + (size_t)(warp - 4) * (32 * 33);   // phase-C buffers are idle now (all MMAs retired)
        float mn = 3.402823466e+38f, mxo = -3.402823466e+38f;
        const int row0 = qt * AQ + quad * 32;
        const int nrows = min(32, T - row0);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int ch = grp * 2 + cc;
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)(O_COL + ch * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = __fmul_rn(__uint_as_float(v[i]), inv);
            __syncwarp();
            for (int rr = 0; rr < nrows; ++rr) {
                const float val = stg[rr * 33 + lane];
                args.out[((long long)b * T + row0 + rr) * (args.H * DK) + h * DK + ch * 32 + lane] = val;
                mn = fminf(mn, val); mxo = fmaxf(mxo, val);
            }
            __syncwarp();
        }
        if (args.minmax_keys && nrows > 0) {
            mn = lb_warp_min(mn); mxo = lb_warp_max(mxo);
            if (lane == 0) lb_mm_update(args.minmax_keys, b, mn, mxo);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
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
// f32 4-D tensor, innermost dim contiguous, 128B swizzle, zero OOB fill
int make_map_f32_4d_uncached(CUtensorMap* map, const void* ptr, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                    const unsigned box[4]) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { lb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return LELE_B200_ERR_CUDA; }
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lb_set_error("cuTensorMapEncodeTiled(f32 4d) failed (%d)", (int)r); return LELE_B200_ERR_CUDA; }
    return LELE_B200_OK;
}
lele_b200_ctx* g_ctx_for_maps = nullptr;
int make_map_f32_4d(CUtensorMap* map, const void* ptr, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                    const unsigned box[4]) {
    lele_b200_ctx* ctx = g_ctx_for_maps;
    unsigned long long h = lb_hash_mix(0x66333234ull, (unsigned long long)(uintptr_t)ptr);
    for (int i = 0; i < 4; ++i) h = lb_hash_mix(lb_hash_mix(h, dims[i]), box[i]);
    for (int i = 0; i < 3; ++i) h = lb_hash_mix(h, strides_bytes[i]);
    auto it = ctx->tmaps.find(h);
    if (it != ctx->tmaps.end()) { memcpy(map, it->second.data(), sizeof(CUtensorMap)); return LELE_B200_OK; }
    int rc = make_map_f32_4d_uncached(map, ptr, dims, strides_bytes, box);
    if (rc) return rc;
    std::vector<unsigned char> blob(sizeof(CUtensorMap));
    memcpy(blob.data(), map, sizeof(CUtensorMap));
    ctx->tmaps.emplace(h, std::move(blob));
    return LELE_B200_OK;
}
int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }
}  // namespace

bool lb_attention_tc_supported(int T, int d, int H) { return H > 0 && d == H * DK && T >= 1 && T <= 2 * NH; }

size_t lb_attention_tc_scratch_bytes(int B, int T, int d, int H) {
    const size_t Tp = (size_t)((T + 3) / 4 * 4);
    return sizeof(float) * (4 * (size_t)B * T * d + 2 * (size_t)B * H * DK * Tp) + 6 * 256;
}

// qkv [B*T, 3d] -> att [B*T, d]; optional fused per-clip min/max keys [B][2]
int lb_attention_tc(lele_b200_ctx* ctx, const float* qkv, int B, int T, int d, int H, float qscale, void* scratch, float* att,
                    unsigned* minmax_keys) {
    LB_REQUIRE(lb_attention_tc_supported(T, d, H), "attention_tc: unsupported geometry T=%d d=%d H=%d", T, d, H);
    const long long M = (long long)B * T;
    const int Tp = (T + 3) / 4 * 4;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    uint8_t* p = (uint8_t*)scratch;
    float* q_hi = (float*)p; p += al(sizeof(float) * M * d);
    float* q_lo = (float*)p; p += al(sizeof(float) * M * d);
    float* k_hi = (float*)p; p += al(sizeof(float) * M * d);
    float* k_lo = (float*)p; p += al(sizeof(float) * M * d);
    float* vt_hi = (float*)p; p += al(sizeof(float) * (size_t)B * H * DK * Tp);
    float* vt_lo = (float*)p;
    attn_split_qk_kernel<<<grid_for(M * d), 256, 0, ctx->stream>>>(qkv, M, d, qscale, q_hi, q_lo, k_hi, k_lo);
    LB_LAUNCH_CHECK(ctx);
    attn_split_vt_kernel<<<dim3(lb_ceil_div(Tp, 32), DK / 32, B * H), 256, 0, ctx->stream>>>(qkv, T, Tp, d, H, vt_hi, vt_lo);
    LB_LAUNCH_CHECK(ctx);

    // q/k: [B][T][H][DK] -> dims (DK, H, T, B)
    const unsigned long long dqk[4] = {(unsigned long long)DK, (unsigned long long)H, (unsigned long long)T, (unsigned long long)B};
    const unsigned long long sqk[3] = {(unsigned long long)DK * 4, (unsigned long long)d * 4, (unsigned long long)T * d * 4};
    const unsigned bq[4] = {KC, 1, AQ, 1}, bk[4] = {KC, 1, NH, 1};
    // v^T: [B][H][DK][Tp] -> dims (Tp, DK, H, B)
    const unsigned long long dv[4] = {(unsigned long long)Tp, (unsigned long long)DK, (unsigned long long)H, (unsigned long long)B};
    const unsigned long long sv[3] = {(unsigned long long)Tp * 4, (unsigned long long)DK * Tp * 4, (unsigned long long)H * DK * Tp * 4};
    const unsigned bv[4] = {KC, DK, 1, 1};
    CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
    int rc;
    g_ctx_for_maps = ctx;
    if ((rc = make_map_f32_4d(&mqh, q_hi, dqk, sqk, bq))) return rc;
    if ((rc = make_map_f32_4d(&mql, q_lo, dqk, sqk, bq))) return rc;
    if ((rc = make_map_f32_4d(&mkh, k_hi, dqk, sqk, bk))) return rc;
    if ((rc = make_map_f32_4d(&mkl, k_lo, dqk, sqk, bk))) return rc;
    if ((rc = make_map_f32_4d(&mvh, vt_hi, dv, sv, bv))) return rc;
    if ((rc = make_map_f32_4d(&mvl, vt_lo, dv, sv, bv))) return rc;
    AttnArgs a;
    a.B = B; a.T = T; a.H = H; a.n_qtiles = lb_ceil_div(T, AQ); a.n_kchunks = lb_ceil_div(T, KC); a.rows_per_slice = T;
    a.out = att; a.minmax_keys = minmax_keys;
    static thread_local bool attr_done = false;
    if (!attr_done) { LB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr_done = true; }
    attn_tc_kernel<<<B * H * a.n_qtiles, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(mqh, mql, mkh, mkl, mvh, mvl, a);
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
