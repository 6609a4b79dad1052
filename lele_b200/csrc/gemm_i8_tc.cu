// gemm_i8_tc.cu -- the dominant kernel: exact u8 x u8 -> s32 GEMM on the 5th-gen tensor cores
// (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by TMA with 128B swizzle) with
// lele's quantised-linear epilogue fused (zero-point corrections, per-slice activation scale x
// per-channel weight scale, bias, ReLU, optional residual adds / min-max / argmax).
//
// Replaces fused_dq_gemm_avx2 + gemm_2rows_avx2 (src/kernels/avx/quantization.rs:225,1603) --
// not a port: those are VPMADDUBSW row kernels; this is a persistent warp-specialised
// TMA -> tcgen05 -> TMEM -> epilogue pipeline.  Integer core is exact (the AVX2 i16 saturation,
// avx/quantization.rs:1597-1600, is not reproduced: the documented formula :926-936 is the contract):
//    acc[i,j] = sum_k a[i,k] w[k,j]                      (tcgen05, s32)
//    int      = acc - w_zp*rowsum[i] - a_zp[i]*colsum[j] + K*a_zp[i]*w_zp
//    y        = f32(int) * (a_scale[i] * w_scale[j]) + bias[j]   (separate mul / add, :1417-1423)
//
// Tiling: CTA tile 128 x 256 x 128B-K, 3 shared-memory stages (48 KB each), 2 TMEM accumulator stages (2 x 256 columns = all 512),
// 576 threads: 16 epilogue warps (TMEM lane quadrant = warp % 4, 64-column group = warp / 4), then the two single-thread control
// roles -- TMA producer (+ TMEM allocation) and MMA issuer -- in the CTA's LAST two warps (the issue arbiter serves the highest warp
// id first).  Persistent: grid = #SMs, tiles walked n-fastest so concurrently running CTAs share A rows in L2.
//
// Epilogue variants (compile-time MODE): plain / min-max / arg-max / residual adds / QKV (+ V^T) / fused output quantiser.  Outputs
// leave through a swizzled per-warp staging tile and TMA tensor stores; an in-place residual (x = x + ...) is a TMA REDUCE-ADD
// store (the L2 adds), the FSMN residual tile is fetched by TMA into the staging tile.  Further variants, all bit-identical:
//   gemm_i8_fused_q_kernel  FFN1 in one pass: the dequantised tile waits in TMEM for the clip's max, then is quantised (below)
//   AF                      the A operand quantised in the kernel from f32 + per-clip keys, resident across the n-blocks (opt-in)
//   mc = 1 / CG2            2-CTA clusters: weight tile by TMA multicast / cta_group::2 pair MMAs 256 x 256 (opt-ins, measured equal)
#include "gemm_i8_tc.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int BM = 128, BN = 256, BK = 128;          // BK in bytes == elements (u8)
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK;               // 16 KB
constexpr int B_STAGE_BYTES = BN * BK;               // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_EPI_WARPS = 16;
constexpr int FIRST_EPI_WARP = 2;                    // warp 0: TMA producer (+ TMEM alloc), warp 1: MMA issuer
// 18 warps -> 5 on two of the four SM sub-partitions (16 K registers each) -> ptxas caps every thread at 96 registers.
// setmaxnreg re-balancing (control warpgroup down, epilogue warps up to 104-112) was tried: ptxas then spills far more
// (600-1600 B per thread), so the cap stays and the residual epilogues live with ~100-300 B of spills.
constexpr int NUM_THREADS = (FIRST_EPI_WARP + NUM_EPI_WARPS) * 32;   // 576
constexpr int TMEM_COLS = 512;
constexpr int EPI_TILE_BYTES = 32 * 128;                       // per-warp 32x32 f32 staging tile, 128B-swizzled (1024 B aligned)
constexpr int EPI_META_BYTES = 3 * 64 * 4;                     // zp*colsum / scale / bias of the warp's 64 columns
constexpr int EPI_BYTES = NUM_EPI_WARPS * (EPI_TILE_BYTES + EPI_META_BYTES);
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int STAGES_CG2 = 4;                        // cta_group::2: 32 KB per stage (A 16 KB + this CTA's half of B 16 KB)
constexpr int MAX_STAGES = 4;
#ifdef LELE_B200_GEMM_TIMELINE
constexpr bool GEMM_DBG = true;    // role wait counters (clock64) printed by CTAs 0 / 77 when LELE_B200_GEMM_DBG=1; costs ~12 registers
#else
constexpr bool GEMM_DBG = false;
#endif
constexpr int UMMA_K = 32;                           // bytes per tcgen05.mma for 8-bit operands

#ifdef LELE_B200_CTRL_WARPS_LOW
__device__ __forceinline__ int lb_logical_warp(int pw) { return pw; }
#else
__device__ __forceinline__ int lb_logical_warp(int pw) { return pw < NUM_EPI_WARPS ? pw + FIRST_EPI_WARP : pw - NUM_EPI_WARPS; }
#endif

// ---- PTX wrappers ----------------------------------------------------------
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
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);   // yield the issue slot to the epilogue warps sharing this SM sub-partition
        if (clock64() - t0 > 4000000000ll) { printf("lele_b200 gemm_i8_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// B operand shared by a 2-CTA cluster: each CTA fetches half of the tile and the TMA unit delivers it to BOTH CTAs' shared memory
// (same CTA-relative offset), signalling each CTA's own mbarrier -- the weight tile crosses the L2 -> SM fabric once per pair.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// ---- cta_group::2: the two CTAs of a cluster (two SMs of one TPC) execute ONE 256 x 256 MMA; each holds its own 128 rows of A and
// half of B (128 of the 256 columns) in shared memory, so the operand bytes an SM reads and TMA writes per MMA are 2/3 of the 1-CTA case.
// TMA loads of either CTA signal the LEADER's (rank 0) barrier: the barrier address with the CTA-rank bit cleared (cute Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t IDESC) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {   // arrives on the barrier at this offset in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {                // arrive on the leader CTA's copy of a barrier
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(0));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (SM100 UMMA; cute/arch/mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=S32 (c_format 2 @4), A=U8 (0 @7), B=U8 (0 @10) or S8 (1 @10: the weight stored as w - 128),
// K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t idesc_for(bool b_signed, int m = BM) {
    return (2u << 4) | (0u << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t IDESC) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival delivered to the barrier at this offset in every CTA of the mask (a shared-memory slot that a peer's multicast
// load refills may only be released when both CTAs' MMAs have consumed it)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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

struct KernelArgs {
    int M, N, K;
    int vec_io;
    int mc;             // 0 = independent CTAs; 1 = 2-CTA clusters along M share each B tile by TMA multicast; 2 = cta_group::2: the
                        // cluster's two CTAs run one 256 x 256 MMA (grid = 2 x resident clusters in both cluster modes)
    int red;            // TMA_OUT epilogues: the tile is ADDED to `out` (cp.reduce ... .add) -- the in-place residual x = x + (...)
    int dbg;            // LELE_B200_GEMM_DBG=1: CTAs 0 and 77 print where their producer / MMA / epilogue roles waited (clock64)
    int num_m_blocks, num_n_blocks, num_k_blocks;
    int rps;            // rows per slice (0x7fffffff when the epilogue has no slices)
    float inv_rps;      // 1.0f / rps: slice index by float multiply + one exact correction step (no integer division)
    LbI8Epilogue ep;
};

// Epilogue specialisations (compile-time, so the inner loops carry no uniform branches)
enum EpiMode { EPI_PLAIN = 0, EPI_MINMAX, EPI_ARGMAX, EPI_R1, EPI_R2, EPI_R12, EPI_QKV, EPI_QUANT };

// n / d for 0 <= n < 2^22 with inv = 1.0f / d: float estimate + one exact correction (an integer division costs ~35
// instructions and the epilogue needs several per tile per warp)
__device__ __forceinline__ int div_by_rps(int n, int d, float inv) {
    int q = __float2int_rz(__fmul_rn(__int2float_rn(n), inv));
    const int r = n - q * d;
    q += (r >= d) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
    int4 v; asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts_s32(uint32_t addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void sts_v4f(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// out[tile] += staging tile, the f32 addition done by the L2 (cp.reduce.async.bulk.tensor .add): an in-place residual
// (x = x + linear(..)) never enters the SM.  One IEEE round-to-nearest addition per element, operands commute, so the
// result is bit-identical to __fadd_rn(x, v) in the epilogue (checked by the bit-exact network tests).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// round-half-even + clamp to [0, 255] in one conversion (float -> integer cvt saturates to the destination range, NaN -> 0)
__device__ __forceinline__ unsigned cvt_sat_u8(float v) { unsigned u; asm("cvt.rni.u8.f32 %0, %1;" : "=r"(u) : "f"(v)); return u; }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// TMA_OUT: the 32x32 f32 sub-tile a warp finished is written by one cp.async.bulk.tensor store from the
// swizzled staging tile (no per-element store phase); used when the epilogue has no residual operand.
// RELU (compile-time; PLAIN / MINMAX / QUANT only): the per-element epilogue is bound by the half-rate ALU pipe (IADD3,
// I2FP, FMNMX ...), so a ReLU that is not asked for -- or is implied (QUANT: unsigned saturation; MINMAX: max(relu(t)) =
// max(max t, 0), min(relu(t)) >= 0) -- must not cost an FMNMX per element.
// WSIGNED: the weight operand is s8 (w - 128, prepare_weights with w_zp == 128): no per-row zero-point term in the epilogue.
// CG2 (compile-time: a kernel that contains cta_group::2 instructions can only be launched as a cluster): the two CTAs of a cluster run
// one 256 x 256 MMA per k-step (see umma_i8_2sm); instantiated for the plain / TMA-store epilogue (FFN2).
// AF (compile-time; K <= 512, one m-block per CTA): the A operand is produced IN the kernel from the f32 activation and its per-clip
// min / max keys -- the reference's dynamic quantiser (dq_to_u8_rowsums_avx2) fused into the GEMM that consumes it.  The 16 epilogue
// warps, idle until the first accumulator is ready, quantise the CTA's 128 x K block straight into the 128B-swizzled K-major layout TMA
// would have produced; it stays resident while the CTA walks the n-blocks (only weight tiles stream).  No u8 tensor, no row-parameter
// arrays, no separate quantiser launch.  Instantiated for the out-projection (EPI_R1, TMA store, s8 weights).
constexpr int AF_KB = 4;                              // resident k-blocks of the A block (K <= 512)
constexpr int AF_STAGES = 2;                          // B ring of the AF variant
static_assert(AF_KB * A_STAGE_BYTES + AF_STAGES * B_STAGE_BYTES <= STAGES * STAGE_BYTES, "the AF layout fits the operand area");
template <int MODE, bool TMA_OUT, bool RELU, bool WSIGNED, bool CG2 = false, bool AF = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_lo, const KernelArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    unsigned long long cta_t0 = 0;
    if (GEMM_DBG && args.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_t0));
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;   // SWIZZLE_128B needs 1024 B alignment
    uint8_t* smem = smem_raw + pad;
    constexpr int NST = AF ? AF_STAGES : (CG2 ? STAGES_CG2 : STAGES);
    static_assert(!(AF && CG2), "AF and CG2 are separate variants");
    constexpr int ETB = EPI_TILE_BYTES;
    constexpr int B_STRIDE = CG2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;      // bytes between two stages of the B ring
    static_assert(STAGES_CG2 * (A_STAGE_BYTES + B_STAGE_BYTES / 2) <= STAGES * STAGE_BYTES, "the cta_group::2 ring fits the operand area");
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (AF ? AF_KB : NST) * A_STAGE_BYTES;
    uint8_t* epi_base = smem + STAGES * STAGE_BYTES;                   // per-warp staging tiles, then column metadata
    uint64_t* bars = (uint64_t*)(epi_base + EPI_BYTES);
    uint64_t* full_bar = bars;                     // [NST]  TMA -> MMA
    uint64_t* empty_bar = bars + MAX_STAGES;       // [NST]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;   // [2]    MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;          // [2]    epilogue -> MMA
    uint32_t* tmem_base_smem = (uint32_t*)(tmem_empty + 2);
    uint64_t* a_ready = bars + 14;                 // AF: the CTA's quantised A block is in shared memory (16 epilogue warps arrive)
    uint64_t* res_bar = bars + 16;                 // [NUM_EPI_WARPS] EPI_R1 + TMA_OUT: the warp's residual sub-tile landed (TMA load into its staging tile)
    constexpr bool R1T = (MODE == EPI_R1) && TMA_OUT;

    // Roles by LOGICAL warp index (0 = TMA producer, 1 = MMA issuer, 2.. = epilogue).  Physically the two single-thread control roles sit
    // in the LAST two warps of the CTA: the SM sub-partition's issue arbiter serves the highest warp id first, and a TMA / MMA issue that
    // waits behind four busy epilogue warps stalls the whole pipeline (build with -DLELE_B200_CTRL_WARPS_LOW for round 1's placement).
    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = lb_logical_warp(pw);
    // tile walk: tile = first + i * stride -> (m, n_blk) = (tile / nnb, tile % nnb); m is the m-block, or with multicast clusters the
    // PAIR of m-blocks the cluster works on (the CTA's own m-block is 2 m + its rank: both CTAs walk the same n-blocks in step)
    const int mc = CG2 ? 2 : (args.mc == 1 ? 1 : 0);
    const int crank = mc ? (int)(blockIdx.x & 1) : 0;
    // (AF: CTA = one m-block, its tiles are that m-block's n-blocks in order)
    const int tile_first = AF ? (int)blockIdx.x * args.num_n_blocks : (mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
    const int tile_stride = AF ? 1 : (mc ? (int)(gridDim.x >> 1) : (int)gridDim.x);
    const int num_tiles_all = (mc ? (args.num_m_blocks + 1) / 2 : args.num_m_blocks) * args.num_n_blocks;
    const int num_tiles = AF ? min(num_tiles_all, tile_first + args.num_n_blocks) : num_tiles_all;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); if (TMA_OUT) prefetch_tmap(&tmap_out); if (R1T) prefetch_tmap(&tmap_lo); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], mc == 1 ? 2 : 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], mc == 2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS); }
        if (R1T) for (int s = 0; s < NUM_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
        if (AF) mbar_init(a_ready, NUM_EPI_WARPS);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (CG2) cluster_sync_all();   // both CTAs of the pair are running before the paired allocation
    if (warp == 0) {   // whole warp: allocate all 512 TMEM columns (1 CTA/SM by construction)
        if (CG2) { // (the same warp of both CTAs, same destination offset)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if (mc) cluster_sync_all();          // the peer's barriers are initialised before any multicast load / remote arrival can reach them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    // PDL: everything above is independent of the previous kernel's output; everything below reads it
    lb_pdl_launch_dependents();
    lb_pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            long long w_empty = 0; const long long t_begin = clock64();
            for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
                const int m_blk = (tile / args.num_n_blocks) * (mc ? 2 : 1) + crank, n_blk = tile % args.num_n_blocks;
                for (int kb = 0; kb < args.num_k_blocks; ++kb) {
                    if (GEMM_DBG && args.dbg) { const long long t0 = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1); w_empty += clock64() - t0; }
                    else mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (AF) {                                      // the A block is resident: only the weight tile streams
                        mbar_expect_tx(&full_bar[stage], B_STAGE_BYTES);
                        tma_load_2d(smem_b + stage * B_STRIDE, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (CG2) {
                        // each CTA fetches its own A rows and its half of B into its own shared memory; both signal the leader's barrier,
                        // which the leader arms for the pair's 64 KB
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES / 2));
                        tma_load_2d_2sm(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
                        tma_load_2d_2sm(smem_b + stage * B_STRIDE, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN + crank * (BN / 2));
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
                    if (mc)   // tmap_b's box is half a tile here: rows [rank * 128, +128) of the n-block, delivered to both CTAs
                        tma_load_2d_mc(smem_b + stage * B_STRIDE + crank * (B_STAGE_BYTES / 2), &tmap_b, &full_bar[stage], kb * BK,
                                       n_blk * BN + crank * (BN / 2), (uint16_t)3);
                    else tma_load_2d(smem_b + stage * B_STRIDE, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
            if (GEMM_DBG && args.dbg && (blockIdx.x == 0 || blockIdx.x == 77))
                printf("GEMMDBG blk %d TMA: total %lld wait_empty %lld\n", blockIdx.x, clock64() - t_begin, w_empty);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread; with cta_group::2 the leader CTA's, for the pair) =====================
        if (lane == 0 && !(CG2 && crank != 0)) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            long long w_tmem = 0, w_full = 0; const long long t_begin = clock64();
            if (AF && tile_first < num_tiles) { mbar_wait(a_ready, 0); tc_fence_after(); }   // the quantised A block is in place
            for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
                if (GEMM_DBG && args.dbg) { const long long t0 = clock64(); mbar_wait(&tmem_empty[acc], acc_phase ^ 1); w_tmem += clock64() - t0; }
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < args.num_k_blocks; ++kb) {
                    if (GEMM_DBG && args.dbg) { const long long t0 = clock64(); mbar_wait(&full_bar[stage], phase); w_full += clock64() - t0; }
                    else mbar_wait(&full_bar[stage], phase);            // TMA bytes landed
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + (AF ? kb : stage) * A_STAGE_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * B_STRIDE));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance along K inside the 128B swizzle atom: +32 B == +2 in the (>>4) address field
                        if (CG2)
                            umma_i8_2sm(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)),
                                        (kb > 0 || k > 0) ? 1u : 0u, idesc_for(WSIGNED, 2 * BM));
                        else
                            umma_i8(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)),
                                    (kb > 0 || k > 0) ? 1u : 0u, idesc_for(WSIGNED));
                    }
                    if (CG2) umma_commit_2sm(&empty_bar[stage], (uint16_t)3);    // both CTAs' slots are free when the pair's MMAs retire
                    else if (mc) umma_commit_mc(&empty_bar[stage], (uint16_t)3);     // ... in both CTAs (the peer refills half of it)
                    else umma_commit(&empty_bar[stage]);           // frees the smem slot when the MMAs retire
                    if (kb == args.num_k_blocks - 1) { if (CG2) umma_commit_2sm(&tmem_full[acc], (uint16_t)3); else umma_commit(&tmem_full[acc]); }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (GEMM_DBG && args.dbg && (blockIdx.x == 0 || blockIdx.x == 77))
                printf("GEMMDBG blk %d MMA: total %lld wait_operands %lld wait_epilogue %lld (tiles %d, k-blocks %d)\n", blockIdx.x, clock64() - t_begin, w_full, w_tmem,
                       (num_tiles - tile_first + tile_stride - 1) / tile_stride, args.num_k_blocks);
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ===================== epilogue: 16 warps, each owns 32 rows x 64 columns of every tile =====================
        // phase 1 (thread = row): TMEM -> exact integer corrections, scale, bias, ReLU (+ min/max, arg-max) -> the warp's
        //          staging tile: 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7) == the TMA SWIZZLE_128B
        //          layout, so both the row-wise 16 B writes and the column-wise 4 B reads are bank-conflict free
        // then     TMA_OUT: one elected lane issues a 32x32 tensor store (clipped at M / N by the hardware)
        //          else   : phase 2 (lane = column) reads the tile transposed -> residual adds -> 128-byte coalesced stores
        const int ew = warp - FIRST_EPI_WARP;
        const int quad = pw & 3;            // TMEM lane quadrant this (physical) warp may access
        const int cgrp = ew >> 2;           // which 64-column group of the tile
        const LbI8Epilogue& ep = args.ep;
        const uint32_t tile_s = smem_u32(epi_base) + (uint32_t)ew * ETB;
        const uint32_t meta_s = smem_u32(epi_base) + NUM_EPI_WARPS * ETB + (uint32_t)ew * EPI_META_BYTES;   // zp*colsum[64] | scale[64] | bias[64]
        const uint32_t my_row_s = tile_s + (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int N = args.N, M = args.M;
        const float FMAX = 3.402823466e+38f;
        const bool vec_io = !TMA_OUT && args.vec_io;       // rows 16-byte aligned: float4 residual loads / output stores
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t res_phase = 0;
        long long w_acc = 0, t_busy0 = 0, busy = 0, d_ld = 0, d_math = 0; const long long t_begin = clock64();
        int pf_rs = 0, pf_zpa = 0, pf_cs[2] = {0, 0}; float pf_sa = 0.0f, pf_ws[2] = {0.0f, 0.0f}, pf_bi[2] = {0.0f, 0.0f};
        unsigned pf_qkey = 0;                                          // EPI_QUANT: one min/max key slot of the warp's clip A (lanes 0-15) / A+1 (16-31)
        const int nnb = args.num_n_blocks;
        const int rps = args.rps; const float inv_rps = args.inv_rps;
        const int last_slice = div_by_rps(M - 1, rps, inv_rps);
        // tile coordinates advance incrementally (tile += gridDim.x): no per-tile integer divisions
        // (m_idx = the m-block, or the cluster's pair of m-blocks; the CTA's own m-block follows from it)
        const int step_m = tile_stride / nnb, step_n = tile_stride % nnb;
        int m_idx = tile_first / nnb, n_blk = tile_first % nnb;
        const int mmul = mc ? 2 : 1;
        int m_blk = m_idx * mmul + crank;
        int af_zpa = 0; float af_sa = 0.0f;                            // AF: the row's activation zero point / scale, derived from the keys
        if (AF && tile_first < num_tiles) {
            // (scale, zp, 1 / scale) of one clip from its 8 (min, max) key slots: lanes 0-15 load the 16 keys, xor-shuffles reduce them
            auto clip_params = [&](int clip, float& scale, float& zp, float& inv) {
                unsigned k = __ldg(ep.a_keys + (size_t)clip * LB_MM_SLOTS * 2 + (lane & 15));
#pragma unroll
                for (int of = 2; of <= 8; of <<= 1) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, k, of);
                    k = (lane & 1) ? max(k, o) : min(k, o);
                }
                const float mn = lb_fkey_inv(__shfl_sync(0xffffffffu, k, 0)), mx = lb_fkey_inv(__shfl_sync(0xffffffffu, k, 1));
                const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);          // dq_params (quant.cu)
                scale = __fdiv_rn(fmaxf(__fsub_rn(amax, amin), 1e-5f), 255.0f);
                zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, scale)), 0.0f), 255.0f);
                inv = __fdiv_rn(1.0f, scale);
            };
            // ---- the quantiser: warp ew takes rows ew, ew + 16, ... of the block; a lane takes 16 consecutive k (64 B of f32 -> one
            //      16-byte chunk of the swizzled K-major tile of its k-block) ----
            const int m0 = (int)blockIdx.x * BM;
            const int kch = args.K >> 4;                               // 16-element chunks per row (K % 128 == 0, K <= 512)
            int cur_clip = -1; float c_scale = 0.0f, c_zp = 0.0f, c_inv = 0.0f;
#pragma unroll 1
            for (int r = ew; r < BM; r += NUM_EPI_WARPS) {
                const int grow = m0 + r;
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (grow < M) {                                        // warp-uniform
                    const int clip = div_by_rps(grow, rps, inv_rps);
                    if (clip != cur_clip) { clip_params(clip, c_scale, c_zp, c_inv); cur_clip = clip; }
                    if (lane < kch) {
                        const float4* src = reinterpret_cast<const float4*>(ep.a_f32 + (size_t)grow * args.K + lane * 16);
                        const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
                        auto q4 = [&](const float4& v) {               // clamp(rint(fma(x, 1 / scale, zp)), 0, 255), packed little-endian
                            const unsigned a = cvt_sat_u8(__fmaf_rn(v.x, c_inv, c_zp)), b = cvt_sat_u8(__fmaf_rn(v.y, c_inv, c_zp));
                            const unsigned c = cvt_sat_u8(__fmaf_rn(v.z, c_inv, c_zp)), d = cvt_sat_u8(__fmaf_rn(v.w, c_inv, c_zp));
                            return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
                        };
                        pk = make_uint4(q4(v0), q4(v1), q4(v2), q4(v3));
                    }
                }
                if (lane < kch) {
                    const uint32_t dst = smem_u32(smem_a) + (uint32_t)(lane >> 3) * A_STAGE_BYTES + (uint32_t)r * 128u + ((((uint32_t)lane & 7u) ^ ((uint32_t)r & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
                }
            }
            fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core's operand fetch
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            // ---- the epilogue's per-row activation parameters: the warp's 32 rows lie in clip A (that of its first row) or A + 1 ----
            {
                const int fr = m0 + quad * 32;
                const int clip_a = min(div_by_rps(min(fr, M - 1), rps, inv_rps), last_slice);
                float sA, zA, iA, sB, zB, iB;
                clip_params(clip_a, sA, zA, iA);
                clip_params(min(clip_a + 1, last_slice), sB, zB, iB);
                const bool in_a_row = fr + lane < (clip_a + 1) * rps;
                af_sa = in_a_row ? sA : sB; af_zpa = (int)(in_a_row ? zA : zB);
            }
        }
        auto fetch_meta = [&](int t, int mb, int nb) {
            if (t >= num_tiles) return;
            const int rw = mb * BM + quad * 32 + lane;
            if (MODE == EPI_QUANT) {
                const int sl = min(div_by_rps(mb * BM + quad * 32, rps, inv_rps) + (lane >> 4), last_slice);
                pf_qkey = __ldg(ep.q_keys + (size_t)sl * LB_MM_SLOTS * 2 + (lane & 15));
            }
            pf_rs = 0; pf_zpa = 0; pf_sa = 0.0f;
            if (AF) { if (rw < M) { pf_zpa = af_zpa; pf_sa = af_sa; } }
            else if (rw < M) { if (!WSIGNED) pf_rs = __ldg(ep.rowsum + rw); pf_zpa = __ldg(ep.row_zp + rw); pf_sa = __ldg(ep.row_scale + rw); }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int cc = min(nb * BN + cgrp * 64 + hh * 32 + lane, N - 1);
                pf_cs[hh] = __ldg(ep.colsum + cc); pf_ws[hh] = __ldg(ep.w_scale + cc); pf_bi[hh] = __ldg(ep.bias + cc);
            }
        };
        fetch_meta(tile_first, m_blk, n_blk);
        for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
            const int first_row = m_blk * BM + quad * 32;
            const int row = first_row + lane;
            const bool row_ok = row < M;
            const int gcol_w = n_blk * BN + cgrp * 64;                 // first global column of this warp
            // this tile's row / column metadata was fetched while the previous tile drained (pf_*), so no global-load
            // latency sits between two accumulators
            const int rs = pf_rs, zpa = pf_zpa; const float sa = pf_sa;
            const int row_corr = WSIGNED ? 0 : args.K * zpa * ep.w_zp - ep.w_zp * rs;   // s8 weights carry their zero point: no row term
            // Per-warp column metadata, pre-combined with the activation parameters of the clip that owns the warp's
            // first row ("A"): zcA[c] = zp_A * colsum[c], csA[c] = scale_A * w_scale[c].  A warp's 32 rows touch a
            // second clip only at clip boundaries (1 warp in ~8 for T' = 271); those rows take the uncombined path.
            const int slice_first = div_by_rps(first_row, rps, inv_rps);
            const bool in_a = row_ok && row < (slice_first + 1) * rps;     // rows ascend: the warp's rows are in slice A or A+1
            const bool all_a = __all_sync(0xffffffffu, !row_ok || in_a);
            const int zpa_a = __shfl_sync(0xffffffffu, zpa, 0);
            const float sa_a = __shfl_sync(0xffffffffu, sa, 0);
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = hh * 32 + lane;
                sts_s32(meta_s + 4 * c, all_a ? zpa_a * pf_cs[hh] : pf_cs[hh]);               // straddling warps keep raw metadata
                sts_f32(meta_s + 256 + 4 * c, all_a ? __fmul_rn(sa_a, pf_ws[hh]) : pf_ws[hh]);
                sts_f32(meta_s + 512 + 4 * c, pf_bi[hh]);
            }
            float q_inv = 0.0f, q_zp = 0.0f, q_scale = 0.0f; int q_sum = 0;
            if (MODE == EPI_QUANT) {
                // lanes 0-15 hold the 8 (min, max) key slots of clip A, lanes 16-31 those of clip A+1: reduce over the slots
                unsigned k = pf_qkey;
#pragma unroll
                for (int of = 2; of <= 8; of <<= 1) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, k, of);
                    k = (lane & 1) ? max(k, o) : min(k, o);
                }
                const int src = in_a ? 0 : 16;
                const float mn = lb_fkey_inv(__shfl_sync(0xffffffffu, k, src)), mx = lb_fkey_inv(__shfl_sync(0xffffffffu, k, src + 1));
                const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);          // dq_params (quant.cu)
                q_scale = __fdiv_rn(fmaxf(__fsub_rn(amax, amin), 1e-5f), 255.0f);
                q_zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, q_scale)), 0.0f), 255.0f);
                q_inv = __fdiv_rn(1.0f, q_scale);
            }
            {   // next tile's coordinates and metadata (consumed one iteration later)
                int mi_n = m_idx + step_m, nb_n = n_blk + step_n;
                if (nb_n >= nnb) { nb_n -= nnb; ++mi_n; }
                fetch_meta(tile + tile_stride, mi_n * mmul + crank, nb_n);
            }
            float best_t = -3.402823466e+38f; int best_c = -1;
            float vmin = FMAX, vmax = -FMAX;                           // this row's min / max over the warp's 64 columns
            const int nrows = min(32, M - first_row);
            __syncwarp();
            if (GEMM_DBG && args.dbg) { const long long t0 = clock64(); mbar_wait(&tmem_full[acc], acc_phase); t_busy0 = clock64(); w_acc += t_busy0 - t0; }
            else mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int chunk = 0; chunk < 2; ++chunk) {
                const int col0 = cgrp * 64 + chunk * 32;               // column inside the tile
                const int gcol0 = n_blk * BN + col0;                   // global column
                uint32_t r[32];
                long long t_c0 = 0;
                const bool skip_chunk = gcol0 >= N || nrows <= 0;      // warp-uniform
                if (R1T && !skip_chunk && lane == 0) {
                    // the residual sub-tile (add1) is fetched by TMA into the warp's staging tile -- in the layout the output
                    // leaves it in -- while the accumulator is loaded and dequantised: no residual registers, no exposed latency
                    tma_store_wait_read();                             // the previous store has read the staging tile
                    mbar_expect_tx(&res_bar[ew], EPI_TILE_BYTES);
                    tma_load_2d_s(tile_s, &tmap_lo, &res_bar[ew], gcol0, first_row);
                }
                if (GEMM_DBG && args.dbg) t_c0 = clock64();
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col0), r);
                if (GEMM_DBG && args.dbg) { const long long t1 = clock64(); d_ld += t1 - t_c0; t_c0 = t1; }
                if (skip_chunk) continue;
                // Residual epilogues: 128-bit accesses, lane = (row % 4, 16-byte column chunk), 4 rows x 128 B per warp
                // instruction, and every residual load of the sub-tile is issued before the accumulator math so that
                // 4 KB (8 KB for two residuals) per warp is in flight while phase 1 runs.
                constexpr bool HAS_R1 = (MODE == EPI_R1 || MODE == EPI_R12), HAS_R2 = (MODE == EPI_R2 || MODE == EPI_R12);
                const int vq = lane & 7, vr = lane >> 3;
                const bool vcol_ok = gcol0 + vq * 4 < N;               // N % 4 == 0 on this path: a float4 is all-in or all-out
                const long long vbase = (long long)(first_row + vr) * N + gcol0 + vq * 4;
                float4 res1[8], res2[8];
                if ((HAS_R1 || HAS_R2) && vec_io) {
                    if (MODE == EPI_R1) {
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (vcol_ok && it * 4 + vr < nrows) res1[it] = __ldg(reinterpret_cast<const float4*>(ep.add1 + vbase + (long long)it * 4 * N));
                    }
                    if (MODE == EPI_R2) {
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (vcol_ok && it * 4 + vr < nrows) res2[it] = __ldg(reinterpret_cast<const float4*>(ep.add2 + vbase + (long long)it * 4 * N));
                    }
                }
                // ---- phase 1 ----
                const bool full_cols = gcol0 + 32 <= N;
                // r[] <- the 32 finished f32 values of this row (accumulator registers are reused in place).  Four
                // compile-time variants: FULL = all 32 columns exist (no per-element column predicate: VIADD + ISETP + FSEL
                // per element would double the ALU-pipe work that bounds this loop), STRADDLE = the warp's rows span two
                // clips (raw column metadata, combined per row here; 1 warp in ~8 for T' = 271).
                auto phase1 = [&](auto full_tag, auto straddle_tag) {
                    constexpr bool FULL = decltype(full_tag)::value, STRADDLE = decltype(straddle_tag)::value;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int4 zc4 = lds_v4(meta_s + 4 * (chunk * 32 + q * 4));
                        const int4 cs4 = lds_v4(meta_s + 256 + 4 * (chunk * 32 + q * 4));
                        const int4 bi4 = lds_v4(meta_s + 512 + 4 * (chunk * 32 + q * 4));
                        const int zc[4] = {zc4.x, zc4.y, zc4.z, zc4.w};
                        const float cs[4] = {__int_as_float(cs4.x), __int_as_float(cs4.y), __int_as_float(cs4.z), __int_as_float(cs4.w)};
                        const float bi[4] = {__int_as_float(bi4.x), __int_as_float(bi4.y), __int_as_float(bi4.z), __int_as_float(bi4.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = q * 4 + e;
                            const int iv = (WSIGNED ? (int)r[idx] : (int)r[idx] + row_corr) - (STRADDLE ? zpa * zc[e] : zc[e]);
                            float t = __fmul_rn((float)iv, STRADDLE ? __fmul_rn(sa, cs[e]) : cs[e]);
                            t = ep.has_bias ? __fadd_rn(t, bi[e]) : t;
                            const bool col_ok = FULL || gcol0 + idx < N;
                            if (MODE == EPI_ARGMAX) {                  // columns ascend, ">=" keeps the last maximum (tokenizer.rs:55)
                                if (col_ok && t >= best_t) { best_t = t; best_c = gcol0 + idx; }
                            }
                            if (MODE == EPI_MINMAX) {
                                if (col_ok) { if (!RELU) vmin = fminf(vmin, t); vmax = fmaxf(vmax, t); }
                            }
                            if (RELU && MODE != EPI_MINMAX && MODE != EPI_QUANT) t = fmaxf(t, 0.0f);
                            r[idx] = __float_as_uint(t);
                        }
                    }
                };
                if (full_cols) { if (all_a) phase1(std::true_type{}, std::false_type{}); else phase1(std::true_type{}, std::true_type{}); }
                else           { if (all_a) phase1(std::false_type{}, std::false_type{}); else phase1(std::false_type{}, std::true_type{}); }
                if (GEMM_DBG && args.dbg) { const long long t1 = clock64(); d_math += t1 - t_c0; t_c0 = t1; }
                if (MODE == EPI_QUANT) {
                    // the consumer's dynamic quantiser fused here: u8 = clamp(rint(fma(y, 1/scale, zp)), 0, 255) with the clip's
                    // (scale, zp) from the min/max a previous max-only pass of this same GEMM produced; ReLU is implied by the
                    // unsigned saturation (a ReLU output has min >= 0, hence zp == 0).  32 B per row -> 1 KB staging -> TMA store.
                    unsigned pk[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const unsigned u0 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[q * 4 + 0]), q_inv, q_zp));
                        const unsigned u1 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[q * 4 + 1]), q_inv, q_zp));
                        const unsigned u2 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[q * 4 + 2]), q_inv, q_zp));
                        const unsigned u3 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[q * 4 + 3]), q_inv, q_zp));
                        pk[q] = __byte_perm(__byte_perm(u0, u1, 0x0040), __byte_perm(u2, u3, 0x0040), 0x5410);
                        q_sum = (int)__dp4a(pk[q], 0x01010101u, (unsigned)q_sum);
                    }
                    if (lane == 0) tma_store_wait_read();              // the previous store has read the staging tile (waited for after the
                    __syncwarp();                                      // math, so that the read overlaps it)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile_s + (uint32_t)lane * 32u), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile_s + (uint32_t)lane * 32u + 16u), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&tmap_out, tile_s, gcol0, first_row);
                    continue;
                }
                if (!ep.out) continue;                                 // arg-max ids only / max-only pass: nothing to write
                if (RELU && MODE == EPI_MINMAX) {                      // (the min/max above ran on the unclamped values)
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.0f));
                }
                // QKV projection with skip_v_out: the v columns leave the SM only as V^T (below), not through the staging tile
                const bool v_only_vt = MODE == EPI_QKV && ep.skip_v_out && gcol0 >= 2 * (N / 3);
                if (R1T) {
                    mbar_wait(&res_bar[ew], res_phase);                // add1's sub-tile is in the staging tile: add it row-wise, in place
                    res_phase ^= 1;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t ad = my_row_s + (((uint32_t)q ^ sw) << 4);
                        const int4 raw = lds_v4(ad);
                        sts_v4f(ad, __fadd_rn(__uint_as_float(r[q * 4 + 0]), __int_as_float(raw.x)), __fadd_rn(__uint_as_float(r[q * 4 + 1]), __int_as_float(raw.y)),
                                __fadd_rn(__uint_as_float(r[q * 4 + 2]), __int_as_float(raw.z)), __fadd_rn(__uint_as_float(r[q * 4 + 3]), __int_as_float(raw.w)));
                    }
                } else if (!v_only_vt) {
                    if (TMA_OUT) {                                     // the previous store must have read the staging tile (waited for
                        if (lane == 0) tma_store_wait_read();          // after the math, so that the read overlaps it)
                        __syncwarp();
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        sts_v4f(my_row_s + (((uint32_t)q ^ sw) << 4), __uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
                }
                if (TMA_OUT) {
                    if (!v_only_vt) {
                        fence_proxy_async();                           // generic-proxy smem writes -> visible to the TMA engine
                        __syncwarp();
                        if (lane == 0) {
                            if ((MODE == EPI_R1 || MODE == EPI_PLAIN) && args.red) tma_reduce_add_2d(&tmap_out, tile_s, gcol0, first_row);
                            else tma_store_2d(&tmap_out, tile_s, gcol0, first_row);
                        }
                    }
                    if (MODE == EPI_QKV) {
                        // operand preparation for the attention kernel: the v columns are also written transposed (V^T, keys
                        // contiguous = the K-major B operand of P.V); r[] holds the 32 finished f32 values of this row.
                        // (q / k are consumed straight out of `out`; every tf32 lo residual is computed on chip by attn_tc.cu.)
                        const int dmodel = N / 3;
                        if (gcol0 >= 2 * dmodel && row_ok) {   // 32 consecutive keys (lanes) per store instruction
                            const int bb = div_by_rps(row, rps, inv_rps), tt = row - bb * rps;
                            const int cv = gcol0 - 2 * dmodel;
                            float* vt = ep.vt + ((size_t)bb * (dmodel >> 7) * 128 + cv) * (size_t)ep.vt_tp + tt;
#pragma unroll
                            for (int idx = 0; idx < 32; ++idx) vt[(size_t)idx * ep.vt_tp] = __uint_as_float(r[idx]);
                            if (tt == rps - 1) {           // last key of the clip: zero the [T, Tp) padding of its V^T rows
                                for (int pd = 1; pd <= ep.vt_tp - rps; ++pd)
#pragma unroll
                                    for (int idx = 0; idx < 32; ++idx) vt[(size_t)idx * ep.vt_tp + pd] = 0.0f;
                            }
                        }
                    }
                } else {
                    __syncwarp();
                    // ---- phase 2 ----
                    const int col = gcol0 + lane;
                    if (vec_io) {
                        if (MODE == EPI_R12) {      // two residuals do not fit beside the accumulator registers: loaded here
                                                    // (r[] is dead), in two halves of 16 rows (4 KB per warp in flight)
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                                for (int it = hf * 4; it < hf * 4 + 4; ++it)
                                    if (vcol_ok && it * 4 + vr < nrows) {
                                        res1[it] = __ldg(reinterpret_cast<const float4*>(ep.add1 + vbase + (long long)it * 4 * N));
                                        res2[it] = __ldg(reinterpret_cast<const float4*>(ep.add2 + vbase + (long long)it * 4 * N));
                                    }
#pragma unroll
                                for (int it = hf * 4; it < hf * 4 + 4; ++it) {
                                    const int rr = it * 4 + vr;
                                    if (vcol_ok && rr < nrows) {
                                        const int4 raw = lds_v4(tile_s + (uint32_t)rr * 128u + (((uint32_t)vq ^ (uint32_t)(rr & 7)) << 4));
                                        float4 v = make_float4(__int_as_float(raw.x), __int_as_float(raw.y), __int_as_float(raw.z), __int_as_float(raw.w));
                                        v.x = __fadd_rn(v.x, res1[it].x); v.y = __fadd_rn(v.y, res1[it].y); v.z = __fadd_rn(v.z, res1[it].z); v.w = __fadd_rn(v.w, res1[it].w);
                                        v.x = __fadd_rn(res2[it].x, v.x); v.y = __fadd_rn(res2[it].y, v.y); v.z = __fadd_rn(res2[it].z, v.z); v.w = __fadd_rn(res2[it].w, v.w);
                                        *reinterpret_cast<float4*>(ep.out + vbase + (long long)it * 4 * N) = v;
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int rr = it * 4 + vr;
                                if (vcol_ok && rr < nrows) {
                                    const int4 raw = lds_v4(tile_s + (uint32_t)rr * 128u + (((uint32_t)vq ^ (uint32_t)(rr & 7)) << 4));
                                    float4 v = make_float4(__int_as_float(raw.x), __int_as_float(raw.y), __int_as_float(raw.z), __int_as_float(raw.w));
                                    if (HAS_R1) { v.x = __fadd_rn(v.x, res1[it].x); v.y = __fadd_rn(v.y, res1[it].y); v.z = __fadd_rn(v.z, res1[it].z); v.w = __fadd_rn(v.w, res1[it].w); }
                                    if (HAS_R2) { v.x = __fadd_rn(res2[it].x, v.x); v.y = __fadd_rn(res2[it].y, v.y); v.z = __fadd_rn(res2[it].z, v.z); v.w = __fadd_rn(res2[it].w, v.w); }
                                    *reinterpret_cast<float4*>(ep.out + vbase + (long long)it * 4 * N) = v;
                                }
                            }
                        }
                    } else if (col < N) {
                        const long long base = (long long)first_row * N + col;
                        float* outp = ep.out + base;
                        const float* a1 = (MODE == EPI_R1 || MODE == EPI_R12) ? ep.add1 + base : nullptr;
                        const float* a2 = (MODE == EPI_R2 || MODE == EPI_R12) ? ep.add2 + base : nullptr;
                        const uint32_t rd_s = tile_s + (uint32_t)(lane & 3) * 4u;
                        const uint32_t chunk_l = (uint32_t)(lane >> 2);
#pragma unroll 8
                        for (int rr = 0; rr < nrows; ++rr) {
                            float v = lds_f32(rd_s + (uint32_t)rr * 128u + ((chunk_l ^ (uint32_t)(rr & 7)) << 4));
                            const long long o = (long long)rr * N;
                            if (MODE == EPI_R1 || MODE == EPI_R12) v = __fadd_rn(v, __ldg(a1 + o));
                            if (MODE == EPI_R2 || MODE == EPI_R12) v = __fadd_rn(__ldg(a2 + o), v);
                            outp[o] = v;
                        }
                    }
                    __syncwarp();
                }
            }
            // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG2) mbar_arrive_rank0(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }   // (cta_group::2: the leader's
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }                                                                // MMA thread waits for both CTAs' epilogues)
            if (GEMM_DBG && args.dbg) busy += clock64() - t_busy0;

            if (MODE == EPI_QUANT && row_ok) {
                if (gcol_w < N) atomicAdd(ep.q_rowsum + row, q_sum);                       // 64 of the row's N columns
                if (n_blk == 0 && cgrp == 0) { ep.q_row_scale[row] = q_scale; ep.q_row_zp[row] = (int)q_zp; }
            }
            if (MODE == EPI_MINMAX && ep.q_rowsum && row_ok && n_blk == 0 && cgrp == 0) ep.q_rowsum[row] = 0;   // max-only pass: arm the next pass's row sums
            if (MODE == EPI_MINMAX && RELU && vmax > -FMAX) { vmin = 0.0f; vmax = fmaxf(vmax, 0.0f); }   // any value >= 0 gives amin = 0
            if (MODE == EPI_MINMAX && nrows > 0) {
                // rows of clip A (the one owning the warp's first row) and of clip A+1 reduce separately; the reduction runs
                // on the order-preserving integer keys the atomics use anyway (one REDUX per value instead of a 5-step shuffle tree)
                const unsigned kmin = lb_fkey(vmin), kmax = lb_fkey(vmax);
                const unsigned mnA = __reduce_min_sync(0xffffffffu, in_a ? kmin : LB_KEY_MIN_INIT), mxA = __reduce_max_sync(0xffffffffu, in_a ? kmax : LB_KEY_MAX_INIT);
                if (lane == 0 && mnA <= mxA) lb_mm_update_keys(ep.minmax_keys, slice_first, mnA, mxA);
                if (!all_a) {
                    const bool in_b = row_ok && !in_a;
                    const unsigned mnB = __reduce_min_sync(0xffffffffu, in_b ? kmin : LB_KEY_MIN_INIT), mxB = __reduce_max_sync(0xffffffffu, in_b ? kmax : LB_KEY_MAX_INIT);
                    if (lane == 0 && mnB <= mxB) lb_mm_update_keys(ep.minmax_keys, slice_first + 1, mnB, mxB);
                }
            }
            if (MODE == EPI_ARGMAX && row_ok && best_c >= 0) atomicMax(ep.argmax_keys + row, ((unsigned long long)lb_fkey_argmax(best_t) << 32) | (unsigned)best_c);
            m_idx += step_m; n_blk += step_n;
            if (n_blk >= nnb) { n_blk -= nnb; ++m_idx; }
            m_blk = m_idx * mmul + crank;
        }
        if (TMA_OUT && lane == 0) tma_store_wait_all();                // staging smem must outlive the bulk stores
        if (GEMM_DBG && args.dbg && (blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (ew == 0 || ew == 15))
            printf("GEMMDBG blk %d EPI warp %d: total %lld wait_accumulator %lld drain %lld (tmem ld %lld, math %lld)\n", blockIdx.x, ew, clock64() - t_begin, w_acc, busy, d_ld, d_math);
    }

    tc_fence_before();
    if (mc) cluster_sync_all();          // no CTA leaves while its peer can still signal its barriers
    else __syncthreads();
    if (GEMM_DBG && args.dbg && threadIdx.x == 0) {
        unsigned long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("GEMMCTA mode %d blk %d sm %u t0 %llu t1 %llu\n", MODE, blockIdx.x, smid, cta_t0, t1);
    }
    if (warp == 0) {
        __syncwarp();
        tc_fence_after();
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}


// =====================================================================================================================
// Fused "linear -> ReLU -> dynamic quantiser" (SenseVoice FFN1): ONE pass over the GEMM.
//
// The reference's dynamic quantiser needs the per-clip min / max of the whole [T, N] output before it can quantise the
// first element (avx/quantization.rs:112-140), so round 1 ran this GEMM twice (a max-only pass, then a quantising pass:
// 2.3 ms of every 23 ms step spent on MMAs whose results were thrown away).  Here the accumulators WAIT IN TMEM instead:
//   * the rows are walked in GROUPS of G whole clips; a group is one wave of the persistent grid (CTA = one 128-row block
//     x one 256-column block of the group, the same (m, n) block in every group), so a CTA's i-th tile belongs to group i;
//   * max pass  me(g): tcgen05.ld the finished accumulator, dequantise (exact integer correction, scale, bias), reduce
//     the per-clip min / max (warp REDUX -> shared-memory atomics -> one pair of global atomics per CTA and clip), write
//     the dequantised f32 values BACK into the same TMEM columns (tcgen05.st), then arrive on the clip's counter;
//   * quantising pass qe(g): once the counters of the warp's clips show every contributing CTA (acquire load), derive
//     (scale, zp) from the keys, tcgen05.ld the f32 values, quantise, TMA-store the u8 tile (the next GEMM's A operand)
//     and hand the accumulator stage back to the MMA warp.
//   The 16 epilogue warps run  me(0) me(1) qe(0) qe(1) me(2) me(3) ...  over the two TMEM stages, so the wait for the other
//   CTAs' maxima is covered by the next group's max pass and the MMAs of group g+2 run under qe(g+1) / me(g+2).
// Requires every CTA of the grid to be co-resident (grid <= #SMs, 1 CTA per SM by shared memory): the host refuses shapes
// that do not fit and the runner uses it only when nothing else shares the device (one lane).  Spins are bounded (trap).
// =====================================================================================================================
constexpr int FQ_TILE_BYTES = 2048;                            // per-warp 32 rows x 64 u8 staging tile (one TMA store per tile)
constexpr int FQ_META_BYTES = 64 * 4;                          // bias of the warp's 64 columns; the per-pass tables zcA | csA | zcB | csB (1 KB) live in the
                                                               // staging tile, which is idle during a max pass (its last TMA store is waited for first)
constexpr int FQ_MAX_SLOTS = BM / 32 + 1;                      // clips a 128-row tile can touch (T >= 32)
constexpr int FQ_EPI_BYTES = NUM_EPI_WARPS * (FQ_TILE_BYTES + FQ_META_BYTES);
constexpr int FQ_SMEM_BYTES = STAGES * STAGE_BYTES + FQ_EPI_BYTES + 1024 /*align slack*/ + 512 /*barriers, CTA key slots*/;
// K <= 512: the CTA's 256 x K weight block (the same in every group) is loaded ONCE and stays in shared memory; only the
// 16 KB activation tiles stream (64 KB instead of 192 KB of L2 -> SM traffic per tile, which is what paces the MMAs)
constexpr int FQ_RES_KB = 4;
constexpr int FQ_RES_SMEM_BYTES = FQ_RES_KB * B_STAGE_BYTES + STAGES * A_STAGE_BYTES + FQ_EPI_BYTES + 1024 + 512;
static_assert(FQ_SMEM_BYTES <= 232448 && FQ_RES_SMEM_BYTES <= 232448, "shared memory budget (fused quantiser)");

struct FusedArgs {
    int M, N, K;
    int nnb, nkb;
    int T; float inv_T;   // rows per clip
    int GT;               // rows per group (G clips)
    int n_groups, n_clips;
    int dbg;              // LELE_B200_GEMM_DBG=1 in a -DLELE_B200_GEMM_TIMELINE build: CTAs 0 / 77 print where their epilogue waited
    int prefetch;         // early look-up of the quantiser parameters (LELE_B200_FFN_PREFETCH=0: always the blocking wait)
    LbI8Epilogue ep;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory"); }

template <bool RELU, bool RESB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_i8_fused_q_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_out, const FusedArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;               // RESB: the FQ_RES_KB resident k-blocks of the weight block
    uint8_t* epi_base = smem_b + (RESB ? FQ_RES_KB : STAGES) * B_STAGE_BYTES;   // 16 staging tiles, then 16 metadata blocks
    uint64_t* bars = (uint64_t*)(epi_base + FQ_EPI_BYTES);
    uint64_t* full_bar = bars;                     // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;           // [STAGES] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * STAGES;       // [2]      MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;          // [2]      epilogue (quantising pass) -> MMA
    uint64_t* b_full = tmem_empty + 2;             // RESB: the resident weight block landed
    uint32_t* tmem_base_smem = (uint32_t*)(b_full + 1);
    unsigned* cta_keys = (unsigned*)(tmem_base_smem + 2);          // [2 parities][FQ_MAX_SLOTS][min, max]

    // Roles by LOGICAL warp index (0 = TMA producer, 1 = MMA issuer, 2.. = epilogue).  Physically the two single-thread control roles sit
    // in the LAST two warps of the CTA: the SM sub-partition's issue arbiter serves the highest warp id first, and a TMA / MMA issue that
    // waits behind four busy epilogue warps stalls the whole pipeline (build with -DLELE_B200_CTRL_WARPS_LOW for round 1's placement).
    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = lb_logical_warp(pw);
    const int nnb = args.nnb;
    const int mb = (int)blockIdx.x / nnb, nb = (int)blockIdx.x % nnb;     // the CTA's block of every group
    const int M = args.M, GT = args.GT;
    // groups in which this CTA's m-block has rows: all but possibly the last (shorter) one
    const int last_rows = M - (args.n_groups - 1) * GT;
    const int n_act = (mb * BM < last_rows) ? args.n_groups : args.n_groups - 1;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], NUM_EPI_WARPS); }
        for (int i = 0; i < 2 * FQ_MAX_SLOTS; ++i) { cta_keys[2 * i] = LB_KEY_MIN_INIT; cta_keys[2 * i + 1] = LB_KEY_MAX_INIT; }
        mbar_init(b_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    lb_pdl_launch_dependents();
    lb_pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (RESB && n_act > 0) {                               // weights do not depend on the previous kernel, but the wait above is cheap
                mbar_expect_tx(b_full, (uint32_t)args.nkb * B_STAGE_BYTES);
                for (int kb = 0; kb < args.nkb; ++kb) tma_load_2d(smem_b + kb * B_STAGE_BYTES, &tmap_b, b_full, kb * BK, nb * BN);
            }
            for (int g = 0; g < n_act; ++g) {
                const int row0 = g * GT + mb * BM;
                for (int kb = 0; kb < args.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], RESB ? A_STAGE_BYTES : STAGE_BYTES);
                    tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, row0);
                    if (!RESB) tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_b, &full_bar[stage], kb * BK, nb * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (RESB && n_act > 0) { mbar_wait(b_full, 0); tc_fence_after(); }
            for (int g = 0; g < n_act; ++g) {
                const int acc = g & 1; const uint32_t acc_phase = (uint32_t)(g >> 1) & 1u;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);            // the quantising pass of group g - 2 drained this stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < args.nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + (RESB ? kb : stage) * B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_i8(tmem_d, adesc + (uint64_t)(k * (UMMA_K >> 4)), bdesc + (uint64_t)(k * (UMMA_K >> 4)),
                                (kb > 0 || k > 0) ? 1u : 0u, idesc_for(true));
                    umma_commit(&empty_bar[stage]);
                    if (kb == args.nkb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: 16 warps, each owns 32 rows x 64 columns of the CTA's block =====================
        const int ew = warp - FIRST_EPI_WARP;
        const int quad = pw & 3;
        const int cgrp = ew >> 2;
        const LbI8Epilogue& ep = args.ep;
        const uint32_t tile_s = smem_u32(epi_base) + (uint32_t)ew * FQ_TILE_BYTES;
        const uint32_t meta_s = tile_s;                                // zcA | csA | zcB | csB during a max pass
        const uint32_t bias_s = smem_u32(epi_base) + NUM_EPI_WARPS * FQ_TILE_BYTES + (uint32_t)ew * FQ_META_BYTES;
        const int T = args.T; const float inv_T = args.inv_T;
        const int N = args.N;
        const int gcol_w = nb * BN + cgrp * 64;
        const bool cols_ok = gcol_w < N;                               // N % 64 == 0 (host): a warp's columns exist or not
        const float FMAX = 3.402823466e+38f;
        // column metadata: the CTA keeps its n-block, so the warp's 64 columns are the same in every group
        int cs_raw[2]; float ws_raw[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int cc = min(gcol_w + hh * 32 + lane, N - 1);
            cs_raw[hh] = __ldg(ep.colsum + cc); ws_raw[hh] = __ldg(ep.w_scale + cc);
            sts_f32(bias_s + 4 * (hh * 32 + lane), __ldg(ep.bias + cc));
        }
        long long w_mma = 0, w_cnt = 0, w_bar = 0, t_me = 0, t_qe = 0; const long long t_begin = clock64();
        struct Geo { int first_row, nrows, sl_a, clip0, gend, g0; bool row_ok, in_a, all_a; };
        auto geometry = [&](int g) {
            Geo q;
            q.g0 = g * GT; q.gend = min(q.g0 + GT, M);
            const int tile_r0 = q.g0 + mb * BM;
            q.first_row = tile_r0 + quad * 32;
            q.nrows = max(0, min(32, q.gend - q.first_row));
            q.row_ok = lane < q.nrows;
            q.clip0 = div_by_rps(tile_r0, T, inv_T);
            q.sl_a = div_by_rps(min(q.first_row, q.gend - 1), T, inv_T);
            q.in_a = q.first_row + lane < (q.sl_a + 1) * T;
            q.all_a = __all_sync(0xffffffffu, !q.row_ok || q.in_a);
            return q;
        };

        // the row's activation parameters are fetched one group ahead (no global-load latency between two passes)
        int pf_zpa = 0; float pf_sa = 0.0f;
        auto fetch_rows = [&](int g) {
            pf_zpa = 0; pf_sa = 0.0f;
            if (g >= n_act) return;
            const int row = g * GT + mb * BM + quad * 32 + lane;
            if (row < min(g * GT + GT, M)) { pf_zpa = __ldg(ep.row_zp + row); pf_sa = __ldg(ep.row_scale + row); }
        };
        fetch_rows(0);

        // The quantising pass of a group needs its clips' arrival counters (an acquire load) and then their key slots: two dependent
        // L2 round trips in front of every tile.  They are issued EARLY instead -- at the end of the pass that runs before it, when the
        // warp would otherwise sit at the CTA barrier and the group's last arrivals are a whole pass old -- and kept in registers
        // (pre_*); if the clips were not complete yet, the quantising pass falls back to the blocking wait.
        int pre_g = -1; float pre_scale = 0.0f, pre_zp = 0.0f, pre_inv = 0.0f;
        auto clip_expect = [&](const Geo& q, int cl) {         // CTAs that arrive for clip cl: nnb column blocks x the m-blocks the clip spans
            const int a = max(cl * T, q.g0) - q.g0, b = min((cl + 1) * T, q.gend) - q.g0;
            return (((b - 1) >> 7) - (a >> 7) + 1) * nnb;
        };
        auto clip_params = [&](const Geo& q, float& q_scale, float& q_zp, float& q_inv) {
            // lanes 0-15 hold the 8 (min, max) key slots of clip A, lanes 16-31 those of clip A+1
            const int sl = min(q.sl_a + (lane >> 4), args.n_clips - 1);
            unsigned k = __ldcg(ep.fq_keys + (size_t)sl * LB_MM_SLOTS * 2 + (lane & 15));
#pragma unroll
            for (int of = 2; of <= 8; of <<= 1) {
                const unsigned o = __shfl_xor_sync(0xffffffffu, k, of);
                k = (lane & 1) ? max(k, o) : min(k, o);
            }
            const int src = q.in_a ? 0 : 16;
            const float mn = lb_fkey_inv(__shfl_sync(0xffffffffu, k, src)), mx = lb_fkey_inv(__shfl_sync(0xffffffffu, k, src + 1));
            const float amax = fmaxf(mx, 0.0f), amin = fminf(mn, 0.0f);          // dq_params (quant.cu)
            q_scale = __fdiv_rn(fmaxf(__fsub_rn(amax, amin), 1e-5f), 255.0f);
            q_zp = fminf(fmaxf(roundf(__fdiv_rn(-amin, q_scale)), 0.0f), 255.0f);
            q_inv = __fdiv_rn(1.0f, q_scale);
        };
        auto prefetch_params = [&](int g) {
            pre_g = -1;
            if (g < 0 || g >= n_act || !args.prefetch) return;
            const Geo q = geometry(g);
            if (!(q.nrows > 0 && cols_ok)) return;
            int done = 1;
            if (lane == 0)
                for (int cl = q.sl_a; cl <= q.sl_a + (q.all_a ? 0 : 1); ++cl)
                    if (ld_acquire_gpu(ep.fq_counters + cl) < clip_expect(q, cl)) done = 0;
            if (!__shfl_sync(0xffffffffu, done, 0)) return;
            clip_params(q, pre_scale, pre_zp, pre_inv);
            pre_g = g;
        };

        // ---- max pass of group g ----
        auto max_pass = [&](int g, int prefetch_g) {
            const long long t_me0 = GEMM_DBG ? clock64() : 0;
            const int s = g & 1; const uint32_t ph = (uint32_t)(g >> 1) & 1u;
            const Geo q = geometry(g);
            const int zpa = pf_zpa; const float sa = pf_sa;
            fetch_rows(g + 1);
            const int lastl = max(q.nrows - 1, 0);
            const int zpa_a = __shfl_sync(0xffffffffu, zpa, 0), zpa_b = __shfl_sync(0xffffffffu, zpa, lastl);
            const float sa_a = __shfl_sync(0xffffffffu, sa, 0), sa_b = __shfl_sync(0xffffffffu, sa, lastl);
            if (lane == 0) tma_store_wait_read();                  // the tables share the staging tile: its last store must have read it
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                       // tables combined with the clip's activation parameters
                const int c = hh * 32 + lane;
                sts_s32(meta_s + 4 * c, zpa_a * cs_raw[hh]);
                sts_f32(meta_s + 256 + 4 * c, __fmul_rn(sa_a, ws_raw[hh]));
                sts_s32(meta_s + 512 + 4 * c, zpa_b * cs_raw[hh]);   // rows past a clip boundary read the second table
                sts_f32(meta_s + 768 + 4 * c, __fmul_rn(sa_b, ws_raw[hh]));
            }
            __syncwarp();
            const uint32_t my_meta = meta_s + (q.in_a ? 0u : 512u);
            float vmax = -FMAX, vmin = FMAX;
            if (q.nrows > 0 && cols_ok) {
                const long long t_w0 = GEMM_DBG ? clock64() : 0;
                mbar_wait(&tmem_full[s], ph);
                if (GEMM_DBG) w_mma += clock64() - t_w0;
                tc_fence_after();
#pragma unroll 1
                for (int chunk = 0; chunk < 2; ++chunk) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * BN + cgrp * 64 + chunk * 32);
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr, r);
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const int4 zc4 = lds_v4(my_meta + 16 * (chunk * 8 + qq));
                        const int4 cs4 = lds_v4(my_meta + 256 + 16 * (chunk * 8 + qq));
                        const int4 bi4 = lds_v4(bias_s + 16 * (chunk * 8 + qq));
                        const int zc[4] = {zc4.x, zc4.y, zc4.z, zc4.w};
                        const float cs[4] = {__int_as_float(cs4.x), __int_as_float(cs4.y), __int_as_float(cs4.z), __int_as_float(cs4.w)};
                        const float bi[4] = {__int_as_float(bi4.x), __int_as_float(bi4.y), __int_as_float(bi4.z), __int_as_float(bi4.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = qq * 4 + e;
                            float t = __fmul_rn((float)((int)r[idx] - zc[e]), cs[e]);
                            t = ep.has_bias ? __fadd_rn(t, bi[e]) : t;
                            vmax = fmaxf(vmax, t);
                            if (!RELU) vmin = fminf(vmin, t);
                            r[idx] = __float_as_uint(t);
                        }
                    }
                    tmem_st_32x32b_x32(taddr, r);                  // the dequantised values wait in TMEM for the clip's scale
                }
                tmem_st_wait();
                if (RELU) { vmin = 0.0f; vmax = fmaxf(vmax, 0.0f); }     // min / max of the ReLU output
                const unsigned kmin = lb_fkey(vmin), kmax = lb_fkey(vmax);
                unsigned* slot = cta_keys + ((size_t)(g & 1) * FQ_MAX_SLOTS + (q.sl_a - q.clip0)) * 2;
                const bool a_ok = q.row_ok && q.in_a;
                const unsigned mnA = __reduce_min_sync(0xffffffffu, a_ok ? kmin : LB_KEY_MIN_INIT), mxA = __reduce_max_sync(0xffffffffu, a_ok ? kmax : LB_KEY_MAX_INIT);
                if (lane == 0) { atomicMin(slot, mnA); atomicMax(slot + 1, mxA); }
                if (!q.all_a) {
                    const bool b_ok = q.row_ok && !q.in_a;
                    const unsigned mnB = __reduce_min_sync(0xffffffffu, b_ok ? kmin : LB_KEY_MIN_INIT), mxB = __reduce_max_sync(0xffffffffu, b_ok ? kmax : LB_KEY_MAX_INIT);
                    if (lane == 0) { atomicMin(slot + 2, mnB); atomicMax(slot + 3, mxB); }
                }
            }
            prefetch_params(prefetch_g);                           // (the previous group's quantiser parameters, see above)
            const long long t_b0 = GEMM_DBG ? clock64() : 0;
            epi_bar_sync();                                        // every warp's keys are in the CTA slots
            if (GEMM_DBG) { w_bar += clock64() - t_b0; t_me += clock64() - t_me0; }
            if (ew == 0 && lane < FQ_MAX_SLOTS) {
                const int tile_r0 = q.g0 + mb * BM, tile_end = min(tile_r0 + BM, q.gend);
                const int clip = q.clip0 + lane;
                if (tile_r0 < tile_end && clip * T < tile_end) {   // the block has rows of this clip: publish, then arrive
                    unsigned* slot = cta_keys + ((size_t)(g & 1) * FQ_MAX_SLOTS + lane) * 2;
                    const unsigned mn = slot[0], mx = slot[1];
                    slot[0] = LB_KEY_MIN_INIT; slot[1] = LB_KEY_MAX_INIT;
                    lb_mm_update_keys(ep.fq_keys, clip, mn, mx);
                    __threadfence();
                    atomicAdd(ep.fq_counters + clip, 1);
                }
            }
        };

        // ---- quantising pass of group g ----
        auto quant_pass = [&](int g, int prefetch_g) {
            const int s = g & 1;
            const Geo q = geometry(g);
            float q_inv = 0.0f, q_zp = 0.0f, q_scale = 0.0f;
            const long long t_qe0 = GEMM_DBG ? clock64() : 0;
            if (q.nrows > 0 && cols_ok) {
                if (pre_g == g) { q_scale = pre_scale; q_zp = pre_zp; q_inv = pre_inv; }
                else {
                    if (lane == 0) {
                        for (int cl = q.sl_a; cl <= q.sl_a + (q.all_a ? 0 : 1); ++cl) {
                            const int expect = clip_expect(q, cl);
                            if (ld_acquire_gpu(ep.fq_counters + cl) < expect) {
                                const long long t0 = clock64();
                                while (ld_acquire_gpu(ep.fq_counters + cl) < expect) {
                                    __nanosleep(64);
                                    if (clock64() - t0 > 4000000000ll) { printf("lele_b200 gemm_i8 fused quantiser: clip %d never completed (block %d)\n", cl, blockIdx.x); __trap(); }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    clip_params(q, q_scale, q_zp, q_inv);
                }
                if (GEMM_DBG) w_cnt += clock64() - t_qe0;
                if (q.nrows == 32) {                               // the previous tile's store has read the staging tile (long ago)
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
#pragma unroll 1
                for (int chunk = 0; chunk < 2; ++chunk) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * BN + cgrp * 64 + chunk * 32);
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr, r);
                    unsigned pk[8];
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const unsigned u0 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[qq * 4 + 0]), q_inv, q_zp));
                        const unsigned u1 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[qq * 4 + 1]), q_inv, q_zp));
                        const unsigned u2 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[qq * 4 + 2]), q_inv, q_zp));
                        const unsigned u3 = cvt_sat_u8(__fmaf_rn(__uint_as_float(r[qq * 4 + 3]), q_inv, q_zp));
                        pk[qq] = __byte_perm(__byte_perm(u0, u1, 0x0040), __byte_perm(u2, u3, 0x0040), 0x5410);
                    }
                    if (q.nrows == 32) {                           // staging: row = lane, 64 B per row
                        const uint32_t dst = tile_s + (uint32_t)lane * 64u + (uint32_t)chunk * 32u;
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16u), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                    } else if (q.row_ok) {                         // last rows of a group: the rows below belong to the next group
                        uint4* dst = reinterpret_cast<uint4*>(ep.q_out + (size_t)(q.first_row + lane) * N + gcol_w + chunk * 32);
                        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                if (q.nrows == 32) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&tmap_out, tile_s, gcol_w, q.first_row);     // 32 rows x 64 columns
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[s]);            // the MMA warp may overwrite this accumulator stage
            if (q.row_ok && nb == 0 && cgrp == 0) { ep.q_row_scale[q.first_row + lane] = q_scale; ep.q_row_zp[q.first_row + lane] = (int)q_zp; }
            prefetch_params(prefetch_g);
            if (GEMM_DBG) t_qe += clock64() - t_qe0;
        };

        for (int p = 0; p < n_act; p += 2) {                       // me(p) me(p+1) qe(p) qe(p+1): one copy of each pass in the instruction cache
            const int pe = min(p + 2, n_act);
#pragma unroll 1
            for (int g = p; g < pe; ++g) max_pass(g, g > p ? g - 1 : -1);        // me(p+1) ends with the look-up for qe(p)
#pragma unroll 1
            for (int g = p; g < pe; ++g) quant_pass(g, g + 1 < pe ? g + 1 : -1);  // qe(p) ends with the look-up for qe(p+1)
        }
        if (lane == 0) tma_store_wait_all();
        if (GEMM_DBG && args.dbg && (blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (ew == 0 || ew == 15))
            printf("FQDBG blk %d warp %d: total %lld | max passes %lld (wait MMA %lld, CTA barrier %lld) | quant passes %lld (wait clips %lld) | groups %d\n",
                   blockIdx.x, ew, clock64() - t_begin, t_me, w_mma, w_bar, t_qe, w_cnt, n_act);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// u8 [rows, cols] row-major, box = [box_rows, 128 B], 128B swizzle, OOB -> zero fill
int make_tmap_u8(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { lb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return LELE_B200_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lb_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)r, rows, cols); return LELE_B200_ERR_CUDA; }
    return LELE_B200_OK;
}

int cached_tmap_u8(lele_b200_ctx* ctx, CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
    const unsigned long long key[10] = {0x75386d61ull, (unsigned long long)(uintptr_t)ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)box_rows};
    if (lb_tmap_lookup(ctx, key, map)) return LELE_B200_OK;
    int rc = make_tmap_u8(map, ptr, rows, cols, box_rows);
    if (rc) return rc;
    lb_tmap_store(ctx, key, map);
    return LELE_B200_OK;
}

// f32 [rows, cols] row-major output, box = 32 x 32 (one epilogue warp's sub-tile), 128B swizzle
int cached_tmap_out_f32(lele_b200_ctx* ctx, CUtensorMap* map, const void* ptr, long long rows, long long cols, long long pitch_cols) {
    const unsigned long long key[10] = {0x6f757466ull, (unsigned long long)(uintptr_t)ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_cols};
    if (lb_tmap_lookup(ctx, key, map)) return LELE_B200_OK;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { lb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return LELE_B200_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_cols * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lb_set_error("cuTensorMapEncodeTiled(out) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols); return LELE_B200_ERR_CUDA; }
    lb_tmap_store(ctx, key, map);
    return LELE_B200_OK;
}
// u8 [rows, cols] row-major output, box = 32 x 32 (one epilogue warp's quantised sub-tile), no swizzle
int cached_tmap_out_u8(lele_b200_ctx* ctx, CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_cols = 32) {
    const unsigned long long key[10] = {0x6f757538ull, (unsigned long long)(uintptr_t)ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)box_cols};
    if (lb_tmap_lookup(ctx, key, map)) return LELE_B200_OK;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { lb_set_error("cuTensorMapEncodeTiled entry point unavailable"); return LELE_B200_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lb_set_error("cuTensorMapEncodeTiled(u8 out) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols); return LELE_B200_ERR_CUDA; }
    lb_tmap_store(ctx, key, map);
    return LELE_B200_OK;
}
// how many 2-CTA clusters of the GEMM kernel (one CTA per SM) the device keeps resident at once: GPCs with an odd number of SMs
// leave one SM without a partner, and a persistent grid must not exceed what is co-resident
int mc_resident_clusters(lele_b200_ctx* ctx) {
    static int cached[64] = {0};                      // per device; -1 = unavailable
    const int dev = ctx->device;
    if (dev < 0 || dev >= 64) return 0;
    if (cached[dev] != 0) return cached[dev] > 0 ? cached[dev] : 0;
    const void* fn = (const void*)gemm_i8_tc_kernel<EPI_PLAIN, true, false, true>;
    if (lb_func_smem(ctx, fn, SMEM_BYTES)) { cached[dev] = -1; return 0; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * ctx->num_sms); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); cached[dev] = -1; return 0; }
    if (n > ctx->num_sms / 2) n = ctx->num_sms / 2;
    cached[dev] = n;
    return n;
}
}  // namespace

int lb_gemm_i8_tc(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K, const LbI8Epilogue& ep) {
    LB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_i8_tc: empty problem");
    LB_REQUIRE(K % 16 == 0, "gemm_i8_tc: K=%d must be a multiple of 16 (TMA row pitch)", K);
    LB_REQUIRE((((uintptr_t)A | (uintptr_t)Wt) & 15) == 0, "gemm_i8_tc: operands must be 16-byte aligned");
    const bool af = ep.a_f32 != nullptr;               // fused input quantiser: A is produced in the kernel
    LB_REQUIRE(af || A, "gemm_i8_tc: no A operand");
    LB_REQUIRE(!af || lb_gemm_i8_afuse_supported(ctx, M, N, K, ep), "gemm_i8_tc: fused input quantiser not available for M=%d N=%d K=%d", M, N, K);
    LB_REQUIRE(!ep.minmax_keys || ep.rows_per_slice >= 32, "gemm_i8_tc: fused min/max needs rows_per_slice >= 32 (a warp's 32 rows may span at most two slices)");
    LB_REQUIRE(!ep.argmax_keys || (!ep.add1 && !ep.add2), "gemm_i8_tc: fused arg-max cannot be combined with residual adds");
    LB_REQUIRE(!ep.w_signed || ep.w_zp == 128, "gemm_i8_tc: a signed weight operand implies zero point 128");
    CUtensorMap ta, tb;
    int rc = cached_tmap_u8(ctx, &tb, Wt, N, K, BN);
    if (rc) return rc;
    if (af) ta = tb;                                   // (unused by that variant)
    else if ((rc = cached_tmap_u8(ctx, &ta, A, M, K, BM))) return rc;
    KernelArgs args;
    args.M = M; args.N = N; args.K = K;
    args.num_m_blocks = lb_ceil_div(M, BM);
    args.num_n_blocks = lb_ceil_div(N, BN);
    args.num_k_blocks = lb_ceil_div(K, BK);
    args.ep = ep;
    args.rps = ep.rows_per_slice > 0 ? ep.rows_per_slice : 0x7fffffff;
    args.inv_rps = 1.0f / (float)args.rps;
    LB_REQUIRE(M < (1 << 22), "gemm_i8_tc: M=%d exceeds the 2^22 rows the epilogue's slice arithmetic is exact for", M);
    args.dbg = getenv("LELE_B200_GEMM_DBG") ? 1 : 0;
    args.vec_io = (N % 4 == 0) && ((((uintptr_t)ep.out | (uintptr_t)ep.add1 | (uintptr_t)ep.add2) & 15) == 0) && !getenv("LELE_B200_GEMM_NO_VEC_IO");
    int tiles = args.num_m_blocks * args.num_n_blocks;
    int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
    // In-place residual (x = x + (linear [+ add1]): add2 == out): the tile is reduce-added into `out` by the TMA engine / L2, so the
    // residual stream is neither loaded into nor stored from registers; what is left is the plain (or + add1) epilogue with a TMA store.
    const bool tma_ok = N % 4 == 0 && (((uintptr_t)ep.out | (uintptr_t)ep.add1) & 15) == 0 && !getenv("LELE_B200_GEMM_NO_TMA_STORE");
    const bool r1_flag = lb_env_flag("LELE_B200_GEMM_R1_TMA", 1);
    const bool red = ep.out && ep.add2 == ep.out && tma_ok && !ep.relu && !ep.minmax_keys && !ep.argmax_keys && !ep.q_out && !ep.vt &&
                     (!ep.add1 || r1_flag) && lb_env_flag("LELE_B200_GEMM_RED", 1);
    const bool r1_tma = ep.out && ep.add1 && (!ep.add2 || red) && tma_ok && r1_flag;
    args.red = red ? 1 : 0;
    int mode = EPI_PLAIN;
    if (ep.q_out) mode = EPI_QUANT;
    else if (ep.vt) mode = EPI_QKV;
    else if (ep.argmax_keys) mode = EPI_ARGMAX;
    else if (ep.minmax_keys) mode = EPI_MINMAX;
    else if (ep.add1 && ep.add2 && !red) mode = EPI_R12;
    else if (ep.add1) mode = EPI_R1;
    else if (ep.add2 && !red) mode = EPI_R2;
    LB_REQUIRE(!(ep.minmax_keys && (ep.add1 || ep.add2)), "gemm_i8_tc: fused min/max with residual adds is not instantiated");
    LB_REQUIRE(ep.out || mode == EPI_ARGMAX || mode == EPI_MINMAX || mode == EPI_QUANT, "gemm_i8_tc: no output requested");
    LB_REQUIRE(!ep.relu || mode == EPI_PLAIN || mode == EPI_MINMAX || mode == EPI_QUANT, "gemm_i8_tc: ReLU is only instantiated for the plain / min-max / quantising epilogues");
    if (mode == EPI_QUANT) {
        LB_REQUIRE(N % 32 == 0 && ep.q_rowsum && ep.q_row_scale && ep.q_row_zp && ep.q_keys && ep.rows_per_slice >= 32 && !ep.add1 && !ep.add2 &&
                   !ep.minmax_keys && !ep.argmax_keys && !ep.out && (((uintptr_t)ep.q_out) & 15) == 0,
                   "gemm_i8_tc: fused output quantiser needs N %% 32 == 0, rows_per_slice >= 32 and the q_* buffers, and excludes the other fusions");
    }
    // residual-free epilogues whose rows are 16-byte aligned store through TMA (no per-element store phase)
    if (mode == EPI_QKV) {
        LB_REQUIRE(N % 384 == 0 && ep.vt && ep.rows_per_slice > 0 && ep.vt_tp >= ep.rows_per_slice && !ep.add1 && !ep.add2 &&
                   !ep.minmax_keys && !ep.argmax_keys && !getenv("LELE_B200_GEMM_NO_TMA_STORE"),
                   "gemm_i8_tc: fused attention-operand epilogue needs N = 3 * heads * 128 and the V^T buffers");
    }
    const bool tma_out = ((mode == EPI_PLAIN || mode == EPI_MINMAX || mode == EPI_QKV) && ep.out && N % 4 == 0 && (((uintptr_t)ep.out) & 15) == 0 &&
                          !getenv("LELE_B200_GEMM_NO_TMA_STORE")) || (mode == EPI_R1 && r1_tma);
    LB_REQUIRE(!red || tma_out, "gemm_i8_tc: the in-place residual epilogue needs the TMA store path");
    CUtensorMap tout = ta, tlo = ta;
    if (mode == EPI_QUANT) {
        rc = cached_tmap_out_u8(ctx, &tout, ep.q_out, M, N);
        if (rc) return rc;
    }
    if (tma_out) {
        // QKV projection: the v third of the row is consumed only through V^T (attention) -- the tensor map stops at 2d columns when
        // the caller says nobody reads v from `out` (ep.skip_v_out), so those TMA stores are clipped by the hardware and never reach HBM
        rc = cached_tmap_out_f32(ctx, &tout, ep.out, M, (mode == EPI_QKV && ep.skip_v_out) ? (N / 3) * 2 : N, N);
        if (rc) return rc;
    }
    if (mode == EPI_QKV) {
        LB_REQUIRE(tma_out, "gemm_i8_tc: fused attention-operand epilogue needs a 16-byte aligned output");
    }
    if (mode == EPI_R1 && tma_out) {          // add1's sub-tiles are fetched by TMA into the epilogue warps' staging tiles
        rc = cached_tmap_out_f32(ctx, &tlo, ep.add1, M, N, N);
        if (rc) return rc;
    }
    // 2-CTA clusters along M that share each weight tile by TMA multicast (32 instead of 48 KB per CTA and k-block over the L2 -> SM
    // fabric).  Built for FFN2 (K = 2048), whose MMA thread waits for operands 40 % of the time -- measured: no gain (30.2 vs 31.4 us per
    // launch), because the limiter is the SM's own shared-memory bandwidth, not the fabric: an SS-mode 128 x 256 x 32 int8 MMA reads 12 KB
    // of operands per 128 cycles (96 B/clk) while TMA writes another 96 B/clk, against 128 B/clk per SM.  Multicast does not change either
    // figure (cta_group::2 would: half of B per CTA).  Opt-in: LELE_B200_GEMM_MC=1 (every shape), results are bit-identical.
    args.mc = 0;
    {
        const char* e = getenv("LELE_B200_GEMM_MC");
        // cta_group::2 pair MMA (256 x 256 per cluster, 4-stage ring of 32 KB): LELE_B200_GEMM_CG2=1, plain / TMA-store epilogue only.
        // Measured on FFN2 (K = 2048): 31.1 us per launch with and without, step 19.30 vs 19.37 ms -- like the multicast variant it
        // changes nothing, so neither the L2 -> SM fabric nor shared-memory bandwidth paces that mainloop (both variants cut them by a
        // third); what remains is per-k-block latency of the single producer thread (2 TMA issues + barrier round trip ~ 560 cycles,
        // role timeline in profiles/r02c_gemm_role_timelines.log).  Opt-in; bit-identical (switch test).
        const char* e2 = getenv("LELE_B200_GEMM_CG2");
        const bool cg2_inst = mode == EPI_PLAIN && tma_out && ep.w_signed && !ep.relu;   // = the cta_group::2 instantiation below
        const bool want2 = cg2_inst && e2 && e2[0] != '0';
        const bool want = (e && e[0] != '0') || want2;
        const int pair_tiles = ((args.num_m_blocks + 1) / 2) * args.num_n_blocks;
        if (want && !af && args.num_m_blocks >= 2) {
            const int ncl = mc_resident_clusters(ctx);
            if (ncl >= 1) {
                args.mc = want2 ? 2 : 1;
                grid = 2 * (pair_tiles < ncl ? pair_tiles : ncl);
                if ((rc = cached_tmap_u8(ctx, &tb, Wt, N, K, BN / 2))) return rc;      // half-tile boxes: each CTA of a pair fetches one
            }
        }
    }
    // the shared-memory opt-in is recorded per context (= per device), once per instantiation
#define LB_LAUNCH_MODE4(MD, TM, RL, WS)                                                                                 \
    {                                                                                                                   \
        if ((rc = lb_func_smem(ctx, (const void*)gemm_i8_tc_kernel<MD, TM, RL, WS>, SMEM_BYTES))) return rc;            \
        LB_CHECK_CUDA(lb_launch_pdl(gemm_i8_tc_kernel<MD, TM, RL, WS>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, ctx->stream, args.mc ? 2 : 1, ta, tb, tout, tlo, args)); \
    }
#define LB_LAUNCH_MODE3(MD, TM, RL) { if (ep.w_signed) LB_LAUNCH_MODE4(MD, TM, RL, true) else LB_LAUNCH_MODE4(MD, TM, RL, false) }
#define LB_LAUNCH_MODE(MD, TM) LB_LAUNCH_MODE3(MD, TM, false)
#define LB_LAUNCH_RELU(MD, TM) { if (ep.relu) LB_LAUNCH_MODE3(MD, TM, true) else LB_LAUNCH_MODE3(MD, TM, false) }
    switch (mode) {
        case EPI_PLAIN:
            if (args.mc == 2) {
                if ((rc = lb_func_smem(ctx, (const void*)gemm_i8_tc_kernel<EPI_PLAIN, true, false, true, true>, SMEM_BYTES))) return rc;
                LB_CHECK_CUDA(lb_launch_pdl(gemm_i8_tc_kernel<EPI_PLAIN, true, false, true, true>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, ctx->stream, 2, ta, tb, tout, tlo, args));
            } else if (tma_out) LB_LAUNCH_RELU(EPI_PLAIN, true) else LB_LAUNCH_RELU(EPI_PLAIN, false)
            break;
        case EPI_MINMAX: if (tma_out) LB_LAUNCH_RELU(EPI_MINMAX, true) else LB_LAUNCH_RELU(EPI_MINMAX, false) break;
        case EPI_ARGMAX: LB_LAUNCH_MODE(EPI_ARGMAX, false) break;
        case EPI_R1:
            if (af) {
                LB_REQUIRE(tma_out && args.mc == 0, "gemm_i8_tc: the fused input quantiser needs the TMA-store epilogue");
                grid = args.num_m_blocks;              // one m-block per CTA, its A block resident
                if ((rc = lb_func_smem(ctx, (const void*)gemm_i8_tc_kernel<EPI_R1, true, false, true, false, true>, SMEM_BYTES))) return rc;
                LB_CHECK_CUDA(lb_launch_pdl(gemm_i8_tc_kernel<EPI_R1, true, false, true, false, true>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, ctx->stream, 1, ta, tb, tout, tlo, args));
            } else if (tma_out) LB_LAUNCH_MODE(EPI_R1, true) else LB_LAUNCH_MODE(EPI_R1, false)
            break;
        case EPI_R2: LB_LAUNCH_MODE(EPI_R2, false) break;
        case EPI_R12: LB_LAUNCH_MODE(EPI_R12, false) break;
        case EPI_QKV: LB_LAUNCH_MODE(EPI_QKV, true) break;
        case EPI_QUANT: LB_LAUNCH_RELU(EPI_QUANT, true) break;
    }
#undef LB_LAUNCH_MODE
#undef LB_LAUNCH_MODE3
#undef LB_LAUNCH_MODE4
#undef LB_LAUNCH_RELU
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}

bool lb_gemm_i8_afuse_supported(lele_b200_ctx* ctx, long long M, int N, int K, const LbI8Epilogue& ep) {
    // opt-in (LELE_B200_GEMM_AFUSE=1): measured 30.6 us against 24.4 + 12.7 us for the separate quantiser launch, but the replayed step does
    // not move (19.22 vs 19.19 ms: the small launch costs ~6 us in-step, what the fused block adds to the GEMM), so the default keeps the
    // GEMM class free of quantiser work
    if (!lb_env_flag("LELE_B200_GEMM_AFUSE", 0) || getenv("LELE_B200_FORCE_SIMT") || getenv("LELE_B200_GEMM_NO_TMA_STORE")) return false;
    if (!ep.a_f32 || !ep.a_keys || !ep.w_signed || ep.relu || !ep.out || !ep.add1 || (ep.add2 && ep.add2 != ep.out)) return false;
    if (ep.minmax_keys || ep.argmax_keys || ep.q_out || ep.vt || ep.fq_keys || ep.rows_per_slice < 32) return false;
    if (K % BK != 0 || K > AF_KB * BK || N % 4 != 0 || M <= 0 || M > (long long)ctx->num_sms * BM || M >= (1 << 22)) return false;
    if ((((uintptr_t)ep.a_f32 | (uintptr_t)ep.out | (uintptr_t)ep.add1) & 15) != 0) return false;
    if (ep.add2 && !lb_env_flag("LELE_B200_GEMM_RED", 1)) return false;
    return lb_env_flag("LELE_B200_GEMM_R1_TMA", 1);
}

// The fused linear -> (ReLU) -> dynamic quantiser GEMM (gemm_i8_fused_q_kernel).  Needs the s8 weight operand, whole clips
// of T >= 32 rows, N % 64 == 0 and a grid (one group of clips per wave) that fits the device.
bool lb_gemm_i8_fused_q_supported(lele_b200_ctx* ctx, long long M, int N, int K, int T, int w_signed) {
    if (!w_signed || T < 32 || M <= 0 || M % T != 0 || N % 64 != 0 || K % 16 != 0 || M >= (1 << 22)) return false;
    const int nnb = lb_ceil_div(N, BN);
    const int mb_fit = ctx->num_sms / nnb;                 // m-blocks of one group that fit the grid
    return mb_fit >= 1 && (long long)mb_fit * BM >= T;     // at least one whole clip per group
}
int lb_gemm_i8_tc_fused_q(lele_b200_ctx* ctx, const uint8_t* A, const uint8_t* Wt, int M, int N, int K, const LbI8Epilogue& ep) {
    const int T = ep.rows_per_slice;
    LB_REQUIRE(lb_gemm_i8_fused_q_supported(ctx, M, N, K, T, ep.w_signed), "gemm_i8 fused quantiser: unsupported shape M=%d N=%d K=%d T=%d", M, N, K, T);
    LB_REQUIRE(ep.q_out && ep.q_row_scale && ep.q_row_zp && ep.fq_keys && ep.fq_counters && (((uintptr_t)ep.q_out) & 15) == 0,
               "gemm_i8 fused quantiser: q_out / q_row_scale / q_row_zp / fq_keys / fq_counters are required");
    LB_REQUIRE((((uintptr_t)A | (uintptr_t)Wt) & 15) == 0, "gemm_i8 fused quantiser: operands must be 16-byte aligned");
    CUtensorMap ta, tb, tout;
    int rc = cached_tmap_u8(ctx, &ta, A, M, K, BM);
    if (rc) return rc;
    if ((rc = cached_tmap_u8(ctx, &tb, Wt, N, K, BN))) return rc;
    if ((rc = cached_tmap_out_u8(ctx, &tout, ep.q_out, M, N, 64))) return rc;
    FusedArgs args;
    args.M = M; args.N = N; args.K = K;
    args.nnb = lb_ceil_div(N, BN); args.nkb = lb_ceil_div(K, BK);
    args.T = T; args.inv_T = 1.0f / (float)T;
    args.n_clips = M / T;
    int G = (int)(((long long)(ctx->num_sms / args.nnb) * BM) / T);     // whole clips per group (= per wave of the grid)
    if (G > args.n_clips) G = args.n_clips;
    args.GT = G * T;
    args.n_groups = lb_ceil_div(args.n_clips, G);
    args.ep = ep;
    args.dbg = getenv("LELE_B200_GEMM_DBG") ? 1 : 0;
    args.prefetch = lb_env_flag("LELE_B200_FFN_PREFETCH", 0) ? 1 : 0;   // measured: 49.5 -> 51.1 us per launch (the extra acquire loads cost more than the hidden latency): opt-in
    const int grid = lb_ceil_div(args.GT, BM) * args.nnb;
    LB_REQUIRE(grid <= ctx->num_sms, "gemm_i8 fused quantiser: grid %d exceeds the %d SMs (every CTA must be resident)", grid, ctx->num_sms);
    const bool resb = args.nkb <= FQ_RES_KB && lb_env_flag("LELE_B200_FFN_RESB", 1);
#define LB_LAUNCH_FQ(RL, RB)                                                                                                  \
    {                                                                                                                         \
        const int smem_bytes = RB ? FQ_RES_SMEM_BYTES : FQ_SMEM_BYTES;                                                        \
        if ((rc = lb_func_smem(ctx, (const void*)gemm_i8_fused_q_kernel<RL, RB>, smem_bytes))) return rc;                     \
        LB_CHECK_CUDA(lb_launch_pdl(gemm_i8_fused_q_kernel<RL, RB>, dim3(grid), dim3(NUM_THREADS), smem_bytes, ctx->stream, 1, ta, tb, tout, args)); \
    }
    if (ep.relu) { if (resb) LB_LAUNCH_FQ(true, true) else LB_LAUNCH_FQ(true, false) }
    else         { if (resb) LB_LAUNCH_FQ(false, true) else LB_LAUNCH_FQ(false, false) }
#undef LB_LAUNCH_FQ
    LB_LAUNCH_CHECK(ctx);
    return LELE_B200_OK;
}
