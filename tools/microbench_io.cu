// Micro-benchmark (GPU box): per-SM ingest / egress ceilings that bound the GEMM epilogue and the TMA pipelines.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench_io tools/microbench_io.cu && gpurun_out/microbench_io
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_stg128(float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)i);
    for (; i < n4; i += st) out[i] = v;
}
__global__ void k_stg128_reps(float4* out, size_t n4, int reps) {
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        float4 v = make_float4(1.f, 2.f, 3.f, (float)r);
        for (size_t i = i0; i < n4; i += st) out[i] = v;
    }
}
// each warp writes 4 rows x 128 B per instruction with a 2 KB row pitch (the residual-epilogue store pattern)
__global__ void k_stg128_rows(float4* out, size_t rows, int reps) {
    const int lane = threadIdx.x & 31;
    const size_t w0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (int r = 0; r < reps; ++r) {
        float4 v = make_float4(1.f, 2.f, 3.f, (float)r);
        for (size_t blk = w0; blk < rows / 4 * 4; blk += nw) {          // blk -> (row group of 4, 128 B column chunk)
            const size_t rg = blk / 4, cc = blk % 4;                       // 4 column chunks of 128 B per 512 B ... pitch 2 KB = 16 chunks; use 4
            out[((rg * 4 + (lane >> 3)) * 2048 + cc * 128 + (lane & 7) * 16) / 16] = v;
        }
    }
}
__global__ void k_ldg128(const float4* in, size_t n4, float* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i + 3 * st < n4; i += 4 * st) {
        float4 a = in[i], b = in[i + st], c = in[i + 2 * st], d = in[i + 3 * st];
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) *sink = acc;
}
// bulk smem -> global stores: each warp owns a 4 KB staging tile and streams it out repeatedly (the GEMM epilogue pattern)
__global__ void __launch_bounds__(512) k_bulk_store(uint8_t* out, size_t bytes_per_cta, int chunk, int reps = 1) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t* mine = sm + (size_t)warp * chunk;
    for (int i = lane * 16; i < chunk; i += 32 * 16) *reinterpret_cast<float4*>(mine + i) = make_float4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint8_t* base = out + (size_t)blockIdx.x * bytes_per_cta;
    if (lane == 0) {
        for (int rp = 0; rp < reps; ++rp)
        for (size_t off = (size_t)warp * chunk; off + chunk <= bytes_per_cta; off += (size_t)nw * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"((uint32_t)__cvta_generic_to_shared(mine)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
// bulk global -> smem loads, `depth` chunks in flight per CTA (one producer thread, the TMA mainloop pattern)
__global__ void __launch_bounds__(128) k_bulk_load(const uint8_t* in, size_t bytes_per_cta, int chunk, int depth, int reps = 1) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t* base = in + (size_t)blockIdx.x * bytes_per_cta;
        const size_t n1 = bytes_per_cta / chunk, n = n1 * reps;
        for (size_t i = 0; i < n + depth; ++i) {
            const int s = (int)(i % depth);
            const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]);
            if (i >= (size_t)depth) {   // wait for the load issued `depth` iterations ago
                const uint32_t par = (uint32_t)(((i / depth) - 1) & 1);
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            }
            if (i < n) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(sm + (size_t)s * chunk)), "l"(base + (i % n1) * chunk), "r"(chunk), "r"(b) : "memory");
            }
        }
    }
}

// tensor-map loads (SWIZZLE_128B, 128-byte inner box) -- the GEMM operand pattern: per k-block one A box (128 rows) + one B box (256 rows)
__global__ void __launch_bounds__(128) k_tensor_load(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int kblocks, int tiles, int depth, int a_rows, int b_rows) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[16];
    uint8_t* smb = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (a_rows + b_rows) * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t n = (size_t)tiles * kblocks;
        for (size_t i = 0; i < n + depth; ++i) {
            const int s = (int)(i % depth);
            const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]);
            if (i >= (size_t)depth) {
                const uint32_t par = (uint32_t)(((i / depth) - 1) & 1);
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            }
            if (i < n) {
                const int t = (int)(i / kblocks), kb = (int)(i % kblocks);
                const int row_a = ((blockIdx.x * tiles + t) * a_rows) % 16384, row_b = ((blockIdx.x * 7 + t) % 8) * b_rows;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(stage_bytes) : "memory");
                uint8_t* dst = smb + (size_t)s * stage_bytes;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(&ta), "r"(b), "r"(kb * 128), "r"(row_a) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(dst + a_rows * 128)), "l"(&tb), "r"(b), "r"(kb * 128), "r"(row_b) : "memory");
            }
        }
    }
}
// same pattern, but the A box is issued by warp 0 and the B box by warp 1 (each arms its own mbarrier)
__global__ void __launch_bounds__(128) k_tensor_load2(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int kblocks, int tiles, int depth, int a_rows, int b_rows) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[2][16];
    uint8_t* smb = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (a_rows + b_rows) * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) for (int w = 0; w < 2; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[w][i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < 2) {
        const size_t n = (size_t)tiles * kblocks;
        const int my_bytes = (w == 0 ? a_rows : b_rows) * 128;
        for (size_t i = 0; i < n + depth; ++i) {
            const int s = (int)(i % depth);
            const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[w][s]);
            if (i >= (size_t)depth) {
                const uint32_t par = (uint32_t)(((i / depth) - 1) & 1);
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            }
            if (i < n) {
                const int t = (int)(i / kblocks), kb = (int)(i % kblocks);
                const int row_a = ((blockIdx.x * tiles + t) * a_rows) % 16384, row_b = ((blockIdx.x * 7 + t) % 8) * b_rows;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(my_bytes) : "memory");
                uint8_t* dst = smb + (size_t)s * stage_bytes;
                if (w == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(&ta), "r"(b), "r"(kb * 128), "r"(row_a) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(dst + a_rows * 128)), "l"(&tb), "r"(b), "r"(kb * 128), "r"(row_b) : "memory");
            }
        }
    }
}
// issue-only cost: N tensor loads back to back into the same stage, one wait at the end
__global__ void __launch_bounds__(128) k_tensor_issue(const __grid_constant__ CUtensorMap ta, int n, int a_rows, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* smb = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * a_rows * 128) : "memory");
        const long long t0 = clock64();
        for (int i = 0; i < n; ++i)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(smb + (size_t)(i % 4) * a_rows * 128)), "l"(&ta), "r"(b), "r"((i % 16) * 128), "r"((int)blockIdx.x * a_rows) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    }
}
// nreq requests of `rows` rows each per stage, all from map ta (box rows = `rows`)
__global__ void __launch_bounds__(128) k_tensor_loadN(const __grid_constant__ CUtensorMap ta, int iters, int depth, int nreq, int rows) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[16];
    uint8_t* smb = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = nreq * rows * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < iters + depth; ++i) {
            const int s = i % depth;
            const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[s]);
            if (i >= depth) {
                const uint32_t par = (uint32_t)(((i / depth) - 1) & 1);
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            }
            if (i < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(stage_bytes) : "memory");
                for (int q = 0; q < nreq; ++q) {
                    const int row = ((blockIdx.x * 64 + i * nreq + q) * rows) % 16384;
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(smb + (size_t)s * stage_bytes + (size_t)q * rows * 128)), "l"(&ta), "r"(b), "r"((i % 16) * 128), "r"(row) : "memory");
                }
            }
        }
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(CUtensorMap* m, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
    cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols}; cuuint32_t box[2] = {128, box_rows}; cuuint32_t es[2] = {1, 1};
    return ((EncodeTiledFn)p)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main() {
    const size_t bytes = (size_t)148 * 2 * 1024 * 1024;   // 296 MB > L2
    uint8_t *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 1, bytes));
    float* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch, size_t nbytes) {
        for (int i = 0; i < 2; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        cudaError_t e = cudaGetLastError();
        printf("%-60s %8.1f us  %7.2f TB/s  %6.1f B/clk/SM@1.9GHz %s\n", name, ms * 1e3, nbytes / ms / 1e9, nbytes / (ms * 1e-3) / 148 / 1.9e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    for (size_t sz : {(size_t)36 << 20, (size_t)144 << 20, bytes}) {
        printf("--- %zu MB ---\n", sz >> 20);
        time("STG.128 grid 148x8 x256", [&] { k_stg128<<<148 * 8, 256>>>((float4*)a, sz / 16); }, sz);
        time("LDG.128 grid 148x8 x256 (4 in flight)", [&] { k_ldg128<<<148 * 8, 256>>>((const float4*)b, sz / 16, sink); }, sz);
        const size_t per_cta = sz / 148 / 65536 * 65536;
        cudaFuncSetAttribute(k_bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192);
        time("bulk store 4 KB x16 warps, 1 CTA/SM", [&] { k_bulk_store<<<148, 512, 16 * 4096>>>(a, per_cta, 4096); }, per_cta * 148);
        time("bulk store 8 KB x16 warps, 1 CTA/SM", [&] { k_bulk_store<<<148, 512, 16 * 8192>>>(a, per_cta, 8192); }, per_cta * 148);
        time("bulk store 4 KB x8 warps, 1 CTA/SM", [&] { k_bulk_store<<<148, 256, 8 * 4096>>>(a, per_cta, 4096); }, per_cta * 148);
        cudaFuncSetAttribute(k_bulk_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        time("bulk load 16 KB x3 deep, 1 CTA/SM", [&] { k_bulk_load<<<148, 128, 3 * 16384>>>(b, per_cta, 16384, 3); }, per_cta * 148);
        time("bulk load 16 KB x9 deep, 1 CTA/SM", [&] { k_bulk_load<<<148, 128, 9 * 16384>>>(b, per_cta, 16384, 9); }, per_cta * 148);
        time("bulk load 32 KB x6 deep, 1 CTA/SM", [&] { k_bulk_load<<<148, 128, 6 * 32768>>>(b, per_cta, 32768, 6); }, per_cta * 148);
        time("bulk load 16 KB x12 deep, 1 CTA/SM", [&] { k_bulk_load<<<148, 128, 12 * 16384>>>(b, per_cta, 16384, 12); }, per_cta * 148);
    }
    // L2-resident re-read: 36 MB read repeatedly (the weights / x case)
    printf("--- 36 MB, warm L2 (read twice back to back) ---\n");
    time("bulk load 16 KB x9 deep, warm", [&] { k_bulk_load<<<148, 128, 9 * 16384>>>(b, ((size_t)36 << 20) / 148 / 65536 * 65536, 16384, 9); }, ((size_t)36 << 20) / 148 / 65536 * 65536 * 148);
    time("LDG.128 warm", [&] { k_ldg128<<<148 * 8, 256>>>((const float4*)b, ((size_t)36 << 20) / 16, sink); }, (size_t)36 << 20);
    {
        const size_t pc = ((size_t)36 << 20) / 148 / 65536 * 65536; const int R = 16;
        printf("--- 36 MB region, %d passes inside one kernel (L2-resident) ---\n", R);
        time("bulk load 16 KB x3 deep", [&] { k_bulk_load<<<148, 128, 3 * 16384>>>(b, pc, 16384, 3, R); }, pc * 148 * R);
        time("bulk load 16 KB x9 deep", [&] { k_bulk_load<<<148, 128, 9 * 16384>>>(b, pc, 16384, 9, R); }, pc * 148 * R);
        time("bulk load 32 KB x6 deep", [&] { k_bulk_load<<<148, 128, 6 * 32768>>>(b, pc, 32768, 6, R); }, pc * 148 * R);
        time("bulk load 48 KB x3 deep", [&] { k_bulk_load<<<148, 128, 3 * 49152>>>(b, pc / 49152 * 49152, 49152, 3, R); }, pc / 49152 * 49152 * 148 * R);
        time("bulk load 48 KB x4 deep", [&] { k_bulk_load<<<148, 128, 4 * 49152>>>(b, pc / 49152 * 49152, 49152, 4, R); }, pc / 49152 * 49152 * 148 * R);
        time("bulk load 64 KB x3 deep", [&] { k_bulk_load<<<148, 128, 3 * 65536>>>(b, pc, 65536, 3, R); }, pc * 148 * R);
        time("bulk load 96 KB x2 deep", [&] { k_bulk_load<<<148, 128, 2 * 98304>>>(b, pc / 98304 * 98304, 98304, 2, R); }, pc / 98304 * 98304 * 148 * R);
        {   // tensor-map operand pattern: A [16384+, 2048] u8 (32 MB), B [2048, 2048] u8 (4 MB): K = 2048 -> 16 k-blocks per tile
            CUtensorMap ta, tb;
            cudaFuncSetAttribute(k_tensor_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (!make_map(&ta, b, 16384 + 256, 2048, 128) && !make_map(&tb, a, 2048, 2048, 256)) {
                const int tiles = 8, kb = 16;
                time("tensor load A128+B256 rows x128B swizzle, 3 stages", [&] { k_tensor_load<<<148, 128, 3 * 49152 + 1024>>>(ta, tb, kb, tiles, 3, 128, 256); }, (size_t)148 * tiles * kb * 49152);
                time("tensor load A128+B256 rows x128B swizzle, 4 stages", [&] { k_tensor_load<<<148, 128, 4 * 49152 + 1024>>>(ta, tb, kb, tiles, 4, 128, 256); }, (size_t)148 * tiles * kb * 49152);
                cudaFuncSetAttribute(k_tensor_load2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                time("tensor load A128+B256, 3 stages, A and B issued by 2 warps", [&] { k_tensor_load2<<<148, 128, 3 * 49152 + 1024>>>(ta, tb, kb, tiles, 3, 128, 256); }, (size_t)148 * tiles * kb * 49152);
                cudaFuncSetAttribute(k_tensor_loadN, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                for (int rows : {64, 128, 256}) {
                    CUtensorMap tr; make_map(&tr, b, 16384 + 256, 2048, rows);
                    for (int nreq : {1, 2, 3, 4}) {
                        const int sb = nreq * rows * 128; if (sb * 3 > 196608) continue;
                        char nm[128]; snprintf(nm, sizeof nm, "tensorN: %d req x %d rows (%d KB stage), 3 stages", nreq, rows, sb >> 10);
                        time(nm, [&] { k_tensor_loadN<<<148, 128, 3 * sb + 1024>>>(tr, 256, 3, nreq, rows); }, (size_t)148 * 256 * sb);
                    }
                }
                CUtensorMap tb2; make_map(&tb2, a, 2048, 2048, 128);
                time("tensor load A128+B128, 5 stages, 2 warps", [&] { k_tensor_load2<<<148, 128, 5 * 32768 + 1024>>>(ta, tb2, kb, tiles, 5, 128, 128); }, (size_t)148 * tiles * kb * 32768);
                {
                    long long* cyc; cudaMalloc(&cyc, 16); long long h[2];
                    cudaFuncSetAttribute(k_tensor_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                    for (int n : {1, 4, 16}) {
                        k_tensor_issue<<<148, 128, 4 * 16384 + 1024>>>(ta, n, 128, cyc); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
                        k_tensor_issue<<<148, 128, 4 * 16384 + 1024>>>(ta, n, 128, cyc); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
                        printf("issue %2d x 16 KB tensor loads (148 CTAs): issue loop %lld cyc, until landed %lld cyc\n", n, h[0], h[1]);
                    }
                }
                time("tensor load A128+B128 rows x128B swizzle, 5 stages", [&] { k_tensor_load<<<148, 128, 5 * 32768 + 1024>>>(ta, tb2, kb, tiles, 5, 128, 128); }, (size_t)148 * tiles * kb * 32768);
            } else printf("tensor map creation failed\n");
        }
        time("STG.128 contiguous, 148x4 x512, L2-resident", [&] { k_stg128_reps<<<148 * 4, 512>>>((float4*)a, ((size_t)36 << 20) / 16, R); }, ((size_t)36 << 20) * R);
        time("STG.128 contiguous, 148x1 x512, L2-resident", [&] { k_stg128_reps<<<148, 512>>>((float4*)a, ((size_t)36 << 20) / 16, R); }, ((size_t)36 << 20) * R);
        time("STG.128 contiguous, 148x2 x1024, L2-resident", [&] { k_stg128_reps<<<148 * 2, 1024>>>((float4*)a, ((size_t)36 << 20) / 16, R); }, ((size_t)36 << 20) * R);
        time("bulk store 4 KB x16 warps", [&] { k_bulk_store<<<148, 512, 16 * 4096>>>(a, pc, 4096, R); }, pc * 148 * R);
        time("bulk store 16 KB x8 warps", [&] { k_bulk_store<<<148, 256, 8 * 16384>>>(a, pc, 16384, R); }, pc * 148 * R);
        time("bulk store 32 KB x4 warps", [&] { k_bulk_store<<<148, 128, 4 * 32768>>>(a, pc, 32768, R); }, pc * 148 * R);
        time("bulk store 1 KB x16 warps", [&] { k_bulk_store<<<148, 512, 16 * 1024>>>(a, pc, 1024, R); }, pc * 148 * R);
        time("bulk store 4 KB x8 warps", [&] { k_bulk_store<<<148, 256, 8 * 4096>>>(a, pc, 4096, R); }, pc * 148 * R);
        time("bulk store 8 KB x16 warps", [&] { k_bulk_store<<<148, 512, 16 * 8192>>>(a, pc, 8192, R); }, pc * 148 * R);
    }
    return 0;
}
