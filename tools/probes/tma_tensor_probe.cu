// Probe: per-SM throughput of TENSOR bulk copies (cp.async.bulk.tensor.3d, one box = all K rows of a slice of
// one batch column of probs[K][B][X]) against the 1-D row copies of tma_probe.cu.  Dev tool, GPU box only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_tensor_probe tma_tensor_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// CTA b walks columns b, b+grid, ...; a column is NBOX boxes of {inner, 1, K}.  Warp 0 lane 0 produces into a
// ring of NSLOT box slots (full/empty mbarriers), the other warps consume (touch or not) and release.
template <bool TOUCH>
__global__ void k_tensor(const __grid_constant__ CUtensorMap map, int K, int64_t B, int inner, int nbox, int nslot, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = (blockDim.x >> 5) - 1;
    const uint32_t box_bytes = (uint32_t)K * inner * 4;
    const uint32_t slot_bytes = (box_bytes + 1023) & ~1023u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)nslot * slot_bytes);
    uint64_t* empty = full + nslot;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nslot; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ncols = ((int64_t)blockIdx.x < B) ? (B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t ntask = ncols * nbox;
    if (warp == NW) {
        if (lane == 0) {
            for (int64_t t = 0; t < ntask; ++t) {
                const int s = (int)(t % nslot);
                const uint32_t n = (uint32_t)(t / nslot);
                if (n > 0) mbar_wait(&empty[s], (n - 1) & 1);
                const int64_t col = blockIdx.x + (t / nbox) * gridDim.x;
                const int q = (int)(t % nbox);
                mbar_expect_tx(&full[s], box_bytes);
                tma_load_3d(smem + (size_t)s * slot_bytes, &map, q * inner, (int)col, 0, &full[s]);
            }
        }
        return;
    }
    float acc = 0.f;
    for (int64_t t = 0; t < ntask; ++t) {
        const int s = (int)(t % nslot);
        const uint32_t n = (uint32_t)(t / nslot);
        mbar_wait(&full[s], n & 1);
        if (TOUCH) {
            const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * slot_bytes);
            const int n16 = box_bytes / 16;
            for (int v = warp * 32 + lane; v < n16; v += NW * 32) {
                float4 q = p[v];
                acc += q.x + q.y + q.z + q.w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 12345.f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 50, X = 784;
    const int64_t B = 1024;
    const int64_t total = (int64_t)K * B * X * 4;
    char* src;
    float* sink;
    cudaMalloc(&src, total);
    cudaMalloc(&sink, 4);
    cudaMemset(src, 0, total);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeFn encode = (EncodeFn)fn;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    struct Cfg { int grid, nw, inner, nslot, touch; };
    Cfg cfgs[] = {{148, 8, 112, 9, 0}, {148, 8, 196, 5, 0}, {148, 8, 112, 9, 1}, {148, 25, 112, 9, 1}, {37, 8, 112, 9, 0},
                  {37, 8, 196, 5, 0},  {148, 8, 112, 4, 0}, {148, 8, 56, 16, 0}, {37, 8, 56, 16, 0},   {74, 8, 112, 9, 0}};
    for (auto& c : cfgs) {
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)B, (cuuint64_t)K};
        cuuint64_t strides[2] = {(cuuint64_t)X * 4, (cuuint64_t)B * X * 4};
        cuuint32_t box[3] = {(cuuint32_t)c.inner, 1, (cuuint32_t)K};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int nbox = X / c.inner;
        const uint32_t box_bytes = K * c.inner * 4, slot_bytes = (box_bytes + 1023) & ~1023u;
        const size_t smem = (size_t)c.nslot * slot_bytes + c.nslot * 16 + 64;
        auto kern = c.touch ? k_tensor<true> : k_tensor<false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int i = 0; i < 3; ++i) kern<<<c.grid, (c.nw + 1) * 32, smem>>>(map, K, B, c.inner, nbox, c.nslot, sink);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; ++i) kern<<<c.grid, (c.nw + 1) * 32, smem>>>(map, K, B, c.inner, nbox, c.nslot, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        const double us = ms * 1e3 / reps;
        printf("tensor grid %3d warps %2d box {%3d,1,%d} = %5u B x %2d slots touch %d smem %6zu: %7.1f us %7.1f GB/s (%5.1f GB/s per SM) %s\n",
               c.grid, c.nw, c.inner, K, box_bytes, c.nslot, c.touch, smem, us, total / us / 1e3, total / us / 1e3 / c.grid,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
