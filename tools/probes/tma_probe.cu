// Probe: per-SM throughput of 1-D bulk async copies (cp.async.bulk + mbarrier) global -> shared, as used by
// the ring kernel, against plain 128-bit loads.  Dev tool, GPU box only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// every warp owns D slots; task t of warp w in CTA b reads chunk (t * gridDim.x * NW + b * NW + w) of `bytes`
template <bool TOUCH>
__global__ void k_tma(const char* __restrict__ src, int64_t nchunks, int bytes, int stride, int D, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const int slot_bytes = (bytes + 127) & ~127;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NW * D * slot_bytes);
    if (lane == 0)
        for (int i = 0; i < D; ++i) mbar_init(&bars[warp * D + i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t per = (int64_t)gridDim.x * NW;
    const int64_t first = (int64_t)blockIdx.x * NW + warp;
    int64_t next = first;
    int pos = 0;
    auto issue = [&]() {
        if (next >= nchunks) return;
        const int s = warp * D + pos;
        mbar_expect_tx(&bars[s], bytes);
        bulk_load(smem + (size_t)s * slot_bytes, src + next * (int64_t)stride, bytes, &bars[s]);
        next += per;
        if (++pos == D) pos = 0;
    };
    if (lane == 0)
        for (int i = 0; i < D; ++i) issue();
    int cpos = 0;
    uint32_t phase = 0;
    float acc = 0.f;
    for (int64_t c = first; c < nchunks; c += per) {
        const int s = warp * D + cpos;
        mbar_wait(&bars[s], phase);
        if (++cpos == D) { cpos = 0; phase ^= 1; }
        if (TOUCH) {
            const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * slot_bytes);
            for (int v = lane; v < bytes / 16; v += 32) {
                float4 q = p[v];
                acc += q.x + q.y + q.z + q.w;
            }
        }
        __syncwarp();
        if (lane == 0) issue();
    }
    if (acc == 12345.f) sink[0] = acc;
}

// same chunks with plain loads: a warp streams its chunk with 128-bit loads, 8 in flight per lane
__global__ void k_ldg(const char* __restrict__ src, int64_t nchunks, int bytes, int stride, float* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const int64_t per = (int64_t)gridDim.x * NW;
    float acc = 0.f;
    const int n16 = bytes / 16;
    for (int64_t c = (int64_t)blockIdx.x * NW + warp; c < nchunks; c += per) {
        const float4* p = reinterpret_cast<const float4*>(src + c * (int64_t)stride);
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int v = lane + 32 * u;
            q[u] = __ldcs(p + (v < n16 ? v : lane));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += q[u].x + q[u].y + q[u].z + q[u].w;
    }
    if (acc == 12345.f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int64_t total = (int64_t)160563200;  // 50 * 1024 * 784 * 4
    char* src;
    float* sink;
    cudaMalloc(&src, total + 4096);
    cudaMalloc(&sink, 4);
    cudaMemset(src, 0, total + 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    struct Cfg { int grid, nw, D, bytes, stride, touch; };
    Cfg cfgs[] = {
        {148, 25, 2, 3136, 3136, 0}, {148, 25, 2, 3136, 3136, 1}, {148, 25, 2, 3072, 3072, 0}, {148, 25, 2, 3200, 3200, 0},
        {148, 25, 1, 3136, 3136, 0}, {148, 12, 4, 3136, 3136, 0}, {148, 25, 2, 1568, 1568, 0}, {148, 16, 2, 6272, 6272, 0},
        {37, 25, 2, 3136, 3136, 0},  {37, 25, 2, 3072, 3072, 0},  {74, 25, 2, 3136, 3136, 0},  {148, 8, 2, 12544, 12544, 0},
        {148, 4, 2, 25088, 25088, 0}, {37, 4, 2, 25088, 25088, 0},
    };
    for (auto& c : cfgs) {
        const int64_t nchunks = total / c.stride;
        const int slot = (c.bytes + 127) & ~127;
        const size_t smem = (size_t)c.nw * c.D * slot + (size_t)c.nw * c.D * 8 + 128;
        auto kern = c.touch ? k_tma<true> : k_tma<false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int i = 0; i < 3; ++i) kern<<<c.grid, c.nw * 32, smem>>>(src, nchunks, c.bytes, c.stride, c.D, sink);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; ++i) kern<<<c.grid, c.nw * 32, smem>>>(src, nchunks, c.bytes, c.stride, c.D, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        const double us = ms * 1e3 / reps;
        printf("tma  grid %3d warps %2d D %d bytes %5d touch %d smem %6zu: %7.1f us  %7.1f GB/s  (%5.1f GB/s per SM) %s\n", c.grid,
               c.nw, c.D, c.bytes, c.touch, smem, us, nchunks * (double)c.bytes / us / 1e3,
               nchunks * (double)c.bytes / us / 1e3 / c.grid, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    for (int grid : {148, 296, 592, 37}) {
        for (int nw : {8, 16, 32}) {
            const int64_t nchunks = total / 3136;
            for (int i = 0; i < 3; ++i) k_ldg<<<grid, nw * 32>>>(src, nchunks, 3136, 3136, sink);
            cudaEventRecord(e0);
            const int reps = 20;
            for (int i = 0; i < reps; ++i) k_ldg<<<grid, nw * 32>>>(src, nchunks, 3136, 3136, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1e3 / reps;
            printf("ldg  grid %3d warps %2d: %7.1f us  %7.1f GB/s\n", grid, nw, us, total / us / 1e3);
        }
    }
    return 0;
}
