// Probe: does the ADDRESS PATTERN of the fused kernel (each SM owns batch columns b, b+148, ...; its 25 warps
// read / write the K = 50 rows [k][b][0..X) of that column, 3136 contiguous bytes each, 3.2 MB apart) cost HBM
// bandwidth against a linear sweep of the same bytes?  No shared memory, no synchronisation: only the pattern.
// Dev tool, GPU box only.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int K = 50, X = 784, X4 = X / 4;
constexpr int64_t B = 1024;

// mode bit0: read, bit1: write.  column-major walk as in the kernel.
template <int MODE>
__global__ void k_cols(const float4* __restrict__ src, float4* __restrict__ dst, float* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    float acc = 0.f;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        for (int k = warp; k < K; k += NW) {
            const int64_t row = ((int64_t)k * B + b) * X4;
            float4 q[7];
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int v = lane + 32 * u;
                if (MODE & 1) q[u] = __ldcs(src + row + (v < X4 ? v : lane));
                else q[u] = make_float4((float)v, 1.f, 2.f, 3.f);
            }
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int v = lane + 32 * u;
                if (MODE & 2) {
                    if (v < X4) __stcs(dst + row + v, q[u]);
                } else {
                    acc += q[u].x + q[u].y + q[u].z + q[u].w;
                }
            }
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}

// linear walk: chunk c of 3136 bytes goes to warp (c mod total warps): neighbouring warps, neighbouring chunks
template <int MODE>
__global__ void k_linear(const float4* __restrict__ src, float4* __restrict__ dst, float* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const int64_t nchunks = (int64_t)K * B;
    float acc = 0.f;
    for (int64_t c = (int64_t)blockIdx.x * NW + warp; c < nchunks; c += (int64_t)gridDim.x * NW) {
        const int64_t row = c * X4;
        float4 q[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            const int v = lane + 32 * u;
            if (MODE & 1) q[u] = __ldcs(src + row + (v < X4 ? v : lane));
            else q[u] = make_float4((float)v, 1.f, 2.f, 3.f);
        }
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            const int v = lane + 32 * u;
            if (MODE & 2) {
                if (v < X4) __stcs(dst + row + v, q[u]);
            } else {
                acc += q[u].x + q[u].y + q[u].z + q[u].w;
            }
        }
    }
    if (acc == 12345.f) sink[0] = acc;
}

template <typename F>
static void run(const char* name, F launch, double bytes) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / reps;
    printf("%-44s %7.1f us  %7.1f GB/s  %s\n", name, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int64_t n4 = (int64_t)K * B * X4;
    float4 *src, *dst;
    float* sink;
    cudaMalloc(&src, n4 * 16);
    cudaMalloc(&dst, n4 * 16);
    cudaMalloc(&sink, 4);
    cudaMemset(src, 0, n4 * 16);
    const double one = (double)n4 * 16;
    for (int warps : {25, 32}) {
        for (int ctas : {1, 2}) {
            const int grid = 148 * ctas, threads = warps * 32;
            char name[96];
            snprintf(name, sizeof name, "columns  read+write  %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_cols<3><<<grid, threads>>>(src, dst, sink); }, 2 * one);
            snprintf(name, sizeof name, "linear   read+write  %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_linear<3><<<grid, threads>>>(src, dst, sink); }, 2 * one);
            snprintf(name, sizeof name, "columns  write only  %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_cols<2><<<grid, threads>>>(src, dst, sink); }, one);
            snprintf(name, sizeof name, "linear   write only  %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_linear<2><<<grid, threads>>>(src, dst, sink); }, one);
            snprintf(name, sizeof name, "columns  read only   %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_cols<1><<<grid, threads>>>(src, dst, sink); }, one);
            snprintf(name, sizeof name, "linear   read only   %d CTA/SM x %d warps", ctas, warps);
            run(name, [&] { k_linear<1><<<grid, threads>>>(src, dst, sink); }, one);
        }
    }
    run("cudaMemcpyAsync D2D", [&] { cudaMemcpyAsync(dst, src, n4 * 16, cudaMemcpyDeviceToDevice); }, 2 * one);
    return 0;
}
