// Probe: per-SM throughput of the three ways an SM can write dprobs[K][B][X] (one batch column = K rows of X floats
// per step): (a) 128-bit register stores (STG.128, with and without the .cs hint), (b) 3-D tensor bulk stores of a
// box {inner, 1, K} from shared memory issued by one thread, (c) 1-D bulk stores of whole rows issued per warp.
// The fused kernel's phase B (resident column -> dprobs) runs at ~17 B/clk per SM whether 37 or 148 CTAs are
// active; this probe tells whether that is the STG path's own limit.  Dev tool, GPU box only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_probe store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int HINT>  // 0: st.global.v4, 1: st.global.cs.v4, 2: st.global.wt? (L1 no allocate) -> use .cg
__device__ __forceinline__ void stg(float* p, float4 v) {
    if (HINT == 0) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (HINT == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (HINT == 2) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// (a) CTA b writes columns b, b+grid, ...; warp w writes rows w, w+NW, ... of each column, 7 float4 per lane and row
template <int HINT>
__global__ void k_stg(float* out, int K, int64_t B, int X) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const int X4 = X / 4;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x)
        for (int k = warp; k < K; k += NW) {
            float* row = out + ((int64_t)k * B + b) * X;
            const float4 v = make_float4((float)k, (float)b, 1.f, 2.f);
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const int i = lane + 32 * u;
                if (i < X4) stg<HINT>(row + 4 * i, v);
            }
        }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// (b) one thread per CTA stores NBOX boxes per column from NSLOT rotating slots, at most DEPTH groups reading smem
__global__ void k_tensor_store(const __grid_constant__ CUtensorMap map, int K, int64_t B, int inner, int nbox, int nslot) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t box_bytes = (uint32_t)K * inner * 4, slot_bytes = (box_bytes + 127) & ~127u;
    for (int i = threadIdx.x; i < (int)(nslot * slot_bytes / 4); i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int64_t b = blockIdx.x; b < B; b += gridDim.x)
            for (int q = 0; q < nbox; ++q) {
                tma_store_3d(&map, q * inner, (int)b, 0, smem + (size_t)s * slot_bytes);
                bulk_commit();
                bulk_wait_read<2>();  // the slot three stores back is free again
                if (++s == nslot) s = 0;
            }
        bulk_wait_all<0>();
    }
}

// (c) every warp stores its rows with 1-D bulk copies from its own two row slots
__global__ void k_bulk1d_store(float* out, int K, int64_t B, int X) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)X * 4;
    for (int i = threadIdx.x; i < NW * 2 * X; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (lane == 0) {
        int s = 0;
        for (int64_t b = blockIdx.x; b < B; b += gridDim.x)
            for (int k = warp; k < K; k += NW) {
                float* row = out + ((int64_t)k * B + b) * X;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row),
                             "r"(smem_u32(smem + (size_t)(warp * 2 + s) * row_bytes)), "r"(row_bytes)
                             : "memory");
                bulk_commit();
                bulk_wait_read<1>();
                s ^= 1;
            }
        bulk_wait_all<0>();
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 50, X = 784;
    const int64_t B = 1024;
    const int64_t total = (int64_t)K * B * X * 4;
    float* dst;
    cudaMalloc(&dst, total);
    cudaMemset(dst, 0, total);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    EncodeFn encode = (EncodeFn)fn;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    auto report = [&](const char* name, int grid, float ms, int reps) {
        const double us = ms * 1e3 / reps;
        const double per_sm = total / us / 1e3 / grid;
        printf("%-44s grid %3d: %7.1f us  %7.1f GB/s  %5.1f GB/s per SM  %5.1f B/clk per SM @%d MHz  %s\n", name, grid, us,
               total / us / 1e3, per_sm, per_sm * 1e3 / (clk_khz / 1e3), clk_khz / 1000, cudaGetErrorString(cudaGetLastError()));
    };
    const int reps = 10;
    float ms;
    for (int grid : {37, 148}) {
        for (int nw : {25, 8}) {
#define RUN_STG(H, NAME)                                                                   \
    for (int i = 0; i < 2; ++i) k_stg<H><<<grid, nw * 32>>>(dst, K, B, X);                   \
    cudaEventRecord(e0);                                                                   \
    for (int i = 0; i < reps; ++i) k_stg<H><<<grid, nw * 32>>>(dst, K, B, X);                \
    cudaEventRecord(e1);                                                                   \
    cudaEventSynchronize(e1);                                                              \
    cudaEventElapsedTime(&ms, e0, e1);                                                     \
    { char nm[64]; snprintf(nm, 64, "%s, %d warps", NAME, nw); report(nm, grid, ms, reps); }
            RUN_STG(0, "STG.128")
            RUN_STG(1, "STG.128 .cs")
            RUN_STG(2, "STG.128 .cg")
        }
        for (int inner : {112, 196, 392}) {
            CUtensorMap map;
            cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)B, (cuuint64_t)K};
            cuuint64_t strides[2] = {(cuuint64_t)X * 4, (cuuint64_t)B * X * 4};
            cuuint32_t box[3] = {(cuuint32_t)inner, 1, (cuuint32_t)K};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int nbox = X / inner, nslot = 4;
            const uint32_t slot_bytes = ((uint32_t)K * inner * 4 + 127) & ~127u;
            const size_t smem = (size_t)nslot * slot_bytes;
            if (smem > 227 * 1024) continue;
            cudaFuncSetAttribute(k_tensor_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int i = 0; i < 2; ++i) k_tensor_store<<<grid, 128, smem>>>(map, K, B, inner, nbox, nslot);
            cudaEventRecord(e0);
            for (int i = 0; i < reps; ++i) k_tensor_store<<<grid, 128, smem>>>(map, K, B, inner, nbox, nslot);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            char nm[64];
            snprintf(nm, 64, "tensor store box {%d,1,%d}, 1 thread", inner, K);
            report(nm, grid, ms, reps);
        }
        {
            const int nw = 25;
            const size_t smem = (size_t)nw * 2 * X * 4;
            cudaFuncSetAttribute(k_bulk1d_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int i = 0; i < 2; ++i) k_bulk1d_store<<<grid, nw * 32, smem>>>(dst, K, B, X);
            cudaEventRecord(e0);
            for (int i = 0; i < reps; ++i) k_bulk1d_store<<<grid, nw * 32, smem>>>(dst, K, B, X);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            report("1-D bulk store of 3136 B rows, 25 warps", grid, ms, reps);
        }
    }
    return 0;
}
