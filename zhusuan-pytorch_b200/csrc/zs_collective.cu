// Peer-memory all-reduce (SUM, float32) over NVLink / NVSwitch: the one exchange of the data-parallel hot path
// (the scalar objective and the replicated networks' parameter gradients).
//
// The reference has no distributed code; a data-parallel user of it would call torch.distributed.all_reduce.  For
// the 5 MB gradient buffer of the example VAE an NCCL all-reduce costs ~30 us per call on a B200 box -- launch and
// protocol latency, not bandwidth -- next to a ~85 us step.  This kernel is the exchange written for that size:
//   * every rank's buffer lives in peer-mapped ("symmetric") device memory, all ranks pass the same table of base
//     pointers;
//   * rank r owns slice r of the buffer: it loads slice r of EVERY rank's buffer over NVLink (all loads of an
//     element in flight together), adds them up in rank order (the result is bit-identical on all ranks, whatever
//     the world size) and stores the sum into slice r of every rank's buffer -- a reduce-scatter and an all-gather
//     in one pass, each byte crossing NVLink once in each direction;
//   * two flag barriers per CTA (peers' data ready / peers' pushes landed), system-scope flags;
//     a CTA of rank r only ever talks to the same-numbered CTA of the other ranks, so nothing is grid-wide.  Flags
//     carry a per-CTA epoch kept in device memory, so a captured CUDA graph can replay the launch.
// No NCCL kernel, no staging copy, no host involvement.
#include "zs_common.cuh"

namespace zs {

struct PeerTable {
    float* buf[ZS_MAX_PEERS];
    unsigned* flags[ZS_MAX_PEERS];
};

// Flags are written and polled with relaxed system-scope accesses; the ordering they need is provided around them:
//   barrier 0 publishes nothing (what the peers read was written by EARLIER kernels of this stream, complete and
//             visible at the home L2 before this kernel started) and the data loads that follow are volatile, i.e. they
//             are served by the owner's L2, never by a stale local cache line;
//   barrier 1 the signalling threads fence at system scope after the CTA barrier that follows the pushes (the pushes
//             must have landed before the flag does); what was pushed into this rank's buffer is read by LATER
//             kernels of this stream.
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The scalar that rides along (the step's objective): buf[extra_index] = *extra_src on this rank before anybody reads
// the buffer, so the caller needs no copy kernel to put it into the bucket.  Every CTA's thread 0 stores the same
// value ahead of the CTA barrier that precedes its ready flags (whichever CTA of a peer reads the element has then
// seen it); the fence orders the store before those flags at system scope.
__device__ __forceinline__ void put_extra(float* own, const float* extra_src, long long extra_index) {
    if (extra_src == nullptr) return;
    const float v = *extra_src;
    asm volatile("st.volatile.global.f32 [%0], %1;" ::"l"(own + extra_index), "f"(v) : "memory");
    __threadfence_system();
}

// flag words of one rank: [set][phase 0|1][cta][source rank], then one epoch word per (set, cta)
__host__ __device__ inline size_t flag_index(int set, int phase, int cta, int src) {
    return (((size_t)set * 2 + phase) * ZS_PEER_MAX_CTAS + cta) * ZS_MAX_PEERS + src;
}
__host__ __device__ inline size_t epoch_index(int set, int cta) {
    return (size_t)ZS_PEER_FLAG_SETS * 2 * ZS_PEER_MAX_CTAS * ZS_MAX_PEERS + (size_t)set * ZS_PEER_MAX_CTAS + cta;
}

// WORLD ranks (compile time: 2, 4 or 8; a smaller world leaves table entries unused), U float4 per thread and rank in
// flight.  The first version (one float4 per thread and rank per iteration, 32 CTAs) took ~30 us for 2.7 MB at
// N = 2: every iteration is an NVLink round trip (~2 us) and a thread went through ~20 of them one after the other.
// Here WORLD x U = 16 loads are issued before the first use and 64 CTAs x 512 threads cover a 2.7 MB slice in ONE
// iteration: the launch costs two flag barriers plus about one round trip.
template <int WORLD, int U>
__global__ void __launch_bounds__(ZS_PEER_THREADS)
    k_allreduce_peer(const __grid_constant__ PeerTable tab, int rank, int world, long long first, long long count,
                     int set, const float* __restrict__ extra_src, long long extra_index) {
    const int cta = blockIdx.x, tid = threadIdx.x;
    unsigned* mine = tab.flags[rank];
    __shared__ unsigned s_epoch;
    if (tid == 0) {
        s_epoch = mine[epoch_index(set, cta)] + 1u;
        put_extra(tab.buf[rank], extra_src, extra_index);
    }
    __syncthreads();
    const unsigned e = s_epoch;
    // ---- barrier 0: my earlier work on this stream is done (this kernel is running); so is every peer's
    if (tid < world && tid != rank) {
        st_flag(tab.flags[tid] + flag_index(set, 0, cta, rank), e);
        while ((int)(ld_flag(mine + flag_index(set, 0, cta, tid)) - e) < 0) {}
    }
    __syncthreads();
    // ---- slice `rank` of [first, first + count): float4 units, the last slice takes the remainder
    const long long units = count >> 2;                   // count % 4 == 0 (checked by the launcher)
    const long long per = (units + world - 1) / world;
    const long long lo = (long long)rank * per, hi = lo + per < units ? lo + per : units;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)cta * blockDim.x + tid; i0 < hi; i0 += stride * U) {
        float4 v[WORLD][U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
#pragma unroll
            for (int p = 0; p < WORLD; ++p)
                if (p < world && i < hi) v[p][u] = ld_peer(tab.buf[p] + first + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                float4 s = v[0][u];
#pragma unroll
                for (int p = 1; p < WORLD; ++p)
                    if (p < world) { s.x += v[p][u].x; s.y += v[p][u].y; s.z += v[p][u].z; s.w += v[p][u].w; }
#pragma unroll
                for (int p = 0; p < WORLD; ++p)
                    if (p < world) st_peer(tab.buf[p] + first + 4 * i, s);
            }
        }
    }
    // ---- barrier 1: my pushes are visible everywhere; every peer's pushes into my buffer are too.  The CTA barrier
    // orders every thread's pushes before the signalling threads' system-scope fence (fences are cumulative), so the
    // other ~500 threads do not each pay for one.
    __syncthreads();
    if (tid < world && tid != rank) {
        __threadfence_system();
        st_flag(tab.flags[tid] + flag_index(set, 1, cta, rank), e);
        while ((int)(ld_flag(mine + flag_index(set, 1, cta, tid)) - e) < 0) {}
    }
    __syncthreads();
    if (tid == 0) mine[epoch_index(set, cta)] = e;
}

// ---- NVSwitch multicast (NVLS) variant ------------------------------------------------------------------------
// With the buffer bound to a multicast object (one address that names the same offset in EVERY rank's memory) the
// switch does the work: multimem.ld_reduce makes the switch fetch an element from all ranks and return the SUM,
// multimem.st writes a value into all ranks' copies.  Rank r still owns slice r, but per element it issues ONE load and
// ONE store whatever the world size.  The switch still has to fetch every rank's copy of a slice and to deliver every
// slice to every rank, so per GPU and direction (1 + 1/N) buffer volumes cross the links (the reduce phase mostly
// outbound, the broadcast phase mostly inbound) against 2(N-1)/N for peer loads / stores: more traffic at N = 2, less
// from N = 4 on.  Measured on the 5.4 MB gradient bucket: N = 2 27.8 us against 20.6 us for the peer kernel, N = 4
// 25.4 against 25.1, N = 8 25.8 against 28.9 (tools/allreduce_bench.py); the caller times both and keeps the faster.
// Barriers, flags and epochs are the unicast ones above.  The switch adds in its own fixed order, so every rank
// receives the same bits, but they may differ in the last place from the rank-order sum.
__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void mc_st(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

template <int U>
__global__ void __launch_bounds__(ZS_PEER_THREADS)
    k_allreduce_nvls(const __grid_constant__ PeerTable tab, float* __restrict__ mc, int rank, int world, long long first,
                     long long count, int set, const float* __restrict__ extra_src, long long extra_index) {
    const int cta = blockIdx.x, tid = threadIdx.x;
    unsigned* mine = tab.flags[rank];
    __shared__ unsigned s_epoch;
    if (tid == 0) {
        s_epoch = mine[epoch_index(set, cta)] + 1u;
        put_extra(tab.buf[rank], extra_src, extra_index);
    }
    __syncthreads();
    const unsigned e = s_epoch;
    if (tid < world && tid != rank) {  // barrier 0: every rank's buffer is complete
        st_flag(tab.flags[tid] + flag_index(set, 0, cta, rank), e);
        while ((int)(ld_flag(mine + flag_index(set, 0, cta, tid)) - e) < 0) {}
    }
    __syncthreads();
    const long long units = count >> 2;
    const long long per = (units + world - 1) / world;
    const long long lo = (long long)rank * per, hi = lo + per < units ? lo + per : units;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)cta * blockDim.x + tid; i0 < hi; i0 += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) v[u] = mc_ld_reduce(mc + first + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) mc_st(mc + first + 4 * i, v[u]);
        }
    }
    __syncthreads();
    if (tid < world && tid != rank) {  // barrier 1: every rank's broadcasts have landed everywhere
        __threadfence_system();
        st_flag(tab.flags[tid] + flag_index(set, 1, cta, rank), e);
        while ((int)(ld_flag(mine + flag_index(set, 1, cta, tid)) - e) < 0) {}
    }
    __syncthreads();
    if (tid == 0) mine[epoch_index(set, cta)] = e;
}

}  // namespace zs

using namespace zs;

extern "C" {

int64_t zs_allreduce_peer_flag_bytes(void) {
    return (int64_t)(epoch_index(ZS_PEER_FLAG_SETS - 1, ZS_PEER_MAX_CTAS - 1) + 1) * 4;
}

int zs_allreduce_sum_peer(float* const* bufs_host, void* const* flags_host, int rank, int world, int64_t first,
                          int64_t count, int flag_set, int ctas, const float* extra_src, int64_t extra_index,
                          zs_stream_t stream) {
    ZS_REQUIRE(extra_src == nullptr || (extra_index >= first && extra_index < first + count), ZS_ERR_ARG);
    ZS_REQUIRE(bufs_host && flags_host && world >= 1 && world <= ZS_MAX_PEERS && rank >= 0 && rank < world, ZS_ERR_ARG);
    ZS_REQUIRE(first >= 0 && count >= 0 && flag_set >= 0 && flag_set < ZS_PEER_FLAG_SETS, ZS_ERR_ARG);
    if (first % 4 != 0 || count % 4 != 0) {
        set_last_error_msg("peer all-reduce: first / count must be multiples of 4 floats");
        return ZS_ERR_ALIGN;
    }
    if (count == 0) return ZS_OK;
    if (world == 1) {
        if (extra_src)
            ZS_CUDA_TRY(cudaMemcpyAsync(bufs_host[0] + extra_index, extra_src, sizeof(float), cudaMemcpyDeviceToDevice,
                                        as_stream(stream)));
        return ZS_OK;
    }
    PeerTable tab;
    for (int p = 0; p < ZS_MAX_PEERS; ++p) {
        tab.buf[p] = p < world ? bufs_host[p] : nullptr;
        tab.flags[p] = p < world ? (unsigned*)flags_host[p] : nullptr;
        if (p < world && (tab.buf[p] == nullptr || tab.flags[p] == nullptr || !aligned16(tab.buf[p]))) {
            set_last_error_msg("peer all-reduce: null / unaligned peer pointer");
            return ZS_ERR_ARG;
        }
    }
    if (ctas <= 0) ctas = 64;
    if (ctas > ZS_PEER_MAX_CTAS) ctas = ZS_PEER_MAX_CTAS;
    auto kern = world <= 2 ? k_allreduce_peer<2, 8> : (world <= 4 ? k_allreduce_peer<4, 4> : k_allreduce_peer<8, 2>);
    kern<<<ctas, ZS_PEER_THREADS, 0, as_stream(stream)>>>(tab, rank, world, (long long)first, (long long)count, flag_set,
                                                          extra_src, (long long)extra_index);
    ZS_LAUNCH_CHECK("k_allreduce_peer");
    return ZS_OK;
}

int zs_allreduce_sum_nvls(float* multicast_buf, float* local_buf, void* const* flags_host, int rank, int world,
                          int64_t first, int64_t count, int flag_set, int ctas, const float* extra_src,
                          int64_t extra_index, zs_stream_t stream) {
    ZS_REQUIRE(multicast_buf && local_buf && flags_host && world >= 1 && world <= ZS_MAX_PEERS && rank >= 0 && rank < world,
               ZS_ERR_ARG);
    ZS_REQUIRE(extra_src == nullptr || (extra_index >= first && extra_index < first + count), ZS_ERR_ARG);
    ZS_REQUIRE(first >= 0 && count >= 0 && flag_set >= 0 && flag_set < ZS_PEER_FLAG_SETS, ZS_ERR_ARG);
    if (first % 4 != 0 || count % 4 != 0 || !aligned16(multicast_buf)) {
        set_last_error_msg("multicast all-reduce: first / count must be multiples of 4 floats, the buffer 16-byte aligned");
        return ZS_ERR_ALIGN;
    }
    if (count == 0) return ZS_OK;
    if (world == 1) {
        if (extra_src)
            ZS_CUDA_TRY(cudaMemcpyAsync(local_buf + extra_index, extra_src, sizeof(float), cudaMemcpyDeviceToDevice,
                                        as_stream(stream)));
        return ZS_OK;
    }
    PeerTable tab;
    for (int p = 0; p < ZS_MAX_PEERS; ++p) {
        tab.buf[p] = p == rank ? local_buf : nullptr;
        tab.flags[p] = p < world ? (unsigned*)flags_host[p] : nullptr;
        if (p < world && tab.flags[p] == nullptr) {
            set_last_error_msg("multicast all-reduce: null flag pointer");
            return ZS_ERR_ARG;
        }
    }
    // a slice is 1/world of the buffer and each element costs one load + one store: fewer CTAs than the peer kernel
    if (ctas <= 0) ctas = 32;
    if (ctas > ZS_PEER_MAX_CTAS) ctas = ZS_PEER_MAX_CTAS;
    k_allreduce_nvls<8><<<ctas, ZS_PEER_THREADS, 0, as_stream(stream)>>>(tab, multicast_buf, rank, world, (long long)first,
                                                                       (long long)count, flag_set, extra_src,
                                                                       (long long)extra_index);
    ZS_LAUNCH_CHECK("k_allreduce_nvls");
    return ZS_OK;
}

}  // extern "C"
