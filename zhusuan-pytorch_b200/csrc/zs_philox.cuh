// Philox4x32-10 counter-based RNG (Salmon et al., SC'11) and the uniform / normal
// conversions used by every sampling kernel.  Element i of a call draws word i%4 of
// Philox(counter = (i/4 lo, i/4 hi, offset lo, offset hi), key = (seed lo, seed hi)), so the
// stream is a pure function of (seed, offset, i): independent of launch geometry and of how
// particles / chains are sharded over ranks.  oracle/zs_oracle_impl.h restates it bit-exactly.
#pragma once
#include <stdint.h>

namespace zs {

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t index, uint64_t offset, uint64_t seed) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32);
    uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// u in [0,1): 24 random bits.  P(u < p) == p up to 2^-24 for the Bernoulli draw.
__host__ __device__ __forceinline__ float u01_closed_open(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-8f; }
// u in (0,1): safe under log()
__host__ __device__ __forceinline__ float u01_open(uint32_t r) {
    return (float)(r >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
}

#ifdef __CUDACC__
// ---- graph-safe stream position ------------------------------------------------------------------------------
// A sampling launch whose (seed, offset) are passed by value draws the same noise every time a captured CUDA graph
// replays it.  With a device-side state (zs_rng_state in include/zs_b200.h: {uint64 offset; uint32 arrivals; uint32
// reserved}) the launch reads its stream position from device memory and advances it for the next launch:
//   - every CTA's leader reads state->offset, THEN (after a fence) counts itself in state->arrivals;
//   - the CTA that arrives last therefore knows every other CTA has already read the offset: it resets the
//     arrival count and stores offset + ZS_RNG_TICK.  Nothing waits on anything, so grids larger than the machine
//     (or persistent kernels) are fine; the next launch on the stream sees the advanced offset.
// `advance == false` only reads (the backward of a sample regenerates the forward's noise from a snapshot).
// `snapshot` (may be null) receives the position this launch used.  Must be called by ALL threads of the CTA (it
// contains a __syncthreads) before the first draw; returns by-value offset + device offset.
constexpr unsigned long long ZS_RNG_TICK_DEV = 4ull;  // == ZS_RNG_TICK

__device__ __forceinline__ uint64_t rng_acquire(uint64_t offset, unsigned long long* state, unsigned long long* snapshot,
                                                bool advance) {
    if (state == nullptr) return offset;  // uniform over the grid
    __shared__ unsigned long long s_rng_base;
    const bool leader = threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0;
    unsigned long long base = 0ull;
    if (leader) {
        base = *reinterpret_cast<volatile unsigned long long*>(state);
        s_rng_base = base;
    }
    __syncthreads();
    const uint64_t eff = offset + s_rng_base;
    // the CTA's only read of the device state is the leader's, above: its arrival may follow the barrier, so the
    // other warps do not wait for the fence and the atomic round trip
    if (leader) {
        if (snapshot != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            snapshot[0] = eff;
            snapshot[1] = 0ull;
        }
        if (advance) {
            __threadfence();
            unsigned int* arrivals = reinterpret_cast<unsigned int*>(state + 1);
            const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(arrivals, 1u) == total - 1u) {
                *reinterpret_cast<volatile unsigned int*>(arrivals) = 0u;
                *reinterpret_cast<volatile unsigned long long*>(state) = base + ZS_RNG_TICK_DEV;
            }
        }
    }
    return eff;
}

// The same protocol for a kernel that has work to do between the leader's read of the state and its first draw: the
// leader loads state[0] early (`base`, any time after the kernel started) and every thread calls this afterwards.
// Always advances; contains the CTA barrier (which also publishes whatever the CTA wrote to shared memory before it).
__device__ __forceinline__ uint64_t rng_acquire_peeked(uint64_t offset, unsigned long long* state, unsigned long long base) {
    __shared__ unsigned long long s_rng_base2;
    const bool leader = threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0;
    if (leader) s_rng_base2 = state ? base : 0ull;
    __syncthreads();
    const uint64_t eff = offset + s_rng_base2;
    if (leader && state != nullptr) {
        __threadfence();
        unsigned int* arrivals = reinterpret_cast<unsigned int*>(state + 1);
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(arrivals, 1u) == total - 1u) {
            *reinterpret_cast<volatile unsigned int*>(arrivals) = 0u;
            *reinterpret_cast<volatile unsigned long long*>(state) = base + ZS_RNG_TICK_DEV;
        }
    }
    return eff;
}

// Box-Muller on two words -> two standard normals, written for instruction count: the sampling kernels are
// issue-bound on the noise (ncu round 2: 350 instructions per float4 of latents, two thirds of them RNG), and libm's
// logf / IEEE sqrtf with their range and denormal handling were ~45 of the ~110 Box-Muller instructions per pair.
//   w = -ln(u1):  lg2.approx has an ABSOLUTE error of 2^-22 (in log2 units) on [0.5, 2): harmless unless w itself is
//       tiny, i.e. u1 within 2^-8 of 1, where the radius sqrt(2w) would inherit a large relative error.  There
//       v = 1 - u1 is exact in float and -ln(1 - v) = v + v^2/2 + v^3/3 (next term < 6e-11) is used instead.
//   radius sqrt.approx (relative error 2^-23); angle on (-pi, pi) with the SFU sin / cos (absolute error 2^-21).
// The resulting normals are within ~2e-6 (absolute) of the libm evaluation of the same uniforms, which is how the
// oracle restates it; one definition keeps every kernel's stream identical.
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float& n0, float& n1) {
    const float u1 = u01_open(r0), u2 = u01_open(r1);
    const float v = 1.0f - u1;
    float lg, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    const float w_far = lg * -0.6931471805599453f;
    const float w_near = v * fmaf(v, fmaf(v, 0.33333334f, 0.5f), 1.0f);
    const float w = v < 0.00390625f ? w_near : w_far;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(w + w));
    float s, c;
    __sincosf(6.283185307179586f * u2 - 3.141592653589793f, &s, &c);
    n0 = -rad * c;  // cos(t - pi) = -cos t
    n1 = -rad * s;
}
// the four standard normals of element group q (elements 4q .. 4q+3)
__device__ __forceinline__ void philox_normal4(uint64_t q, uint64_t offset, uint64_t seed, float out[4]) {
    Philox4 r = philox4x32_10(q, offset, seed);
    box_muller(r.x, r.y, out[0], out[1]);
    box_muller(r.z, r.w, out[2], out[3]);
}
__device__ __forceinline__ void philox_uniform4(uint64_t q, uint64_t offset, uint64_t seed, float out[4]) {
    Philox4 r = philox4x32_10(q, offset, seed);
    out[0] = u01_closed_open(r.x);
    out[1] = u01_closed_open(r.y);
    out[2] = u01_closed_open(r.z);
    out[3] = u01_closed_open(r.w);
}
#endif

}  // namespace zs
