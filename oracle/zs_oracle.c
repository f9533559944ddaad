/*
 * zs_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See zs_oracle_impl.h.
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11), restated from the published algorithm; known-answer vectors from the Random123
 * distribution are checked in tests/test_oracle_golden.py.  Same counter/key convention as the
 * kernels (zhusuan-pytorch_b200/csrc/zs_philox.cuh): counter = (i/4 lo, i/4 hi, offset lo,
 * offset hi), key = (seed lo, seed hi). */
static void philox4x32_10(uint32_t ctr[4], uint32_t key[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * ctr[0];
        uint64_t p1 = (uint64_t)M1 * ctr[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += W0; k1 += W1;
    }
}

/* raw known-answer interface: explicit counter and key */
void orc_philox_kat(uint32_t out[4], const uint32_t ctr[4], const uint32_t key[2]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    philox4x32_10(c, k);
    memcpy(out, c, sizeof(c));
}

void orc_philox_raw(uint32_t* out, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t q = 0; q < n / 4; ++q) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(c, k);
        memcpy(out + 4 * q, c, sizeof(c));
    }
}

static float u01_closed_open(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-8f; }
static float u01_open(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f; }

/* uniforms on the OPEN interval (0,1): the noise source of the Logistic / Laplace samplers */
void orc_philox_uniform_open_f32(float* out, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t q = 0; q < (n + 3) / 4; ++q) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(c, k);
        for (int j = 0; j < 4; ++j)
            if (4 * q + j < n) out[4 * q + j] = u01_open(c[j]);
    }
}

void orc_philox_uniform_f32(float* out, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t q = 0; q < (n + 3) / 4; ++q) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(c, k);
        for (int j = 0; j < 4; ++j)
            if (4 * q + j < n) out[4 * q + j] = u01_closed_open(c[j]);
    }
}

/* Box-Muller exactly as zs_philox.cuh: (r0,r1)->(n0,n1), (r2,r3)->(n2,n3) */
void orc_philox_normal_f32(float* out, int64_t n, float mean, float std, uint64_t seed, uint64_t offset) {
    for (int64_t q = 0; q < (n + 3) / 4; ++q) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(c, k);
        float v[4];
        for (int h = 0; h < 2; ++h) {
            float u1 = u01_open(c[2 * h]), u2 = u01_open(c[2 * h + 1]);
            float rad = sqrtf(-2.0f * logf(u1));
            double ang = 2.0 * 3.14159265358979323846 * (double)u2;
            v[2 * h] = rad * (float)cos(ang);
            v[2 * h + 1] = rad * (float)sin(ang);
        }
        for (int j = 0; j < 4; ++j)
            if (4 * q + j < n) out[4 * q + j] = mean + std * v[j];
    }
}

int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define REAL float
#define SUFFIX _f32
#define RLOG logf
#define REXP expf
#define RSQRT sqrtf
#include "zs_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef RLOG
#undef REXP
#undef RSQRT

#define REAL double
#define SUFFIX _f64
#define RLOG log
#define REXP exp
#define RSQRT sqrt
#include "zs_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef RLOG
#undef REXP
#undef RSQRT
