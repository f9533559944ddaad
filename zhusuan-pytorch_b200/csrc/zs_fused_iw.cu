// Fused resident-column kernel: Bernoulli likelihood log-pmf + importance-weighted objective
// (SGVB or VIMCO), forward AND backward, one launch, probs read from HBM once.
//
// Replaces, for a Bernoulli likelihood node under ImportanceWeightedObjective (reference file:line):
//   Bernoulli._log_prob            zhusuan/distributions/bernoulli.py:84-95   (fwd + autograd bwd)
//   StochasticTensor.log_prob      zhusuan/framework/stochastic_tensor.py:160-181 (event sum)
//   ImportanceWeightedObjective.log_joint / sgvb / vimco
//                                  zhusuan/variational/importance_weighted_objective.py:66-77,102-191
//
// Design: one persistent CTA per SM walks batch columns b.  For a column, the K
// particle rows probs[k, b, :] (K*X*4 bytes, 157 KB at K=50, X=784) are staged in shared memory by
// 1-D bulk async copies (cp.async.bulk, one mbarrier per row).  Each warp owns a fixed set of rows:
//   phase A  wait for the row, reduce its log-pmf against x[b,:]            (warp shuffle)
//   sync     warp 0 forms log-weights, softmax weights / VIMCO signal, cost, d/dlogp, d/dlogq
//   phase B  dprobs = g_k * (x/(p+eps) - (1-x)/((1-p)+eps)) from the RESIDENT row -> HBM,
//            then immediately re-arm the row's mbarrier and issue the bulk copy of the same row
//            of the CTA's next column, so next-column loads overlap this column's stores.
// HBM traffic per particle-sample: 4X read + 4X write (+ [K,B] scalars) instead of 8X + 4X for the
// two-pass form; the unfused entry points remain as the general fallback.
#include "zs_common.cuh"

namespace zs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "ZS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra ZS_DONE;\n"
        "bra ZS_WAIT;\n"
        "ZS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (TMA engine, UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct FusedSmemLayout {
    int K, X, Kpad;
    __host__ __device__ FusedSmemLayout(int K_, int X_) : K(K_), X(X_), Kpad((K_ + 3) & ~3) {}
    __host__ __device__ size_t rows_off() const { return 0; }
    __host__ __device__ size_t xrow_off() const { return (size_t)K * X * 4; }
    __host__ __device__ size_t lpx_off() const { return xrow_off() + (size_t)2 * X * 4; }   // [2][Kpad]
    __host__ __device__ size_t rest_off() const { return lpx_off() + (size_t)2 * Kpad * 4; }
    __host__ __device__ size_t lq_off() const { return rest_off() + (size_t)Kpad * 4; }
    __host__ __device__ size_t g_off() const { return lq_off() + (size_t)Kpad * 4; }
    __host__ __device__ size_t xw_off() const { return (g_off() + (size_t)Kpad * 4 + 15) & ~(size_t)15; }  // double[Kpad]
    __host__ __device__ size_t bar_off() const { return (xw_off() + (size_t)Kpad * 8 + 15) & ~(size_t)15; }
    __host__ __device__ size_t total() const { return bar_off() + (size_t)(K + 2) * 8; }
};

// One warp turns the K log-weights of a column into cost, weights and gradients (all in smem/regs).
// As in k_iw_objective, log w and its centring are carried in double (K values: negligible work).
template <int EST>
__device__ __forceinline__ void warp_objective(int lane, int K, int64_t B, int64_t b, const float* s_lpx,
                                               const float* s_other, const float* s_lq, double* s_xw, float* s_g,
                                               float gscale, float* __restrict__ cost, float* __restrict__ dlogp,
                                               float* __restrict__ dlogq, float* __restrict__ logpx_out) {
    const unsigned FULL = 0xffffffffu;
    double m1 = -INFINITY, m2 = -INFINITY, sumd = 0.0;
    int i1 = -1;
    for (int k = lane; k < K; k += 32) {
        // same association as k_iw_objective: (logp - logq) + extra
        double xv = ((double)s_lpx[k] - (double)s_lq[k]) + (double)s_other[k];
        s_xw[k] = xv;
        if (xv > m1 || i1 < 0) {
            m2 = m1; m1 = xv; i1 = k;
        } else if (xv > m2) {
            m2 = xv;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double om1 = __shfl_xor_sync(FULL, m1, o), om2 = __shfl_xor_sync(FULL, m2, o);
        int oi1 = __shfl_xor_sync(FULL, i1, o);
        bool other = (oi1 >= 0) && (i1 < 0 || om1 > m1 || (om1 == m1 && oi1 < i1));
        double loser = other ? m1 : om1;
        m2 = fmax(loser, fmax(m2, om2));
        m1 = other ? om1 : m1;
        i1 = other ? oi1 : i1;
    }
    const float gap = (float)(m1 - m2);

    float S = 0.f, S2 = 0.f;
    for (int k = lane; k < K; k += 32) {
        double xv = s_xw[k];
        S += expf((float)(xv - m1));
        if (EST == ZS_EST_VIMCO) {
            sumd += xv - m1;
            if (k != i1) S2 += expf((float)(xv - m2));
        }
    }
    S = warp_sum(S);
    if (EST == ZS_EST_VIMCO) {
        S2 = warp_sum(S2);
        sumd = warp_sum(sumd);
    }

    double c_acc = 0.0;
    const float invS = 1.0f / S;
    const double km1 = (double)(K - 1);
    for (int k = lane; k < K; k += 32) {
        const double xv = s_xw[k];
        const float e = expf((float)(xv - m1));
        const float wt = e / S;
        c_acc -= (double)wt * xv;
        float gq = wt;
        if (EST == ZS_EST_VIMCO) {
            const float lq = s_lq[k];
            const float mu_m = (float)((sumd - (xv - m1)) / km1);
            float sig;
            if (k == i1 && gap > 1.0f) {
                float Sloo = S2 + expf(mu_m + gap);
                sig = gap + (logf(S) - logf(Sloo));
            } else {
                sig = -log1pf((expf(mu_m) - e) * invS);
            }
            c_acc -= (double)lq * (double)sig;
            gq = wt - sig;
        }
        const float gp = -wt * gscale;
        s_g[k] = gp;
        if (dlogp) dlogp[(int64_t)k * B + b] = gp;
        if (dlogq) dlogq[(int64_t)k * B + b] = gq * gscale;
        if (logpx_out) logpx_out[(int64_t)k * B + b] = s_lpx[k];
    }
    c_acc = warp_sum(c_acc);
    if (lane == 0 && cost) cost[b] = (float)c_acc;
}

template <int EST>
__global__ void __launch_bounds__(1024, 1)
    k_iw_bernoulli_fused(float* __restrict__ cost, float* __restrict__ dprobs, float* __restrict__ dlogp,
                         float* __restrict__ dlogq, float* __restrict__ logpx_out, const float* __restrict__ probs,
                         const float* __restrict__ x, const float* __restrict__ logp_other,
                         const float* __restrict__ logq, int K, int64_t B, int X, float gscale) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FusedSmemLayout L(K, X);
    float* rows = reinterpret_cast<float*>(smem + L.rows_off());
    float* xrow = reinterpret_cast<float*>(smem + L.xrow_off());
    float* s_lpx = reinterpret_cast<float*>(smem + L.lpx_off());
    float* s_rest = reinterpret_cast<float*>(smem + L.rest_off());
    float* s_lq = reinterpret_cast<float*>(smem + L.lq_off());
    float* s_g = reinterpret_cast<float*>(smem + L.g_off());
    double* s_xw = reinterpret_cast<double*>(smem + L.xw_off());
    uint64_t* bar_row = reinterpret_cast<uint64_t*>(smem + L.bar_off());
    uint64_t* bar_x = bar_row + K;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    const int X4 = X >> 2;
    const uint32_t row_bytes = (uint32_t)X * 4u;
    const float LN2 = 0.6931471805599453f;

    if (threadIdx.x == 0) {
        for (int k = 0; k < K + 2; ++k) mbar_init(&bar_row[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t b0 = blockIdx.x;
    if (b0 >= B) return;  // uniform per CTA

    // prologue: first column's rows (each warp loads the rows it owns) and x row
    if (lane == 0) {
        for (int k = warp; k < K; k += NW) {
            mbar_expect_tx(&bar_row[k], row_bytes);
            bulk_load(rows + (size_t)k * X, probs + ((int64_t)k * B + b0) * X, row_bytes, &bar_row[k]);
        }
    }
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_x[0], row_bytes);
        bulk_load(xrow, x + b0 * X, row_bytes, &bar_x[0]);
    }

    int it = 0;
    for (int64_t b = b0; b < B; b += gridDim.x, ++it) {
        const int64_t b_next = b + gridDim.x;
        const bool has_next = b_next < B;
        const float* xb = xrow + (size_t)(it & 1) * X;
        float* lpx = s_lpx + (size_t)(it & 1) * L.Kpad;

        // per-column scalars of the other log-weight terms
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float o = logp_other ? logp_other[(int64_t)k * B + b] : 0.f;
            float q = logq ? logq[(int64_t)k * B + b] : 0.f;
            s_rest[k] = o;
            s_lq[k] = q;
        }

        mbar_wait(&bar_x[it & 1], (uint32_t)((it >> 1) & 1));
        const float4* x4 = reinterpret_cast<const float4*>(xb);
        // binary observations (the MNIST-shaped configs) need one log / one reciprocal per element
        bool binary = true;
        for (int v = lane; v < X4; v += 32) {
            float4 xx = x4[v];
            binary = binary && (xx.x == 0.f || xx.x == 1.f) && (xx.y == 0.f || xx.y == 1.f) &&
                     (xx.z == 0.f || xx.z == 1.f) && (xx.w == 0.f || xx.w == 1.f);
        }
        binary = __all_sync(0xffffffffu, binary);

        // ---- phase A: log-pmf of each owned row -------------------------------------------
        for (int k = warp; k < K; k += NW) {
            mbar_wait(&bar_row[k], (uint32_t)(it & 1));
            const float4* p4 = reinterpret_cast<const float4*>(rows + (size_t)k * X);
            float acc = 0.f;
            if (binary) {
                float mn = 1.0f;
#pragma unroll 4
                for (int v = lane; v < X4; v += 32) {
                    const float4 p = p4[v], xx = x4[v];
                    float a, bb;
                    a = p.x + 1e-8f; bb = (1.0f - p.x) + 1e-8f; mn = fminf(mn, fminf(a, bb)); acc += fast_log2(xx.x == 1.f ? a : bb);
                    a = p.y + 1e-8f; bb = (1.0f - p.y) + 1e-8f; mn = fminf(mn, fminf(a, bb)); acc += fast_log2(xx.y == 1.f ? a : bb);
                    a = p.z + 1e-8f; bb = (1.0f - p.z) + 1e-8f; mn = fminf(mn, fminf(a, bb)); acc += fast_log2(xx.z == 1.f ? a : bb);
                    a = p.w + 1e-8f; bb = (1.0f - p.w) + 1e-8f; mn = fminf(mn, fminf(a, bb)); acc += fast_log2(xx.w == 1.f ? a : bb);
                }
                // the reference's x*log(a) + (1-x)*log(b) is NaN whenever either log argument is negative
                if (mn < 0.f) acc = __int_as_float(0x7fc00000);
            } else {
#pragma unroll 4
                for (int v = lane; v < X4; v += 32) {
                    const float4 p = p4[v], xx = x4[v];
                    acc += xx.x * fast_log2(p.x + 1e-8f) + (1.0f - xx.x) * fast_log2((1.0f - p.x) + 1e-8f);
                    acc += xx.y * fast_log2(p.y + 1e-8f) + (1.0f - xx.y) * fast_log2((1.0f - p.y) + 1e-8f);
                    acc += xx.z * fast_log2(p.z + 1e-8f) + (1.0f - xx.z) * fast_log2((1.0f - p.z) + 1e-8f);
                    acc += xx.w * fast_log2(p.w + 1e-8f) + (1.0f - xx.w) * fast_log2((1.0f - p.w) + 1e-8f);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) lpx[k] = acc * LN2;
        }
        __syncthreads();

        // x row of the next column: its buffer was last read in phase B of the previous column
        if (threadIdx.x == 0 && has_next) {
            mbar_expect_tx(&bar_x[(it + 1) & 1], row_bytes);
            bulk_load(xrow + (size_t)((it + 1) & 1) * X, x + b_next * X, row_bytes, &bar_x[(it + 1) & 1]);
        }
        if (warp == 0)
            warp_objective<EST>(lane, K, B, b, lpx, s_rest, s_lq, s_xw, s_g, gscale, cost, dlogp, dlogq, logpx_out);
        __syncthreads();

        // ---- phase B: dprobs from the resident rows, then refill the row for the next column ----
        for (int k = warp; k < K; k += NW) {
            if (dprobs) {
                const float4* p4 = reinterpret_cast<const float4*>(rows + (size_t)k * X);
                float* drow = dprobs + ((int64_t)k * B + b) * X;
                const float g = s_g[k];
                if (binary) {
                    const float ng = -g;
#pragma unroll 4
                    for (int v = lane; v < X4; v += 32) {
                        const float4 p = p4[v], xx = x4[v];
                        Pack<float> o;
                        o.v[0] = (xx.x == 1.f ? g : ng) * fast_rcp(xx.x == 1.f ? p.x + 1e-8f : (1.0f - p.x) + 1e-8f);
                        o.v[1] = (xx.y == 1.f ? g : ng) * fast_rcp(xx.y == 1.f ? p.y + 1e-8f : (1.0f - p.y) + 1e-8f);
                        o.v[2] = (xx.z == 1.f ? g : ng) * fast_rcp(xx.z == 1.f ? p.z + 1e-8f : (1.0f - p.z) + 1e-8f);
                        o.v[3] = (xx.w == 1.f ? g : ng) * fast_rcp(xx.w == 1.f ? p.w + 1e-8f : (1.0f - p.w) + 1e-8f);
                        st_pack_stream(drow + 4 * v, o);
                    }
                } else {
#pragma unroll 4
                    for (int v = lane; v < X4; v += 32) {
                        const float4 p = p4[v], xx = x4[v];
                        Pack<float> o;
                        o.v[0] = (g * xx.x) * fast_rcp(p.x + 1e-8f) - (g * (1.0f - xx.x)) * fast_rcp((1.0f - p.x) + 1e-8f);
                        o.v[1] = (g * xx.y) * fast_rcp(p.y + 1e-8f) - (g * (1.0f - xx.y)) * fast_rcp((1.0f - p.y) + 1e-8f);
                        o.v[2] = (g * xx.z) * fast_rcp(p.z + 1e-8f) - (g * (1.0f - xx.z)) * fast_rcp((1.0f - p.z) + 1e-8f);
                        o.v[3] = (g * xx.w) * fast_rcp(p.w + 1e-8f) - (g * (1.0f - xx.w)) * fast_rcp((1.0f - p.w) + 1e-8f);
                        st_pack_stream(drow + 4 * v, o);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && has_next) {
                mbar_expect_tx(&bar_row[k], row_bytes);
                bulk_load(rows + (size_t)k * X, probs + ((int64_t)k * B + b_next) * X, row_bytes, &bar_row[k]);
            }
        }
    }
}

static int pick_warps(int K) {
    // each warp owns K/NW rows: prefer an exact divisor so phase A/B are balanced
    int best = 0;
    for (int d = 32; d >= 8; --d)
        if (K % d == 0) { best = d; break; }
    if (best) return best;
    return K < 32 ? K : 32;
}

}  // namespace zs

using namespace zs;

extern "C" {

int64_t zs_iw_bernoulli_fused_smem_bytes(int64_t K, int64_t X) {
    if (K < 1 || X < 1 || K > 1 << 20 || X > 1 << 24) return -1;
    return (int64_t)FusedSmemLayout((int)K, (int)X).total();
}

int zs_iw_bernoulli_fused(int estimator, float* cost, float* dprobs, float* dlogp, float* dlogq, float* logpx_out,
                          const float* probs, const float* x, const float* logp_other, const float* logq, int64_t K,
                          int64_t B, int64_t X, double grad_scale, zs_stream_t stream) {
    ZS_REQUIRE(probs && x && K >= 1 && B >= 0 && X >= 1, ZS_ERR_ARG);
    ZS_REQUIRE(estimator == ZS_EST_SGVB || estimator == ZS_EST_VIMCO, ZS_ERR_ARG);
    ZS_REQUIRE(!(estimator == ZS_EST_VIMCO && (K < 2 || logq == nullptr)), ZS_ERR_ARG);
    if (B == 0) return ZS_OK;
    if (X % 4 != 0 || K < 8 || K > 4096 || X > (1 << 20)) {
        set_last_error_msg("fused kernel needs X % 4 == 0 and 8 <= K <= 4096");
        return ZS_ERR_UNSUPPORTED;
    }
    if (!aligned16(probs) || !aligned16(x) || !aligned16(dprobs)) {
        set_last_error_msg("fused kernel needs 16-byte aligned probs / x / dprobs");
        return ZS_ERR_ALIGN;
    }
    const size_t smem = FusedSmemLayout((int)K, (int)X).total();
    int dev = 0, max_optin = 0;
    ZS_CUDA_TRY(cudaGetDevice(&dev));
    ZS_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)max_optin) {
        set_last_error_msg("fused kernel: K*X rows do not fit in shared memory");
        return ZS_ERR_UNSUPPORTED;
    }
    const int nw = pick_warps((int)K);
    const int threads = nw * 32;
    auto kern = estimator == ZS_EST_SGVB ? k_iw_bernoulli_fused<ZS_EST_SGVB> : k_iw_bernoulli_fused<ZS_EST_VIMCO>;
    ZS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    ZS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > B) grid = B;
    kern<<<(unsigned)grid, threads, smem, as_stream(stream)>>>(cost, dprobs, dlogp, dlogq, logpx_out, probs, x,
                                                               logp_other, logq, (int)K, B, (int)X,
                                                               (float)grad_scale);
    ZS_LAUNCH_CHECK("k_iw_bernoulli_fused");
    return ZS_OK;
}

int64_t zs_iw_step_host_workspace(int64_t K, int64_t B, int64_t X) {
    if (K < 1 || B < 0 || X < 1) return -1;
    // probs, dprobs [K,B,X]; x [B,X]; logp_other, logq, dlogp, dlogq, logpx [K,B]; cost [B]; 256 B slack each
    const int64_t kbx = K * B * X * 4, bx = B * X * 4, kb = K * B * 4, bb = B * 4;
    auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
    return 2 * up(kbx) + up(bx) + 5 * up(kb) + up(bb);
}

int zs_iw_step_host(int estimator, float* cost_host, float* dprobs_host, float* dlogp_host, float* dlogq_host,
                    const float* probs_host, const float* x_host, const float* logp_other_host,
                    const float* logq_host, int64_t K, int64_t B, int64_t X, double grad_scale, void* ws,
                    int64_t ws_bytes, zs_stream_t stream) {
    ZS_REQUIRE(probs_host && x_host && ws && K >= 1 && B >= 1 && X >= 1, ZS_ERR_ARG);
    if (ws_bytes < zs_iw_step_host_workspace(K, B, X)) return ZS_ERR_WORKSPACE;
    ZS_REQUIRE(aligned16(ws), ZS_ERR_ALIGN);
    cudaStream_t st = as_stream(stream);
    const int64_t kbx = K * B * X * 4, bx = B * X * 4, kb = K * B * 4, bb = B * 4;
    auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
    char* p = (char*)ws;
    float* d_probs = (float*)p; p += up(kbx);
    float* d_dprobs = (float*)p; p += up(kbx);
    float* d_x = (float*)p; p += up(bx);
    float* d_other = (float*)p; p += up(kb);
    float* d_logq = (float*)p; p += up(kb);
    float* d_dlogp = (float*)p; p += up(kb);
    float* d_dlogq = (float*)p; p += up(kb);
    float* d_lpx = (float*)p; p += up(kb);
    float* d_cost = (float*)p;

    ZS_CUDA_TRY(cudaMemcpyAsync(d_probs, probs_host, kbx, cudaMemcpyHostToDevice, st));
    ZS_CUDA_TRY(cudaMemcpyAsync(d_x, x_host, bx, cudaMemcpyHostToDevice, st));
    if (logp_other_host) ZS_CUDA_TRY(cudaMemcpyAsync(d_other, logp_other_host, kb, cudaMemcpyHostToDevice, st));
    if (logq_host) ZS_CUDA_TRY(cudaMemcpyAsync(d_logq, logq_host, kb, cudaMemcpyHostToDevice, st));

    int rc = zs_iw_bernoulli_fused(estimator, d_cost, dprobs_host ? d_dprobs : nullptr, d_dlogp, d_dlogq, nullptr,
                                   d_probs, d_x, logp_other_host ? d_other : nullptr, logq_host ? d_logq : nullptr, K,
                                   B, X, grad_scale, stream);
    if (rc == ZS_ERR_UNSUPPORTED) {
        // two-pass form: likelihood log-pmf, objective over [K,B], likelihood backward
        rc = zs_bernoulli_logpmf_fwd(ZS_F32, d_lpx, d_x, ZS_KBCAST, d_probs, ZS_FULL, K, B, X, stream);
        if (rc != ZS_OK) return rc;
        if (!logq_host) ZS_CUDA_TRY(cudaMemsetAsync(d_logq, 0, kb, st));
        rc = zs_iw_objective(ZS_F32, estimator, d_cost, d_dlogp, d_dlogq, d_lpx, d_logq,
                             logp_other_host ? d_other : nullptr, K, B, grad_scale, stream);
        if (rc != ZS_OK) return rc;
        if (dprobs_host)
            rc = zs_bernoulli_logpmf_bwd(ZS_F32, nullptr, d_dprobs, d_dlogp, d_x, ZS_KBCAST, d_probs, ZS_FULL, K, B, X,
                                         stream);
    }
    if (rc != ZS_OK) return rc;
    if (cost_host) ZS_CUDA_TRY(cudaMemcpyAsync(cost_host, d_cost, bb, cudaMemcpyDeviceToHost, st));
    if (dprobs_host) ZS_CUDA_TRY(cudaMemcpyAsync(dprobs_host, d_dprobs, kbx, cudaMemcpyDeviceToHost, st));
    if (dlogp_host) ZS_CUDA_TRY(cudaMemcpyAsync(dlogp_host, d_dlogp, kb, cudaMemcpyDeviceToHost, st));
    if (dlogq_host) ZS_CUDA_TRY(cudaMemcpyAsync(dlogq_host, d_dlogq, kb, cudaMemcpyDeviceToHost, st));
    ZS_CUDA_TRY(cudaStreamSynchronize(st));
    return ZS_OK;
}

}  // extern "C"
