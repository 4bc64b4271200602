// K3 -- GroupNorm(32, C) over NCDHW volumes, split into (statistics) + (normalise [+residual] [+ReLU]).
// Replaces nn.GroupNorm in convbn_3d / hourglass (cmf/models/cmfsm.py:58,269,280) and the residual adds /
// ReLUs that follow it (:285-301, :685-693).  Statistics are per-(b,channel) sum and sum of squares kept
// in double; the conv kernels accumulate them in their epilogue, gn_stats exists for un-fused producers.
// HBM-bound: gn_apply reads x (+residual) once and writes y once.
#include "common.cuh"

namespace cmfb200 {

constexpr int kGnThreads = 256;
constexpr int kGnChunk = kGnThreads * 4 * 8;  // floats per CTA

__device__ __forceinline__ void block_reduce_add2(double s, double ss, double* dst) {
    __shared__ double red[2][kGnThreads / 32];
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = s;
        red[1][warp] = ss;
    }
    __syncthreads();
    if (warp == 0) {
        s = lane < kGnThreads / 32 ? red[0][lane] : 0.0;
        ss = lane < kGnThreads / 32 ? red[1][lane] : 0.0;
        s = warp_sum(s);
        ss = warp_sum(ss);
        if (lane == 0) {
            atomicAdd(dst, s);
            atomicAdd(dst + 1, ss);
        }
    }
}

__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                              long long spatial) {
    const long long bc = blockIdx.y;
    const float* p = x + bc * spatial;
    const long long beg = (long long)blockIdx.x * kGnChunk;
    const long long end = min(spatial, beg + kGnChunk);
    // double from the first element on (see gn_epilogue in conv_common.cuh)
    double s = 0.0, ss = 0.0;
    if ((spatial & 3) == 0) {
        for (long long i = beg + threadIdx.x * 4; i < end; i += kGnThreads * 4) {
            const float4 v = ld_streaming_f4(p + i);
            const double a = v.x, b = v.y, c = v.z, d = v.w;
            s += (a + b) + (c + d);
            ss += (a * a + b * b) + (c * c + d * d);
        }
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kGnThreads) {
            const float v = p[i];
            s += (double)v;
            ss += (double)v * (double)v;
        }
    }
    block_reduce_add2(s, ss, sums + 2 * bc);
}

__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const float* __restrict__ x,
                                                              const double* __restrict__ sums,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ residual, float* __restrict__ y,
                                                              int C, int G, long long spatial, float eps, int relu) {
    const long long bc = blockIdx.y;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    const int cpg = C / G;
    const int g0 = (c / cpg) * cpg;
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < cpg; ++j) {
        s += sums[2 * (b * C + g0 + j)];
        ss += sums[2 * (b * C + g0 + j) + 1];
    }
    const double n = (double)cpg * (double)spatial;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = rsqrt(var + (double)eps);
    const float scale = (float)(rstd * (double)gamma[c]);
    const float shift = (float)((double)beta[c] - mean * rstd * (double)gamma[c]);

    const float* px = x + bc * spatial;
    const float* pr = residual ? residual + bc * spatial : nullptr;
    float* py = y + bc * spatial;
    const long long beg = (long long)blockIdx.x * kGnChunk;
    const long long end = min(spatial, beg + kGnChunk);
    if ((spatial & 3) == 0) {
        for (long long i = beg + threadIdx.x * 4; i < end; i += kGnThreads * 4) {
            float4 v = *reinterpret_cast<const float4*>(px + i);
            v.x = fmaf(v.x, scale, shift);
            v.y = fmaf(v.y, scale, shift);
            v.z = fmaf(v.z, scale, shift);
            v.w = fmaf(v.w, scale, shift);
            if (pr) {
                const float4 r = *reinterpret_cast<const float4*>(pr + i);
                v.x += r.x;
                v.y += r.y;
                v.z += r.z;
                v.w += r.w;
            }
            if (relu) {
                v.x = fmaxf(v.x, 0.f);
                v.y = fmaxf(v.y, 0.f);
                v.z = fmaxf(v.z, 0.f);
                v.w = fmaxf(v.w, 0.f);
            }
            *reinterpret_cast<float4*>(py + i) = v;
        }
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kGnThreads) {
            float v = fmaf(px[i], scale, shift);
            if (pr) v += pr[i];
            if (relu) v = fmaxf(v, 0.f);
            py[i] = v;
        }
    }
}

// ---- backward (training): y = gamma * xhat + beta, out = [relu](y [+ residual]),  xhat = (x - mean) * rstd.
// With g' = grad_out * [out > 0] (ReLU mask, when the forward applied one):
//   S1[b,c] = sum g',  S2[b,c] = sum g' * xhat                       (gn_bwd_stats: per-(b,channel) doubles)
//   dx = rstd * (gamma_c g' - A - xhat Bq),  A = sum_{c in group} gamma_c S1 / N,  Bq = sum gamma_c S2 / N
//   d_residual = g',  d_gamma[c] = sum_b S2[b,c],  d_beta[c] = sum_b S1[b,c]   (the last two on the host: C values)
// Replaces the mask multiply + aten::native_group_norm_backward (3-4 ATen launches per layer).
struct GnGroupStat {
    float mean_rstd, rstd;  // xhat = fma(x, rstd, -mean*rstd)
};

__device__ __forceinline__ GnGroupStat gn_group_stat(const double* __restrict__ sums, long long b, int c, int C, int cpg,
                                                     long long spatial, float eps) {
    const int g0 = (c / cpg) * cpg;
    double s = 0.0, ss = 0.0;
    for (int j = 0; j < cpg; ++j) {
        s += sums[2 * (b * C + g0 + j)];
        ss += sums[2 * (b * C + g0 + j) + 1];
    }
    const double n = (double)cpg * (double)spatial;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = rsqrt(var + (double)eps);
    return GnGroupStat{(float)(-mean * rstd), (float)rstd};
}

__global__ void __launch_bounds__(kGnThreads) gn_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                  const float* __restrict__ out,
                                                                  const double* __restrict__ sums,
                                                                  double* __restrict__ bsums, int C, int G,
                                                                  long long spatial, float eps) {
    const long long bc = blockIdx.y;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    const GnGroupStat st = gn_group_stat(sums, b, c, C, C / G, spatial, eps);
    const float* pg = g + bc * spatial;
    const float* px = x + bc * spatial;
    const float* po = out ? out + bc * spatial : nullptr;
    const long long beg = (long long)blockIdx.x * kGnChunk;
    const long long end = min(spatial, beg + kGnChunk);
    double s1 = 0.0, s2 = 0.0;
    if ((spatial & 3) == 0) {
        for (long long i = beg + threadIdx.x * 4; i < end; i += kGnThreads * 4) {
            float4 gv = *reinterpret_cast<const float4*>(pg + i);
            const float4 xv = *reinterpret_cast<const float4*>(px + i);
            if (po) {
                const float4 ov = *reinterpret_cast<const float4*>(po + i);
                gv.x = ov.x > 0.f ? gv.x : 0.f;
                gv.y = ov.y > 0.f ? gv.y : 0.f;
                gv.z = ov.z > 0.f ? gv.z : 0.f;
                gv.w = ov.w > 0.f ? gv.w : 0.f;
            }
            const float h0 = fmaf(xv.x, st.rstd, st.mean_rstd), h1 = fmaf(xv.y, st.rstd, st.mean_rstd);
            const float h2 = fmaf(xv.z, st.rstd, st.mean_rstd), h3 = fmaf(xv.w, st.rstd, st.mean_rstd);
            s1 += ((double)gv.x + (double)gv.y) + ((double)gv.z + (double)gv.w);
            s2 += ((double)gv.x * h0 + (double)gv.y * h1) + ((double)gv.z * h2 + (double)gv.w * h3);
        }
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kGnThreads) {
            float gv = pg[i];
            if (po && !(po[i] > 0.f)) gv = 0.f;
            s1 += (double)gv;
            s2 += (double)gv * (double)fmaf(px[i], st.rstd, st.mean_rstd);
        }
    }
    block_reduce_add2(s1, s2, bsums + 2 * bc);
}

__global__ void __launch_bounds__(kGnThreads) gn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                  const float* __restrict__ out,
                                                                  const double* __restrict__ sums,
                                                                  const double* __restrict__ bsums,
                                                                  const float* __restrict__ gamma, float* __restrict__ dx,
                                                                  float* __restrict__ dres, int C, int G,
                                                                  long long spatial, float eps) {
    const long long bc = blockIdx.y;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    const int cpg = C / G;
    const GnGroupStat st = gn_group_stat(sums, b, c, C, cpg, spatial, eps);
    const int g0 = (c / cpg) * cpg;
    double a = 0.0, bq = 0.0;
    for (int j = 0; j < cpg; ++j) {
        a += (double)gamma[g0 + j] * bsums[2 * (b * C + g0 + j)];
        bq += (double)gamma[g0 + j] * bsums[2 * (b * C + g0 + j) + 1];
    }
    const double n = (double)cpg * (double)spatial;
    const float k_g = st.rstd * gamma[c];
    const float k_a = (float)(-(double)st.rstd * a / n);
    const float k_b = (float)(-(double)st.rstd * bq / n);
    const float* pg = g + bc * spatial;
    const float* px = x + bc * spatial;
    const float* po = out ? out + bc * spatial : nullptr;
    float* pd = dx + bc * spatial;
    float* pr = dres ? dres + bc * spatial : nullptr;
    const long long beg = (long long)blockIdx.x * kGnChunk;
    const long long end = min(spatial, beg + kGnChunk);
    if ((spatial & 3) == 0) {
        for (long long i = beg + threadIdx.x * 4; i < end; i += kGnThreads * 4) {
            float4 gv = *reinterpret_cast<const float4*>(pg + i);
            const float4 xv = *reinterpret_cast<const float4*>(px + i);
            if (po) {
                const float4 ov = *reinterpret_cast<const float4*>(po + i);
                gv.x = ov.x > 0.f ? gv.x : 0.f;
                gv.y = ov.y > 0.f ? gv.y : 0.f;
                gv.z = ov.z > 0.f ? gv.z : 0.f;
                gv.w = ov.w > 0.f ? gv.w : 0.f;
            }
            float4 d;
            d.x = fmaf(gv.x, k_g, fmaf(fmaf(xv.x, st.rstd, st.mean_rstd), k_b, k_a));
            d.y = fmaf(gv.y, k_g, fmaf(fmaf(xv.y, st.rstd, st.mean_rstd), k_b, k_a));
            d.z = fmaf(gv.z, k_g, fmaf(fmaf(xv.z, st.rstd, st.mean_rstd), k_b, k_a));
            d.w = fmaf(gv.w, k_g, fmaf(fmaf(xv.w, st.rstd, st.mean_rstd), k_b, k_a));
            *reinterpret_cast<float4*>(pd + i) = d;
            if (pr) *reinterpret_cast<float4*>(pr + i) = gv;
        }
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kGnThreads) {
            float gv = pg[i];
            if (po && !(po[i] > 0.f)) gv = 0.f;
            pd[i] = fmaf(gv, k_g, fmaf(fmaf(px[i], st.rstd, st.mean_rstd), k_b, k_a));
            if (pr) pr[i] = gv;
        }
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_gn_stats(const float* x, double* gn_sums, int B, int C, long long spatial, void* stream) {
    CMF_REQUIRE(x && gn_sums, "gn_stats: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && spatial > 0, "gn_stats: non-positive dimension");
    CMF_REQUIRE((long long)B * C <= 65535, "gn_stats: B*C exceeds grid limit");
    dim3 grid((unsigned)cdiv(spatial, kGnChunk), (unsigned)(B * C));
    gn_stats_kernel<<<grid, kGnThreads, 0, (cudaStream_t)stream>>>(x, gn_sums, spatial);
    CMF_LAUNCH_CHECK("gn_stats_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_gn_apply(const float* x, const double* gn_sums, const float* gamma, const float* beta,
                                const float* residual, float* y, int B, int C, int G, long long spatial, float eps,
                                int relu, void* stream) {
    CMF_REQUIRE(x && gn_sums && gamma && beta && y, "gn_apply: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && G > 0 && spatial > 0, "gn_apply: non-positive dimension");
    CMF_REQUIRE(C % G == 0, "gn_apply: C=%d not divisible by G=%d", C, G);
    CMF_REQUIRE((long long)B * C <= 65535, "gn_apply: B*C exceeds grid limit");
    dim3 grid((unsigned)cdiv(spatial, kGnChunk), (unsigned)(B * C));
    gn_apply_kernel<<<grid, kGnThreads, 0, (cudaStream_t)stream>>>(x, gn_sums, gamma, beta, residual, y, C, G, spatial,
                                                                    eps, relu);
    CMF_LAUNCH_CHECK("gn_apply_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_gn_bwd(const float* grad_out, const float* x, const float* out_or_null, const double* gn_sums,
                              const float* gamma, double* bwd_sums, float* dx, float* dres_or_null, int B, int C, int G,
                              long long spatial, float eps, void* stream) {
    CMF_REQUIRE(grad_out && x && gn_sums && gamma && bwd_sums && dx, "gn_bwd: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && G > 0 && spatial > 0, "gn_bwd: non-positive dimension");
    CMF_REQUIRE(C % G == 0, "gn_bwd: C=%d not divisible by G=%d", C, G);
    CMF_REQUIRE((long long)B * C <= 65535, "gn_bwd: B*C exceeds grid limit");
    cudaStream_t st = (cudaStream_t)stream;
    CMF_CUDA(cudaMemsetAsync(bwd_sums, 0, (size_t)B * C * 2 * sizeof(double), st));
    dim3 grid((unsigned)cdiv(spatial, kGnChunk), (unsigned)(B * C));
    gn_bwd_stats_kernel<<<grid, kGnThreads, 0, st>>>(grad_out, x, out_or_null, gn_sums, bwd_sums, C, G, spatial, eps);
    CMF_LAUNCH_CHECK("gn_bwd_stats_kernel");
    gn_bwd_apply_kernel<<<grid, kGnThreads, 0, st>>>(grad_out, x, out_or_null, gn_sums, bwd_sums, gamma, dx, dres_or_null,
                                                     C, G, spatial, eps);
    CMF_LAUNCH_CHECK("gn_bwd_apply_kernel");
    return CMFB200_OK;
}
