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
