// Fused training loss of the reference drivers (train.py:162-174):
//     mask = (disp < maxdisp) & (disp > 0)
//     loss = 0.5 * smooth_l1(out1[mask], disp[mask]) + 0.7 * smooth_l1(out2[mask], ...) + smooth_l1(out3[mask], ...)   (means)
// The reference materialises the boolean mask, six gathered copies and three reductions (~20 ATen launches and a
// host sync for the dynamic shape of out[mask]).  Here: ONE pass that reads the three outputs and the target once and
// produces the three masked sums and the valid-pixel count (doubles, one atomic set per CTA), and ONE pass for the
// three gradients.  The division by the (global, all-reduced in data-parallel training) count happens on the host
// tensors, so the kernel never synchronises.  HBM-bound: 4 reads (+3 writes backward) per pixel.
#include "common.cuh"

namespace cmfb200 {

__device__ __forceinline__ float smooth_l1(float x) {  // beta = 1 (torch default)
    const float a = fabsf(x);
    return a < 1.f ? 0.5f * x * x : a - 0.5f;
}

__global__ void __launch_bounds__(256) masked_smooth_l1_fwd_kernel(const float* __restrict__ o1, const float* __restrict__ o2,
                                                                   const float* __restrict__ o3, const float* __restrict__ disp,
                                                                   double* __restrict__ sums, long long n, float maxdisp) {
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, cnt = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = disp[i];
        if (d < maxdisp && d > 0.f) {
            s1 += (double)smooth_l1(o1[i] - d);
            s2 += (double)smooth_l1(o2[i] - d);
            s3 += (double)smooth_l1(o3[i] - d);
            cnt += 1.0;
        }
    }
    __shared__ double red[4][8];
    s1 = warp_sum(s1), s2 = warp_sum(s2), s3 = warp_sum(s3), cnt = warp_sum(cnt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[0][warp] = s1, red[1][warp] = s2, red[2][warp] = s3, red[3][warp] = cnt;
    __syncthreads();
    if (threadIdx.x < 4) {
        double a = 0.0;
        for (int k = 0; k < 8; ++k) a += red[threadIdx.x][k];
        atomicAdd(sums + threadIdx.x, a);
    }
}

// g_i = scale[i] * clamp(o_i - d, -1, 1) on valid pixels, 0 elsewhere; scale[i] = weight_i * dLoss / count (device)
__global__ void __launch_bounds__(256) masked_smooth_l1_bwd_kernel(const float* __restrict__ o1, const float* __restrict__ o2,
                                                                   const float* __restrict__ o3, const float* __restrict__ disp,
                                                                   const float* __restrict__ scale, float* __restrict__ g1,
                                                                   float* __restrict__ g2, float* __restrict__ g3, long long n,
                                                                   float maxdisp) {
    const float c1 = scale[0], c2 = scale[1], c3 = scale[2];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = disp[i];
        const bool valid = d < maxdisp && d > 0.f;
        g1[i] = valid ? c1 * fminf(fmaxf(o1[i] - d, -1.f), 1.f) : 0.f;
        g2[i] = valid ? c2 * fminf(fmaxf(o2[i] - d, -1.f), 1.f) : 0.f;
        g3[i] = valid ? c3 * fminf(fmaxf(o3[i] - d, -1.f), 1.f) : 0.f;
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_masked_smooth_l1_fwd(const float* out1, const float* out2, const float* out3, const float* disp,
                                            double* sums4, long long n, float maxdisp, void* stream) {
    CMF_REQUIRE(out1 && out2 && out3 && disp && sums4, "masked_smooth_l1_fwd: null pointer");
    CMF_REQUIRE(n > 0, "masked_smooth_l1_fwd: empty input");
    long long blocks = cdiv(n, 256 * 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    masked_smooth_l1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out1, out2, out3, disp, sums4, n, maxdisp);
    CMF_LAUNCH_CHECK("masked_smooth_l1_fwd_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_masked_smooth_l1_bwd(const float* out1, const float* out2, const float* out3, const float* disp,
                                            const float* scale3, float* g1, float* g2, float* g3, long long n,
                                            float maxdisp, void* stream) {
    CMF_REQUIRE(out1 && out2 && out3 && disp && scale3 && g1 && g2 && g3, "masked_smooth_l1_bwd: null pointer");
    CMF_REQUIRE(n > 0, "masked_smooth_l1_bwd: empty input");
    long long blocks = cdiv(n, 256 * 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    masked_smooth_l1_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out1, out2, out3, disp, scale3, g1, g2,
                                                                                    g3, n, maxdisp);
    CMF_LAUNCH_CHECK("masked_smooth_l1_bwd_kernel");
    return CMFB200_OK;
}
