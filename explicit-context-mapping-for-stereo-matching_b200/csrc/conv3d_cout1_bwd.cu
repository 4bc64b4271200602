// Backward of the classifier tail conv (nn.Conv3d(32, 1, 3, padding=1, bias=False), classifN.2,
// cmf/models/cmfsm.py:624,629,634) for the training path.  ATen/cuDNN needs 3.4 ms per classifier for this
// Cout = 1 shape (batch 8, 48x64x128); both halves are tiny direct kernels:
//   dgrad  dX[b,ci,u]  = sum_tap W[ci,tap] * dY[b, u - off(tap)]     one thread per voxel, all Cin in registers;
//                        bound by writing dX once (Cin x 4 B per voxel)
//   wgrad  dW[ci,tap]  = sum_{b,v} dY[b,v] * X[b,ci,v + off(tap)]    grid (chunk, ci, b), 27 accumulators per thread,
//                        X neighbours come from L1; block reduction + one atomic per (ci,tap) per CTA
#include "common.cuh"

namespace cmfb200 {

constexpr int kC1Threads = 256;

template <int CIN>
__global__ void __launch_bounds__(kC1Threads) conv3d_cout1_dgrad_kernel(const float* __restrict__ wgt,
                                                                        const float* __restrict__ gy,
                                                                        float* __restrict__ dx, int D, int H, int W) {
    __shared__ __align__(16) float sw[27][CIN];
    for (int i = threadIdx.x; i < 27 * CIN; i += kC1Threads) sw[i % 27][i / 27] = wgt[i];  // wgt: [CIN][27]
    __syncthreads();
    const size_t plane = (size_t)H * W, vol = (size_t)D * plane;
    const int b = blockIdx.y;
    const float* g = gy + (size_t)b * vol;
    float* out = dx + (size_t)b * CIN * vol;
    for (size_t i = (size_t)blockIdx.x * kC1Threads + threadIdx.x; i < vol; i += (size_t)gridDim.x * kC1Threads) {
        const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / plane);
        float acc[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const int dd = d - kd + 1;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h - kh + 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ww = w - kw + 1;
                    const bool ok = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
                    const float v = ok ? __ldg(g + (size_t)dd * plane + (size_t)hh * W + ww) : 0.f;
                    const float* wr = sw[(kd * 3 + kh) * 3 + kw];
#pragma unroll
                    for (int c4 = 0; c4 < CIN / 4; ++c4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wr + c4 * 4);
                        acc[c4 * 4 + 0] = fmaf(wv.x, v, acc[c4 * 4 + 0]);
                        acc[c4 * 4 + 1] = fmaf(wv.y, v, acc[c4 * 4 + 1]);
                        acc[c4 * 4 + 2] = fmaf(wv.z, v, acc[c4 * 4 + 2]);
                        acc[c4 * 4 + 3] = fmaf(wv.w, v, acc[c4 * 4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) out[(size_t)c * vol + i] = acc[c];
    }
}

__global__ void __launch_bounds__(kC1Threads) conv3d_cout1_wgrad_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ gy,
                                                                        float* __restrict__ dw, int Cin, int D, int H,
                                                                        int W) {
    __shared__ float red[kC1Threads / 32][27];
    const size_t plane = (size_t)H * W, vol = (size_t)D * plane;
    const int ci = blockIdx.y, b = blockIdx.z;
    const float* g = gy + (size_t)b * vol;
    const float* xc = x + ((size_t)b * Cin + ci) * vol;
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
    if ((W & 3) == 0) {
        // four consecutive-w voxels per thread: per (kd,kh) one aligned float4 + two edge scalars of X feed 12 FMAs
        const size_t nq = vol >> 2;
        for (size_t q = (size_t)blockIdx.x * kC1Threads + threadIdx.x; q < nq; q += (size_t)gridDim.x * kC1Threads) {
            const size_t i = q << 2;
            const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / plane);
            const float4 gv = *reinterpret_cast<const float4*>(g + i);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int dd = d + kd - 1;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = h + kh - 1;
                    if ((unsigned)dd >= (unsigned)D || (unsigned)hh >= (unsigned)H) continue;
                    const float* row = xc + (size_t)dd * plane + (size_t)hh * W + w;
                    const float4 m = __ldg(reinterpret_cast<const float4*>(row));
                    const float lft = (w > 0) ? __ldg(row - 1) : 0.f;
                    const float rgt = (w + 4 < W) ? __ldg(row + 4) : 0.f;
                    const int t = (kd * 3 + kh) * 3;
                    acc[t + 0] = fmaf(gv.x, lft, fmaf(gv.y, m.x, fmaf(gv.z, m.y, fmaf(gv.w, m.z, acc[t + 0]))));
                    acc[t + 1] = fmaf(gv.x, m.x, fmaf(gv.y, m.y, fmaf(gv.z, m.z, fmaf(gv.w, m.w, acc[t + 1]))));
                    acc[t + 2] = fmaf(gv.x, m.y, fmaf(gv.y, m.z, fmaf(gv.z, m.w, fmaf(gv.w, rgt, acc[t + 2]))));
                }
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kC1Threads + threadIdx.x; i < vol; i += (size_t)gridDim.x * kC1Threads) {
            const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / plane);
            const float gv = g[i];
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int dd = d + kd - 1;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = h + kh - 1;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ww = w + kw - 1;
                        const bool ok = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
                        const float v = ok ? __ldg(xc + (size_t)dd * plane + (size_t)hh * W + ww) : 0.f;
                        acc[(kd * 3 + kh) * 3 + kw] = fmaf(gv, v, acc[(kd * 3 + kh) * 3 + kw]);
                    }
                }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        const float s = warp_sum(acc[t]);
        if (lane == 0) red[warp][t] = s;
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        float s = 0.f;
#pragma unroll
        for (int wp = 0; wp < kC1Threads / 32; ++wp) s += red[wp][threadIdx.x];
        atomicAdd(dw + ci * 27 + threadIdx.x, s);
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_conv3d_cout1_bwd(const float* x, const float* weight, const float* grad_y, float* dx, float* dw,
                                        int B, int Cin, int D, int H, int W, void* stream) {
    CMF_REQUIRE(x && weight && grad_y && dx && dw, "conv3d_cout1_bwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && B <= 65535, "conv3d_cout1_bwd: bad shape");
    CMF_REQUIRE(Cin == 32, "conv3d_cout1_bwd: unsupported Cin=%d (supported: 32)", Cin);
    cudaStream_t st = (cudaStream_t)stream;
    const long long vol = (long long)D * H * W;
    dim3 gd((unsigned)min((long long)kNumSMs * 16, cdiv(vol, kC1Threads)), (unsigned)B);
    conv3d_cout1_dgrad_kernel<32><<<gd, kC1Threads, 0, st>>>(weight, grad_y, dx, D, H, W);
    CMF_LAUNCH_CHECK("conv3d_cout1_dgrad_kernel");
    CMF_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cin * 27 * sizeof(float), st));
    dim3 gw((unsigned)min((long long)32, cdiv(vol, kC1Threads * 16)), (unsigned)Cin, (unsigned)B);
    conv3d_cout1_wgrad_kernel<<<gw, kC1Threads, 0, st>>>(x, grad_y, dw, Cin, D, H, W);
    CMF_LAUNCH_CHECK("conv3d_cout1_wgrad_kernel");
    return CMFB200_OK;
}
