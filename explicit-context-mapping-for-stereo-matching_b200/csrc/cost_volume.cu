// K1 -- concat cost volume, fp32 NCDHW, bit-exact with the reference slice-copy loop
// (cmf/models/cmfsm.py:667-682; contract in SURVEY.md appendix A.1).
//
// HBM-bound: reads 2*B*C*h*w*4 bytes once, writes B*2C*D*h*w*4 bytes (zeros included).
// One CTA owns (b, c, a block of ROWS image rows): it stages the ROWS left-feature rows and the ROWS
// right-feature rows in shared memory with two 1-D bulk copies (TMA engine, mbarrier completion), the
// right rows behind a zero prefix so that R[x-d] for x<d reads +0.0 without a branch, and then streams
// the 2*D shifted / masked copies out with 128-bit no-allocate stores.  For fixed (c,d) the ROWS rows
// are contiguous in the output, so every warp store instruction covers one contiguous 512 B run.
#include "common.cuh"

namespace cmfb200 {

constexpr int kCvThreads = 256;
constexpr int kCvRows = 4;

// smem layout (floats): sL[ROWS*w] | per row r: zero prefix [DP] + sR row [w]   (DP = D rounded up to 4)
__global__ void __launch_bounds__(kCvThreads) cost_volume_concat_fwd_kernel(
    const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ cost, int C, int h, int w, int D,
    int DP) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;

    const int y0 = blockIdx.x * kCvRows;
    const int c = blockIdx.y;
    const int b = blockIdx.z;
    const int rows = min(kCvRows, h - y0);
    const int tid = threadIdx.x;
    const int rp = DP + w;  // pitch of a zero-prefixed right row

    float* sL = smem;
    float* sZR = smem + kCvRows * w;

    const size_t in_off = (((size_t)b * C + c) * h + y0) * w;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        // rows of one channel are contiguous in NCHW: one bulk copy for the left block, one per right row
        mbar_arrive_expect_tx(&bar, (uint32_t)(2 * rows * w * sizeof(float)));
        bulk_g2s(sL, L + in_off, (uint32_t)(rows * w * sizeof(float)), &bar);
        for (int r = 0; r < rows; ++r)
            bulk_g2s(sZR + r * rp + DP, R + in_off + (size_t)r * w, (uint32_t)(w * sizeof(float)), &bar);
    }
    for (int i = tid; i < rows * DP; i += kCvThreads) sZR[(i / DP) * rp + (i % DP)] = 0.0f;
    __syncthreads();  // barrier init + zero prefix visible to everyone
    mbar_wait(&bar, 0);

    const int w4 = w >> 2;
    const int per_d = rows * w4;  // float4 per (half, d) owned by this CTA
    const size_t plane = (size_t)h * w;
    float* outL = cost + (((size_t)b * 2 * C + c) * D) * plane + (size_t)y0 * w;
    float* outR = outL + (size_t)C * D * plane;

    const int warp = tid >> 5, lane = tid & 31;
    for (int d = warp; d < D; d += kCvThreads / 32) {
        float* oL = outL + (size_t)d * plane;
        float* oR = outR + (size_t)d * plane;
        for (int i = lane; i < per_d; i += 32) {
            const int r = i / w4;
            const int x = (i - r * w4) << 2;
            float4 l = *reinterpret_cast<const float4*>(sL + r * w + x);
            l.x = (x + 0 >= d) ? l.x : 0.0f;
            l.y = (x + 1 >= d) ? l.y : 0.0f;
            l.z = (x + 2 >= d) ? l.z : 0.0f;
            l.w = (x + 3 >= d) ? l.w : 0.0f;
            const float* pr = sZR + r * rp + DP + x - d;
            float4 rr = make_float4(pr[0], pr[1], pr[2], pr[3]);
            st_streaming_f4(oL + (size_t)r * w + x, l);
            st_streaming_f4(oR + (size_t)r * w + x, rr);
        }
    }
}

// Adjoint.  One CTA per (y, c, b); thread per x.
__global__ void __launch_bounds__(128) cost_volume_concat_bwd_kernel(const float* __restrict__ g,
                                                                      float* __restrict__ dL, float* __restrict__ dR,
                                                                      int C, int h, int w, int D) {
    const int y = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const size_t plane = (size_t)h * w;
    const float* gL = g + (((size_t)b * 2 * C + c) * D) * plane + (size_t)y * w;
    const float* gR = gL + (size_t)C * D * plane;
    const size_t o = (((size_t)b * C + c) * h + y) * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        float sl = 0.f, sr = 0.f;
        const int dl = min(x, D - 1);          // d <= x
        const int dr = min(D - 1, w - 1 - x);  // x + d < w
        for (int d = 0; d <= dl; ++d) sl += gL[(size_t)d * plane + x];
        for (int d = 0; d <= dr; ++d) sr += gR[(size_t)d * plane + x + d];
        dL[o + x] = sl;
        dR[o + x] = sr;
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_cost_volume_concat_fwd(const float* L, const float* R, float* cost, int B, int C, int h, int w,
                                              int D, void* stream) {
    CMF_REQUIRE(L && R && cost, "cost_volume_concat_fwd: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && D > 0, "cost_volume_concat_fwd: non-positive dimension");
    CMF_REQUIRE(w % 4 == 0, "cost_volume_concat_fwd: w=%d must be a multiple of 4", w);
    CMF_REQUIRE(C <= 65535 && B <= 65535, "cost_volume_concat_fwd: C/B exceed grid limits");
    const int DP = (D + 3) & ~3;
    const size_t smem = (size_t)(kCvRows * w + kCvRows * (DP + w)) * sizeof(float);
    CMF_REQUIRE(smem <= 200 * 1024, "cost_volume_concat_fwd: row block does not fit in shared memory (w=%d, D=%d)", w, D);
    if (smem > 48 * 1024)
        CMF_CUDA(cudaFuncSetAttribute(cost_volume_concat_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    dim3 grid((unsigned)cdiv(h, kCvRows), (unsigned)C, (unsigned)B);
    cost_volume_concat_fwd_kernel<<<grid, kCvThreads, smem, (cudaStream_t)stream>>>(L, R, cost, C, h, w, D, DP);
    CMF_LAUNCH_CHECK("cost_volume_concat_fwd_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_cost_volume_concat_bwd(const float* g, float* dL, float* dR, int B, int C, int h, int w, int D,
                                              void* stream) {
    CMF_REQUIRE(g && dL && dR, "cost_volume_concat_bwd: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && D > 0, "cost_volume_concat_bwd: non-positive dimension");
    CMF_REQUIRE(C <= 65535 && B <= 65535, "cost_volume_concat_bwd: C/B exceed grid limits");
    dim3 grid((unsigned)h, (unsigned)C, (unsigned)B);
    cost_volume_concat_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g, dL, dR, C, h, w, D);
    CMF_LAUNCH_CHECK("cost_volume_concat_bwd_kernel");
    return CMFB200_OK;
}
