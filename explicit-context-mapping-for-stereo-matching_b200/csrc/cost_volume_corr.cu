// K1 (correlation form) -- corr[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d] for x >= d, +0.0 otherwise; with
// `normalize` the cosine similarity sum_c L R / max(|L| |R|, 1e-8) of the two feature vectors instead.
//
// The reference's live models build the CONCAT volume (cmf/models/cmfsm.py:667-682, cost_volume.cu); its only
// correlation matching is the cosine similarity of left features with shifted right features in the dead file
// "cmf/models/rstereo # dense volume match.py":309-311.  This kernel is that operation with K1's shift / mask
// convention (R[x-d], zero where x < d), so parity for it is pinned by the CPU restatement in oracle/ only.
//
// HBM-bound and tiny (reads 2*C*h*w floats once, writes D*h*w): a CTA owns one image row, stages the C x w rows of
// both feature maps in shared memory (odd pitch, right row behind a zero prefix of D entries), and every warp walks
// (d, 32-pixel chunk) items with lane = channel: 32 products per lane, then ONE transposing warp-shuffle reduction
// (31 shuffles) leaves the channel sum of pixel x0+l in lane l -> coalesced 128-byte stores.
#include "common.cuh"

namespace cmfb200 {

namespace {
__device__ __forceinline__ float transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float give = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, give, off);
        }
    }
    return v[0];
}
}  // namespace

__global__ void __launch_bounds__(256) cost_volume_corr_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                               float* __restrict__ out, int C, int h, int w, int D,
                                                               int normalize) {
    extern __shared__ float sm[];
    const int y = blockIdx.x, b = blockIdx.y;
    const int wp = ((w + 31) & ~31);              // rows are read in 32-pixel chunks
    const int pl = wp | 1, pr = (D + wp) | 1;     // odd pitches: lane = channel reads are bank-conflict free
    float* sL = sm;                               // [C][pl]
    float* sR = sL + (size_t)C * pl;              // [C][pr], pixel x at column D + x
    float* nL = sR + (size_t)C * pr;              // [wp] |L| per pixel (normalize)
    float* nR = nL + wp;                          // [D + wp] |R| per pixel, same zero prefix
    const size_t plane = (size_t)h * w;
    for (int i = threadIdx.x; i < C * pl; i += 256) {
        const int c = i / pl, x = i - c * pl;
        sL[i] = x < w ? __ldg(L + ((size_t)b * C + c) * plane + (size_t)y * w + x) : 0.f;
    }
    for (int i = threadIdx.x; i < C * pr; i += 256) {
        const int c = i / pr, x = i - c * pr - D;
        sR[i] = (x >= 0 && x < w) ? __ldg(R + ((size_t)b * C + c) * plane + (size_t)y * w + x) : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = wp / 32;
    if (normalize) {
        for (int it = warp; it < 2 * chunks; it += 8) {
            const bool right = it >= chunks;
            const int x0 = (right ? it - chunks : it) * 32;
            float acc = 0.f;
            for (int c0 = 0; c0 < C; c0 += 32) {
                float v[32];
                const float* row = right ? sR + (size_t)(c0 + lane) * pr + D + x0 : sL + (size_t)(c0 + lane) * pl + x0;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (c0 + lane < C) ? row[i] * row[i] : 0.f;
                acc += transpose_sum32(v, lane);
            }
            (right ? nR + D : nL)[x0 + lane] = sqrtf(acc);
        }
        for (int i = threadIdx.x; i < D; i += 256) nR[i] = 0.f;
        __syncthreads();
    }
    const float inv_c = 1.f / (float)C;
    float* dst = out + (size_t)b * D * plane + (size_t)y * w;
    for (int it = warp; it < D * chunks; it += 8) {
        const int d = it / chunks, x0 = (it - d * chunks) * 32;
        float acc = 0.f;
        for (int c0 = 0; c0 < C; c0 += 32) {
            float v[32];
            const float* l = sL + (size_t)(c0 + lane) * pl + x0;
            const float* r = sR + (size_t)(c0 + lane) * pr + D + x0 - d;  // columns < D hold zeros: x < d contributes 0
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (c0 + lane < C) ? l[i] * r[i] : 0.f;
            acc += transpose_sum32(v, lane);
        }
        const int x = x0 + lane;
        if (x < w) {
            float o;
            if (normalize)
                o = x >= d ? acc / fmaxf(nL[x] * nR[D + x - d], 1e-8f) : 0.f;
            else
                o = x >= d ? acc * inv_c : 0.f;
            dst[(size_t)d * plane + x] = o;
        }
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_cost_volume_corr_fwd(const float* L, const float* R, float* corr, int B, int C, int h, int w,
                                            int D, int normalize, void* stream) {
    CMF_REQUIRE(L && R && corr, "cost_volume_corr_fwd: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && D > 0, "cost_volume_corr_fwd: non-positive dimension");
    CMF_REQUIRE(B <= 65535, "cost_volume_corr_fwd: B exceeds grid limit");
    const int wp = (w + 31) & ~31;
    const size_t smem = ((size_t)C * (wp | 1) + (size_t)C * ((D + wp) | 1) + wp + D + wp) * sizeof(float);
    CMF_REQUIRE(smem <= 200 * 1024, "cost_volume_corr_fwd: a feature row pair (C=%d, w=%d, D=%d) does not fit in shared memory",
                C, w, D);
    if (smem > 48 * 1024)
        CMF_CUDA(cudaFuncSetAttribute(cost_volume_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)h, (unsigned)B);
    cost_volume_corr_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(L, R, corr, C, h, w, D, normalize);
    CMF_LAUNCH_CHECK("cost_volume_corr_kernel");
    return CMFB200_OK;
}
