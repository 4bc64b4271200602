// Epilogue of the `bilinear_cmf*` baselines (reference cmf/models/bilinear_cmf.py:418-452 and the _sub_8/_sub_16
// files): the three classifier volumes are accumulated at low resolution (cost2 = c2 + cost1, cost3 = c3 + cost2),
// trilinearly upsampled (align_corners=False) to [B, maxdisp, H, W], soft-maxed over maxdisp and regressed.  The
// reference materialises three [B,maxdisp,H,W] volumes; here one thread per output pixel first interpolates the D' low-res
// planes bilinearly at its (y,x) (kept in shared memory), then walks the maxdisp planes, interpolating linearly along d
// into three online softmax regressions.  Same nesting of the interpolation as ATen: d( h( w ) ).
#include "common.cuh"

namespace cmfb200 {

constexpr int kTlThreads = 128;

struct TlOnline {
    float m, s, t;
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; t = 0.f; }
    __device__ __forceinline__ void push(float v, float d) {
        if (v > m) {
            const float r = expf(m - v);
            s *= r;
            t *= r;
            m = v;
        }
        const float e = expf(v - m);
        s += e;
        t = fmaf(d, e, t);
    }
    __device__ __forceinline__ float result() const { return t / s; }
};

// ATen area_pixel_compute_source_index, align_corners=False: max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void tl_src(float scale, int dst, int n_in, int& i0, int& i1, float& l0, float& l1) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(kTlThreads) trilinear_softargmin_kernel(
    const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3, float* __restrict__ out1,
    float* __restrict__ out2, float* __restrict__ out3, int Dl, int h, int w, int maxdisp, int H, int W) {
    extern __shared__ float sS[];  // [3][Dl][kTlThreads]
    const int b = blockIdx.y;
    const size_t hplane = (size_t)H * W, lplane = (size_t)h * w;
    const size_t pix = (size_t)blockIdx.x * kTlThreads + threadIdx.x;
    if (pix >= hplane) return;  // no barriers below
    const int y = (int)(pix / W), x = (int)(pix - (size_t)y * W);
    int y0, y1, x0, x1;
    float hy, ly, hx, lx;
    tl_src((float)h / (float)H, y, h, y0, y1, hy, ly);
    tl_src((float)w / (float)W, x, w, x0, x1, hx, lx);
    const size_t o00 = (size_t)y0 * w + x0, o01 = (size_t)y0 * w + x1, o10 = (size_t)y1 * w + x0, o11 = (size_t)y1 * w + x1;
    for (int j = 0; j < Dl; ++j) {
        const size_t base = ((size_t)b * Dl + j) * lplane;
        float v[3][4];
        const size_t offs[4] = {o00, o01, o10, o11};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[0][q] = c1[base + offs[q]];
            v[1][q] = __fadd_rn(c2[base + offs[q]], v[0][q]);  // cumulative sums at low resolution, as the reference
            v[2][q] = __fadd_rn(c3[base + offs[q]], v[1][q]);
        }
#pragma unroll
        for (int n = 0; n < 3; ++n)
            sS[(n * Dl + j) * kTlThreads + threadIdx.x] =
                hy * (hx * v[n][0] + lx * v[n][1]) + ly * (hx * v[n][2] + lx * v[n][3]);
    }
    TlOnline o[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) o[n].init();
    const float dscale = (float)Dl / (float)maxdisp;
    for (int d = 0; d < maxdisp; ++d) {
        int j0, j1;
        float t0, t1;
        tl_src(dscale, d, Dl, j0, j1, t0, t1);
        const float fd = (float)d;
#pragma unroll
        for (int n = 0; n < 3; ++n) {
            const float val = t0 * sS[(n * Dl + j0) * kTlThreads + threadIdx.x] + t1 * sS[(n * Dl + j1) * kTlThreads + threadIdx.x];
            o[n].push(val, fd);
        }
    }
    out1[(size_t)b * hplane + pix] = o[0].result();
    out2[(size_t)b * hplane + pix] = o[1].result();
    out3[(size_t)b * hplane + pix] = o[2].result();
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_trilinear_softargmin_fwd(const float* c1, const float* c2, const float* c3, float* out1, float* out2,
                                                float* out3, int B, int Dl, int h, int w, int maxdisp, int H, int W,
                                                void* stream) {
    CMF_REQUIRE(c1 && c2 && c3 && out1 && out2 && out3, "trilinear_softargmin_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Dl > 0 && h > 0 && w > 0 && maxdisp > 0 && H > 0 && W > 0 && B <= 65535,
                "trilinear_softargmin_fwd: bad shape");
    const size_t smem = (size_t)3 * Dl * kTlThreads * sizeof(float);
    CMF_REQUIRE(smem <= 200 * 1024, "trilinear_softargmin_fwd: D'=%d too large", Dl);
    CMF_CUDA(cudaFuncSetAttribute(trilinear_softargmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)cdiv((long long)H * W, kTlThreads), (unsigned)B);
    trilinear_softargmin_kernel<<<grid, kTlThreads, smem, (cudaStream_t)stream>>>(c1, c2, c3, out1, out2, out3, Dl, h, w,
                                                                                  maxdisp, H, W);
    CMF_LAUNCH_CHECK("trilinear_softargmin_kernel");
    return CMFB200_OK;
}
