// Epilogue of the 1/16-resolution variant cmfsm_sub_16 (cmf/models/cmfsm_sub_16.py:760-850): cost-VOLUME mapping.
// The reference nearest-upsamples each classifier volume to [B, maxdisp, H, W] (805 MB at 540x960), mixes it over the
// five spatial neighbours, builds three more [B, maxdisp, H, W] volumes from the target-image weights shifted by the
// disparity, mixes along the disparity axis and regresses with a softmax over maxdisp planes -- three times.
// Here nothing of that is materialised: one thread per full-resolution pixel.
//   * The upsampled cost only depends on (d/scale, cell), so the spatially mixed value is F_n[j] for j < D' = maxdisp/scale:
//         F_n[j] = w_c c_n[j,cell] + w_r c_n[j,cell+x] + w_l c_n[j,cell-x] + w_t c_n[j,cell-y] + w_b c_n[j,cell+y]
//     (c_2 += c_1, c_3 += c_2 accumulated first; neighbours outside the image add nothing) -- 3 x D' values per pixel,
//     parked in shared memory.
//   * For d = 0..maxdisp-1:  v = F[d/s] t(x-d) + F[d/s+1] l(x-d) [d < maxdisp-s] + F[d/s-1] r(x-d) [d >= s], with the
//     target weights t, r, l read at column x-d (1 where x < d), pushed into three online softmax regressions.
// Accumulation order and non-fused mul/add follow the reference's statement order.
#include "common.cuh"

namespace cmfb200 {

constexpr int kVmThreads = 128;

struct VmOnline {
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

__global__ void __launch_bounds__(kVmThreads) volume_mapping_kernel(
    const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
    const float* __restrict__ w5, const float* __restrict__ w3, float* __restrict__ out1, float* __restrict__ out2,
    float* __restrict__ out3, int Dl, int h, int w, int scale) {
    extern __shared__ float sF[];  // [3][Dl][kVmThreads]
    const int b = blockIdx.y;
    const int H = h * scale, W = w * scale;
    const size_t hplane = (size_t)H * W, lplane = (size_t)h * w;
    const size_t pix = (size_t)blockIdx.x * kVmThreads + threadIdx.x;
    if (pix >= hplane) return;  // no barriers below
    const int y = (int)(pix / W), x = (int)(pix - (size_t)y * W);
    const int cy = y / scale, cx = x / scale;
    const float* pw = w5 + (size_t)b * 5 * hplane + pix;
    const float wc = pw[0], wr = pw[hplane], wl = pw[2 * hplane], wt = pw[3 * hplane], wb = pw[4 * hplane];
    const bool has_r = cx + 1 < w, has_l = cx > 0, has_t = cy > 0, has_b = cy + 1 < h;
    const size_t cell = (size_t)cy * w + cx;
    for (int j = 0; j < Dl; ++j) {
        const size_t off = ((size_t)b * Dl + j) * lplane + cell;
        float acc[3];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const bool ok = k == 0 || (k == 1 && has_r) || (k == 2 && has_l) || (k == 3 && has_t) || (k == 4 && has_b);
            if (!ok) continue;
            const long long sh = k == 1 ? 1 : (k == 2 ? -1 : (k == 3 ? -(long long)w : (k == 4 ? (long long)w : 0)));
            const float wk = k == 0 ? wc : (k == 1 ? wr : (k == 2 ? wl : (k == 3 ? wt : wb)));
            const float v1 = c1[off + sh];
            const float v2 = __fadd_rn(c2[off + sh], v1);  // cost2 = up(c2) + cost1, cost3 = up(c3) + cost2
            const float v3 = __fadd_rn(c3[off + sh], v2);
            if (k == 0) {
                acc[0] = __fmul_rn(v1, wk);
                acc[1] = __fmul_rn(v2, wk);
                acc[2] = __fmul_rn(v3, wk);
            } else {
                acc[0] = __fadd_rn(acc[0], __fmul_rn(v1, wk));
                acc[1] = __fadd_rn(acc[1], __fmul_rn(v2, wk));
                acc[2] = __fadd_rn(acc[2], __fmul_rn(v3, wk));
            }
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) sF[(n * Dl + j) * kVmThreads + threadIdx.x] = acc[n];
    }
    const int maxdisp = Dl * scale;
    const float* pt = w3 + (size_t)b * 3 * hplane + (size_t)y * W;  // target weights of this row: centre, right, left
    VmOnline o[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) o[n].init();
    for (int j = 0; j < Dl; ++j) {
        float f0[3], fp[3], fm[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) {
            f0[n] = sF[(n * Dl + j) * kVmThreads + threadIdx.x];
            fp[n] = j + 1 < Dl ? sF[(n * Dl + j + 1) * kVmThreads + threadIdx.x] : 0.f;
            fm[n] = j > 0 ? sF[(n * Dl + j - 1) * kVmThreads + threadIdx.x] : 0.f;
        }
        for (int dd = 0; dd < scale; ++dd) {
            const int d = j * scale + dd;
            float vt = 1.f, vr = 1.f, vl = 1.f;
            if (x >= d) {
                vt = pt[x - d];
                vr = pt[hplane + x - d];
                vl = pt[2 * hplane + x - d];
            }
            const float fd = (float)d;
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                float v = __fmul_rn(f0[n], vt);
                if (d < maxdisp - scale) v = __fadd_rn(v, __fmul_rn(fp[n], vl));
                if (d >= scale) v = __fadd_rn(v, __fmul_rn(fm[n], vr));
                o[n].push(v, fd);
            }
        }
    }
    out1[(size_t)b * hplane + pix] = o[0].result();
    out2[(size_t)b * hplane + pix] = o[1].result();
    out3[(size_t)b * hplane + pix] = o[2].result();
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_volume_mapping_fwd(const float* c1, const float* c2, const float* c3, const float* weights5,
                                          const float* weights3, float* out1, float* out2, float* out3, int B, int Dl,
                                          int h, int w, int scale, void* stream) {
    CMF_REQUIRE(c1 && c2 && c3 && weights5 && weights3 && out1 && out2 && out3, "volume_mapping_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Dl > 0 && h > 0 && w > 0 && scale > 0 && B <= 65535, "volume_mapping_fwd: bad shape");
    const size_t smem = (size_t)3 * Dl * kVmThreads * sizeof(float);
    CMF_REQUIRE(smem <= 200 * 1024, "volume_mapping_fwd: D'=%d too large", Dl);
    CMF_CUDA(cudaFuncSetAttribute(volume_mapping_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long hplane = (long long)h * scale * w * scale;
    dim3 grid((unsigned)cdiv(hplane, kVmThreads), (unsigned)B);
    volume_mapping_kernel<<<grid, kVmThreads, smem, (cudaStream_t)stream>>>(c1, c2, c3, weights5, weights3, out1, out2, out3,
                                                                            Dl, h, w, scale);
    CMF_LAUNCH_CHECK("volume_mapping_kernel");
    return CMFB200_OK;
}
