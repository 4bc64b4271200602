// SPP tail of the 2-D feature extractor (reference feature_extraction.forward, cmf/models/cmfsm.py:207-233):
// four average pools (64/32/16/8, stride = kernel, floor mode) of the 128-channel 1/4-resolution map, and -- after
// the per-branch 1x1 conv + GroupNorm + ReLU -- bilinear upsampling (align_corners=False) of the four 32-channel
// branch maps back to 1/4 resolution, concatenated with layer2's and layer4's outputs into the 320-channel input
// of `lastconv`.  Two kernels replace 4 avg_pool2d + 4 F.interpolate + torch.cat:
//   spp_pool:             one pass over the input produces the 8x8 pool; the 16/32/64 pools are 2x2 averages of the
//                         previous level (floor-mode windows nest: floor(H/2k)*2 <= floor(H/k)).
//   spp_upsample_concat:  writes cat = [raw(64) | skip(128) | up(b4) | up(b3) | up(b2) | up(b1)] in one pass.
// HBM-bound (reads 128 + 64 + 128 channels once, writes 320).
#include "common.cuh"

namespace cmfb200 {

// level 0: [BC][H][W] -> [BC][H/8][W/8]; one warp per output pixel row segment: thread per output element
__global__ void spp_pool8_kernel(const float* __restrict__ x, float* __restrict__ p8, int H, int W, int H8, int W8) {
    const size_t bc = blockIdx.y;
    const float* src = x + bc * (size_t)H * W;
    const int n = H8 * W8;
    // one warp per output element: 64 inputs, 2 per lane, coalesced 32-byte row reads
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int oy = warp / W8, ox = warp - oy * W8;
    const int r = lane >> 2, c = (lane & 3) * 2;  // 8 rows x 4 lane-pairs
    const float2 v = *reinterpret_cast<const float2*>(src + (size_t)(oy * 8 + r) * W + ox * 8 + c);
    const float s = warp_sum(v.x + v.y);
    if (lane == 0) p8[bc * n + warp] = s * (1.0f / 64.0f);
}

// level k -> 2k: 2x2 average, floor mode
__global__ void spp_pool2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int Hi, int Wi, int Ho, int Wo,
                                   int BC) {
    const int n = Ho * Wo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BC * n; i += gridDim.x * blockDim.x) {
        const int bc = i / n, r = i - bc * n;
        const int oy = r / Wo, ox = r - oy * Wo;
        const float* p = in + (size_t)bc * Hi * Wi + (size_t)(2 * oy) * Wi + 2 * ox;
        out[i] = ((p[0] + p[1]) + (p[Wi] + p[Wi + 1])) * 0.25f;
    }
}

struct BranchMap {
    const float* ptr;  // [B][32][h][w]
    int h, w;
};

__device__ __forceinline__ float bilinear_at(const float* __restrict__ m, int h, int w, float sy, float sx) {
    // PyTorch upsample_bilinear2d, align_corners=False: negative source coordinates clamp to 0
    sy = fmaxf(sy, 0.f);
    sx = fmaxf(sx, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    return hy * (hx * m[y0 * w + x0] + lx * m[y0 * w + x1]) + ly * (hx * m[y1 * w + x0] + lx * m[y1 * w + x1]);
}

// cat[b][0:64] = raw, [64:192] = skip, [192:224] = up(branch4), [224:256] = up(b3), [256:288] = up(b2), [288:320] = up(b1)
__global__ void spp_upsample_concat_kernel(const float* __restrict__ raw, const float* __restrict__ skip, BranchMap b4,
                                           BranchMap b3, BranchMap b2, BranchMap b1, float* __restrict__ cat, int H,
                                           int W, int H_full, int y_off, int raw_c) {
    const int b = blockIdx.z, c = blockIdx.y;  // c in [0, raw_c + 128 + 128)
    const int ctot = raw_c + 256, cplain = raw_c + 128;
    const size_t plane = (size_t)H * W;
    float* dst = cat + ((size_t)b * ctot + c) * plane;
    if (c < cplain) {  // plain copies (128-bit when the plane allows)
        const float* src = c < raw_c ? raw + ((size_t)b * raw_c + c) * plane : skip + ((size_t)b * 128 + (c - raw_c)) * plane;
        if ((plane & 3) == 0) {
            for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < plane; i += (size_t)gridDim.x * blockDim.x * 4)
                *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
        } else {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
                dst[i] = src[i];
        }
        return;
    }
    const int k = (c - cplain) >> 5, ch = (c - cplain) & 31;
    const BranchMap bm = k == 0 ? b4 : (k == 1 ? b3 : (k == 2 ? b2 : b1));
    const float* m = bm.ptr + ((size_t)b * 32 + ch) * bm.h * bm.w;
    const float ry = (float)bm.h / (float)H_full, rx = (float)bm.w / (float)W;  // branch maps cover the FULL image
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        dst[i] = bilinear_at(m, bm.h, bm.w, ry * ((float)(y + y_off) + 0.5f) - 0.5f, rx * ((float)x + 0.5f) - 0.5f);
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_spp_pool_fwd(const float* x, float* p8, float* p16, float* p32, float* p64, int B, int C, int H,
                                    int W, void* stream) {
    CMF_REQUIRE(x && p8 && p16 && p32 && p64, "spp_pool_fwd: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && H >= 64 && W >= 64, "spp_pool_fwd: the 64x64 pool needs H,W >= 64 (got %dx%d)", H, W);
    CMF_REQUIRE(W % 2 == 0, "spp_pool_fwd: W must be even");
    CMF_REQUIRE((long long)B * C <= 65535, "spp_pool_fwd: B*C exceeds grid limit");
    cudaStream_t st = (cudaStream_t)stream;
    const int BC = B * C;
    const int H8 = H / 8, W8 = W / 8, H16 = H / 16, W16 = W / 16, H32 = H / 32, W32 = W / 32, H64 = H / 64, W64 = W / 64;
    dim3 g8((unsigned)cdiv((long long)H8 * W8 * 32, 256), (unsigned)BC);
    spp_pool8_kernel<<<g8, 256, 0, st>>>(x, p8, H, W, H8, W8);
    CMF_LAUNCH_CHECK("spp_pool8_kernel");
    spp_pool2x2_kernel<<<(unsigned)cdiv((long long)BC * H16 * W16, 256), 256, 0, st>>>(p8, p16, H8, W8, H16, W16, BC);
    CMF_LAUNCH_CHECK("spp_pool2x2_kernel");
    spp_pool2x2_kernel<<<(unsigned)cdiv((long long)BC * H32 * W32, 256), 256, 0, st>>>(p16, p32, H16, W16, H32, W32, BC);
    CMF_LAUNCH_CHECK("spp_pool2x2_kernel");
    spp_pool2x2_kernel<<<(unsigned)cdiv((long long)BC * H64 * W64, 256), 256, 0, st>>>(p32, p64, H32, W32, H64, W64, BC);
    CMF_LAUNCH_CHECK("spp_pool2x2_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_spp_upsample_concat_fwd(const float* raw, const float* skip, const float* b4, const float* b3,
                                               const float* b2, const float* b1, float* cat, int B, int H, int W,
                                               int H_full, int y_off, void* stream) {
    CMF_REQUIRE(raw && skip && b4 && b3 && b2 && b1 && cat, "spp_upsample_concat_fwd: null pointer");
    CMF_REQUIRE(B > 0 && H > 0 && H_full >= 64 && W >= 64 && B <= 65535, "spp_upsample_concat_fwd: bad shape");
    CMF_REQUIRE(y_off >= 0 && y_off + H <= H_full, "spp_upsample_concat_fwd: rows [%d,%d) outside the image height %d", y_off, y_off + H, H_full);
    const BranchMap m4{b4, H_full / 8, W / 8}, m3{b3, H_full / 16, W / 16}, m2{b2, H_full / 32, W / 32}, m1{b1, H_full / 64, W / 64};
    dim3 grid((unsigned)min((long long)16, cdiv((long long)H * W, 1024)), 320, (unsigned)B);
    spp_upsample_concat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, skip, m4, m3, m2, m1, cat, H, W, H_full, y_off, 64);
    CMF_LAUNCH_CHECK("spp_upsample_concat_kernel");
    return CMFB200_OK;
}

// Same kernel with explicit branch-map sizes and raw-channel count (cmfsm_sub_8: pools 8/16/32/4 on the 1/8 map, 64 raw
// channels, cmfsm_sub_8.py:152-170, 207-231; cmfsm_sub_16: pools 8/16/2/4 on the 1/16 map, 128 raw channels):
// cat = [raw(raw_c) | skip(128) | up(ba) | up(bb) | up(bc) | up(bd)].
extern "C" int cmfb200_spp_upsample_concat_sized_fwd(const float* raw, const float* skip, const float* ba,
                                                     const float* bb, const float* bc, const float* bd, float* cat,
                                                     int B, int raw_c, int H, int W, int ha, int wa, int hb, int wb, int hc,
                                                     int wc, int hd, int wd, void* stream) {
    CMF_REQUIRE(raw && skip && ba && bb && bc && bd && cat, "spp_upsample_concat_sized_fwd: null pointer");
    CMF_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (raw_c == 64 || raw_c == 128), "spp_upsample_concat_sized_fwd: bad shape");
    CMF_REQUIRE(ha > 0 && wa > 0 && hb > 0 && wb > 0 && hc > 0 && wc > 0 && hd > 0 && wd > 0,
                "spp_upsample_concat_sized_fwd: empty branch map");
    const BranchMap ma{ba, ha, wa}, mb{bb, hb, wb}, mc{bc, hc, wc}, md{bd, hd, wd};
    dim3 grid((unsigned)min((long long)16, cdiv((long long)H * W, 1024)), (unsigned)(raw_c + 256), (unsigned)B);
    spp_upsample_concat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, skip, ma, mb, mc, md, cat, H, W, H, 0, raw_c);
    CMF_LAUNCH_CHECK("spp_upsample_concat_kernel");
    return CMFB200_OK;
}
