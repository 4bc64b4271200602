// K5 -- context-mapping weights: nine neighbour logits from the similarity MLP 66-32-16-8-1 (1x1 convs,
// LeakyReLU 0.01) followed by a softmax over the nine neighbours.
// Replaces eight_related_context_mapping.forward + similarity_measure1 (cmf/models/cmfsm.py:443-593,
// 304-358); closed form in SURVEY.md appendix A.3.
//
// The first layer is linear, so  W0 . [lr ; hr ; code] = W0[:,0:32].lr(cell) + W0[:,32:64].hr(pixel) +
// W0[:,64:66].code  : the lr part is evaluated once per low-res cell (block + 1-cell halo, shared memory),
// the hr part once per pixel (registers) and both are shared by the nine neighbours -- 3.6x fewer MACs
// than evaluating the MLP nine times, and none of the reference's 9 concatenated [B,66,H,W] tensors exist.
// One thread per full-resolution pixel; a CTA covers 2 x 8 low-res cells (8 x 32 pixels at scale 4).
#include "common.cuh"

namespace cmfb200 {

constexpr int kK5Threads = 256;
constexpr int kK5CellsY = 2, kK5CellsX = 8;
constexpr int kK5HaloY = kK5CellsY + 2, kK5HaloX = kK5CellsX + 2;
constexpr int kK5Halo = kK5HaloY * kK5HaloX;  // 40
constexpr int kPad = 36;                       // padded stride of a 32-vector: conflict-free 128-bit reads

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.01f * v; }

// positional codes (matrix_generation, cmfsm.py:391-428): kind 0 = off, 1 = inc, 2 = dec
__device__ __forceinline__ float pos_code(int kind, int i, int s) {
    if (kind == 1) return (float)(i + 1);
    if (kind == 2) return (float)(s - i);
    const int half = s >> 1;
    return (float)(i < half ? i - half : i - half + 1);
}

__global__ void __launch_bounds__(kK5Threads) ctxmap_weights_kernel(
    const float* __restrict__ lr, const float* __restrict__ hr, const float* __restrict__ w0,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
    float* __restrict__ out, int h, int w, int scale, int vy0, int vy1) {
    __shared__ __align__(16) float sW0lr[32][32];  // [in][out]
    __shared__ __align__(16) float sW0hr[32][32];  // [in][out]
    __shared__ __align__(16) float sW0c[2][32];    // code channels 64,65
    __shared__ __align__(16) float sW1[32][16];    // [in][out]
    __shared__ __align__(16) float sW2[16][8];
    __shared__ __align__(16) float sW3[8];
    __shared__ __align__(16) float sLr[32][kK5Halo];
    __shared__ __align__(16) float sAlr[kK5Halo][kPad];
    __shared__ float sLogit[9][kK5Threads];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int cy0 = blockIdx.y * kK5CellsY, cx0 = blockIdx.x * kK5CellsX;
    const int H = h * scale, W = w * scale;
    const size_t lplane = (size_t)h * w, hplane = (size_t)H * W;

    // ---- stage weights (transposed to [in][out]) and the lr halo block
    for (int i = tid; i < 32 * 66; i += kK5Threads) {
        const int o = i / 66, c = i - o * 66;
        const float v = w0[i];
        if (c < 32) sW0lr[c][o] = v;
        else if (c < 64) sW0hr[c - 32][o] = v;
        else sW0c[c - 64][o] = v;
    }
    for (int i = tid; i < 16 * 32; i += kK5Threads) sW1[i % 32][i / 32] = w1[i];
    if (tid < 8 * 16) sW2[tid % 16][tid / 16] = w2[tid];
    if (tid < 8) sW3[tid] = w3[tid];
    for (int i = tid; i < 32 * kK5Halo; i += kK5Threads) {
        const int c = i / kK5Halo, t = i - c * kK5Halo;
        const int cy = cy0 + t / kK5HaloX - 1, cx = cx0 + t % kK5HaloX - 1;
        float v = 0.f;
        if (cy >= 0 && cy < h && cx >= 0 && cx < w) v = lr[((size_t)b * 32 + c) * lplane + (size_t)cy * w + cx];
        sLr[c][t] = v;
    }
    __syncthreads();
    // ---- lr half of layer 0, once per halo cell
    for (int i = tid; i < kK5Halo * 32; i += kK5Threads) {
        const int t = i >> 5, o = i & 31;
        float a = 0.f;
#pragma unroll 8
        for (int c = 0; c < 32; ++c) a = fmaf(sW0lr[c][o], sLr[c][t], a);
        sAlr[t][o] = a;
    }
    __syncthreads();

    // ---- per pixel
    const int tile_w = kK5CellsX * scale;  // pixels per tile row
    const int tile_px = kK5CellsY * scale * tile_w;
    for (int p = tid; p < tile_px; p += kK5Threads) {
        const int ty = p / tile_w, tx = p - ty * tile_w;
        const int y = cy0 * scale + ty, x = cx0 * scale + tx;
        if (y >= H || x >= W) continue;  // (no barriers below)
        const int ly = ty / scale + 1, lx = tx / scale + 1;
        const int cy = cy0 + ly - 1, cx = cx0 + lx - 1;
        const int py = y % scale, px = x % scale;

        // hr half of layer 0
        float ahr[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) ahr[o] = 0.f;
        const float* ph = hr + (size_t)b * 32 * hplane + (size_t)y * W + x;
#pragma unroll 4
        for (int c = 0; c < 32; ++c) {
            const float v = ph[c * hplane];
#pragma unroll
            for (int o4 = 0; o4 < 8; ++o4) {
                const float4 wv = *reinterpret_cast<const float4*>(&sW0hr[c][o4 * 4]);
                ahr[o4 * 4 + 0] = fmaf(wv.x, v, ahr[o4 * 4 + 0]);
                ahr[o4 * 4 + 1] = fmaf(wv.y, v, ahr[o4 * 4 + 1]);
                ahr[o4 * 4 + 2] = fmaf(wv.z, v, ahr[o4 * 4 + 2]);
                ahr[o4 * 4 + 3] = fmaf(wv.w, v, ahr[o4 * 4 + 3]);
            }
        }

        // neighbours in the reference order c,l,r,t,b,lt,rt,lb,rb; code kinds per SURVEY.md A.3
        // (the diagonals reuse the axis encodings, cmfsm.py:459-462)
        for (int k = 0; k < 9; ++k) {
            const int dy = (k == 3 || k == 5 || k == 6) ? -1 : ((k == 4 || k == 7 || k == 8) ? 1 : 0);
            const int dx = (k == 1 || k == 5 || k == 7) ? -1 : ((k == 2 || k == 6 || k == 8) ? 1 : 0);
            const int kx = (k == 1 || k == 5) ? 2 : ((k == 2 || k == 6) ? 1 : 0);  // code over x (channel 64)
            const int ky = (k == 3 || k == 7) ? 2 : ((k == 4 || k == 8) ? 1 : 0);  // code over y (channel 65)
            const int ny = cy + dy, nx = cx + dx;
            float logit = -100.0f;
            if (ny >= vy0 && ny < vy1 && nx >= 0 && nx < w) {  // [vy0,vy1): cell rows inside the IMAGE (row bands pass halos)
                const float p0 = pos_code(kx, px, scale), p1 = pos_code(ky, py, scale);
                const float* alr = &sAlr[(ly + dy) * kK5HaloX + lx + dx][0];
                float h1[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) h1[o] = 0.f;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 a = *reinterpret_cast<const float4*>(alr + i4 * 4);
                    const float4 q0 = *reinterpret_cast<const float4*>(&sW0c[0][i4 * 4]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&sW0c[1][i4 * 4]);
                    float h0[4];
                    h0[0] = leaky(fmaf(q1.x, p1, fmaf(q0.x, p0, a.x + ahr[i4 * 4 + 0])));
                    h0[1] = leaky(fmaf(q1.y, p1, fmaf(q0.y, p0, a.y + ahr[i4 * 4 + 1])));
                    h0[2] = leaky(fmaf(q1.z, p1, fmaf(q0.z, p0, a.z + ahr[i4 * 4 + 2])));
                    h0[3] = leaky(fmaf(q1.w, p1, fmaf(q0.w, p0, a.w + ahr[i4 * 4 + 3])));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int o4 = 0; o4 < 4; ++o4) {
                            const float4 wv = *reinterpret_cast<const float4*>(&sW1[i4 * 4 + j][o4 * 4]);
                            h1[o4 * 4 + 0] = fmaf(wv.x, h0[j], h1[o4 * 4 + 0]);
                            h1[o4 * 4 + 1] = fmaf(wv.y, h0[j], h1[o4 * 4 + 1]);
                            h1[o4 * 4 + 2] = fmaf(wv.z, h0[j], h1[o4 * 4 + 2]);
                            h1[o4 * 4 + 3] = fmaf(wv.w, h0[j], h1[o4 * 4 + 3]);
                        }
                    }
                }
                float h2[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) h2[o] = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v = leaky(h1[i]);
#pragma unroll
                    for (int o = 0; o < 8; ++o) h2[o] = fmaf(sW2[i][o], v, h2[o]);
                }
                logit = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) logit = fmaf(sW3[i], leaky(h2[i]), logit);
            }
            sLogit[k][tid] = logit;
        }
        float m = sLogit[0][tid];
#pragma unroll
        for (int k = 1; k < 9; ++k) m = fmaxf(m, sLogit[k][tid]);
        float e[9], s = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            e[k] = expf(sLogit[k][tid] - m);
            s += e[k];
        }
        float* po = out + (size_t)b * 9 * hplane + (size_t)y * W + x;
#pragma unroll
        for (int k = 0; k < 9; ++k) po[k * hplane] = e[k] / s;
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_ctxmap_weights_fwd(const float* lr, const float* hr, const float* w0, const float* w1,
                                          const float* w2, const float* w3, float* weights9, int B, int h, int w,
                                          int scale, int valid_y0, int valid_y1, void* stream) {
    CMF_REQUIRE(lr && hr && w0 && w1 && w2 && w3 && weights9, "ctxmap_weights_fwd: null pointer");
    CMF_REQUIRE(B > 0 && h > 0 && w > 0, "ctxmap_weights_fwd: non-positive dimension");
    CMF_REQUIRE(scale >= 2 && scale % 2 == 0, "ctxmap_weights_fwd: odd scale %d (the reference exit()s)", scale);
    CMF_REQUIRE(B <= 65535, "ctxmap_weights_fwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kK5CellsX), (unsigned)cdiv(h, kK5CellsY), (unsigned)B);
    CMF_REQUIRE(valid_y0 >= 0 && valid_y1 <= h && valid_y0 < valid_y1, "ctxmap_weights_fwd: bad valid row range [%d,%d) for h=%d", valid_y0, valid_y1, h);
    ctxmap_weights_kernel<<<grid, kK5Threads, 0, (cudaStream_t)stream>>>(lr, hr, w0, w1, w2, w3, weights9, h, w, scale,
                                                                          valid_y0, valid_y1);
    CMF_LAUNCH_CHECK("ctxmap_weights_kernel");
    return CMFB200_OK;
}
