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

// Neighbour sets.  VARIANT 0 = cmfsm (eight_related_context_mapping, cmfsm.py:443-593): nine neighbours in the order
// c,l,r,t,b,lt,rt,lb,rb, logit -100 where the neighbour cell is outside the image, output = softmax.
// VARIANT 1 = cmfsm_sub_8 (six_related_context_mapping, cmfsm_sub_8.py:440-572, reference-image half): five
// neighbours c,r,l,t,b, a trailing LeakyReLU on the MLP output, logit 0 outside the image, output = softmax * logit.
template <int VARIANT>
struct K5Set;
template <>
struct K5Set<0> {
    static constexpr int NK = 9;
    __device__ static int dy(int k) { return (k == 3 || k == 5 || k == 6) ? -1 : ((k == 4 || k == 7 || k == 8) ? 1 : 0); }
    __device__ static int dx(int k) { return (k == 1 || k == 5 || k == 7) ? -1 : ((k == 2 || k == 6 || k == 8) ? 1 : 0); }
    // code kinds per SURVEY.md A.3 (the diagonals reuse the axis encodings, cmfsm.py:459-462)
    __device__ static int kx(int k) { return (k == 1 || k == 5) ? 2 : ((k == 2 || k == 6) ? 1 : 0); }
    __device__ static int ky(int k) { return (k == 3 || k == 7) ? 2 : ((k == 4 || k == 8) ? 1 : 0); }
};
template <>
struct K5Set<1> {
    static constexpr int NK = 5;
    __device__ static int dy(int k) { return k == 3 ? -1 : (k == 4 ? 1 : 0); }
    __device__ static int dx(int k) { return k == 1 ? 1 : (k == 2 ? -1 : 0); }
    __device__ static int kx(int k) { return k == 1 ? 2 : (k == 2 ? 1 : 0); }  // right: dec, left: inc (sub8.py:497,516)
    __device__ static int ky(int k) { return k == 3 ? 2 : (k == 4 ? 1 : 0); }  // top: dec, bottom: inc (sub8.py:534,546)
};

template <>
struct K5Set<2> {  // target-image half of six_related_context_mapping: centre, right, left (cmfsm_sub_16.py:488-573)
    static constexpr int NK = 3;
    __device__ static int dy(int) { return 0; }
    __device__ static int dx(int k) { return k == 1 ? 1 : (k == 2 ? -1 : 0); }
    __device__ static int kx(int k) { return k == 1 ? 2 : (k == 2 ? 1 : 0); }
    __device__ static int ky(int) { return 0; }
};

template <int VARIANT>
__global__ void __launch_bounds__(kK5Threads) ctxmap_weights_kernel(
    const float* __restrict__ lr, const float* __restrict__ hr, const float* __restrict__ w0,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
    float* __restrict__ out, int h, int w, int scale, int vy0, int vy1) {
    using NS = K5Set<VARIANT>;
    constexpr int NK = NS::NK;
    __shared__ __align__(16) float sW0lr[32][32];  // [in][out]
    __shared__ __align__(16) float sW0hr[32][32];  // [in][out]
    __shared__ __align__(16) float sW0c[2][32];    // code channels 64,65
    __shared__ __align__(16) float sW1[32][16];    // [in][out]
    __shared__ __align__(16) float sW2[16][8];
    __shared__ __align__(16) float sW3[8];
    __shared__ __align__(16) float sLr[32][kK5Halo];
    __shared__ __align__(16) float sAlr[kK5Halo][kPad];
    __shared__ float sLogit[NK][kK5Threads];

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

        for (int k = 0; k < NK; ++k) {
            const int dy = NS::dy(k), dx = NS::dx(k);
            const int kx = NS::kx(k), ky = NS::ky(k);  // code over x (channel 64) / over y (channel 65)
            const int ny = cy + dy, nx = cx + dx;
            float logit = VARIANT == 0 ? -100.0f : 0.0f;
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
                if (VARIANT != 0) logit = leaky(logit);
            }
            sLogit[k][tid] = logit;
        }
        float m = sLogit[0][tid];
#pragma unroll
        for (int k = 1; k < NK; ++k) m = fmaxf(m, sLogit[k][tid]);
        float e[NK], s = 0.f;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            e[k] = expf(sLogit[k][tid] - m);
            s += e[k];
        }
        float* po = out + (size_t)b * NK * hplane + (size_t)y * W + x;
#pragma unroll
        for (int k = 0; k < NK; ++k) po[k * hplane] = VARIANT == 0 ? e[k] / s : (e[k] / s) * sLogit[k][tid];
    }
}


// ------------------------------------------------------------------------------------------------
// K5 backward (training): gradient of the nine softmax weights w.r.t. the MLP weights and both feature maps.
//
// Per pixel and valid neighbour the MLP is re-evaluated (same factorisation as the forward) and back-propagated
// in registers.  With  a0 = A_lr(cell_k) + A_hr(pixel) + W0[:,64:66].code  the kernel emits
//   dAhr [B,32,H,W]  = sum_k da0            (one thread owns one pixel: plain stores)
//   dAlr [B,32,h,w] += da0 at cell_k        (shared-memory accumulators per tile, then global atomics)
//   dW1, dW2, dW3, dW0[:,64:66]             (sums of outer products over all samples)
// and the host finishes the two LINEAR maps A_hr = W0[:,32:64].hr, A_lr = W0[:,0:32].lr (d hr, d lr, dW0) as 1x1
// GEMMs.  Outer products: the 32 samples of a warp stage (left, right) vectors in a per-warp shared tile and lane j
// accumulates column j over the 32 samples, so nothing is reduced with shuffles and no per-sample gradient ever
// goes to HBM.  softmax backward uses the saved forward output: dlogit_k = P_k (g_k - sum_j g_j P_j).
// Tile = 2 x 8 cells, scale 4; a warp = two cells x 16 pixels.
// ------------------------------------------------------------------------------------------------
constexpr int kB5TStride = 33;
constexpr int kB5WarpFloats = 32 * kB5TStride + 32 * 16;  // T[32][33] + U[32][16]
constexpr int kB5NumGrad = 512 + 128 + 8 + 64;            // dW1[16][32], dW2[8][16], dW3[8], dWcode[32][2]

struct K5BwdSmem {
    float W0lr[32][32];   // [in][out]
    float W0hr[32][32];   // [in][out]
    float W0c[2][32];
    float W1[32][16];     // [in][out]   forward
    float W1t[16][32];    // [out][in]   backward (dh0 = W1^T da1)
    float W2[16][8];      // [in][out]
    float W2t[8][16];     // [out][in]
    float W3[8];
    float Lr[32][kK5Halo];
    float Alr[kK5Halo][kPad];
    float dAlr[kK5Halo][32];
    float scratch[8][kB5WarpFloats];
};

__global__ void __launch_bounds__(kK5Threads, 1) ctxmap_weights_bwd_kernel(
    const float* __restrict__ lr, const float* __restrict__ hr, const float* __restrict__ w0,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
    const float* __restrict__ prob, const float* __restrict__ gout, float* __restrict__ dAhr,
    float* __restrict__ dAlr, float* __restrict__ dWbuf, int B, int h, int w, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) uint8_t k5b_raw[];
    K5BwdSmem& S = *reinterpret_cast<K5BwdSmem*>(k5b_raw);
    constexpr int scale = 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = h * scale, W = w * scale;
    const size_t lplane = (size_t)h * w, hplane = (size_t)H * W;

    for (int i = tid; i < 32 * 66; i += kK5Threads) {
        const int o = i / 66, c = i - o * 66;
        const float v = w0[i];
        if (c < 32) S.W0lr[c][o] = v;
        else if (c < 64) S.W0hr[c - 32][o] = v;
        else S.W0c[c - 64][o] = v;
    }
    for (int i = tid; i < 16 * 32; i += kK5Threads) {
        S.W1[i % 32][i / 32] = w1[i];
        S.W1t[i / 32][i % 32] = w1[i];
    }
    if (tid < 8 * 16) {
        S.W2[tid % 16][tid / 16] = w2[tid];
        S.W2t[tid / 16][tid % 16] = w2[tid];
    }
    if (tid < 8) S.W3[tid] = w3[tid];

    float* T = S.scratch[warp];           // [32][33]
    float* U = T + 32 * kB5TStride;       // [32][16]
    float acc1[16], acc2[8], acc3 = 0.f, accC[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) acc1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = 0.f;

    // thread -> pixel: warp = (cell row, cell pair), lane = (cell of the pair, 4x4 pixel)
    const int cyl = warp >> 2, cxl = (warp & 3) * 2 + (lane >> 4);
    const int py = (lane & 15) >> 2, px = lane & 3;
    const int ly = cyl + 1, lx = cxl + 1;

    const int tiles = tiles_x * tiles_y * B;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y);
        const int r = tile - b * tiles_x * tiles_y;
        const int cy0 = (r / tiles_x) * kK5CellsY, cx0 = (r % tiles_x) * kK5CellsX;
        __syncthreads();  // previous tile fully flushed (and the weights are staged on the first pass)
        for (int i = tid; i < 32 * kK5Halo; i += kK5Threads) {
            const int c = i / kK5Halo, t = i - c * kK5Halo;
            const int cy = cy0 + t / kK5HaloX - 1, cx = cx0 + t % kK5HaloX - 1;
            float v = 0.f;
            if (cy >= 0 && cy < h && cx >= 0 && cx < w) v = lr[((size_t)b * 32 + c) * lplane + (size_t)cy * w + cx];
            S.Lr[c][t] = v;
        }
        for (int i = tid; i < kK5Halo * 32; i += kK5Threads) S.dAlr[i >> 5][i & 31] = 0.f;
        __syncthreads();
        for (int i = tid; i < kK5Halo * 32; i += kK5Threads) {
            const int t = i >> 5, o = i & 31;
            float a = 0.f;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) a = fmaf(S.W0lr[c][o], S.Lr[c][t], a);
            S.Alr[t][o] = a;
        }
        __syncthreads();

        const int cy = cy0 + cyl, cx = cx0 + cxl;
        const bool in_img = (cy < h) && (cx < w);
        const int y = cy * scale + py, x = cx * scale + px;
        const size_t pix = in_img ? (size_t)y * W + x : 0;

        // hr half of layer 0 + softmax backward
        float ahr[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) ahr[o] = 0.f;
        float dot = 0.f;  // sum_j g_j P_j of the softmax backward
        {
            const float* ph = hr + (size_t)b * 32 * hplane + pix;
#pragma unroll 4
            for (int c = 0; c < 32; ++c) {
                const float v = in_img ? ph[c * hplane] : 0.f;
#pragma unroll
                for (int o4 = 0; o4 < 8; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&S.W0hr[c][o4 * 4]);
                    ahr[o4 * 4 + 0] = fmaf(wv.x, v, ahr[o4 * 4 + 0]);
                    ahr[o4 * 4 + 1] = fmaf(wv.y, v, ahr[o4 * 4 + 1]);
                    ahr[o4 * 4 + 2] = fmaf(wv.z, v, ahr[o4 * 4 + 2]);
                    ahr[o4 * 4 + 3] = fmaf(wv.w, v, ahr[o4 * 4 + 3]);
                }
            }
            if (in_img) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    dot = fmaf(prob[((size_t)b * 9 + k) * hplane + pix], gout[((size_t)b * 9 + k) * hplane + pix], dot);
            }
        }
        float dah[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) dah[o] = 0.f;

#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            const int dy = (k == 3 || k == 5 || k == 6) ? -1 : ((k == 4 || k == 7 || k == 8) ? 1 : 0);
            const int dx = (k == 1 || k == 5 || k == 7) ? -1 : ((k == 2 || k == 6 || k == 8) ? 1 : 0);
            const int kx = (k == 1 || k == 5) ? 2 : ((k == 2 || k == 6) ? 1 : 0);
            const int ky = (k == 3 || k == 7) ? 2 : ((k == 4 || k == 8) ? 1 : 0);
            const int ny = cy + dy, nx = cx + dx;
            const bool valid = in_img && ny >= 0 && ny < h && nx >= 0 && nx < w;
            float dlk = 0.f;  // invalid neighbours: the logit is the constant -100
            if (valid) {
                const size_t gi = ((size_t)b * 9 + k) * hplane + pix;
                dlk = prob[gi] * (gout[gi] - dot);
            }
            const float p0 = pos_code(kx, px, scale), p1 = pos_code(ky, py, scale);
            const int ncell = (ly + dy) * kK5HaloX + lx + dx;

            // ---- forward re-evaluation
            float h0[32], h1[16], h2[8];
            {
                const float* alr = &S.Alr[ncell][0];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 a = *reinterpret_cast<const float4*>(alr + i4 * 4);
                    const float4 q0 = *reinterpret_cast<const float4*>(&S.W0c[0][i4 * 4]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&S.W0c[1][i4 * 4]);
                    h0[i4 * 4 + 0] = leaky(fmaf(q1.x, p1, fmaf(q0.x, p0, a.x + ahr[i4 * 4 + 0])));
                    h0[i4 * 4 + 1] = leaky(fmaf(q1.y, p1, fmaf(q0.y, p0, a.y + ahr[i4 * 4 + 1])));
                    h0[i4 * 4 + 2] = leaky(fmaf(q1.z, p1, fmaf(q0.z, p0, a.z + ahr[i4 * 4 + 2])));
                    h0[i4 * 4 + 3] = leaky(fmaf(q1.w, p1, fmaf(q0.w, p0, a.w + ahr[i4 * 4 + 3])));
                }
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) h1[o] = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
#pragma unroll
                for (int o4 = 0; o4 < 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&S.W1[i][o4 * 4]);
                    h1[o4 * 4 + 0] = fmaf(wv.x, h0[i], h1[o4 * 4 + 0]);
                    h1[o4 * 4 + 1] = fmaf(wv.y, h0[i], h1[o4 * 4 + 1]);
                    h1[o4 * 4 + 2] = fmaf(wv.z, h0[i], h1[o4 * 4 + 2]);
                    h1[o4 * 4 + 3] = fmaf(wv.w, h0[i], h1[o4 * 4 + 3]);
                }
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) h1[o] = leaky(h1[o]);
#pragma unroll
            for (int o = 0; o < 8; ++o) h2[o] = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 wa = *reinterpret_cast<const float4*>(&S.W2[i][0]);
                const float4 wb = *reinterpret_cast<const float4*>(&S.W2[i][4]);
                h2[0] = fmaf(wa.x, h1[i], h2[0]); h2[1] = fmaf(wa.y, h1[i], h2[1]);
                h2[2] = fmaf(wa.z, h1[i], h2[2]); h2[3] = fmaf(wa.w, h1[i], h2[3]);
                h2[4] = fmaf(wb.x, h1[i], h2[4]); h2[5] = fmaf(wb.y, h1[i], h2[5]);
                h2[6] = fmaf(wb.z, h1[i], h2[6]); h2[7] = fmaf(wb.w, h1[i], h2[7]);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) h2[o] = leaky(h2[o]);

            // ---- dW3 += dl * h2          (lane: column lane%8, samples of quarter lane/8)
#pragma unroll
            for (int o = 0; o < 8; ++o) T[lane * kB5TStride + o] = h2[o];
            U[lane * 16] = dlk;
            __syncwarp();
            {
                const int j = lane & 7, q = lane >> 3;
#pragma unroll
                for (int s_ = 0; s_ < 8; ++s_) acc3 = fmaf(U[(q * 8 + s_) * 16], T[(q * 8 + s_) * kB5TStride + j], acc3);
            }
            __syncwarp();
            // ---- layer 2 backward: da2 = W3 * dl * leaky'(a2);  dW2 += da2 (x) h1
            float da2[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) da2[o] = S.W3[o] * dlk * (h2[o] > 0.f ? 1.f : 0.01f);
#pragma unroll
            for (int o = 0; o < 16; ++o) T[lane * kB5TStride + o] = h1[o];
#pragma unroll
            for (int o = 0; o < 8; ++o) U[lane * 16 + o] = da2[o];
            __syncwarp();
            {
                const int j = lane & 15, hf = lane >> 4;
#pragma unroll 4
                for (int s_ = 0; s_ < 16; ++s_) {
                    const int sm = hf * 16 + s_;
                    const float hv = T[sm * kB5TStride + j];
                    const float4 ua = *reinterpret_cast<const float4*>(&U[sm * 16]);
                    const float4 ub = *reinterpret_cast<const float4*>(&U[sm * 16 + 4]);
                    acc2[0] = fmaf(ua.x, hv, acc2[0]); acc2[1] = fmaf(ua.y, hv, acc2[1]);
                    acc2[2] = fmaf(ua.z, hv, acc2[2]); acc2[3] = fmaf(ua.w, hv, acc2[3]);
                    acc2[4] = fmaf(ub.x, hv, acc2[4]); acc2[5] = fmaf(ub.y, hv, acc2[5]);
                    acc2[6] = fmaf(ub.z, hv, acc2[6]); acc2[7] = fmaf(ub.w, hv, acc2[7]);
                }
            }
            __syncwarp();
            // ---- layer 1 backward: da1 = (W2^T da2) * leaky'(a1);  dW1 += da1 (x) h0
            float da1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) da1[i] = 0.f;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 wv = *reinterpret_cast<const float4*>(&S.W2t[o][i4 * 4]);
                    da1[i4 * 4 + 0] = fmaf(wv.x, da2[o], da1[i4 * 4 + 0]);
                    da1[i4 * 4 + 1] = fmaf(wv.y, da2[o], da1[i4 * 4 + 1]);
                    da1[i4 * 4 + 2] = fmaf(wv.z, da2[o], da1[i4 * 4 + 2]);
                    da1[i4 * 4 + 3] = fmaf(wv.w, da2[o], da1[i4 * 4 + 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) da1[i] *= (h1[i] > 0.f ? 1.f : 0.01f);
#pragma unroll
            for (int o = 0; o < 32; ++o) T[lane * kB5TStride + o] = h0[o];
#pragma unroll
            for (int o = 0; o < 16; ++o) U[lane * 16 + o] = da1[o];
            __syncwarp();
#pragma unroll 2
            for (int s_ = 0; s_ < 32; ++s_) {
                const float hv = T[s_ * kB5TStride + lane];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 u = *reinterpret_cast<const float4*>(&U[s_ * 16 + i4 * 4]);
                    acc1[i4 * 4 + 0] = fmaf(u.x, hv, acc1[i4 * 4 + 0]);
                    acc1[i4 * 4 + 1] = fmaf(u.y, hv, acc1[i4 * 4 + 1]);
                    acc1[i4 * 4 + 2] = fmaf(u.z, hv, acc1[i4 * 4 + 2]);
                    acc1[i4 * 4 + 3] = fmaf(u.w, hv, acc1[i4 * 4 + 3]);
                }
            }
            __syncwarp();
            // ---- layer 0 backward: da0 = (W1^T da1) * leaky'(a0)   (h0 is overwritten by da0)
            {
                float d0[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) d0[i] = 0.f;
#pragma unroll
                for (int o = 0; o < 16; ++o) {
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 wv = *reinterpret_cast<const float4*>(&S.W1t[o][i4 * 4]);
                        d0[i4 * 4 + 0] = fmaf(wv.x, da1[o], d0[i4 * 4 + 0]);
                        d0[i4 * 4 + 1] = fmaf(wv.y, da1[o], d0[i4 * 4 + 1]);
                        d0[i4 * 4 + 2] = fmaf(wv.z, da1[o], d0[i4 * 4 + 2]);
                        d0[i4 * 4 + 3] = fmaf(wv.w, da1[o], d0[i4 * 4 + 3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = d0[i] * (h0[i] > 0.f ? 1.f : 0.01f);
                    dah[i] += v;
                    T[lane * kB5TStride + i] = v;
                }
            }
            U[lane * 16 + 0] = p0;
            U[lane * 16 + 1] = p1;
            __syncwarp();
            {
                // lane j: dWcode[j][0..1] += da0[s][j] * code[s];  cell sums of da0 for dAlr (lanes 0-15 / 16-31)
                float sa = 0.f, sb = 0.f;
#pragma unroll 4
                for (int s_ = 0; s_ < 16; ++s_) {
                    const float v = T[s_ * kB5TStride + lane];
                    sa += v;
                    accC[0] = fmaf(v, U[s_ * 16 + 0], accC[0]);
                    accC[1] = fmaf(v, U[s_ * 16 + 1], accC[1]);
                }
#pragma unroll 4
                for (int s_ = 16; s_ < 32; ++s_) {
                    const float v = T[s_ * kB5TStride + lane];
                    sb += v;
                    accC[0] = fmaf(v, U[s_ * 16 + 0], accC[0]);
                    accC[1] = fmaf(v, U[s_ * 16 + 1], accC[1]);
                }
                const int cell_a = (cyl + 1 + dy) * kK5HaloX + (warp & 3) * 2 + 1 + dx;  // neighbour of the pair's first cell
                atomicAdd(&S.dAlr[cell_a][lane], sa);
                atomicAdd(&S.dAlr[cell_a + 1][lane], sb);
            }
            __syncwarp();
        }

        if (in_img) {
            float* pd = dAhr + (size_t)b * 32 * hplane + pix;
#pragma unroll
            for (int o = 0; o < 32; ++o) pd[o * hplane] = dah[o];
        }
        __syncthreads();
        for (int i = tid; i < kK5Halo * 32; i += kK5Threads) {
            const int t = i >> 5, o = i & 31;
            const int gy = cy0 + t / kK5HaloX - 1, gx = cx0 + t % kK5HaloX - 1;
            const float v = S.dAlr[t][o];
            if (gy >= 0 && gy < h && gx >= 0 && gx < w && v != 0.f)
                atomicAdd(dAlr + ((size_t)b * 32 + o) * lplane + (size_t)gy * w + gx, v);
        }
    }

    // ---- weight-gradient partials of this CTA -> global
    __syncthreads();
    float* red = reinterpret_cast<float*>(&S.scratch[0][0]);
    for (int i = tid; i < kB5NumGrad; i += kK5Threads) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(&red[i * 32 + lane], acc1[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&red[512 + i * 16 + (lane & 15)], acc2[i]);
    atomicAdd(&red[640 + (lane & 7)], acc3);
    atomicAdd(&red[648 + lane * 2 + 0], accC[0]);
    atomicAdd(&red[648 + lane * 2 + 1], accC[1]);
    __syncthreads();
    for (int i = tid; i < kB5NumGrad; i += kK5Threads) atomicAdd(dWbuf + i, red[i]);
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
    ctxmap_weights_kernel<0><<<grid, kK5Threads, 0, (cudaStream_t)stream>>>(lr, hr, w0, w1, w2, w3, weights9, h, w, scale,
                                                                             valid_y0, valid_y1);
    CMF_LAUNCH_CHECK("ctxmap_weights_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_ctxmap_weights5_fwd(const float* lr, const float* hr, const float* w0, const float* w1,
                                           const float* w2, const float* w3, float* weights5, int B, int h, int w,
                                           int scale, void* stream) {
    CMF_REQUIRE(lr && hr && w0 && w1 && w2 && w3 && weights5, "ctxmap_weights5_fwd: null pointer");
    CMF_REQUIRE(B > 0 && h > 0 && w > 0, "ctxmap_weights5_fwd: non-positive dimension");
    CMF_REQUIRE(scale >= 2 && scale % 2 == 0, "ctxmap_weights5_fwd: odd scale %d (the reference exit()s)", scale);
    CMF_REQUIRE(B <= 65535, "ctxmap_weights5_fwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kK5CellsX), (unsigned)cdiv(h, kK5CellsY), (unsigned)B);
    ctxmap_weights_kernel<1><<<grid, kK5Threads, 0, (cudaStream_t)stream>>>(lr, hr, w0, w1, w2, w3, weights5, h, w, scale,
                                                                             0, h);
    CMF_LAUNCH_CHECK("ctxmap_weights_kernel<1>");
    return CMFB200_OK;
}

extern "C" int cmfb200_ctxmap_weights_bwd(const float* lr, const float* hr, const float* w0, const float* w1,
                                          const float* w2, const float* w3, const float* weights9,
                                          const float* grad_weights9, float* d_ahr, float* d_alr, float* d_wbuf, int B,
                                          int h, int w, int scale, void* stream) {
    CMF_REQUIRE(lr && hr && w0 && w1 && w2 && w3 && weights9 && grad_weights9 && d_ahr && d_alr && d_wbuf,
                "ctxmap_weights_bwd: null pointer");
    CMF_REQUIRE(B > 0 && h > 0 && w > 0, "ctxmap_weights_bwd: non-positive dimension");
    CMF_REQUIRE(scale == 4, "ctxmap_weights_bwd: only scale 4 (cmfsm) is implemented, got %d", scale);
    cudaStream_t st = (cudaStream_t)stream;
    CMF_CUDA(cudaMemsetAsync(d_alr, 0, (size_t)B * 32 * h * w * sizeof(float), st));
    CMF_CUDA(cudaMemsetAsync(d_wbuf, 0, kB5NumGrad * sizeof(float), st));
    CMF_CUDA(cudaFuncSetAttribute(ctxmap_weights_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(K5BwdSmem)));
    const int tiles_x = (int)cdiv(w, kK5CellsX), tiles_y = (int)cdiv(h, kK5CellsY);
    const long long tiles = (long long)tiles_x * tiles_y * B;
    CMF_REQUIRE(tiles < (1ll << 31), "ctxmap_weights_bwd: too many tiles");
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(tiles < 2ll * sms ? tiles : 2ll * sms);
    ctxmap_weights_bwd_kernel<<<grid, kK5Threads, sizeof(K5BwdSmem), st>>>(lr, hr, w0, w1, w2, w3, weights9,
                                                                            grad_weights9, d_ahr, d_alr, d_wbuf, B, h, w,
                                                                            tiles_x, tiles_y);
    CMF_LAUNCH_CHECK("ctxmap_weights_bwd_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_ctxmap_weights3_fwd(const float* lr, const float* hr, const float* w0, const float* w1,
                                           const float* w2, const float* w3, float* weights3, int B, int h, int w,
                                           int scale, void* stream) {
    CMF_REQUIRE(lr && hr && w0 && w1 && w2 && w3 && weights3, "ctxmap_weights3_fwd: null pointer");
    CMF_REQUIRE(B > 0 && h > 0 && w > 0, "ctxmap_weights3_fwd: non-positive dimension");
    CMF_REQUIRE(scale >= 2 && scale % 2 == 0, "ctxmap_weights3_fwd: odd scale %d (the reference exit()s)", scale);
    CMF_REQUIRE(B <= 65535, "ctxmap_weights3_fwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kK5CellsX), (unsigned)cdiv(h, kK5CellsY), (unsigned)B);
    ctxmap_weights_kernel<2><<<grid, kK5Threads, 0, (cudaStream_t)stream>>>(lr, hr, w0, w1, w2, w3, weights3, h, w, scale,
                                                                             0, h);
    CMF_LAUNCH_CHECK("ctxmap_weights_kernel<2>");
    return CMFB200_OK;
}
