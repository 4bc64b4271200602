// K4 -- fused epilogue: softmax over disparity + soft-argmin regression for the three cumulative cost
// volumes, x`scale` nearest upsampling and the learned 9-neighbour context mapping, one launch for all
// three outputs.  Replaces cmf/models/cmfsm.py:703-769 (+ disparityregression :111-123); closed form in
// SURVEY.md appendix A.4.
//
// HBM-bound: reads 3*B*D*h*w*4 (classifier volumes) + 9*B*H*W*4 (weights), writes 3*B*H*W*4.
// One CTA owns a TCY x TCX block of low-res cells.  Phase 1: one thread per cell of the block plus a
// 1-cell halo runs three online softmax/regressions down the disparity axis (cost2 = c2 + cost1,
// cost3 = c3 + cost2 formed in registers) and parks p_1..3 in shared memory.  Phase 2: each thread
// produces 4 horizontally adjacent output pixels (one cell wide) per output with 128-bit loads of the
// nine weight planes, accumulating in the reference's neighbour order without FMA contraction.
#include "common.cuh"

namespace cmfb200 {

constexpr int kK4Threads = 256;
constexpr int kTCY = 8, kTCX = 16;
constexpr int kHY = kTCY + 2, kHX = kTCX + 2;

struct Online {
    float m, s, t;
    __device__ __forceinline__ void init() {
        m = -INFINITY;
        s = 0.f;
        t = 0.f;
    }
    __device__ __forceinline__ void push(float v, float d) {
        if (v > m) {
            const float r = expf(m - v);  // expf(-inf) = 0 on the first element
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

// VARIANT 0 = cmfsm: cumulative costs, nine weights in the order c,l,r,t,b,lt,rt,lb,rb (cmfsm.py:703-769).
// VARIANT 1 = cmfsm_sub_8: three independent regressions (no cumulative sums), five weights in the order c,r,l,t,b
// (cmfsm_sub_8.py:757-802).
template <int VARIANT>
__global__ void __launch_bounds__(kK4Threads) softargmin_ctxmap_kernel(
    const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
    const float* __restrict__ wts, float* __restrict__ out1, float* __restrict__ out2, float* __restrict__ out3,
    float* __restrict__ pred_lr, int B, int D, int h, int w, int scale) {
    __shared__ float sp[3][kHY][kHX];
    const int b = blockIdx.z;
    const int cy0 = blockIdx.y * kTCY, cx0 = blockIdx.x * kTCX;
    const size_t plane = (size_t)h * w;

    // ---- phase 1: soft-argmin of the block + halo
    for (int t = threadIdx.x; t < kHY * kHX; t += kK4Threads) {
        const int hy = t / kHX, hx = t - hy * kHX;
        const int cy = cy0 + hy - 1, cx = cx0 + hx - 1;
        float p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (cy >= 0 && cy < h && cx >= 0 && cx < w) {
            const size_t off = (size_t)b * D * plane + (size_t)cy * w + cx;
            Online o1, o2, o3;
            o1.init();
            o2.init();
            o3.init();
            for (int d = 0; d < D; ++d) {
                const float v1 = c1[off + d * plane];
                const float v2 = VARIANT == 0 ? __fadd_rn(c2[off + d * plane], v1) : c2[off + d * plane];
                const float v3 = VARIANT == 0 ? __fadd_rn(c3[off + d * plane], v2) : c3[off + d * plane];
                const float fd = (float)d;
                o1.push(v1, fd);
                o2.push(v2, fd);
                o3.push(v3, fd);
            }
            p1 = o1.result();
            p2 = o2.result();
            p3 = o3.result();
            if (pred_lr && hy >= 1 && hy <= kTCY && hx >= 1 && hx <= kTCX) {
                const size_t o = (size_t)b * plane + (size_t)cy * w + cx;
                pred_lr[o] = p1;
                pred_lr[(size_t)B * plane + o] = p2;
                pred_lr[2 * (size_t)B * plane + o] = p3;
            }
        }
        sp[0][hy][hx] = p1;
        sp[1][hy][hx] = p2;
        sp[2][hy][hx] = p3;
    }
    __syncthreads();

    // ---- phase 2: mapped upsampling.  Neighbour order of the reference (c,l,r,t,b,lt,rt,lb,rb).
    const int H = h * scale, W = w * scale;
    const size_t oplane = (size_t)H * W;
    const int tile_h = kTCY * scale, tile_w4 = (kTCX * scale) >> 2;
    const float fs = (float)scale;
    constexpr int NK = VARIANT == 0 ? 9 : 5;
    const int dys[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};
    const int dxs[9] = {0, VARIANT == 0 ? -1 : 1, VARIANT == 0 ? 1 : -1, 0, 0, -1, 1, -1, 1};
    for (int i = threadIdx.x; i < tile_h * tile_w4; i += kK4Threads) {
        const int ty = i / tile_w4, tx = (i - ty * tile_w4) << 2;
        const int y = cy0 * scale + ty, x = cx0 * scale + tx;
        if (y >= H || x >= W) continue;
        const int ly = ty / scale + 1, lx = tx / scale + 1;  // halo coordinates of the centre cell
        const int cy = cy0 + ly - 1, cx = cx0 + lx - 1;
        const float* wp = wts + (size_t)b * NK * oplane + (size_t)y * W + x;
        float4 acc[3];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int ny = cy + dys[k], nx = cx + dxs[k];
            if (ny < 0 || ny >= h || nx < 0 || nx >= w) continue;  // the reference adds nothing there
            const float4 wk = ld_streaming_f4(wp + k * oplane);
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const float up = __fmul_rn(fs, sp[o][ly + dys[k]][lx + dxs[k]]);
                if (k == 0) {
                    acc[o] = make_float4(__fmul_rn(up, wk.x), __fmul_rn(up, wk.y), __fmul_rn(up, wk.z),
                                         __fmul_rn(up, wk.w));
                } else {
                    acc[o].x = __fadd_rn(acc[o].x, __fmul_rn(up, wk.x));
                    acc[o].y = __fadd_rn(acc[o].y, __fmul_rn(up, wk.y));
                    acc[o].z = __fadd_rn(acc[o].z, __fmul_rn(up, wk.z));
                    acc[o].w = __fadd_rn(acc[o].w, __fmul_rn(up, wk.w));
                }
            }
        }
        const size_t oo = (size_t)b * oplane + (size_t)y * W + x;
        st_streaming_f4(out1 + oo, acc[0]);
        st_streaming_f4(out2 + oo, acc[1]);
        st_streaming_f4(out3 + oo, acc[2]);
    }
}


// ------------------------------------------------------------------------------------------------
// K4 backward (training, cmfsm): gradients of the three mapped outputs w.r.t. the raw classifier volumes and the nine
// weights.  One launch, no atomics (gather formulation), same tiling as the forward.
//   phase 1  p_n, running max m_n and normaliser s_n of the three cumulative volumes for the block + 1-cell halo
//   phase 2  per output pixel: dw_k = scale * sum_n g_n p_n[cell+off_k]   (0 where the neighbour cell is outside)
//   phase 3  per interior cell c and neighbour k: sum over the pixels of cell c-off_k of g_n w_k  -> dp_n[c]
//   phase 4  per interior cell: dcost_n[d] = softmax_n[d] (d - p_n) dp_n;  dc3 = dcost3, dc2 = dcost2 + dcost3,
//            dc1 = dcost1 + dcost2 + dcost3 (the cumulative sums of the forward)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kK4Threads) softargmin_ctxmap_bwd_kernel(
    const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
    const float* __restrict__ wts, const float* __restrict__ g1, const float* __restrict__ g2,
    const float* __restrict__ g3, float* __restrict__ dc1, float* __restrict__ dc2, float* __restrict__ dc3,
    float* __restrict__ dw, int B, int D, int h, int w, int scale) {
    __shared__ float sp[3][kHY][kHX];
    __shared__ float sm[3][kTCY][kTCX], ss[3][kTCY][kTCX];
    __shared__ float sdp[3][kTCY * kTCX][9];
    const int b = blockIdx.z;
    const int cy0 = blockIdx.y * kTCY, cx0 = blockIdx.x * kTCX;
    const size_t plane = (size_t)h * w;
    const int H = h * scale, W = w * scale;
    const size_t oplane = (size_t)H * W;
    const float fs = (float)scale;
    const int dys[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};
    const int dxs[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};

    // ---- phase 1
    for (int t = threadIdx.x; t < kHY * kHX; t += kK4Threads) {
        const int hy = t / kHX, hx = t - hy * kHX;
        const int cy = cy0 + hy - 1, cx = cx0 + hx - 1;
        float p[3] = {0.f, 0.f, 0.f};
        if (cy >= 0 && cy < h && cx >= 0 && cx < w) {
            const size_t off = (size_t)b * D * plane + (size_t)cy * w + cx;
            Online o[3];
            o[0].init(); o[1].init(); o[2].init();
            for (int d = 0; d < D; ++d) {
                const float v1 = c1[off + d * plane];
                const float v2 = __fadd_rn(c2[off + d * plane], v1);
                const float v3 = __fadd_rn(c3[off + d * plane], v2);
                const float fd = (float)d;
                o[0].push(v1, fd); o[1].push(v2, fd); o[2].push(v3, fd);
            }
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                p[n] = o[n].result();
                if (hy >= 1 && hy <= kTCY && hx >= 1 && hx <= kTCX) {
                    sm[n][hy - 1][hx - 1] = o[n].m;
                    ss[n][hy - 1][hx - 1] = o[n].s;
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) sp[n][hy][hx] = p[n];
    }
    __syncthreads();

    // ---- phase 2: dweights of the tile's pixels (4 adjacent pixels = one cell's row segment per thread)
    const int tile_h = kTCY * scale, tile_w4 = (kTCX * scale) >> 2;
    for (int i = threadIdx.x; i < tile_h * tile_w4; i += kK4Threads) {
        const int ty = i / tile_w4, tx = (i - ty * tile_w4) << 2;
        const int y = cy0 * scale + ty, x = cx0 * scale + tx;
        if (y >= H || x >= W) continue;
        const int ly = ty / scale + 1, lx = tx / scale + 1;
        const int cy = cy0 + ly - 1, cx = cx0 + lx - 1;
        const size_t oo = (size_t)b * oplane + (size_t)y * W + x;
        const float4 ga = *reinterpret_cast<const float4*>(g1 + oo);
        const float4 gb = *reinterpret_cast<const float4*>(g2 + oo);
        const float4 gc = *reinterpret_cast<const float4*>(g3 + oo);
        float* pd = dw + (size_t)b * 9 * oplane + (size_t)y * W + x;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int ny = cy + dys[k], nx = cx + dxs[k];
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ny >= 0 && ny < h && nx >= 0 && nx < w) {
                const float u1 = fs * sp[0][ly + dys[k]][lx + dxs[k]], u2 = fs * sp[1][ly + dys[k]][lx + dxs[k]];
                const float u3 = fs * sp[2][ly + dys[k]][lx + dxs[k]];
                r.x = fmaf(ga.x, u1, fmaf(gb.x, u2, gc.x * u3));
                r.y = fmaf(ga.y, u1, fmaf(gb.y, u2, gc.y * u3));
                r.z = fmaf(ga.z, u1, fmaf(gb.z, u2, gc.z * u3));
                r.w = fmaf(ga.w, u1, fmaf(gb.w, u2, gc.w * u3));
            }
            *reinterpret_cast<float4*>(pd + k * oplane) = r;
        }
    }

    // ---- phase 3: dp_n[c] = scale * sum_k sum_{pixels of cell c-off_k} g_n w_k   (thread per (cell, k) pair)
    for (int i = threadIdx.x; i < kTCY * kTCX * 9; i += kK4Threads) {
        const int k = i % 9, cell = i / 9;
        const int ly = cell / kTCX, lx = cell - ly * kTCX;
        const int cy = cy0 + ly, cx = cx0 + lx;          // the cell that receives the gradient
        const int sy = cy - dys[k], sx = cx - dxs[k];    // the cell whose pixels use (cy,cx) as neighbour k
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (cy < h && cx < w && sy >= 0 && sy < h && sx >= 0 && sx < w) {
            const size_t base = (size_t)b * oplane + (size_t)(sy * scale) * W + (size_t)sx * scale;
            const float* pw = wts + ((size_t)b * 9 + k) * oplane + (size_t)(sy * scale) * W + (size_t)sx * scale;
            for (int py = 0; py < scale; ++py)
                for (int px = 0; px < scale; px += 4) {
                    const size_t o = (size_t)py * W + px;
                    const float4 wv = *reinterpret_cast<const float4*>(pw + o);
                    const float4 ga = *reinterpret_cast<const float4*>(g1 + base + o);
                    const float4 gb = *reinterpret_cast<const float4*>(g2 + base + o);
                    const float4 gc = *reinterpret_cast<const float4*>(g3 + base + o);
                    a1 += (ga.x * wv.x + ga.y * wv.y) + (ga.z * wv.z + ga.w * wv.w);
                    a2 += (gb.x * wv.x + gb.y * wv.y) + (gb.z * wv.z + gb.w * wv.w);
                    a3 += (gc.x * wv.x + gc.y * wv.y) + (gc.z * wv.z + gc.w * wv.w);
                }
        }
        sdp[0][cell][k] = a1;
        sdp[1][cell][k] = a2;
        sdp[2][cell][k] = a3;
    }
    __syncthreads();

    // ---- phase 4: softmax backward down the disparity axis (thread per interior cell)
    for (int cell = threadIdx.x; cell < kTCY * kTCX; cell += kK4Threads) {
        const int ly = cell / kTCX, lx = cell - ly * kTCX;
        const int cy = cy0 + ly, cx = cx0 + lx;
        if (cy >= h || cx >= w) continue;
        float dp[3], pn[3], mn[3], rs[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) a += sdp[n][cell][k];
            dp[n] = a * fs;
            pn[n] = sp[n][ly + 1][lx + 1];
            mn[n] = sm[n][ly][lx];
            rs[n] = 1.f / ss[n][ly][lx];
        }
        const size_t off = (size_t)b * D * plane + (size_t)cy * w + cx;
        for (int d = 0; d < D; ++d) {
            const float v1 = c1[off + d * plane];
            const float v2 = __fadd_rn(c2[off + d * plane], v1);
            const float v3 = __fadd_rn(c3[off + d * plane], v2);
            const float fd = (float)d;
            const float t1 = expf(v1 - mn[0]) * rs[0] * (fd - pn[0]) * dp[0];
            const float t2 = expf(v2 - mn[1]) * rs[1] * (fd - pn[1]) * dp[1];
            const float t3 = expf(v3 - mn[2]) * rs[2] * (fd - pn[2]) * dp[2];
            dc3[off + d * plane] = t3;
            dc2[off + d * plane] = t2 + t3;
            dc1[off + d * plane] = t1 + (t2 + t3);
        }
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_softargmin_ctxmap_fwd(const float* c1, const float* c2, const float* c3, const float* weights9,
                                             float* out1, float* out2, float* out3, float* pred_lr, int B, int D, int h,
                                             int w, int scale, void* stream) {
    CMF_REQUIRE(c1 && c2 && c3 && weights9 && out1 && out2 && out3, "softargmin_ctxmap_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, "softargmin_ctxmap_fwd: non-positive dimension");
    CMF_REQUIRE(scale >= 4 && scale % 4 == 0, "softargmin_ctxmap_fwd: scale=%d must be a positive multiple of 4", scale);
    CMF_REQUIRE(B <= 65535, "softargmin_ctxmap_fwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kTCX), (unsigned)cdiv(h, kTCY), (unsigned)B);
    softargmin_ctxmap_kernel<0><<<grid, kK4Threads, 0, (cudaStream_t)stream>>>(c1, c2, c3, weights9, out1, out2, out3,
                                                                                pred_lr, B, D, h, w, scale);
    CMF_LAUNCH_CHECK("softargmin_ctxmap_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_softargmin_ctxmap5_fwd(const float* c1, const float* c2, const float* c3, const float* weights5,
                                              float* out1, float* out2, float* out3, float* pred_lr, int B, int D, int h,
                                              int w, int scale, void* stream) {
    CMF_REQUIRE(c1 && c2 && c3 && weights5 && out1 && out2 && out3, "softargmin_ctxmap5_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, "softargmin_ctxmap5_fwd: non-positive dimension");
    CMF_REQUIRE(scale >= 4 && scale % 4 == 0, "softargmin_ctxmap5_fwd: scale=%d must be a positive multiple of 4", scale);
    CMF_REQUIRE(B <= 65535, "softargmin_ctxmap5_fwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kTCX), (unsigned)cdiv(h, kTCY), (unsigned)B);
    softargmin_ctxmap_kernel<1><<<grid, kK4Threads, 0, (cudaStream_t)stream>>>(c1, c2, c3, weights5, out1, out2, out3,
                                                                                pred_lr, B, D, h, w, scale);
    CMF_LAUNCH_CHECK("softargmin_ctxmap_kernel<1>");
    return CMFB200_OK;
}

extern "C" int cmfb200_softargmin_ctxmap_bwd(const float* c1, const float* c2, const float* c3, const float* weights9,
                                             const float* g1, const float* g2, const float* g3, float* dc1, float* dc2,
                                             float* dc3, float* dweights9, int B, int D, int h, int w, int scale,
                                             void* stream) {
    CMF_REQUIRE(c1 && c2 && c3 && weights9 && g1 && g2 && g3 && dc1 && dc2 && dc3 && dweights9,
                "softargmin_ctxmap_bwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, "softargmin_ctxmap_bwd: non-positive dimension");
    CMF_REQUIRE(scale >= 4 && scale % 4 == 0, "softargmin_ctxmap_bwd: scale=%d must be a positive multiple of 4", scale);
    CMF_REQUIRE(B <= 65535, "softargmin_ctxmap_bwd: B exceeds grid limit");
    dim3 grid((unsigned)cdiv(w, kTCX), (unsigned)cdiv(h, kTCY), (unsigned)B);
    softargmin_ctxmap_bwd_kernel<<<grid, kK4Threads, 0, (cudaStream_t)stream>>>(c1, c2, c3, weights9, g1, g2, g3, dc1, dc2,
                                                                                dc3, dweights9, B, D, h, w, scale);
    CMF_LAUNCH_CHECK("softargmin_ctxmap_bwd_kernel");
    return CMFB200_OK;
}
