// K2 (fp32 parity path) -- the classifier tail conv 3x3x3, Cin -> 1 (classifN.2, cmf/models/cmfsm.py:624,629,634).
// With a single output channel the generic register tile (channels x voxels) degenerates to 12 FMAs per 5 shared-memory
// loads (4.8 TMAC/s measured).  Here a thread owns 4 (d) x 4 (h) outputs of one column w: per input channel it walks
// the 6 x 6 (plane, row) pairs of its neighbourhood, loads the three columns w-1..w+1 of each (consecutive lanes ->
// consecutive banks) and feeds up to 27 FMAs from them, 432 FMAs per 108 conflict-free loads, all 27 weights in
// registers.  The halo'd input block of one channel ([10][10][40] floats) arrives as ONE TMA box with zero fill at
// the borders -- no per-thread staging arithmetic.  (The box starts at w0-4: the innermost TMA coordinate of an
// fp32 map has to stay 16-byte aligned -- w0-1 faults with "illegal instruction".)  4-stage mbarrier ring over the
// input channels; 3 CTAs of 128 threads per SM.  fp32 FMA accumulation order per output: ci -> kd -> kh -> kw.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {
constexpr int kQThreads = 128;
constexpr int kQTD = 8, kQTH = 8, kQTW = 32;
constexpr int kQPD = kQTD + 2, kQPH = kQTH + 2, kQPW = 40;      // box columns w0-4 .. w0+35 (w0-1 .. w0+32 are used)
constexpr int kQPatch = kQPD * kQPH * kQPW;                     // floats per channel
constexpr int kQStage = ((kQPatch * 4 + 127) & ~127) / 4;       // TMA destinations stay 128-byte aligned
constexpr int kQNS = 4;
}

__global__ void __launch_bounds__(kQThreads, 3)
    conv3d_cout1_fp32_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ wp,
                             float* __restrict__ y, int Cin, int D, int H, int W, int tiles_w, int tiles_h, int hoff) {
    extern __shared__ uint8_t q1_raw[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q1_raw) + 127) & ~uintptr_t(127));
    float* sIn = smem;                                   // [NS][kQStage]
    float* sWt = smem + kQNS * kQStage;                  // [Cin][28] : the 27 taps of a channel, padded
    uint64_t* full = reinterpret_cast<uint64_t*>(sWt + Cin * 28);

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * kQTW, h0 = tile_y * kQTH, d0 = blockIdx.y * kQTD, b = blockIdx.z;
    const int lane = tid & 31, hgrp = (tid >> 5) & 1, slab = tid >> 6;  // outputs: d0+4*slab+dd, h0+4*hgrp+hh, w0+lane

    if (tid == 0) {
        for (int s = 0; s < kQNS; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < Cin * 27; i += kQThreads) {  // wp: [Cin][27] (Cout = 1), tap = (kd*3+kh)*3+kw
        const int ci = i / 27;
        sWt[ci * 28 + (i - ci * 27)] = wp[i];
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < kQNS && s < Cin; ++s) {
            mbar_arrive_expect_tx(full + s, kQPatch * 4);
            tma_load_5d(sIn + s * kQStage, &tmap_x, full + s, w0 - 4, h0 - 1 + hoff, d0 - 1, b * Cin + s, 0);
        }
    }

    float acc[4][4];  // [dd][hh]
#pragma unroll
    for (int dd = 0; dd < 4; ++dd)
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) acc[dd][hh] = 0.f;

    for (int ci = 0; ci < Cin; ++ci) {
        const int s = ci % kQNS;
        float wk[28];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(sWt + ci * 28 + 4 * j);
            wk[4 * j + 0] = t.x; wk[4 * j + 1] = t.y; wk[4 * j + 2] = t.z; wk[4 * j + 3] = t.w;
        }
        mbar_wait(full + s, (ci / kQNS) & 1);
        const float* base = sIn + s * kQStage + ((slab * 4) * kQPH + hgrp * 4) * kQPW + lane + 3;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const float* q = base + (p * kQPH + r) * kQPW;
                const float v0 = q[0], v1 = q[1], v2 = q[2];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int dd = p - kd;
                    if (dd < 0 || dd > 3) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int hh = r - kh;
                        if (hh < 0 || hh > 3) continue;
                        const int t = (kd * 3 + kh) * 3;
                        acc[dd][hh] = fmaf(wk[t + 2], v2, fmaf(wk[t + 1], v1, fmaf(wk[t], v0, acc[dd][hh])));
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with stage s
        if (tid == 0 && ci + kQNS < Cin) {
            mbar_arrive_expect_tx(full + s, kQPatch * 4);
            tma_load_5d(sIn + s * kQStage, &tmap_x, full + s, w0 - 4, h0 - 1 + hoff, d0 - 1, b * Cin + ci + kQNS, 0);
        }
    }

    const int ow = w0 + lane;
    if (ow < W) {
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            const int od = d0 + slab * 4 + dd;
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
                const int oh = h0 + hgrp * 4 + hh;
                if (od < D && oh < H) y[(((size_t)b * D + od) * H + oh) * W + ow] = acc[dd][hh];
            }
        }
    }
}

// used by cmfb200_conv3d_k3_fwd / _rows_fwd for Cout == 1; returns -1 when the shape does not qualify (caller falls back).
// Row window (row bands): x has H_in rows, the H output rows read input rows m - 1 + hoff + kh (rows outside x are zero).
int conv3d_cout1_fp32_dispatch(const float* x, const float* wp, float* y, int B, int Cin, int D, int H_in, int W, int hoff,
                               int H, cudaStream_t st) {
    if ((W & 3) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 || Cin > 64 || B > 65535)
        return -1;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W, (cuuint64_t)H_in, (cuuint64_t)D, (cuuint64_t)B * Cin, 1};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 4, (cuuint64_t)H_in * W * 4, (cuuint64_t)D * H_in * W * 4,
                                (cuuint64_t)B * Cin * D * H_in * W * 4};
    const cuuint32_t box[5] = {kQPW, kQPH, kQPD, 1, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "conv3d_cout1_fp32", CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return rc;
    const size_t smem = (size_t)kQNS * kQStage * 4 + (size_t)Cin * 28 * 4 + kQNS * 8 + 128;
    CMF_CUDA(cudaFuncSetAttribute(conv3d_cout1_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_w = (int)cdiv(W, kQTW), tiles_h = (int)cdiv(H, kQTH);
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)cdiv(D, kQTD), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535, "conv3d_cout1_fp32: depth too large");
    conv3d_cout1_fp32_kernel<<<grid, kQThreads, smem, st>>>(tmap, wp, y, Cin, D, H, W, tiles_w, tiles_h, hoff);
    CMF_LAUNCH_CHECK("conv3d_cout1_fp32_kernel");
    return CMFB200_OK;
}

}  // namespace cmfb200
