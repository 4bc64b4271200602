// Auxiliary kernels and entry points of the bf16 / C8 pipeline of the 3-D aggregation (BASELINE config 4):
// weight packing, K1 in C8, GroupNorm apply on C8 (+ parity-split copy), layout converters, and the
// `cmfb200_conv3d_igemm_bf16_fwd` entry (the tcgen05 kernel itself is conv3d_igemm_kdstack.cu).
//
// Activation layout "C8": bf16 [B][C/8][D][H][W][8] -- a voxel's 8-channel group is one 16-byte unit and the
// voxels of a channel group are dense.  That makes every im2col row of every tap a 16-byte unit at a constant
// pitch, which is exactly the no-swizzle K-major UMMA canonical layout ((8,m),(8,2)):((16B,SBO),(1,LBO)); a
// rank-5 TMA box load brings a halo'd activation block into shared memory with the conv padding zero-filled.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

// ---- auxiliary kernels of the bf16 / C8 pipeline -----------------------------------------------------
// weights: conv [Cout][Cin][27] (or deconv [Cin][Cout][27]) fp32 -> bf16 [27][Cin/8][Cout][8]
__global__ void pack_igemm_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int Cout, int Cin,
                                         int transposed) {
    const int n = 27 * Cin * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = i & 7;
        const int co = (i >> 3) % Cout;
        const int chunk = (i / (8 * Cout)) % (Cin / 8);
        const int tap = i / (Cin * Cout);
        const int ci = chunk * 8 + j;
        const size_t src = transposed ? ((size_t)ci * Cout + co) * 27 + tap : ((size_t)co * Cin + ci) * 27 + tap;
        p[i] = __float2bfloat16_rn(w[src]);
    }
}

// K1 in C8/bf16: cost[b][chunk][d][y][x][8]; chunks 0..C/8-1 = left features masked by x>=d, the rest = right
// features shifted by d.  L,R fp32 NCHW.  Same scheme as the fp32 kernel (cost_volume.cu): a CTA stages the rows of
// one 8-channel group once in shared memory (converted to bf16, 16 B per voxel, right rows behind a zero prefix)
// and streams the D shifted / masked copies with 128-bit stores; for fixed (chunk, d) its rows are contiguous.
constexpr int kCv8Rows = 2;
__global__ void __launch_bounds__(256) cost_volume_c8_bf16_kernel(const float* __restrict__ L,
                                                                  const float* __restrict__ R,
                                                                  __nv_bfloat16* __restrict__ cost, int C, int h, int w,
                                                                  int D, int DP) {
    extern __shared__ uint4 sv[];  // [rows][DP + w]
    const int nc = C / 8;
    const int y0 = blockIdx.x * kCv8Rows, chunk = blockIdx.y, b = blockIdx.z;
    const int rows = min(kCv8Rows, h - y0);
    const bool right = chunk >= nc;
    const int c0 = (right ? chunk - nc : chunk) * 8;
    const size_t plane = (size_t)h * w;
    const int pitch = DP + w;
    const float* src = (right ? R : L) + ((size_t)b * C + c0) * plane + (size_t)y0 * w;
    for (int i = threadIdx.x; i < rows * w; i += 256) {  // rows are contiguous: i == r*w + x
        __nv_bfloat162 p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(__ldg(src + (2 * e) * plane + i), __ldg(src + (2 * e + 1) * plane + i));
        sv[(i / w) * pitch + DP + (i % w)] = *reinterpret_cast<const uint4*>(p);
    }
    for (int i = threadIdx.x; i < rows * DP; i += 256) sv[(i / DP) * pitch + (i % DP)] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* out = reinterpret_cast<uint4*>(cost) + (((size_t)b * 2 * nc + chunk) * D) * plane + (size_t)y0 * w;
    for (int d = warp; d < D; d += 8) {
        uint4* o = out + (size_t)d * plane;
        for (int i = lane; i < rows * w; i += 32) {
            const int r = i / w, x = i - r * w;
            uint4 v = sv[r * pitch + DP + (right ? x - d : x)];
            if (!right && x < d) v = make_uint4(0, 0, 0, 0);
            asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(o + i), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
        }
    }
}

// GroupNorm apply on C8/bf16 (+ residual C8/bf16) (+ ReLU); one CTA column per (b, chunk)
__global__ void __launch_bounds__(256, 3) gn_apply_c8_bf16_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const double* __restrict__ sums,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               const __nv_bfloat16* __restrict__ residual,
                                                               __nv_bfloat16* __restrict__ y,
                                                               __nv_bfloat16* __restrict__ y_split,
                                                               float* __restrict__ y_f32, int C, int G, int D, int H,
                                                               int W, float eps, int relu) {
    const long long spatial = (long long)D * H * W;
    __shared__ float sscale[8], sshift[8];
    const int nc = C / 8;
    const int chunk = blockIdx.y % nc;
    const long long b = blockIdx.y / nc;
    if (threadIdx.x < 8) {
        const int c = chunk * 8 + threadIdx.x;
        const int cpg = C / G;
        const int g0 = (c / cpg) * cpg;
        double s = 0.0, ss = 0.0;
        for (int j = 0; j < cpg; ++j) {
            s += sums[2 * (b * C + g0 + j)];
            ss += sums[2 * (b * C + g0 + j) + 1];
        }
        const double n = (double)cpg * (double)spatial;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double rstd = rsqrt(var + (double)eps);
        sscale[threadIdx.x] = (float)(rstd * (double)gamma[c]);
        sshift[threadIdx.x] = (float)((double)beta[c] - mean * rstd * (double)gamma[c]);
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = sscale[j];
        sh[j] = sshift[j];
    }
    const size_t base = (size_t)blockIdx.y * spatial;
    // four independent 16-byte loads (+ residual) in flight per thread: the pass is pure HBM streaming
    constexpr int U = 4;
    const int stride = (int)(gridDim.x * blockDim.x), n_sp = (int)spatial;  // spatial < 2^31 (checked by the entry point)
    const __nv_bfloat16* xb = x + base * 8;
    const __nv_bfloat16* rb = residual ? residual + base * 8 : nullptr;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_sp; i0 += U * stride) {
        uint4 raw[U], rraw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * stride;
            raw[u] = rraw[u] = make_uint4(0, 0, 0, 0);
            if (i < n_sp) {
                raw[u] = *reinterpret_cast<const uint4*>(xb + (size_t)i * 8);
                if (rb) rraw[u] = *reinterpret_cast<const uint4*>(rb + (size_t)i * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * stride;
            if (i >= n_sp) break;
            const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
            const __nv_bfloat162* rin = reinterpret_cast<const __nv_bfloat162*>(&rraw[u]);
            __nv_bfloat162 out[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float f0 = fmaf(__low2float(in[e]), sc[2 * e], sh[2 * e]);
                float f1 = fmaf(__high2float(in[e]), sc[2 * e + 1], sh[2 * e + 1]);
                if (residual) {
                    f0 += __low2float(rin[e]);
                    f1 += __high2float(rin[e]);
                }
                if (relu) {
                    f0 = fmaxf(f0, 0.f);
                    f1 = fmaxf(f1, 0.f);
                }
                out[e] = __floats2bfloat162_rn(f0, f1);
                if (y_f32 != nullptr) {  // un-rounded fp32 [B][C][D][H][W] copy (input of the fp32 classifier tail)
                    y_f32[((size_t)b * C + chunk * 8 + 2 * e) * spatial + i] = f0;
                    y_f32[((size_t)b * C + chunk * 8 + 2 * e + 1) * spatial + i] = f1;
                }
            }
            if (y != nullptr) *reinterpret_cast<uint4*>(y + (base + i) * 8) = *reinterpret_cast<const uint4*>(out);
            if (y_split != nullptr) {  // parity-split copy for a stride-2 consumer: [B][8][C/8][D/2][H/2][W/2][8]
                const int w = i % W, h = (i / W) % H, d = i / (W * H);
                const int par = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
                const size_t dst =
                    (((((size_t)b * 8 + par) * nc + chunk) * (D >> 1) + (d >> 1)) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
                *reinterpret_cast<uint4*>(y_split + dst * 8) = *reinterpret_cast<const uint4*>(out);
            }
        }
    }
}

// layout converters between C8/bf16 [B][C/8][S][8] and dense fp32 [B][C][S]
__global__ void c8_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long spatial) {
    const size_t bc = blockIdx.y;  // (b, chunk)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < spatial; i += (long long)gridDim.x * blockDim.x) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + (bc * spatial + i) * 8);
        const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            y[(bc * 8 + 2 * e) * spatial + i] = __low2float(in[e]);
            y[(bc * 8 + 2 * e + 1) * spatial + i] = __high2float(in[e]);
        }
    }
}
__global__ void f32_to_c8_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long spatial) {
    const size_t bc = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < spatial; i += (long long)gridDim.x * blockDim.x) {
        __nv_bfloat162 out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            out[e] = __floats2bfloat162_rn(x[(bc * 8 + 2 * e) * spatial + i], x[(bc * 8 + 2 * e + 1) * spatial + i]);
        *reinterpret_cast<uint4*>(y + (bc * spatial + i) * 8) = *reinterpret_cast<const uint4*>(out);
    }
}

// depth-stacked persistent schedule (conv3d_igemm_kdstack.cu): 32->32, 64->32, 64->64
int conv3d_igemm_kdstack_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout, int D,
                                  int H, int W, cudaStream_t st);

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_pack_igemm_weight_bf16(const float* weight, void* packed, int Cout, int Cin, int transposed,
                                              void* stream) {
    CMF_REQUIRE(weight && packed, "pack_igemm_weight_bf16: null pointer");
    CMF_REQUIRE(Cout > 0 && Cin > 0 && Cin % 8 == 0, "pack_igemm_weight_bf16: Cin must be a positive multiple of 8");
    const int n = 27 * Cin * Cout;
    pack_igemm_weight_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        weight, reinterpret_cast<__nv_bfloat16*>(packed), Cout, Cin, transposed);
    CMF_LAUNCH_CHECK("pack_igemm_weight_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_conv3d_igemm_bf16_fwd(const void* x_c8, const void* packed_w, void* y_c8, double* gn_sums, int B,
                                             int Cin, int Cout, int D, int H, int W, void* stream) {
    CMF_REQUIRE(x_c8 && packed_w && y_c8, "conv3d_igemm_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_igemm_bf16_fwd: non-positive dimension");
    CMF_REQUIRE((reinterpret_cast<uintptr_t>(x_c8) & 15) == 0, "conv3d_igemm_bf16_fwd: input must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    CMF_REQUIRE(Cout == 32 || (Cin == 64 && Cout == 64),
                "conv3d_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->32, 64->32, 64->64", Cin, Cout);
    return conv3d_igemm_kdstack_dispatch(x_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, D, H, W, st);
}

extern "C" int cmfb200_cost_volume_concat_c8_bf16(const float* L, const float* R, void* cost_c8, int B, int C, int h,
                                                  int w, int D, void* stream) {
    CMF_REQUIRE(L && R && cost_c8, "cost_volume_concat_c8_bf16: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && h > 0 && w > 0 && D > 0, "cost_volume_concat_c8_bf16: bad shape");
    CMF_REQUIRE(B <= 65535, "cost_volume_concat_c8_bf16: B exceeds grid limit");
    const int DP = (D + 3) & ~3;
    const size_t smem = (size_t)kCv8Rows * (DP + w) * 16;
    CMF_REQUIRE(smem <= 200 * 1024, "cost_volume_concat_c8_bf16: row block does not fit in shared memory");
    if (smem > 48 * 1024)
        CMF_CUDA(cudaFuncSetAttribute(cost_volume_c8_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)cdiv(h, kCv8Rows), (unsigned)(2 * (C / 8)), (unsigned)B);
    cost_volume_c8_bf16_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(L, R, reinterpret_cast<__nv_bfloat16*>(cost_c8),
                                                                           C, h, w, D, DP);
    CMF_LAUNCH_CHECK("cost_volume_c8_bf16_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_gn_apply_c8_bf16(const void* x_c8, const double* gn_sums, const float* gamma, const float* beta,
                                        const void* residual_c8, void* y_c8, void* y_split_c8, float* y_f32, int B, int C,
                                        int G, int D, int H, int W, float eps, int relu, void* stream) {
    CMF_REQUIRE(x_c8 && gn_sums && gamma && beta && (y_c8 || y_f32), "gn_apply_c8_bf16: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && G > 0 && C % G == 0 && D > 0 && H > 0 && W > 0, "gn_apply_c8_bf16: bad shape");
    CMF_REQUIRE(y_split_c8 == nullptr || ((D % 2 == 0) && (H % 2 == 0) && (W % 2 == 0)),
                "gn_apply_c8_bf16: the parity-split copy needs even D,H,W (got %d,%d,%d)", D, H, W);
    const long long spatial = (long long)D * H * W;
    CMF_REQUIRE(spatial < (1LL << 31) - 4LL * 256 * kNumSMs * 3, "gn_apply_c8_bf16: volume too large for 32-bit positions");
    CMF_REQUIRE((long long)B * (C / 8) <= 65535, "gn_apply_c8_bf16: B*C/8 exceeds grid limit");
    // three CTAs per SM (one wave) over all (b, chunk) columns, each thread walking four positions per iteration
    const long long cols = (long long)B * (C / 8);
    const long long gx = std::max(1LL, std::min(cdiv(spatial, 256 * 4), cdiv((long long)kNumSMs * 3, cols)));
    dim3 grid((unsigned)gx, (unsigned)cols);
    gn_apply_c8_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_c8), gn_sums, gamma, beta,
        reinterpret_cast<const __nv_bfloat16*>(residual_c8), reinterpret_cast<__nv_bfloat16*>(y_c8),
        reinterpret_cast<__nv_bfloat16*>(y_split_c8), y_f32, C, G, D, H, W, eps, relu);
    CMF_LAUNCH_CHECK("gn_apply_c8_bf16_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_c8_bf16_to_f32(const void* x_c8, float* y, int B, int C, long long spatial, void* stream) {
    CMF_REQUIRE(x_c8 && y, "c8_bf16_to_f32: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && spatial > 0 && (long long)B * (C / 8) <= 65535, "c8_bf16_to_f32: bad shape");
    dim3 grid((unsigned)min((long long)kNumSMs * 8, cdiv(spatial, 256)), (unsigned)(B * (C / 8)));
    c8_bf16_to_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_c8), y,
                                                                   spatial);
    CMF_LAUNCH_CHECK("c8_bf16_to_f32_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_f32_to_c8_bf16(const float* x, void* y_c8, int B, int C, long long spatial, void* stream) {
    CMF_REQUIRE(x && y_c8, "f32_to_c8_bf16: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && spatial > 0 && (long long)B * (C / 8) <= 65535, "f32_to_c8_bf16: bad shape");
    dim3 grid((unsigned)min((long long)kNumSMs * 8, cdiv(spatial, 256)), (unsigned)(B * (C / 8)));
    f32_to_c8_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y_c8), spatial);
    CMF_LAUNCH_CHECK("f32_to_c8_bf16_kernel");
    return CMFB200_OK;
}

