// K2 (throughput path) -- 3x3x3 stride-1 convolution of the aggregation network as an implicit GEMM on the
// 5th-generation tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in TMEM), operands
// staged by TMA, one elected thread issuing the MMAs.  Replaces nn.Conv3d(k3,s1,p1) of convbn_3d /
// hourglass.conv2,conv4 / classif.0 (cmf/models/cmfsm.py:49-58, 248-259, 604-634) in bf16 mode.
//
// Activation layout "C8": bf16 [B][C/8][D][H][W][8] -- a voxel's 8-channel group is one 16-byte unit and the
// voxels of a channel group are dense.  That makes every im2col row of every tap a 16-byte unit at a constant
// pitch, which is exactly the no-swizzle K-major UMMA canonical layout ((8,m),(8,2)):((16B,SBO),(1,LBO)):
//   * ONE rank-5 TMA box load {(8+2)*8 ch, 16+2, BD+2, C/8, 1} brings the halo'd activation block of a CTA
//     into shared memory as [C/8][BD+2][18][10] x 16 B; out-of-volume coordinates (the conv padding) are
//     zero-filled by the TMA unit;
//   * the A operand of tap (kd,kh,kw), depth slice mt, k-step kc is just a descriptor on that block:
//     start = base + 2kc*chunk + (((mt+kd)*18 + kh)*10 + kw)*16 B, SBO = 160 B (next h line), LBO = chunk:
//     128 GEMM rows = 16 h-lines x 8 w-voxels.  No im2col buffer, every activation byte is fetched once per CTA.
//   * weights are pre-packed per tap as [C/8][Cout][8] bf16 (same canonical layout, SBO = 128 B,
//     LBO = Cout*16 B) and streamed through a small ring with 1-D bulk copies.
// Work per CTA: BD depth slices x 16 x 8 voxels x all Cout; accumulators: BD tiles of 128 lanes x Cout
// columns in TMEM.  Warp roles: w0 = TMA producer, w1 = TMEM allocator + MMA issuer, w2-5 = epilogue
// (tcgen05.ld -> bf16 C8 store + GroupNorm sum/sum-of-squares, reduced per CTA, one double atomic per channel).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

constexpr int kIgTW = 8, kIgTH = 16;           // output tile in w, h (128 GEMM rows)
constexpr int kIgPW = kIgTW + 2, kIgPH = kIgTH + 2;

template <int CIN, int COUT, int BD, int NS>
struct IgCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int PD = BD + 2;
    static constexpr int VOX = PD * kIgPH * kIgPW;
    static constexpr int CHUNK_BYTES = VOX * 16;
    static constexpr int A_BYTES = NC * CHUNK_BYTES;
    static constexpr int TAP_BYTES = CIN * COUT * 2;
    static constexpr int TMEM_COLS = (BD * COUT <= 32) ? 32 : (BD * COUT <= 64) ? 64 : (BD * COUT <= 128) ? 128
                                     : (BD * COUT <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = A_BYTES + NS * TAP_BYTES + 1024 /*barriers, tmem ptr, reduction scratch*/
                                      + 4 * COUT * 2 * 8 + 1024 /*alignment slack*/;
    static_assert(BD * COUT <= 512, "accumulators exceed TMEM");
    static_assert(CIN % 16 == 0 && COUT % 16 == 0 && COUT <= 256, "UMMA shape");
};

// ---- the kernel -------------------------------------------------------------------------------------
template <int CIN, int COUT, int BD, int NS>
__global__ void __launch_bounds__(kIgThreads, 2)
    conv3d_igemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                             __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int D, int H, int W,
                             int tiles_w) {
    using G = IgCfg<CIN, COUT, BD, NS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sW = smem + G::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + NS * G::TAP_BYTES);
    uint64_t* barA = bars;            // activation block landed
    uint64_t* barD = bars + 1;        // all MMAs retired, accumulators final
    uint64_t* full = bars + 2;        // [NS] weight tap landed
    uint64_t* empty = bars + 2 + NS;  // [NS] weight tap consumed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * NS);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][COUT][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * kIgTW, h0 = tile_y * kIgTH, d0 = blockIdx.y * BD;
    const int b = blockIdx.z;

    if (threadIdx.x == 0) {
        mbar_init(barA, 1);
        mbar_init(barD, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(G::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: one box for the halo'd activation block, then the 27 weight taps
            mbar_arrive_expect_tx(barA, G::A_BYTES);
            tma_load_5d(sA, &tmap_x, barA, (w0 - 1) * 8, h0 - 1, d0 - 1, 0, b);
            for (int tap = 0; tap < 27; ++tap) {
                const int s = tap % NS;
                if (tap >= NS) mbar_wait(empty + s, ((tap / NS) - 1) & 1);
                mbar_arrive_expect_tx(full + s, G::TAP_BYTES);
                bulk_g2s(sW + s * G::TAP_BYTES, wpk + (size_t)tap * CIN * COUT, G::TAP_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer (single thread)
            // InstrDescriptor: c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1, K-major A and B, N>>3 [17,23), M>>4 [24,29)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(sA), w_0 = smem_u32(sW);
            mbar_wait(barA, 0);
            tc_fence_after();
            for (int tap = 0; tap < 27; ++tap) {
                const int s = tap % NS;
                const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                mbar_wait(full + s, (tap / NS) & 1);
                tc_fence_after();
#pragma unroll
                for (int mt = 0; mt < BD; ++mt) {
                    const uint32_t arow = a0 + ((((mt + kd) * kIgPH + kh) * kIgPW) + kw) * 16;
#pragma unroll
                    for (int kc = 0; kc < CIN / 16; ++kc) {
                        const uint64_t ad = umma_desc(arow + 2 * kc * G::CHUNK_BYTES, G::CHUNK_BYTES, kIgPW * 16);
                        const uint64_t bd = umma_desc(w_0 + s * G::TAP_BYTES + 2 * kc * (COUT * 16), COUT * 16, 128);
                        umma_bf16(tmem_base + mt * COUT, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(empty + s);  // frees the weight slot once these MMAs have read it
            }
            umma_commit(barD);
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int h = h0 + (row >> 3), w = w0 + (row & 7);
        const bool hw_ok = (h < H) && (w < W);
        const size_t plane = (size_t)H * W;
        mbar_wait(barD, 0);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < COUT / 32; ++half) {
            float s[32], ss[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                s[c] = 0.f;
                ss[c] = 0.f;
            }
#pragma unroll 1
            for (int mt = 0; mt < BD; ++mt) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + mt * COUT + half * 32, v);
                const int d = d0 + mt;
                if (hw_ok && d < D) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 p[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float f0 = __uint_as_float(v[j * 8 + 2 * e]), f1 = __uint_as_float(v[j * 8 + 2 * e + 1]);
                            p[e] = __floats2bfloat162_rn(f0, f1);
                            // statistics of the values actually stored (bf16-rounded), so GroupNorm is self-consistent
                            const float r0 = __low2float(p[e]), r1 = __high2float(p[e]);
                            s[j * 8 + 2 * e] += r0;
                            ss[j * 8 + 2 * e] = fmaf(r0, r0, ss[j * 8 + 2 * e]);
                            s[j * 8 + 2 * e + 1] += r1;
                            ss[j * 8 + 2 * e + 1] = fmaf(r1, r1, ss[j * 8 + 2 * e + 1]);
                        }
                        const int chunk = half * 4 + j;
                        __nv_bfloat16* dst = y + ((((size_t)b * (COUT / 8) + chunk) * D + d) * plane + (size_t)h * W + w) * 8;
                        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(p);
                    }
                }
            }
            if (gn_sums != nullptr) {  // lane l ends up with the warp total of column half*32 + l
                sred[(quad * COUT + half * 32 + lane) * 2 + 0] = (double)warp_transpose_sum32(s, lane);
                sred[(quad * COUT + half * 32 + lane) * 2 + 1] = (double)warp_transpose_sum32(ss, lane);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (gn_sums != nullptr && threadIdx.x < COUT * 2) {
        const int c = threadIdx.x >> 1, which = threadIdx.x & 1;
        double a = 0.0;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) a += sred[(qd * COUT + c) * 2 + which];
        atomicAdd(gn_sums + ((size_t)b * COUT + c) * 2 + which, a);
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(G::TMEM_COLS)
                     : "memory");
    }
}

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
__global__ void __launch_bounds__(256) gn_apply_c8_bf16_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const double* __restrict__ sums,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               const __nv_bfloat16* __restrict__ residual,
                                                               __nv_bfloat16* __restrict__ y,
                                                               __nv_bfloat16* __restrict__ y_split, int C, int G,
                                                               int D, int H, int W, float eps, int relu) {
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
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < spatial; i += (long long)gridDim.x * blockDim.x) {
        uint4 raw = *reinterpret_cast<const uint4*>(x + (base + i) * 8);
        const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(&raw);
        uint4 rraw = make_uint4(0, 0, 0, 0);
        if (residual) rraw = *reinterpret_cast<const uint4*>(residual + (base + i) * 8);
        const __nv_bfloat162* rin = reinterpret_cast<const __nv_bfloat162*>(&rraw);
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
        }
        *reinterpret_cast<uint4*>(y + (base + i) * 8) = *reinterpret_cast<const uint4*>(out);
        if (y_split != nullptr) {  // parity-split copy for a stride-2 consumer: [B][8][C/8][D/2][H/2][W/2][8]
            const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / ((long long)W * H));
            const int par = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
            const size_t dst = (((((size_t)b * 8 + par) * nc + chunk) * (D >> 1) + (d >> 1)) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
            *reinterpret_cast<uint4*>(y_split + dst * 8) = *reinterpret_cast<const uint4*>(out);
        }
    }
}

// classifier tail conv (Cin -> 1, 3x3x3, pad 1) straight from C8/bf16 to the fp32 volume K4 consumes
// (classifN.2, cmf/models/cmfsm.py:624,629,634).  N=1 has no tensor-core shape; the op is bound by reading the
// input once (106 MB at config 2): one thread per output voxel, 128-bit loads that hit L1 for the 27-fold reuse,
// fp32 accumulation; weights [Cin][27] fp32 staged in shared memory as [Cin/8][27][8].
__global__ void __launch_bounds__(256) conv3d_c8_cout1_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const float* __restrict__ wgt, float* __restrict__ y,
                                                              int NC, int D, int H, int W) {
    extern __shared__ __align__(16) float sw[];  // [NC][27][8]
    for (int i = threadIdx.x; i < NC * 27 * 8; i += 256) {
        const int j = i & 7, tap = (i >> 3) % 27, chunk = i / (27 * 8);
        sw[i] = wgt[(chunk * 8 + j) * 27 + tap];
    }
    __syncthreads();
    const size_t plane = (size_t)H * W, vol = (size_t)D * plane;
    const int b = blockIdx.y;
    const uint4* xb = reinterpret_cast<const uint4*>(x) + (size_t)b * NC * vol;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < vol; i += (size_t)gridDim.x * 256) {
        const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / plane);
        float acc = 0.f;
        for (int chunk = 0; chunk < NC; ++chunk) {
            const uint4* xc = xb + (size_t)chunk * vol;
            const float* wc = sw + chunk * 27 * 8;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int dd = d + kd - 1;
                if ((unsigned)dd >= (unsigned)D) continue;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = h + kh - 1;
                    if ((unsigned)hh >= (unsigned)H) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ww = w + kw - 1;
                        if ((unsigned)ww >= (unsigned)W) continue;
                        const uint4 raw = __ldg(xc + (size_t)dd * plane + (size_t)hh * W + ww);
                        const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
                        const float4 w0 = *reinterpret_cast<const float4*>(wc + ((kd * 3 + kh) * 3 + kw) * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(wc + ((kd * 3 + kh) * 3 + kw) * 8 + 4);
                        acc = fmaf(__low2float(v[0]), w0.x, acc);
                        acc = fmaf(__high2float(v[0]), w0.y, acc);
                        acc = fmaf(__low2float(v[1]), w0.z, acc);
                        acc = fmaf(__high2float(v[1]), w0.w, acc);
                        acc = fmaf(__low2float(v[2]), w1.x, acc);
                        acc = fmaf(__high2float(v[2]), w1.y, acc);
                        acc = fmaf(__low2float(v[3]), w1.z, acc);
                        acc = fmaf(__high2float(v[3]), w1.w, acc);
                    }
                }
            }
        }
        y[(size_t)b * vol + i] = acc;
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

// ---- host side ----------------------------------------------------------------------------------------
template <int CIN, int COUT, int BD, int NS>
static int launch_igemm(const void* x, const void* wpk, void* y, double* gn, int B, int D, int H, int W,
                        cudaStream_t st) {
    using G = IgCfg<CIN, COUT, BD, NS>;
    EncodeTiledFn encode = get_encode_fn();
    CMF_REQUIRE(encode != nullptr, "conv3d_igemm: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)G::NC * D * H * W * 16};
    const cuuint32_t box[5] = {kIgPW * 8, kIgPH, (cuuint32_t)G::PD, (cuuint32_t)G::NC, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CMF_REQUIRE(r == CUDA_SUCCESS, "conv3d_igemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    auto kern = conv3d_igemm_bf16_kernel<CIN, COUT, BD, NS>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(W, kIgTW), tiles_h = (int)cdiv(H, kIgTH);
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)cdiv(D, BD), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d_igemm: grid too large");
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(wpk),
                                                  reinterpret_cast<__nv_bfloat16*>(y), gn, D, H, W, tiles_w);
    CMF_LAUNCH_CHECK("conv3d_igemm_bf16_kernel");
    return CMFB200_OK;
}

// persistent schedule (conv3d_igemm_persistent.cu); the kernel above is kept as the simple reference schedule
int conv3d_igemm_persistent_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                     int D, int H, int W, cudaStream_t st);

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
    static const bool simple_schedule = getenv("CMFB200_IGEMM_SIMPLE") != nullptr;  // A/B switch for profiling
    if (!simple_schedule) return conv3d_igemm_persistent_dispatch(x_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, D, H, W, st);
    if (Cin == 32 && Cout == 32) return launch_igemm<32, 32, 4, 4>(x_c8, packed_w, y_c8, gn_sums, B, D, H, W, st);
    if (Cin == 64 && Cout == 32) return launch_igemm<64, 32, 2, 4>(x_c8, packed_w, y_c8, gn_sums, B, D, H, W, st);
    if (Cin == 64 && Cout == 64) return launch_igemm<64, 64, 2, 2>(x_c8, packed_w, y_c8, gn_sums, B, D, H, W, st);
    CMF_REQUIRE(false, "conv3d_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->32, 64->32, 64->64", Cin,
                Cout);
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
                                        const void* residual_c8, void* y_c8, void* y_split_c8, int B, int C, int G,
                                        int D, int H, int W, float eps, int relu, void* stream) {
    CMF_REQUIRE(x_c8 && gn_sums && gamma && beta && y_c8, "gn_apply_c8_bf16: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && G > 0 && C % G == 0 && D > 0 && H > 0 && W > 0, "gn_apply_c8_bf16: bad shape");
    CMF_REQUIRE(y_split_c8 == nullptr || ((D % 2 == 0) && (H % 2 == 0) && (W % 2 == 0)),
                "gn_apply_c8_bf16: the parity-split copy needs even D,H,W (got %d,%d,%d)", D, H, W);
    const long long spatial = (long long)D * H * W;
    CMF_REQUIRE((long long)B * (C / 8) <= 65535, "gn_apply_c8_bf16: B*C/8 exceeds grid limit");
    dim3 grid((unsigned)min((long long)kNumSMs * 8, cdiv(spatial, 256)), (unsigned)(B * (C / 8)));
    gn_apply_c8_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_c8), gn_sums, gamma, beta,
        reinterpret_cast<const __nv_bfloat16*>(residual_c8), reinterpret_cast<__nv_bfloat16*>(y_c8),
        reinterpret_cast<__nv_bfloat16*>(y_split_c8), C, G, D, H, W, eps, relu);
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

extern "C" int cmfb200_conv3d_c8_cout1_fwd(const void* x_c8, const float* weight, float* y, int B, int Cin, int D, int H,
                                           int W, void* stream) {
    CMF_REQUIRE(x_c8 && weight && y, "conv3d_c8_cout1_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && Cin % 8 == 0 && D > 0 && H > 0 && W > 0 && B <= 65535, "conv3d_c8_cout1_fwd: bad shape");
    const long long vol = (long long)D * H * W;
    const size_t smem = (size_t)Cin * 27 * sizeof(float);
    dim3 grid((unsigned)min((long long)kNumSMs * 16, cdiv(vol, 256)), (unsigned)B);
    conv3d_c8_cout1_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_c8), weight,
                                                                       y, Cin / 8, D, H, W);
    CMF_LAUNCH_CHECK("conv3d_c8_cout1_kernel");
    return CMFB200_OK;
}
