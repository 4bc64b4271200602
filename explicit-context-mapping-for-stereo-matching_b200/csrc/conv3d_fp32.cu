// K2 (fp32 parity path) -- 3x3x3 convolutions / transposed convolutions of the aggregation network on the
// CUDA cores with exact fp32 FMA accumulation, NCDHW.  Replaces nn.Conv3d / nn.ConvTranspose3d inside
// convbn_3d, hourglass and classif (cmf/models/cmfsm.py:49-58, 240-303, 604-634).  This is the mode the
// fp32 parity gate runs in (the reference's fp32 result moves by 2e-3 px between thread counts, SURVEY.md
// section 0.7, so tensor-core operand rounding is not an option here); the bf16 tcgen05 implicit GEMM in
// conv3d_igemm_kdstack.cu is the throughput mode.
//
// Register-tiled direct convolution: a CTA owns a TD x TH x 32 block of output voxels and ALL output
// channels; per chunk of CC input channels the halo'd input patch and the [CC][27][COUT] weight slice
// are staged in shared memory (cp.async, 2-stage ring: chunk i+1 is prefetched while chunk i is consumed;
// per-thread staging slots computed once); a thread owns CPT output channels x 4 consecutive-w voxels (32 fp32
// accumulators for CPT=8).  Per (ci,kd,kh) it issues 2 vector loads of input + 6 broadcast loads of weights
// for 96 FMAs, i.e. the inner loop is FMA-issue bound, not shared-memory bound.
// GroupNorm statistics (sum, sum of squares per (b,channel)) are reduced in the epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace cmfb200 {

// ------------------------------------------------------------------------------------------------
// forward convolution, stride S in {1,2}
// ------------------------------------------------------------------------------------------------
template <int COUT, int CPT, int S, int CC, int TW>
__global__ void __launch_bounds__(kConvThreads, 2)
    conv3d_k3_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y,
                     double* __restrict__ gn_sums, int Cin, int D, int H, int W, int Do, int Ho, int Wo, int tiles_w,
                     int hoff) {
    using T = ConvTile<COUT, CPT, TW>;
    constexpr int PD = (T::TD - 1) * S + 3, PH = (T::TH - 1) * S + 3, PW = (TW - 1) * S + 3;
    constexpr int PWP = (PW + 3) & ~3;
    constexpr int PATCH = PD * PH * PWP;  // floats per input channel
    constexpr int WSL = 27 * COUT;        // weight floats per input channel
    constexpr int NI = (kVPT - 1) * S + 3;
    constexpr int NI4 = (NI + 3) / 4;
    constexpr int NSLOT = (PD * PH * PW + kConvThreads - 1) / kConvThreads;
    constexpr int STAGE = CC * (PATCH + WSL);  // floats per pipeline stage
    // stride 2: a patch row is stored de-interleaved, [even columns | odd columns] (odd half at HALF): output v reads
    // columns 2v, 2v+1, 2v+2 = E[v], O[v], E[v+1], so a thread's inputs are one aligned float4 + one scalar of E and one
    // aligned float4 of O at a 16-byte lane pitch (conflict-free) instead of three float4 at a 32-byte pitch (2-way).
    constexpr int HALF = (((PW + 1) / 2) + 3) & ~3;
    static_assert(S == 1 || HALF + PW / 2 <= PWP, "de-interleaved row does not fit the row pitch");
    static_assert((TW / kVPT - 1) * kVPT * S + NI4 * 4 <= PWP, "vector over-read leaves the patch row");
    static_assert((CC * WSL) % 4 == 0 && (CC * PATCH) % 4 == 0, "stage slices must stay 16-byte aligned");

    extern __shared__ __align__(16) float smem[];  // 2 stages of [CC][PD][PH][PWP] + [CC][27][COUT]

    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * TW, h0 = tile_y * T::TH, d0 = blockIdx.y * T::TD;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid / T::NQ;
    const int q = tid % T::NQ;
    const int qx = q % (TW / kVPT);
    const int row = q / (TW / kVPT);
    const int th = row % T::TH, td = row / T::TH;

    // accumulators as channel pairs: one FFMA2 (sm_100 packed fp32x2 FMA, two IEEE round-to-nearest FMAs) updates
    // channels (2j, 2j+1) of one voxel -- half the issue slots of scalar FFMA, bit-identical results
    constexpr int CP2 = (CPT + 1) / 2;
    float2 acc2[CP2][kVPT];
#pragma unroll
    for (int c = 0; c < CP2; ++c)
#pragma unroll
        for (int v = 0; v < kVPT; ++v) acc2[c][v] = make_float2(0.f, 0.f);

    const size_t in_plane = (size_t)H * W;
    const size_t in_vol = (size_t)D * in_plane;
    const float* xb = x + (size_t)b * Cin * in_vol;
    const int di0 = d0 * S - 1, hi0 = h0 * S - 1 + hoff, wi0 = w0 * S - 1;  // hoff: row window (row-band sharding)

    // ---- staging slots of this thread (identical for every input channel): computed once
    int goff[NSLOT], soff[NSLOT];
    bool ok[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
        const int e = tid + j * kConvThreads;
        const int pw = e % PW;
        const int r = e / PW;
        const int ph = r % PH, pd = r / PH;
        const int di = di0 + pd, hi = hi0 + ph, wi = wi0 + pw;
        const bool in_patch = e < PD * PH * PW;
        ok[j] = in_patch && (unsigned)di < (unsigned)D && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        goff[j] = ok[j] ? (di * H + hi) * W + wi : 0;
        const int pcol = (S == 2) ? ((pw & 1) ? HALF + (pw >> 1) : (pw >> 1)) : pw;
        soff[j] = in_patch ? (pd * PH + ph) * PWP + pcol : -1;
    }

    auto stage = [&](int c0, int buf) {
        float* sIn = smem + buf * STAGE;
        float* sW = sIn + CC * PATCH;
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
            const float* src = xb + (size_t)(c0 + ci) * in_vol;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j)
                if (soff[j] >= 0) cp_async_4_zfill(sIn + ci * PATCH + soff[j], src + goff[j], ok[j]);
        }
        const float* wsrc = wp + (size_t)c0 * WSL;  // contiguous [CC][27][COUT] slice
        for (int i = tid * 4; i < CC * WSL; i += kConvThreads * 4) cp_async_16(sW + i, wsrc + i);
        cp_async_commit();
    };

    // zero the alignment tail of every patch row once (never written by cp.async, read by the vector loads);
    // the de-interleaved stride-2 rows have no slot that is read but not written
    if constexpr (PWP > PW && S == 1) {
        for (int i = tid; i < 2 * CC * PD * PH * (PWP - PW); i += kConvThreads) {
            const int t = i % (PWP - PW);
            const int r = i / (PWP - PW);  // (buf, ci, pd*PH+ph) flattened
            const int prow = r % (PD * PH), ci = (r / (PD * PH)) % CC, buf = r / (PD * PH * CC);
            smem[buf * STAGE + ci * PATCH + prow * PWP + PW + t] = 0.f;
        }
    }

    const int nchunks = Cin / CC;
    stage(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            stage((ch + 1) * CC, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sIn = smem + buf * STAGE;
        const float* sW = sIn + CC * PATCH;
#pragma unroll 1
        for (int ci = 0; ci < CC; ++ci) {
            const float* pin = sIn + ci * PATCH + (td * S * PH + th * S) * PWP + qx * kVPT * (S == 2 ? 1 : S);
            const float* pw_ = sW + ci * WSL + cg * CPT;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float* prow = pin + (kd * PH + kh) * PWP;
                    float in[NI4 * 4];
                    if constexpr (S == 2) {
                        const float4 e = *reinterpret_cast<const float4*>(prow);
                        const float e4 = prow[4];
                        const float4 o = *reinterpret_cast<const float4*>(prow + HALF);
                        // in[2v + kw]: even positions from E, odd from O
                        in[0] = e.x; in[1] = o.x; in[2] = e.y; in[3] = o.y; in[4] = e.z; in[5] = o.z; in[6] = e.w;
                        in[7] = o.w; in[8] = e4;
                    } else {
#pragma unroll
                        for (int j = 0; j < NI4; ++j) {
                            const float4 a = *reinterpret_cast<const float4*>(prow + 4 * j);
                            in[4 * j + 0] = a.x; in[4 * j + 1] = a.y; in[4 * j + 2] = a.z; in[4 * j + 3] = a.w;
                        }
                    }
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float* wt = pw_ + ((kd * 3 + kh) * 3 + kw) * COUT;
                        float2 w2[CP2];
                        if constexpr (CPT == 8) {
                            const float4 w0v = *reinterpret_cast<const float4*>(wt);
                            const float4 w1v = *reinterpret_cast<const float4*>(wt + 4);
                            w2[0] = make_float2(w0v.x, w0v.y); w2[1] = make_float2(w0v.z, w0v.w);
                            w2[2] = make_float2(w1v.x, w1v.y); w2[3] = make_float2(w1v.z, w1v.w);
                        } else {
                            w2[0] = make_float2(wt[0], 0.f);
                        }
#pragma unroll
                        for (int v = 0; v < kVPT; ++v) {
                            const float2 i2 = make_float2(in[v * S + kw], in[v * S + kw]);
#pragma unroll
                            for (int c = 0; c < CP2; ++c) acc2[c][v] = __ffma2_rn(w2[c], i2, acc2[c][v]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone done with `buf` before the next prefetch overwrites it
    }

    float acc[CPT][kVPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int v = 0; v < kVPT; ++v) acc[c][v] = (c & 1) ? acc2[c >> 1][v].y : acc2[c >> 1][v].x;

    // ---- epilogue: store + GroupNorm partial sums
    const int od = d0 + td, oh = h0 + th, ow = w0 + qx * kVPT;
    const bool row_ok = (od < Do) && (oh < Ho);
    const size_t out_plane = (size_t)Ho * Wo;
    double s[CPT], ss[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        s[c] = 0.0;
        ss[c] = 0.0;
    }
    if (row_ok && ow < Wo) {
        const bool vec = ((Wo & 3) == 0);  // then ow+3 < Wo and the address is 16-byte aligned
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int co = cg * CPT + c;
            float* py = y + (((size_t)b * COUT + co) * Do + od) * out_plane + (size_t)oh * Wo + ow;
            if (vec) {
                *reinterpret_cast<float4*>(py) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
#pragma unroll
                for (int v = 0; v < kVPT; ++v) {
                    s[c] += (double)acc[c][v];
                    ss[c] = fma((double)acc[c][v], (double)acc[c][v], ss[c]);
                }
            } else {
#pragma unroll
                for (int v = 0; v < kVPT; ++v)
                    if (ow + v < Wo) {
                        py[v] = acc[c][v];
                        s[c] += (double)acc[c][v];
                        ss[c] = fma((double)acc[c][v], (double)acc[c][v], ss[c]);
                    }
            }
        }
    }
    if (gn_sums != nullptr) gn_epilogue<COUT, CPT>(s, ss, cg, smem, gn_sums, b);
}

// ------------------------------------------------------------------------------------------------
// stride-1 convolution, TWO output rows per thread (the layers that carry 90 % of the 3-D MACs).
// With FFMA2 the scalar kernel above became shared-memory bound (ncu: 86 % of the LSU wavefront peak vs 74 % FMA
// pipe): rows r and r+1 share the four input rows r..r+3, so per (ci,kd) a thread issues 8 vector loads of input +
// 18 broadcast loads of weights for 576 FMAs (26 loads) instead of 3 x (2+6) = 24 loads for 288 FMAs.
// Tile: 2 depth slices x TH rows x TW columns, TW in {32, 16}.
// ------------------------------------------------------------------------------------------------
template <int COUT, int TW>
struct Conv3dR2Cfg {
    static constexpr int CPT = 8;
    static constexpr int NCG = COUT / CPT;
    static constexpr int NQ = kConvThreads / NCG;
    static constexpr int NTR = NQ / (TW / kVPT);  // thread rows (each = one row pair)
    static constexpr int TD = 2;
    static constexpr int TH = NTR;                // NTR/2 pairs per depth slice x 2 rows
    static constexpr int PD = TD + 2, PH = TH + 2, PW = TW + 2;
    static constexpr int PWP = (TW == 16) ? 24 : 36;  // 24: rows 2 apart land 16 banks apart (conflict-free LDS.128)
    static constexpr int PATCH = PD * PH * PWP;
    static constexpr int WSL = 27 * COUT;
    static constexpr int NSLOT = (PD * PH * PW + kConvThreads - 1) / kConvThreads;
    static_assert(NTR % 2 == 0 && (TW / kVPT - 1) * kVPT + 8 <= PWP, "bad R2 tile");
};

template <int COUT, int CC, int TW>
__global__ void __launch_bounds__(kConvThreads, 2)
    conv3d_k3_r2_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y,
                        double* __restrict__ gn_sums, int Cin, int D, int H, int W, int tiles_w, int hoff, int Ho) {
    using G = Conv3dR2Cfg<COUT, TW>;
    constexpr int CPT = G::CPT;
    constexpr int STAGE = CC * (G::PATCH + G::WSL);
    extern __shared__ __align__(16) float smem[];

    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * TW, h0 = tile_y * G::TH, d0 = blockIdx.y * G::TD;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid / G::NQ;
    const int q = tid % G::NQ;
    const int qx = q % (TW / kVPT);
    const int tr = q / (TW / kVPT);
    const int td = tr / (G::NTR / 2);
    const int r0 = 2 * (tr % (G::NTR / 2));  // first output row of this thread inside the tile; second = r0 + 1

    const size_t in_plane = (size_t)H * W;
    const size_t in_vol = (size_t)D * in_plane;
    const float* xb = x + (size_t)b * Cin * in_vol;
    const int di0 = d0 - 1, hi0 = h0 - 1 + hoff, wi0 = w0 - 1;  // hoff / Ho: row window (row-band sharding)

    int goff[G::NSLOT], soff[G::NSLOT];
    bool ok[G::NSLOT];
#pragma unroll
    for (int j = 0; j < G::NSLOT; ++j) {
        const int e = tid + j * kConvThreads;
        const int pw = e % G::PW;
        const int r = e / G::PW;
        const int ph = r % G::PH, pd = r / G::PH;
        const int di = di0 + pd, hi = hi0 + ph, wi = wi0 + pw;
        const bool in_patch = e < G::PD * G::PH * G::PW;
        ok[j] = in_patch && (unsigned)di < (unsigned)D && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        goff[j] = ok[j] ? (di * H + hi) * W + wi : 0;
        soff[j] = in_patch ? (pd * G::PH + ph) * G::PWP + pw : -1;
    }
    auto stage = [&](int c0, int buf) {
        float* sIn = smem + buf * STAGE;
        float* sW = sIn + CC * G::PATCH;
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
            const float* src = xb + (size_t)(c0 + ci) * in_vol;
#pragma unroll
            for (int j = 0; j < G::NSLOT; ++j)
                if (soff[j] >= 0) cp_async_4_zfill(sIn + ci * G::PATCH + soff[j], src + goff[j], ok[j]);
        }
        const float* wsrc = wp + (size_t)c0 * G::WSL;
        for (int i = tid * 4; i < CC * G::WSL; i += kConvThreads * 4) cp_async_16(sW + i, wsrc + i);
        cp_async_commit();
    };

    float2 acc2[2][CPT / 2][kVPT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < CPT / 2; ++c)
#pragma unroll
            for (int v = 0; v < kVPT; ++v) acc2[r][c][v] = make_float2(0.f, 0.f);

    // zero the alignment tail of every patch row once (never written by cp.async, read by the vector loads)
    for (int i = tid; i < 2 * CC * G::PD * G::PH * (G::PWP - G::PW); i += kConvThreads) {
        const int t = i % (G::PWP - G::PW);
        const int r = i / (G::PWP - G::PW);
        const int prow = r % (G::PD * G::PH), ci = (r / (G::PD * G::PH)) % CC, buf = r / (G::PD * G::PH * CC);
        smem[buf * STAGE + ci * G::PATCH + prow * G::PWP + G::PW + t] = 0.f;
    }

    const int nchunks = Cin / CC;
    stage(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            stage((ch + 1) * CC, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sIn = smem + buf * STAGE;
        const float* sW = sIn + CC * G::PATCH;
#pragma unroll 1
        for (int ci = 0; ci < CC; ++ci) {
            const float* pin = sIn + ci * G::PATCH + (td * G::PH + r0) * G::PWP + qx * kVPT;
            const float* pw_ = sW + ci * G::WSL + cg * CPT;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                float in[4][8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* prow = pin + (kd * G::PH + j) * G::PWP;
                    const float4 a = *reinterpret_cast<const float4*>(prow);
                    const float4 c4 = *reinterpret_cast<const float4*>(prow + 4);
                    in[j][0] = a.x; in[j][1] = a.y; in[j][2] = a.z; in[j][3] = a.w;
                    in[j][4] = c4.x; in[j][5] = c4.y; in[j][6] = c4.z; in[j][7] = c4.w;
                }
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float* wt = pw_ + ((kd * 3 + kh) * 3 + kw) * COUT;
                        const float4 w0v = *reinterpret_cast<const float4*>(wt);
                        const float4 w1v = *reinterpret_cast<const float4*>(wt + 4);
                        const float2 w2[CPT / 2] = {make_float2(w0v.x, w0v.y), make_float2(w0v.z, w0v.w),
                                                    make_float2(w1v.x, w1v.y), make_float2(w1v.z, w1v.w)};
#pragma unroll
                        for (int v = 0; v < kVPT; ++v) {
                            const float2 i0 = make_float2(in[kh][v + kw], in[kh][v + kw]);
                            const float2 i1 = make_float2(in[kh + 1][v + kw], in[kh + 1][v + kw]);
#pragma unroll
                            for (int c = 0; c < CPT / 2; ++c) {
                                acc2[0][c][v] = __ffma2_rn(w2[c], i0, acc2[0][c][v]);
                                acc2[1][c][v] = __ffma2_rn(w2[c], i1, acc2[1][c][v]);
                            }
                        }
                    }
            }
        }
        __syncthreads();
    }

    const int od = d0 + td, ow = w0 + qx * kVPT;
    const size_t out_plane = (size_t)Ho * W;
    double s[CPT], ss[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        s[c] = 0.0;
        ss[c] = 0.0;
    }
    const bool vec = ((W & 3) == 0);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oh = h0 + r0 + r;
        if (od < D && oh < Ho && ow < W) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                float a[kVPT];
#pragma unroll
                for (int v = 0; v < kVPT; ++v) a[v] = (c & 1) ? acc2[r][c >> 1][v].y : acc2[r][c >> 1][v].x;
                float* py = y + (((size_t)b * COUT + cg * CPT + c) * D + od) * out_plane + (size_t)oh * W + ow;
                if (vec) *reinterpret_cast<float4*>(py) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
                for (int v = 0; v < kVPT; ++v)
                    if (vec || ow + v < W) {
                        if (!vec) py[v] = a[v];
                        s[c] += (double)a[v];
                        ss[c] = fma((double)a[v], (double)a[v], ss[c]);
                    }
            }
        }
    }
    if (gn_sums != nullptr) gn_epilogue<COUT, CPT>(s, ss, cg, smem, gn_sums, b);
}

// Row window (row-band sharding of one image pair): the input holds halo rows, output row m reads input rows
// m*S - 1 + hoff + kh and only Ho output rows exist.  hoff = 0 / Ho = -1: the ordinary padded convolution.
struct RowWin {
    int hoff = 0, Ho = -1;
};

template <int COUT, int CC, int TW>
static int launch_conv_r2(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int D, int H, int W,
                          cudaStream_t st, RowWin rw = RowWin()) {
    using G = Conv3dR2Cfg<COUT, TW>;
    constexpr size_t smem = 2 * (size_t)CC * (G::PATCH + G::WSL) * sizeof(float);
    static_assert(smem <= 110 * 1024, "two CTAs per SM must fit");
    const int Ho = rw.Ho >= 0 ? rw.Ho : H;
    const int tiles_w = (int)cdiv(W, TW), tiles_h = (int)cdiv(Ho, G::TH);
    auto kern = conv3d_k3_r2_kernel<COUT, CC, TW>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)cdiv(D, G::TD), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d: grid too large");
    kern<<<grid, kConvThreads, smem, st>>>(x, wp, y, gn, Cin, D, H, W, tiles_w, rw.hoff, Ho);
    CMF_LAUNCH_CHECK("conv3d_k3_r2_kernel");
    return CMFB200_OK;
}

// stride-1, Cout in {32,64}: two-rows-per-thread kernel with the tile width that wastes fewer lanes
template <int COUT, int CC>
static int launch_conv_r2_best(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int D, int H, int W,
                               cudaStream_t st, RowWin rw = RowWin()) {
    using G32 = Conv3dR2Cfg<COUT, 32>;
    using G16 = Conv3dR2Cfg<COUT, 16>;
    const long long Hr = rw.Ho >= 0 ? rw.Ho : H;
    const long long c32 = cdiv(W, 32) * 32 * cdiv(Hr, G32::TH) * G32::TH;
    const long long c16 = cdiv(W, 16) * 16 * cdiv(Hr, G16::TH) * G16::TH;
    if (c16 < c32) return launch_conv_r2<COUT, CC, 16>(x, wp, y, gn, B, Cin, D, H, W, st, rw);
    return launch_conv_r2<COUT, CC, 32>(x, wp, y, gn, B, Cin, D, H, W, st, rw);
}

// ------------------------------------------------------------------------------------------------
// transposed convolution k3 s2 p1 op1 (output = 2x input).  od = 2*id - 1 + kd.
// A CTA owns one (pd,ph) output-parity class of a TH x 32 block of INPUT positions at one input depth:
// outputs (od,oh) = (2*id+pd, 2*ih+ph), ow = 2*iw .. 2*iw+1.  Along each of d,h: parity 0 uses tap k=1 of
// input i; parity 1 uses tap k=2 of input i and tap k=0 of input i+1.  A thread owns CPT channels x 8
// consecutive output columns (4 input columns).
// ------------------------------------------------------------------------------------------------
template <int COUT, int CPT, int CC>
__global__ void __launch_bounds__(kConvThreads, 2)
    deconv3d_k3s2_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y,
                         double* __restrict__ gn_sums, int Cin, int D, int H, int W, int tiles_w, int Hc) {
    using T = ConvTile<COUT, CPT>;
    static_assert(T::TD == 1, "deconv tile is one depth slice");
    constexpr int PD = 2, PH = T::TH + 1, PW = kTW + 1;
    constexpr int PWP = (PW + 3) & ~3;  // 36
    constexpr int PATCH = PD * PH * PWP;
    constexpr int WSL = 12 * COUT;  // at most 2 x 2 (d,h) tap pairs x 3 w taps are live for one parity class
    constexpr int NO = 2 * kVPT;    // output columns per thread
    constexpr int NSLOT = (PD * PH * PW + kConvThreads - 1) / kConvThreads;
    constexpr int STAGE = CC * (PATCH + WSL);

    extern __shared__ __align__(16) float smem[];  // 2 stages of [CC][2][PH][PWP] + [CC][nd*nh*3][COUT]

    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int iw0 = tile_x * kTW, ih0 = tile_y * T::TH;
    const int id = blockIdx.y >> 2;
    const int pd = (blockIdx.y >> 1) & 1, ph = blockIdx.y & 1;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid / T::NQ;
    const int q = tid % T::NQ;
    const int qx = q % (kTW / kVPT);
    const int th = q / (kTW / kVPT);

    float2 acc2[CPT / 2][NO];  // channel pairs, updated with FFMA2 (packed fp32x2 FMA)
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c)
#pragma unroll
        for (int v = 0; v < NO; ++v) acc2[c][v] = make_float2(0.f, 0.f);

    const size_t in_plane = (size_t)H * W;
    const size_t in_vol = (size_t)D * in_plane;
    const float* xb = x + (size_t)b * Cin * in_vol;
    const int nd = pd ? 2 : 1, nh = ph ? 2 : 1;
    const int ntap = nd * nh * 3;  // live taps of this parity class

    int goff[NSLOT], soff[NSLOT];
    bool ok[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
        const int e = tid + j * kConvThreads;
        const int pw = e % PW;
        const int r = e / PW;
        const int phh = r % PH, pdd = r / PH;
        const int di = id + pdd, hi = ih0 + phh, wi = iw0 + pw;
        const bool in_patch = e < PD * PH * PW;
        ok[j] = in_patch && di < D && hi < H && wi < W;
        goff[j] = ok[j] ? (di * H + hi) * W + wi : 0;
        soff[j] = in_patch ? (pdd * PH + phh) * PWP + pw : -1;
    }

    auto stage = [&](int c0, int buf) {
        float* sIn = smem + buf * STAGE;
        float* sW = sIn + CC * PATCH;
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
            const float* src = xb + (size_t)(c0 + ci) * in_vol;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j)
                if (soff[j] >= 0) cp_async_4_zfill(sIn + ci * PATCH + soff[j], src + goff[j], ok[j]);
        }
        // live weight rows only: [ci][jd][jh][kw][COUT]; parity 0 uses tap k=1, parity 1 taps k=2 (j=0) and k=0 (j=1)
        constexpr int ROW4 = COUT / 4;
        for (int i = tid; i < CC * ntap * ROW4; i += kConvThreads) {
            const int j4 = i % ROW4;
            int r = i / ROW4;
            const int kw = r % 3;
            r /= 3;
            const int jh = r % nh;
            r /= nh;
            const int jd = r % nd, ci = r / nd;
            const int kd = pd ? (jd == 0 ? 2 : 0) : 1, kh = ph ? (jh == 0 ? 2 : 0) : 1;
            cp_async_16(sW + ci * WSL + ((jd * nh + jh) * 3 + kw) * COUT + j4 * 4,
                        wp + ((size_t)(c0 + ci) * 27 + (kd * 3 + kh) * 3 + kw) * COUT + j4 * 4);
        }
        cp_async_commit();
    };

    // zero the alignment tail of every patch row once
    for (int i = tid; i < 2 * CC * PD * PH * (PWP - PW); i += kConvThreads) {
        const int t = i % (PWP - PW);
        const int r = i / (PWP - PW);
        const int prow = r % (PD * PH), ci = (r / (PD * PH)) % CC, buf = r / (PD * PH * CC);
        smem[buf * STAGE + ci * PATCH + prow * PWP + PW + t] = 0.f;
    }

    const int nchunks = Cin / CC;
    stage(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            stage((ch + 1) * CC, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sIn = smem + buf * STAGE;
        const float* sW = sIn + CC * PATCH;
#pragma unroll 1
        for (int ci = 0; ci < CC; ++ci) {
            const float* pin = sIn + ci * PATCH + th * PWP + qx * kVPT;
            const float* pw_ = sW + ci * WSL + cg * CPT;
#pragma unroll 1
            for (int jd = 0; jd < nd; ++jd) {
#pragma unroll 1
                for (int jh = 0; jh < nh; ++jh) {
                    const float* prow = pin + (jd * PH + jh) * PWP;
                    const float4 a = *reinterpret_cast<const float4*>(prow);
                    const float in[5] = {a.x, a.y, a.z, a.w, prow[4]};
                    const float* wt = pw_ + (jd * nh + jh) * 3 * COUT;
                    float2 wk[3][CPT / 2];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 w0v = *reinterpret_cast<const float4*>(wt + kw * COUT);
                        const float4 w1v = *reinterpret_cast<const float4*>(wt + kw * COUT + 4);
                        wk[kw][0] = make_float2(w0v.x, w0v.y); wk[kw][1] = make_float2(w0v.z, w0v.w);
                        wk[kw][2] = make_float2(w1v.x, w1v.y); wk[kw][3] = make_float2(w1v.z, w1v.w);
                    }
#pragma unroll
                    for (int v = 0; v < kVPT; ++v) {
                        // even column 2*(i+v): tap kw=1 of input v ; odd column: kw=2 of v, kw=0 of v+1
                        const float2 a = make_float2(in[v], in[v]), bnext = make_float2(in[v + 1], in[v + 1]);
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) {
                            acc2[c][2 * v] = __ffma2_rn(wk[1][c], a, acc2[c][2 * v]);
                            acc2[c][2 * v + 1] = __ffma2_rn(wk[2][c], a, acc2[c][2 * v + 1]);
                            acc2[c][2 * v + 1] = __ffma2_rn(wk[0][c], bnext, acc2[c][2 * v + 1]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    float acc[CPT][NO];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int v = 0; v < NO; ++v) acc[c][v] = (c & 1) ? acc2[c >> 1][v].y : acc2[c >> 1][v].x;
    const int Do = 2 * D, Ho = 2 * Hc, Wo = 2 * W;  // Hc: input rows that produce output (row bands: H - 1 halo row)
    const int od = 2 * id + pd, oh = 2 * (ih0 + th) + ph, ow = 2 * (iw0 + qx * kVPT);
    const size_t out_plane = (size_t)Ho * Wo;
    double s[CPT], ss[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        s[c] = 0.0;
        ss[c] = 0.0;
    }
    if (oh < Ho && ow < Wo) {
        const bool vec = ((Wo & 3) == 0) && (ow + NO <= Wo);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int co = cg * CPT + c;
            float* py = y + (((size_t)b * COUT + co) * Do + od) * out_plane + (size_t)oh * Wo + ow;
            if (vec) {
                *reinterpret_cast<float4*>(py) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                *reinterpret_cast<float4*>(py + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
            }
#pragma unroll
            for (int v = 0; v < NO; ++v)
                if (ow + v < Wo) {
                    if (!vec) py[v] = acc[c][v];
                    s[c] += (double)acc[c][v];
                    ss[c] = fma((double)acc[c][v], (double)acc[c][v], ss[c]);
                }
        }
    }
    if (gn_sums != nullptr) gn_epilogue<COUT, CPT>(s, ss, cg, smem, gn_sums, b);
}

// weight packing: conv [Cout][Cin][27] -> [Cin][27][Cout];  deconv [Cin][Cout][27] -> [Cin][27][Cout]
__global__ void pack_conv3d_weight_kernel(const float* __restrict__ w, float* __restrict__ p, int Cout, int Cin,
                                          int transposed) {
    const int n = Cout * Cin * 27;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int co = i % Cout;
        const int t = (i / Cout) % 27;
        const int ci = i / (Cout * 27);
        const size_t src = transposed ? ((size_t)ci * Cout + co) * 27 + t : ((size_t)co * Cin + ci) * 27 + t;
        p[i] = w[src];
    }
}

template <int COUT, int CPT, int S, int CC, int TW>
static int launch_conv(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int D, int H, int W,
                       cudaStream_t st, RowWin rw = RowWin()) {
    using T = ConvTile<COUT, CPT, TW>;
    constexpr int PD = (T::TD - 1) * S + 3, PH = (T::TH - 1) * S + 3, PW = (TW - 1) * S + 3;
    constexpr int PWP = (PW + 3) & ~3;
    constexpr size_t smem = 2 * (size_t)CC * (PD * PH * PWP + 27 * COUT) * sizeof(float);
    static_assert(smem <= 110 * 1024, "two CTAs per SM must fit");
    const int Do = (D - 1) / S + 1, Ho = rw.Ho >= 0 ? rw.Ho : (H - 1) / S + 1, Wo = (W - 1) / S + 1;
    const int tiles_w = (int)cdiv(Wo, TW), tiles_h = (int)cdiv(Ho, T::TH);
    auto kern = conv3d_k3_kernel<COUT, CPT, S, CC, TW>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)cdiv(Do, T::TD), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d: grid too large");
    kern<<<grid, kConvThreads, smem, st>>>(x, wp, y, gn, Cin, D, H, W, Do, Ho, Wo, tiles_w, rw.hoff);
    CMF_LAUNCH_CHECK("conv3d_k3_kernel");
    return CMFB200_OK;
}

// picks the tile width (32 or 16 voxels) that wastes fewer lanes on the ragged right / bottom edge
template <int COUT, int CPT, int S, int CC>
static int launch_conv_best(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int D, int H, int W,
                            cudaStream_t st, RowWin rw = RowWin()) {
    const long long Ho = rw.Ho >= 0 ? rw.Ho : (H - 1) / S + 1, Wo = (W - 1) / S + 1;
    const long long c32 = cdiv(Wo, 32) * 32 * cdiv(Ho, ConvTile<COUT, CPT, 32>::TH) * ConvTile<COUT, CPT, 32>::TH;
    const long long c16 = cdiv(Wo, 16) * 16 * cdiv(Ho, ConvTile<COUT, CPT, 16>::TH) * ConvTile<COUT, CPT, 16>::TH;
    if (c16 < c32) return launch_conv<COUT, CPT, S, CC, 16>(x, wp, y, gn, B, Cin, D, H, W, st, rw);
    return launch_conv<COUT, CPT, S, CC, 32>(x, wp, y, gn, B, Cin, D, H, W, st, rw);
}

template <int COUT, int CPT, int CC>
static int launch_deconv(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int D, int H, int W,
                         cudaStream_t st, int Hc = -1) {
    using T = ConvTile<COUT, CPT>;
    constexpr size_t smem = 2 * (size_t)CC * (2 * (T::TH + 1) * 36 + 12 * COUT) * sizeof(float);
    static_assert(smem <= 110 * 1024, "two CTAs per SM must fit");
    if (Hc < 0) Hc = H;
    const int tiles_w = (int)cdiv(W, kTW), tiles_h = (int)cdiv(Hc, T::TH);
    auto kern = deconv3d_k3s2_kernel<COUT, CPT, CC>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(D * 4), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "deconv3d: grid too large");
    kern<<<grid, kConvThreads, smem, st>>>(x, wp, y, gn, Cin, D, H, W, tiles_w, Hc);
    CMF_LAUNCH_CHECK("deconv3d_k3s2_kernel");
    return CMFB200_OK;
}

int conv3d_cout1_fp32_dispatch(const float* x, const float* wp, float* y, int B, int Cin, int D, int H_in, int W, int hoff,
                               int H_out, cudaStream_t st);  // conv3d_cout1_fp32.cu

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_pack_conv3d_weight(const float* weight, float* packed, int Cout, int Cin, int transposed,
                                          void* stream) {
    CMF_REQUIRE(weight && packed, "pack_conv3d_weight: null pointer");
    CMF_REQUIRE(Cout > 0 && Cin > 0, "pack_conv3d_weight: non-positive dimension");
    const int n = Cout * Cin * 27;
    pack_conv3d_weight_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(weight, packed, Cout, Cin,
                                                                                          transposed);
    CMF_LAUNCH_CHECK("pack_conv3d_weight_kernel");
    return CMFB200_OK;
}

static int conv3d_k3_dispatch(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin, int Cout,
                             int D, int H, int W, int stride, RowWin rw, cudaStream_t st) {
    if (Cout == 32 && stride == 1) return launch_conv_r2_best<32, 4>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st, rw);
    if (Cout == 64 && stride == 1) return launch_conv_r2_best<64, 4>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st, rw);
    if (Cout == 32 && stride == 2) return launch_conv<32, 8, 2, 2, 32>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st, rw);
    if (Cout == 64 && stride == 2) return launch_conv<64, 8, 2, 2, 32>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st, rw);
    if (Cout == 1 && stride == 1) {
        if (gn_sums == nullptr) {  // classifier tail: TMA-staged 4x4-outputs-per-thread kernel (also on a row window)
            const int rc = rw.Ho < 0 ? conv3d_cout1_fp32_dispatch(x, packed_w, y, B, Cin, D, H, W, 0, H, st)
                                     : conv3d_cout1_fp32_dispatch(x, packed_w, y, B, Cin, D, H, W, rw.hoff, rw.Ho, st);
            if (rc >= 0) return rc;
        }
        return launch_conv_best<1, 1, 1, 4>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st, rw);
    }
    CMF_REQUIRE(false, "conv3d_k3_fwd: unsupported (Cout=%d, stride=%d); Cout in {1,32,64}", Cout, stride);
}

extern "C" int cmfb200_conv3d_k3_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin,
                                     int Cout, int D, int H, int W, int stride, void* stream) {
    CMF_REQUIRE(x && packed_w && y, "conv3d_k3_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0, "conv3d_k3_fwd: non-positive dimension");
    CMF_REQUIRE(Cin % 8 == 0, "conv3d_k3_fwd: Cin=%d must be a multiple of 8", Cin);
    CMF_REQUIRE(stride == 1 || stride == 2, "conv3d_k3_fwd: stride=%d not in {1,2}", stride);
    return conv3d_k3_dispatch(x, packed_w, y, gn_sums, B, Cin, Cout, D, H, W, stride, RowWin(), (cudaStream_t)stream);
}

extern "C" int cmfb200_conv3d_k3_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B,
                                          int Cin, int Cout, int D, int H_in, int W, int stride, int h_offset, int H_out,
                                          void* stream) {
    CMF_REQUIRE(x && packed_w && y, "conv3d_k3_rows_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && D > 0 && H_in > 0 && W > 0 && H_out > 0, "conv3d_k3_rows_fwd: non-positive dimension");
    CMF_REQUIRE(Cin % 8 == 0, "conv3d_k3_rows_fwd: Cin=%d must be a multiple of 8", Cin);
    CMF_REQUIRE(stride == 1 || stride == 2, "conv3d_k3_rows_fwd: stride=%d not in {1,2}", stride);
    CMF_REQUIRE(h_offset >= 0 && (H_out - 1) * stride + h_offset - 1 < H_in + 1,
                "conv3d_k3_rows_fwd: row window (offset %d, %d rows, stride %d) leaves the %d input rows", h_offset, H_out,
                stride, H_in);
    RowWin rw;
    rw.hoff = h_offset;
    rw.Ho = H_out;
    return conv3d_k3_dispatch(x, packed_w, y, gn_sums, B, Cin, Cout, D, H_in, W, stride, rw, (cudaStream_t)stream);
}

extern "C" int cmfb200_deconv3d_k3s2_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B,
                                         int Cin, int Cout, int D, int H, int W, void* stream) {
    CMF_REQUIRE(x && packed_w && y, "deconv3d_k3s2_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0, "deconv3d_k3s2_fwd: non-positive dimension");
    CMF_REQUIRE(Cin % 8 == 0, "deconv3d_k3s2_fwd: Cin=%d must be a multiple of 8", Cin);
    CMF_REQUIRE(D * 4 <= 65535, "deconv3d_k3s2_fwd: depth too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 32) return launch_deconv<32, 8, 8>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st);
    if (Cout == 64) return launch_deconv<64, 8, 8>(x, packed_w, y, gn_sums, B, Cin, D, H, W, st);
    CMF_REQUIRE(false, "deconv3d_k3s2_fwd: unsupported Cout=%d (32 or 64)", Cout);
}

extern "C" int cmfb200_deconv3d_k3s2_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B,
                                              int Cin, int Cout, int D, int H_in, int H_compute, int W, void* stream) {
    CMF_REQUIRE(x && packed_w && y, "deconv3d_k3s2_rows_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && D > 0 && H_in > 0 && W > 0, "deconv3d_k3s2_rows_fwd: non-positive dimension");
    CMF_REQUIRE(Cin % 8 == 0, "deconv3d_k3s2_rows_fwd: Cin=%d must be a multiple of 8", Cin);
    CMF_REQUIRE(H_compute > 0 && H_compute <= H_in, "deconv3d_k3s2_rows_fwd: H_compute=%d outside (0, %d]", H_compute, H_in);
    CMF_REQUIRE(D * 4 <= 65535, "deconv3d_k3s2_rows_fwd: depth too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 32) return launch_deconv<32, 8, 8>(x, packed_w, y, gn_sums, B, Cin, D, H_in, W, st, H_compute);
    if (Cout == 64) return launch_deconv<64, 8, 8>(x, packed_w, y, gn_sums, B, Cin, D, H_in, W, st, H_compute);
    CMF_REQUIRE(false, "deconv3d_k3s2_rows_fwd: unsupported Cout=%d (32 or 64)", Cout);
}
