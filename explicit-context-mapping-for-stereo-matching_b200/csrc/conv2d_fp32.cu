// 2-D convolutions of the feature extractor (reference feature_extraction / BasicBlock / convbn,
// cmf/models/cmfsm.py:36-46, 61-85, 126-236) on the CUDA cores with exact fp32 FMA accumulation, NCHW.
// cuDNN's strict-fp32 path on sm_100 falls back to FFT / complex-GEMM kernels (82.7 ms for both images at
// 576x960, profiles/r01_probe_features_cudnn.txt) and TF32 operand rounding breaks fp32 parity in exactly
// this part of the network, hence a hand-written direct convolution.
//
// Same register tiling as the 3-D kernel: a CTA owns TH x 32 output pixels x COUT_TILE output channels
// (grid.y walks channel tiles), a thread owns 8 channels x 4 consecutive-x pixels.  Operands are staged with
// cp.async into a 2-stage ring of CC-input-channel chunks (prefetch chunk i+1 while the FMA loop runs on
// chunk i); the per-thread staging slots (global offset, shared offset, in-image flag) are computed once.
// GroupNorm statistics are reduced in the epilogue.
#include "common.cuh"
#include "conv_common.cuh"

namespace cmfb200 {

template <int KS, int S, int DIL, int COUT_TILE>
struct Conv2dCfg {
    static constexpr int CPT = 8;
    static constexpr int NCG = COUT_TILE / CPT;
    static constexpr int NQ = kConvThreads / NCG;
    static constexpr int TH = NQ / (kTW / kVPT);
    static constexpr int PH = (TH - 1) * S + (KS - 1) * DIL + 1;
    static constexpr int PW = (kTW - 1) * S + (KS - 1) * DIL + 1;
    static constexpr int PWP = (PW + 3) & ~3;
    static constexpr int PATCH = PH * PWP;
    static constexpr int NI = (kVPT - 1) * S + (KS - 1) * DIL + 1;
    static constexpr int NI4 = (NI + 3) / 4;
    static constexpr int NSLOT = (PH * PW + kConvThreads - 1) / kConvThreads;
    static constexpr int WSL = KS * KS * COUT_TILE;  // staged weight floats per input channel
    static_assert((kTW / kVPT - 1) * kVPT * S + NI4 * 4 <= PWP, "vector over-read leaves the patch row");
    static_assert(NQ % 32 == 0, "channel group must be warp-uniform");
};

template <int KS, int S, int DIL, int COUT_TILE, int CC>
__global__ void __launch_bounds__(kConvThreads, 2)
    conv2d_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y,
                  double* __restrict__ gn_sums, int Cin, int Cout, int H, int W, int Ho, int Wo, int tiles_w, int hoff) {
    using G = Conv2dCfg<KS, S, DIL, COUT_TILE>;
    constexpr int CPT = G::CPT;
    constexpr int STAGE = CC * (G::PATCH + G::WSL);  // floats per pipeline stage

    extern __shared__ __align__(16) float smem[];

    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * kTW, h0 = tile_y * G::TH;
    const int cb = blockIdx.y * COUT_TILE;  // first output channel of this CTA
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid / G::NQ;
    const int q = tid % G::NQ;
    const int qx = q % (kTW / kVPT);
    const int th = q / (kTW / kVPT);
    constexpr int PAD = (KS / 2) * DIL;

    // ---- staging slots of this thread (same for every input channel)
    int goff[G::NSLOT], soff[G::NSLOT];
    bool ok[G::NSLOT];
    const int hi0 = h0 * S - PAD + hoff, wi0 = w0 * S - PAD;  // hoff: row window (row-band sharding)
#pragma unroll
    for (int j = 0; j < G::NSLOT; ++j) {
        const int e = tid + j * kConvThreads;
        const int ph = e / G::PW, pw = e - ph * G::PW;
        const int hi = hi0 + ph, wi = wi0 + pw;
        const bool in_patch = e < G::PH * G::PW;
        ok[j] = in_patch && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        goff[j] = ok[j] ? hi * W + wi : 0;
        soff[j] = in_patch ? ph * G::PWP + pw : -1;
    }
    const size_t in_plane = (size_t)H * W;
    const float* xb = x + (size_t)b * Cin * in_plane;

    auto stage = [&](int c0, int buf) {
        float* sIn = smem + buf * STAGE;
        float* sW = sIn + CC * G::PATCH;
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
            const float* src = xb + (size_t)(c0 + ci) * in_plane;
#pragma unroll
            for (int j = 0; j < G::NSLOT; ++j)
                if (soff[j] >= 0) cp_async_4_zfill(sIn + ci * G::PATCH + soff[j], src + goff[j], ok[j]);
        }
        // weights: rows of COUT_TILE floats out of wp[Cin][KS*KS][Cout]
        constexpr int ROW4 = COUT_TILE / 4;
        for (int i = tid; i < CC * KS * KS * ROW4; i += kConvThreads) {
            const int r = i / ROW4, j4 = i - r * ROW4;  // r = ci*KS*KS + tap
            cp_async_16(sW + r * COUT_TILE + j4 * 4, wp + ((size_t)c0 * KS * KS + r) * Cout + cb + j4 * 4);
        }
        cp_async_commit();
    };

    float2 acc2[CPT / 2][kVPT];  // channel pairs, updated with FFMA2 (packed fp32x2 FMA)
#pragma unroll
    for (int c = 0; c < CPT / 2; ++c)
#pragma unroll
        for (int v = 0; v < kVPT; ++v) acc2[c][v] = make_float2(0.f, 0.f);

    // zero the alignment tail of every patch row once (never written by cp.async, read by vector loads)
    if constexpr (G::PWP > G::PW) {
        for (int i = tid; i < 2 * CC * G::PH * (G::PWP - G::PW); i += kConvThreads) {
            const int t = i % (G::PWP - G::PW);
            const int r = i / (G::PWP - G::PW);  // (buf, ci, ph) flattened
            const int ph = r % G::PH, ci = (r / G::PH) % CC, buf = r / (G::PH * CC);
            smem[buf * STAGE + ci * G::PATCH + ph * G::PWP + G::PW + t] = 0.f;
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
        const float* sW = sIn + CC * G::PATCH;
#pragma unroll 1
        for (int ci = 0; ci < CC; ++ci) {
            const float* pin = sIn + ci * G::PATCH + th * S * G::PWP + qx * kVPT * S;
            const float* pwt = sW + ci * G::WSL + cg * CPT;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh) {
                const float* prow = pin + kh * DIL * G::PWP;
                float in[G::NI4 * 4];
#pragma unroll
                for (int j = 0; j < G::NI4; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(prow + 4 * j);
                    in[4 * j + 0] = a.x; in[4 * j + 1] = a.y; in[4 * j + 2] = a.z; in[4 * j + 3] = a.w;
                }
#pragma unroll
                for (int kw = 0; kw < KS; ++kw) {
                    const float* wt = pwt + (kh * KS + kw) * COUT_TILE;
                    const float4 w0v = *reinterpret_cast<const float4*>(wt);
                    const float4 w1v = *reinterpret_cast<const float4*>(wt + 4);
                    const float2 w2[CPT / 2] = {make_float2(w0v.x, w0v.y), make_float2(w0v.z, w0v.w),
                                                make_float2(w1v.x, w1v.y), make_float2(w1v.z, w1v.w)};
#pragma unroll
                    for (int v = 0; v < kVPT; ++v) {
                        const float2 i2 = make_float2(in[v * S + kw * DIL], in[v * S + kw * DIL]);
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) acc2[c][v] = __ffma2_rn(w2[c], i2, acc2[c][v]);
                    }
                }
            }
        }
        __syncthreads();  // everyone done with `buf` before the next iteration's prefetch overwrites it
    }

    // ---- epilogue
    float acc[CPT][kVPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int v = 0; v < kVPT; ++v) acc[c][v] = (c & 1) ? acc2[c >> 1][v].y : acc2[c >> 1][v].x;
    const int oh = h0 + th, ow = w0 + qx * kVPT;
    const size_t out_plane = (size_t)Ho * Wo;
    double s[CPT], ss[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        s[c] = 0.0;
        ss[c] = 0.0;
    }
    if (oh < Ho && ow < Wo) {
        const bool vec = ((Wo & 3) == 0);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int co = cb + cg * CPT + c;
            float* py = y + ((size_t)b * Cout + co) * out_plane + (size_t)oh * Wo + ow;
            if (vec) *reinterpret_cast<float4*>(py) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
#pragma unroll
            for (int v = 0; v < kVPT; ++v)
                if (vec || ow + v < Wo) {
                    if (!vec) py[v] = acc[c][v];
                    s[c] += (double)acc[c][v];
                    ss[c] = fma((double)acc[c][v], (double)acc[c][v], ss[c]);
                }
        }
    }
    if (gn_sums != nullptr) gn_epilogue<COUT_TILE, CPT>(s, ss, cg, smem, gn_sums, b, Cout, cb);
}

// ------------------------------------------------------------------------------------------------
// 3x3 stride-1 (dilation 1 or 2) with TWO output rows per thread: rows r and r+DIL share the four input rows
// r + {0,1,2,3}*DIL, so per input channel a thread issues 8 vector loads of input + 18 broadcast loads of weights
// for 576 FMAs (1:22 instead of 1:12) -- these layers carry 80 % of the feature extractor's MACs.
// ------------------------------------------------------------------------------------------------
template <int DIL, int COUT_TILE, int TW>
struct Conv2dR2Cfg {
    static constexpr int CPT = 8;
    static constexpr int NCG = COUT_TILE / CPT;
    static constexpr int NQ = kConvThreads / NCG;
    static constexpr int TROWS = NQ / (TW / kVPT);  // thread rows
    static constexpr int TH = 2 * TROWS;             // output rows per CTA
    static constexpr int PH = TH + 2 * DIL;
    static constexpr int PW = TW + 2 * DIL;
    // TW=16, DIL=1: pitch 24 puts the two rows of a load phase (2 rows apart) 16 banks apart -> conflict-free LDS.128
    static constexpr int PWP = (TW == 16 && DIL == 1) ? 24 : ((PW + 3) & ~3);
    static constexpr int PATCH = PH * PWP;
    static constexpr int NI4 = (kVPT + 2 * DIL + 3) / 4;  // 2
    static constexpr int NSLOT = (PH * PW + kConvThreads - 1) / kConvThreads;
    static constexpr int WSL = 9 * COUT_TILE;
    static_assert(TROWS % DIL == 0, "row pairing needs TROWS to be a multiple of the dilation");
    static_assert((TW / kVPT - 1) * kVPT + NI4 * 4 <= PWP, "vector over-read leaves the patch row");
};

template <int DIL, int COUT_TILE, int CC, int TW>
__global__ void __launch_bounds__(kConvThreads, 2)
    conv2d_r2_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y,
                     double* __restrict__ gn_sums, int Cin, int Cout, int H, int W, int tiles_w, int hoff, int Ho) {
    using G = Conv2dR2Cfg<DIL, COUT_TILE, TW>;
    constexpr int CPT = G::CPT;
    constexpr int STAGE = CC * (G::PATCH + G::WSL);
    extern __shared__ __align__(16) float smem[];

    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * TW, h0 = tile_y * G::TH;
    const int cb = blockIdx.y * COUT_TILE;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid / G::NQ;
    const int q = tid % G::NQ;
    const int qx = q % (TW / kVPT);
    const int tr = q / (TW / kVPT);
    const int r0 = (tr / DIL) * 2 * DIL + (tr % DIL);  // first output row of this thread (tile-relative); second = r0+DIL

    int goff[G::NSLOT], soff[G::NSLOT];
    bool ok[G::NSLOT];
    const int hi0 = h0 - DIL + hoff, wi0 = w0 - DIL;  // hoff / Ho: row window (row-band sharding)
#pragma unroll
    for (int j = 0; j < G::NSLOT; ++j) {
        const int e = tid + j * kConvThreads;
        const int ph = e / G::PW, pw = e - ph * G::PW;
        const int hi = hi0 + ph, wi = wi0 + pw;
        const bool in_patch = e < G::PH * G::PW;
        ok[j] = in_patch && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        goff[j] = ok[j] ? hi * W + wi : 0;
        soff[j] = in_patch ? ph * G::PWP + pw : -1;
    }
    const size_t in_plane = (size_t)H * W;
    const float* xb = x + (size_t)b * Cin * in_plane;

    auto stage = [&](int c0, int buf) {
        float* sIn = smem + buf * STAGE;
        float* sW = sIn + CC * G::PATCH;
#pragma unroll
        for (int ci = 0; ci < CC; ++ci) {
            const float* src = xb + (size_t)(c0 + ci) * in_plane;
#pragma unroll
            for (int j = 0; j < G::NSLOT; ++j)
                if (soff[j] >= 0) cp_async_4_zfill(sIn + ci * G::PATCH + soff[j], src + goff[j], ok[j]);
        }
        constexpr int ROW4 = COUT_TILE / 4;
        for (int i = tid; i < CC * 9 * ROW4; i += kConvThreads) {
            const int r = i / ROW4, j4 = i - r * ROW4;
            cp_async_16(sW + r * COUT_TILE + j4 * 4, wp + ((size_t)c0 * 9 + r) * Cout + cb + j4 * 4);
        }
        cp_async_commit();
    };

    // channel-pair accumulators updated with FFMA2 (packed fp32x2 FMA, bit-identical to two FFMAs)
    float2 acc2[2][CPT / 2][kVPT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < CPT / 2; ++c)
#pragma unroll
            for (int v = 0; v < kVPT; ++v) acc2[r][c][v] = make_float2(0.f, 0.f);

    if constexpr (G::PWP > G::PW) {
        for (int i = tid; i < 2 * CC * G::PH * (G::PWP - G::PW); i += kConvThreads) {
            const int t = i % (G::PWP - G::PW);
            const int r = i / (G::PWP - G::PW);
            const int ph = r % G::PH, ci = (r / G::PH) % CC, buf = r / (G::PH * CC);
            smem[buf * STAGE + ci * G::PATCH + ph * G::PWP + G::PW + t] = 0.f;
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
        const float* sW = sIn + CC * G::PATCH;
#pragma unroll 1
        for (int ci = 0; ci < CC; ++ci) {
            const float* pin = sIn + ci * G::PATCH + r0 * G::PWP + qx * kVPT;
            const float* pwt = sW + ci * G::WSL + cg * CPT;
            float in[4][G::NI4 * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < G::NI4; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(pin + j * DIL * G::PWP + 4 * k);
                    in[j][4 * k + 0] = a.x; in[j][4 * k + 1] = a.y; in[j][4 * k + 2] = a.z; in[j][4 * k + 3] = a.w;
                }
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float* wt = pwt + (kh * 3 + kw) * COUT_TILE;
                    const float4 w0v = *reinterpret_cast<const float4*>(wt);
                    const float4 w1v = *reinterpret_cast<const float4*>(wt + 4);
                    const float2 w2[CPT / 2] = {make_float2(w0v.x, w0v.y), make_float2(w0v.z, w0v.w),
                                                make_float2(w1v.x, w1v.y), make_float2(w1v.z, w1v.w)};
#pragma unroll
                    for (int v = 0; v < kVPT; ++v) {
                        const float2 i0 = make_float2(in[kh][v + kw * DIL], in[kh][v + kw * DIL]);
                        const float2 i1 = make_float2(in[kh + 1][v + kw * DIL], in[kh + 1][v + kw * DIL]);
#pragma unroll
                        for (int c = 0; c < CPT / 2; ++c) {
                            acc2[0][c][v] = __ffma2_rn(w2[c], i0, acc2[0][c][v]);
                            acc2[1][c][v] = __ffma2_rn(w2[c], i1, acc2[1][c][v]);
                        }
                    }
                }
        }
        __syncthreads();
    }

    float acc[2][CPT][kVPT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
            for (int v = 0; v < kVPT; ++v) acc[r][c][v] = (c & 1) ? acc2[r][c >> 1][v].y : acc2[r][c >> 1][v].x;
    const int ow = w0 + qx * kVPT;
    const size_t out_plane = (size_t)Ho * W;  // stride 1, "same" padding along W; Ho rows (row window)
    double s[CPT], ss[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        s[c] = 0.0;
        ss[c] = 0.0;
    }
    const bool vec = ((W & 3) == 0);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oh = h0 + r0 + r * DIL;
        if (oh < Ho && ow < W) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int co = cb + cg * CPT + c;
                float* py = y + ((size_t)b * Cout + co) * out_plane + (size_t)oh * W + ow;
                if (vec) *reinterpret_cast<float4*>(py) = make_float4(acc[r][c][0], acc[r][c][1], acc[r][c][2], acc[r][c][3]);
#pragma unroll
                for (int v = 0; v < kVPT; ++v)
                    if (vec || ow + v < W) {
                        if (!vec) py[v] = acc[r][c][v];
                        s[c] += (double)acc[r][c][v];
                        ss[c] = fma((double)acc[r][c][v], (double)acc[r][c][v], ss[c]);
                    }
            }
        }
    }
    if (gn_sums != nullptr) gn_epilogue<COUT_TILE, CPT>(s, ss, cg, smem, gn_sums, b, Cout, cb);
}

struct RowWin2 {  // row window for row-band sharding (see conv3d_fp32.cu): hoff = 0 / Ho = -1 = ordinary convolution
    int hoff = 0, Ho = -1;
};

template <int DIL, int COUT_TILE, int CC, int TW>
static int launch_conv2d_r2_tw(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int Cout, int H, int W,
                            cudaStream_t st, RowWin2 rw) {
    using G = Conv2dR2Cfg<DIL, COUT_TILE, TW>;
    constexpr size_t smem = 2 * (size_t)CC * (G::PATCH + G::WSL) * sizeof(float);
    static_assert(smem <= 110 * 1024, "two CTAs per SM must fit");
    const int Ho = rw.Ho >= 0 ? rw.Ho : H;
    const int tiles_w = (int)cdiv(W, TW), tiles_h = (int)cdiv(Ho, G::TH);
    auto kern = conv2d_r2_kernel<DIL, COUT_TILE, CC, TW>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(Cout / COUT_TILE), (unsigned)B);
    CMF_REQUIRE(grid.z <= 65535, "conv2d: batch too large");
    kern<<<grid, kConvThreads, smem, st>>>(x, wp, y, gn, Cin, Cout, H, W, tiles_w, rw.hoff, Ho);
    CMF_LAUNCH_CHECK("conv2d_r2_kernel");
    return CMFB200_OK;
}

template <int DIL, int COUT_TILE, int CC>
static int launch_conv2d_r2(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int Cout, int H, int W,
                            cudaStream_t st, RowWin2 rw) {
    // tile width 32 or 16: whichever wastes fewer lanes on the ragged edges (w = 240 = 15 x 16 = 7.5 x 32)
    const long long Hr = rw.Ho >= 0 ? rw.Ho : H;
    const long long c32 = cdiv(W, 32) * 32 * cdiv(Hr, Conv2dR2Cfg<DIL, COUT_TILE, 32>::TH) * Conv2dR2Cfg<DIL, COUT_TILE, 32>::TH;
    const long long c16 = cdiv(W, 16) * 16 * cdiv(Hr, Conv2dR2Cfg<DIL, COUT_TILE, 16>::TH) * Conv2dR2Cfg<DIL, COUT_TILE, 16>::TH;
    // measured at w=240 (6.7 % fewer lanes): 16-wide tiles are 1.4 % SLOWER here (smaller halo reuse, 2-way bank
    // conflicts on the 80-byte row pitch), so they are only used when they save more than 10 %
    if ((DIL == 1 && c16 < c32) || c16 * 10 < c32 * 9)
        return launch_conv2d_r2_tw<DIL, COUT_TILE, CC, 16>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
    return launch_conv2d_r2_tw<DIL, COUT_TILE, CC, 32>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
}

// weight packing: [Cout][Cin][KS*KS] -> [Cin][KS*KS][Cout]
__global__ void pack_conv2d_weight_kernel(const float* __restrict__ w, float* __restrict__ p, int Cout, int Cin,
                                          int taps) {
    const int n = Cout * Cin * taps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int co = i % Cout;
        const int t = (i / Cout) % taps;
        const int ci = i / (Cout * taps);
        p[i] = w[((size_t)co * Cin + ci) * taps + t];
    }
}

template <int KS, int S, int DIL, int COUT_TILE, int CC>
static int launch_conv2d(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int Cout, int H, int W,
                         cudaStream_t st, RowWin2 rw) {
    using G = Conv2dCfg<KS, S, DIL, COUT_TILE>;
    constexpr size_t smem = 2 * (size_t)CC * (G::PATCH + G::WSL) * sizeof(float);
    static_assert(smem <= 110 * 1024, "two CTAs per SM must fit");
    constexpr int PAD = (KS / 2) * DIL;
    const int Ho = rw.Ho >= 0 ? rw.Ho : (H + 2 * PAD - (KS - 1) * DIL - 1) / S + 1;
    const int Wo = (W + 2 * PAD - (KS - 1) * DIL - 1) / S + 1;
    const int tiles_w = (int)cdiv(Wo, kTW), tiles_h = (int)cdiv(Ho, G::TH);
    auto kern = conv2d_kernel<KS, S, DIL, COUT_TILE, CC>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(Cout / COUT_TILE), (unsigned)B);
    CMF_REQUIRE(grid.z <= 65535, "conv2d: batch too large");
    kern<<<grid, kConvThreads, smem, st>>>(x, wp, y, gn, Cin, Cout, H, W, Ho, Wo, tiles_w, rw.hoff);
    CMF_LAUNCH_CHECK("conv2d_kernel");
    return CMFB200_OK;
}

template <int KS, int S, int DIL>
static int dispatch_conv2d(const float* x, const float* wp, float* y, double* gn, int B, int Cin, int Cout, int H,
                           int W, cudaStream_t st, RowWin2 rw) {
    if (Cin == 3) {
        if constexpr (KS == 3 && S == 1 && DIL == 1) {
            if (Cout == 32) return launch_conv2d<3, 1, 1, 32, 3>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
        }
        CMF_REQUIRE(false, "conv2d: Cin=3 only for the 3x3 s1 stem conv with Cout=32");
    }
    CMF_REQUIRE(Cin % 8 == 0, "conv2d: Cin=%d must be 3 or a multiple of 8", Cin);
    if constexpr (KS == 3 && S == 1 && DIL <= 2) {  // the bulk of the MACs: two output rows per thread
        if (Cout == 32) return launch_conv2d_r2<DIL, 32, 8>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
        if (Cout % 64 == 0) return launch_conv2d_r2<DIL, 64, 8>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
    }
    if (Cout == 32) return launch_conv2d<KS, S, DIL, 32, 8>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
    if (Cout % 64 == 0) return launch_conv2d<KS, S, DIL, 64, 8>(x, wp, y, gn, B, Cin, Cout, H, W, st, rw);
    CMF_REQUIRE(false, "conv2d: Cout=%d must be 32 or a multiple of 64", Cout);
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_pack_conv2d_weight(const float* weight, float* packed, int Cout, int Cin, int ksize,
                                          void* stream) {
    CMF_REQUIRE(weight && packed, "pack_conv2d_weight: null pointer");
    CMF_REQUIRE(Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "pack_conv2d_weight: bad shape");
    const int n = Cout * Cin * ksize * ksize;
    pack_conv2d_weight_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(weight, packed, Cout, Cin,
                                                                                          ksize * ksize);
    CMF_LAUNCH_CHECK("pack_conv2d_weight_kernel");
    return CMFB200_OK;
}

static int conv2d_any(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin, int Cout, int H,
                      int W, int ksize, int stride, int dilation, RowWin2 rw, cudaStream_t st) {
    if (ksize == 3 && stride == 1 && dilation == 1) return dispatch_conv2d<3, 1, 1>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    if (ksize == 3 && stride == 2 && dilation == 1) return dispatch_conv2d<3, 2, 1>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    if (ksize == 3 && stride == 1 && dilation == 2) return dispatch_conv2d<3, 1, 2>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    if (ksize == 3 && stride == 1 && dilation == 4) return dispatch_conv2d<3, 1, 4>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    if (ksize == 1 && stride == 1 && dilation == 1) return dispatch_conv2d<1, 1, 1>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    if (ksize == 1 && stride == 2 && dilation == 1) return dispatch_conv2d<1, 2, 1>(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, st, rw);
    CMF_REQUIRE(false, "conv2d_fwd: unsupported (ksize=%d, stride=%d, dilation=%d)", ksize, stride, dilation);
}

extern "C" int cmfb200_conv2d_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin,
                                  int Cout, int H, int W, int ksize, int stride, int dilation, void* stream) {
    CMF_REQUIRE(x && packed_w && y, "conv2d_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "conv2d_fwd: non-positive dimension");
    return conv2d_any(x, packed_w, y, gn_sums, B, Cin, Cout, H, W, ksize, stride, dilation, RowWin2(), (cudaStream_t)stream);
}

extern "C" int cmfb200_conv2d_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin,
                                       int Cout, int H_in, int W, int ksize, int stride, int dilation, int h_offset,
                                       int H_out, void* stream) {
    CMF_REQUIRE(x && packed_w && y, "conv2d_rows_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H_in > 0 && W > 0 && H_out > 0 && h_offset >= 0,
                "conv2d_rows_fwd: bad dimension");
    RowWin2 rw;
    rw.hoff = h_offset;
    rw.Ho = H_out;
    return conv2d_any(x, packed_w, y, gn_sums, B, Cin, Cout, H_in, W, ksize, stride, dilation, rw, (cudaStream_t)stream);
}
