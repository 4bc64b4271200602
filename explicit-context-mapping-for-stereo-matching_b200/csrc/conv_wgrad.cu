// Weight gradient of every convolution of the network (training backward, train.py:166-181 -> autograd of
// cmf/models/cmfsm.py: convbn / convbn_3d / hourglass / feature_extraction):
//
//     dW[co][ci][kd][kh][kw] = sum_{b, o}  dy[b][co][o] * x[b][ci][ o*stride - pad + k*dilation ]      (x zero-padded)
//
// for nn.Conv2d (KD = 1; 3x3 stride 1 dilation 1/2/4, 3x3 stride 2, 1x1 stride 1/2) and nn.Conv3d (3x3x3 stride 1/2).
// The transposed conv (y = deconv(x, Wt[ci][co][k]), o' = 2i - 1 + k) is the same sum with the roles swapped:
// dWt[ci][co][k] = sum_i x[ci][i] * dy[co][2i - 1 + k], i.e. this kernel with (dy := x, x := dy, stride 2) -- the
// result is laid out as the ConvTranspose3d weight.  Replaces aten::convolution_backward (cuDNN, TF32 by default) with
// strict fp32 FMAs.
//
// A voxel-reduction GEMM: M x N = Cout x (Cin * taps) is small, K = all output positions is huge.  Register tiling on the
// fp32 pipe: a CTA owns a 32 (co) x 32 (ci) x taps block of dW and a contiguous range of (b, d, h, 32-wide w segment)
// "stages"; warp = 4 output channels, lane = input channel, so a thread keeps 4 x taps accumulators (108 for a 3x3x3
// kernel) and per group of 4 output positions does 4*4*taps FMAs for 16 broadcast dy loads + taps/3 * 6..9 x loads
// (7:1 FMA per shared-memory load).  Stages are double-buffered with cp.async (zero fill = the conv padding and the
// ragged tails).  Partial blocks are added to dW with fp32 atomics (dW must be zero on entry).
#include "common.cuh"
#include "conv_common.cuh"

namespace cmfb200 {

namespace {

constexpr int kWgThreads = 256;
constexpr int kWgPitch = 36;  // floats per staged input row: bank-conflict-free for lane = channel (9*36 = 4 mod 32)

template <int KD, int KHW, int STRIDE, int DIL>
struct WgCfg {
    static constexpr int TAPS = KD * KHW * KHW;
    static constexpr int ROWS = KD * KHW;                        // staged input rows per channel
    static constexpr int P = STRIDE == 1 ? 32 : 16;              // output positions per stage
    static constexpr int SPAN = (P - 1) * STRIDE + (KHW - 1) * DIL + 1;  // input columns a stage touches
    static constexpr int X_FLOATS = 32 * ROWS * kWgPitch;
    static constexpr int DY_FLOATS = 32 * P;
    static constexpr int STAGE_FLOATS = X_FLOATS + DY_FLOATS;
    static constexpr int SMEM_BYTES = 2 * STAGE_FLOATS * 4;
    static_assert(SPAN <= kWgPitch, "stage does not fit the row pitch");
};

struct WgDims {
    int B, Cin, Cout, D, H, W, Do, Ho, Wo, segs;  // segs = ceil(Wo / P)
    long long stages;                             // B * Do * Ho * segs
};

}  // namespace

template <int KD, int KHW, int STRIDE, int DIL>
__global__ void __launch_bounds__(kWgThreads, 1)
    conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, const WgDims dm) {
    using G = WgCfg<KD, KHW, STRIDE, DIL>;
    extern __shared__ float smem_wg[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tiles = (dm.Cin + 31) / 32;
    const int co0 = (blockIdx.y / ci_tiles) * 32, ci0 = (blockIdx.y % ci_tiles) * 32;
    const int pad_hw = (KHW / 2) * DIL, pad_d = KD / 2;
    const long long s_beg = dm.stages * blockIdx.x / gridDim.x, s_end = dm.stages * (blockIdx.x + 1) / gridDim.x;
    const size_t in_plane = (size_t)dm.H * dm.W, out_plane = (size_t)dm.Ho * dm.Wo;

    auto stage_load = [&](long long s, int buf) {
        float* sx = smem_wg + buf * G::STAGE_FLOATS;
        float* sdy = sx + G::X_FLOATS;
        long long r = s;
        const int seg = (int)(r % dm.segs);
        r /= dm.segs;
        const int ho = (int)(r % dm.Ho);
        r /= dm.Ho;
        const int dz = (int)(r % dm.Do);
        const int b = (int)(r / dm.Do);
        const int wo0 = seg * G::P;
        const int wi0 = wo0 * STRIDE - pad_hw;
        // input patch: [32 ci][KD*KHW rows][SPAN] (4-byte copies: the row start is not 16-byte aligned in general)
        for (int i = threadIdx.x; i < 32 * G::ROWS * G::SPAN; i += kWgThreads) {
            const int col = i % G::SPAN;
            const int row = (i / G::SPAN) % G::ROWS;
            const int c = i / (G::SPAN * G::ROWS);
            const int kd = row / KHW, kh = row - kd * KHW;
            const int di = dz * (KD == 3 ? STRIDE : 1) - pad_d + kd;
            const int hi = ho * STRIDE - pad_hw + kh * DIL;
            const int wi = wi0 + col;
            const bool ok = (ci0 + c < dm.Cin) && di >= 0 && di < dm.D && hi >= 0 && hi < dm.H && wi >= 0 && wi < dm.W;
            const float* src = ok ? x + (((size_t)b * dm.Cin + ci0 + c) * dm.D + di) * in_plane + (size_t)hi * dm.W + wi : x;
            cp_async_4_zfill(sx + (c * G::ROWS + row) * kWgPitch + col, src, ok);
        }
        // dy segment: [32 co][P]
        for (int i = threadIdx.x; i < 32 * G::P; i += kWgThreads) {
            const int p = i % G::P, c = i / G::P;
            const bool ok = wo0 + p < dm.Wo;
            const float* src = ok ? dy + (((size_t)b * dm.Cout + co0 + c) * dm.Do + dz) * out_plane + (size_t)ho * dm.Wo + wo0 + p
                                  : dy;
            cp_async_4_zfill(sdy + c * G::P + p, src, ok);
        }
        cp_async_commit();
    };

    float acc[4][G::TAPS];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < G::TAPS; ++t) acc[j][t] = 0.f;

    if (s_beg < s_end) stage_load(s_beg, 0);
    for (long long s = s_beg; s < s_end; ++s) {
        const int buf = (int)((s - s_beg) & 1);
        if (s + 1 < s_end) {
            stage_load(s + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* sx = smem_wg + buf * G::STAGE_FLOATS + lane * G::ROWS * kWgPitch;
        const float* sdy = smem_wg + buf * G::STAGE_FLOATS + G::X_FLOATS + warp * 4 * G::P;
#pragma unroll 1
        for (int p0 = 0; p0 < G::P; p0 += 4) {
            float g[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(sdy + j * G::P + p0);  // warp-wide broadcast
                g[j][0] = v.x, g[j][1] = v.y, g[j][2] = v.z, g[j][3] = v.w;
            }
            constexpr int NX = 3 * STRIDE + (KHW - 1) * DIL + 1;  // input columns of 4 outputs
            constexpr int NX4 = (NX + 3) / 4;
#pragma unroll
            for (int row = 0; row < G::ROWS; ++row) {
                float xv[NX4 * 4];
#pragma unroll
                for (int q = 0; q < NX4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(sx + row * kWgPitch + p0 * STRIDE + 4 * q);
                    xv[4 * q] = v.x, xv[4 * q + 1] = v.y, xv[4 * q + 2] = v.z, xv[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int kw = 0; kw < KHW; ++kw)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            acc[j][row * KHW + kw] = fmaf(g[j][p], xv[p * STRIDE + kw * DIL], acc[j][row * KHW + kw]);
            }
        }
        __syncthreads();
    }
    if (ci0 + lane < dm.Cin) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* dst = dw + ((size_t)(co0 + warp * 4 + j) * dm.Cin + ci0 + lane) * G::TAPS;
#pragma unroll
            for (int t = 0; t < G::TAPS; ++t) atomicAdd(dst + t, acc[j][t]);
        }
    }
}

template <int KD, int KHW, int STRIDE, int DIL>
static int launch_wgrad(const float* x, const float* dy, float* dw, WgDims dm, cudaStream_t st) {
    using G = WgCfg<KD, KHW, STRIDE, DIL>;
    dm.segs = (int)cdiv(dm.Wo, G::P);
    dm.stages = (long long)dm.B * dm.Do * dm.Ho * dm.segs;
    auto kern = conv_wgrad_kernel<KD, KHW, STRIDE, DIL>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = (dm.Cout / 32) * (int)cdiv(dm.Cin, 32);
    long long gx = (2LL * sms + tiles - 1) / tiles;  // ~2 waves of one-CTA-per-SM blocks over all (co, ci) tiles
    if (gx > dm.stages) gx = dm.stages;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)tiles);
    kern<<<grid, kWgThreads, G::SMEM_BYTES, st>>>(x, dy, dw, dm);
    CMF_LAUNCH_CHECK("conv_wgrad_kernel");
    return CMFB200_OK;
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_conv_wgrad(const float* x, const float* dy, float* dw, int B, int Cin, int Cout, int D, int H, int W,
                                  int KD, int KHW, int stride, int dilation, void* stream) {
    CMF_REQUIRE(x && dy && dw, "conv_wgrad: null pointer");
    CMF_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "conv_wgrad: non-positive dimension");
    CMF_REQUIRE(Cout % 32 == 0, "conv_wgrad: Cout=%d must be a multiple of 32", Cout);
    CMF_REQUIRE(KD == 1 || KD == 3, "conv_wgrad: KD must be 1 or 3");
    CMF_REQUIRE(KD == 3 || D == 1, "conv_wgrad: a 2-D convolution (KD=1) takes D=1");
    WgDims dm;
    dm.B = B, dm.Cin = Cin, dm.Cout = Cout, dm.D = D, dm.H = H, dm.W = W;
    dm.Do = KD == 3 ? (D - 1) / stride + 1 : 1;
    dm.Ho = (H - 1) / stride + 1;
    dm.Wo = (W - 1) / stride + 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (KD == 3 && KHW == 3 && stride == 1 && dilation == 1) return launch_wgrad<3, 3, 1, 1>(x, dy, dw, dm, st);
    if (KD == 3 && KHW == 3 && stride == 2 && dilation == 1) return launch_wgrad<3, 3, 2, 1>(x, dy, dw, dm, st);
    if (KD == 1 && KHW == 3 && stride == 1 && dilation == 1) return launch_wgrad<1, 3, 1, 1>(x, dy, dw, dm, st);
    if (KD == 1 && KHW == 3 && stride == 1 && dilation == 2) return launch_wgrad<1, 3, 1, 2>(x, dy, dw, dm, st);
    if (KD == 1 && KHW == 3 && stride == 2 && dilation == 1) return launch_wgrad<1, 3, 2, 1>(x, dy, dw, dm, st);
    if (KD == 1 && KHW == 1 && stride == 1) return launch_wgrad<1, 1, 1, 1>(x, dy, dw, dm, st);
    if (KD == 1 && KHW == 1 && stride == 2) return launch_wgrad<1, 1, 2, 1>(x, dy, dw, dm, st);
    CMF_REQUIRE(false, "conv_wgrad: unsupported (KD=%d, k=%d, stride=%d, dilation=%d)", KD, KHW, stride, dilation);
}
