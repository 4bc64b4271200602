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
//
// conv_wgrad_tma_kernel is the same register tiling with the staging moved off the fp32 pipe: when the row pitch of x
// and dy is a multiple of 16 bytes, one elected thread fetches a stage with TMA tile loads (x as a rank-5
// (W, C, H, D, B) view so that the box lands as [kd][kh][ci][36 columns], zero fill outside the tensor = the padding and
// the ragged tails) into a four-deep ring of mbarrier-tracked buffers.  The cp.async kernel spends a third of its issue
// slots on address arithmetic for 4-byte copies; it stays as the path for unaligned row pitches.
#include "common.cuh"
#include "conv_common.cuh"
#include "igemm_common.cuh"

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


constexpr int kWgRing = 4;

// TMA tile loads start on 16-byte boundaries of the innermost dimension (measured: tools/probe_tma_f32.cu -- a start
// column of -1 raises an illegal instruction), and the input window of a stage starts one column left of its outputs.
// So the main box starts at the aligned column wo0*STRIDE - 4 and is 36 wide (pitch 36 = conflict-free float4 reads for
// lane = channel), the four columns a stride-1 3x3 stage needs beyond it come as a second [row][ci][4] box, and the
// threads read whole aligned float4s around the window of 8 output positions at a time.
template <int KD, int KHW, int STRIDE, int DIL>
struct WgTmaCfg {
    using G = WgCfg<KD, KHW, STRIDE, DIL>;
    static constexpr int BOX_H = DIL == 1 ? KHW : 1;            // rows one tile load covers (dilated rows: one load each)
    static constexpr int X_LOADS = KHW / BOX_H;
    static constexpr int LEAD = KHW == 1 ? 0 : 4;                // columns the box starts left of the first output's input
    static constexpr int OFF = LEAD - (KHW / 2) * DIL;           // box column of (output 0, tap 0)
    static constexpr int NXQ = (OFF + 7 * STRIDE + (KHW - 1) * DIL) / 4 + 1;  // float4s around 8 output positions
    static constexpr int QSTEP = 2 * STRIDE;                     // float4s between consecutive groups of 8
    static constexpr int LAST_Q = (G::P / 8 - 1) * QSTEP + NXQ - 1;
    static constexpr bool TAIL = LAST_Q == 9;                    // float4 #9 = columns 36..39: the second box
    static_assert(LAST_Q <= 9 && (TAIL || LAST_Q <= 8), "stage window exceeds the staged columns");
    static constexpr int X_FLOATS = G::ROWS * 32 * kWgPitch;     // [kd][kh][ci][36]
    static constexpr int TAIL_FLOATS = TAIL ? G::ROWS * 32 * 4 : 0;  // [kd][kh][ci][4]
    static constexpr int DY_FLOATS = 32 * G::P;                  // [co][P]
    static constexpr int STAGE_FLOATS = X_FLOATS + TAIL_FLOATS + DY_FLOATS;
    static constexpr int SMEM_BYTES = kWgRing * STAGE_FLOATS * 4 + 64;
    static_assert((X_FLOATS * 4) % 128 == 0 && (TAIL_FLOATS * 4) % 128 == 0 && (STAGE_FLOATS * 4) % 128 == 0,
                  "TMA destinations are 128-byte aligned");
    static_assert((BOX_H * 32 * kWgPitch * 4) % 128 == 0 && (BOX_H * 32 * 16) % 128 == 0, "per-row boxes stay aligned");
    static_assert(KD == 1 || DIL == 1, "dilated 3-D convolutions are not in the network");
    static_assert(G::P % 8 == 0, "threads walk a stage 8 output positions at a time");
};

template <int KD, int KHW, int STRIDE, int DIL>
__global__ void __launch_bounds__(kWgThreads, KD == 1 ? 2 : 1)  // 2-D tiles keep <= 36 accumulators: two CTAs per SM
    conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_tail,
                          const __grid_constant__ CUtensorMap tm_dy, float* __restrict__ dw, const WgDims dm) {
    using G = WgCfg<KD, KHW, STRIDE, DIL>;
    using T = WgTmaCfg<KD, KHW, STRIDE, DIL>;
    extern __shared__ __align__(128) float smem_wg_ring[];  // no static shared memory: the window starts 1 KB-aligned
    float* ring = smem_wg_ring;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kWgRing * T::STAGE_FLOATS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tiles = (dm.Cin + 31) / 32;
    const int co0 = (blockIdx.y / ci_tiles) * 32, ci0 = (blockIdx.y % ci_tiles) * 32;
    const int pad_hw = (KHW / 2) * DIL, pad_d = KD / 2;
    const long long s_beg = dm.stages * blockIdx.x / gridDim.x, s_end = dm.stages * (blockIdx.x + 1) / gridDim.x;
    const int n = (int)(s_end - s_beg);

    // producer state (thread 0): coordinates of the next stage to fetch
    int l_seg = 0, l_ho = 0, l_dz = 0, l_b = 0, l_k = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgRing; ++i) mbar_init(full + i, 1);
        fence_mbar_init();
        long long r = s_beg;
        l_seg = (int)(r % dm.segs), r /= dm.segs;
        l_ho = (int)(r % dm.Ho), r /= dm.Ho;
        l_dz = (int)(r % dm.Do);
        l_b = (int)(r / dm.Do);
    }
    auto fetch = [&]() {  // thread 0 only
        const int buf = l_k % kWgRing;
        float* sx = ring + buf * T::STAGE_FLOATS;
        mbar_arrive_expect_tx(full + buf, T::STAGE_FLOATS * 4);
        const int wo0 = l_seg * G::P;
        const int wi0 = wo0 * STRIDE - T::LEAD, hi0 = l_ho * STRIDE - pad_hw;
        const int di0 = KD == 3 ? l_dz * STRIDE - pad_d : 0;
#pragma unroll
        for (int i = 0; i < T::X_LOADS; ++i) {
            tma_load_5d(sx + i * T::BOX_H * 32 * kWgPitch, &tm_x, full + buf, wi0, ci0, hi0 + i * T::BOX_H * DIL, di0, l_b);
            if (T::TAIL)
                tma_load_5d(sx + T::X_FLOATS + i * T::BOX_H * 32 * 4, &tm_tail, full + buf, wi0 + kWgPitch, ci0,
                            hi0 + i * T::BOX_H * DIL, di0, l_b);
        }
        tma_load_5d(sx + T::X_FLOATS + T::TAIL_FLOATS, &tm_dy, full + buf, wo0, co0, l_ho, l_dz, l_b);
        ++l_k;
        if (++l_seg == dm.segs) {
            l_seg = 0;
            if (++l_ho == dm.Ho) {
                l_ho = 0;
                if (++l_dz == dm.Do) l_dz = 0, ++l_b;
            }
        }
    };

    float acc[4][G::TAPS];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < G::TAPS; ++t) acc[j][t] = 0.f;

    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < kWgRing - 1 && k < n; ++k) fetch();
    for (int k = 0; k < n; ++k) {
        __syncthreads();  // every warp is done with the buffer of stage k-1: refill it with stage k + ring - 1
        if (threadIdx.x == 0 && k + kWgRing - 1 < n) fetch();
        const int buf = k % kWgRing;
        mbar_wait(full + buf, (uint32_t)(k / kWgRing) & 1u);
        const float* sx = ring + buf * T::STAGE_FLOATS + lane * kWgPitch;
        const float* stail = ring + buf * T::STAGE_FLOATS + T::X_FLOATS + lane * 4;
        const float* sdy = ring + buf * T::STAGE_FLOATS + T::X_FLOATS + T::TAIL_FLOATS + warp * 4 * G::P;
#pragma unroll 1
        for (int grp = 0; grp < G::P / 8; ++grp) {
            float g[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 v = *reinterpret_cast<const float4*>(sdy + j * G::P + grp * 8 + 4 * h);  // broadcast
                    g[j][4 * h] = v.x, g[j][4 * h + 1] = v.y, g[j][4 * h + 2] = v.z, g[j][4 * h + 3] = v.w;
                }
            const bool last = grp == G::P / 8 - 1;
#pragma unroll
            for (int row = 0; row < G::ROWS; ++row) {
                float xv[T::NXQ * 4];
#pragma unroll
                for (int q = 0; q < T::NXQ; ++q) {
                    const float* src = sx + row * 32 * kWgPitch + (grp * T::QSTEP + q) * 4;
                    if (T::TAIL && q == T::NXQ - 1 && last) src = stail + row * 32 * 4;
                    const float4 v = *reinterpret_cast<const float4*>(src);
                    xv[4 * q] = v.x, xv[4 * q + 1] = v.y, xv[4 * q + 2] = v.z, xv[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int kw = 0; kw < KHW; ++kw)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int p = 0; p < 8; ++p)
                            acc[j][row * KHW + kw] =
                                fmaf(g[j][p], xv[T::OFF + p * STRIDE + kw * DIL], acc[j][row * KHW + kw]);
            }
        }
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

// fp32 NC(D)HW tensor as the rank-5 (W, C, H, D, B) view the stage boxes are cut from
static int encode_wgrad_map(CUtensorMap* tm, const float* base, int B, int C, int D, int H, int W, int box_w, int box_h,
                            int box_d) {
    const cuuint64_t gdim[5] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)D * H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)H * W * 4,
                                (cuuint64_t)C * D * H * W * 4};
    const cuuint32_t box[5] = {(cuuint32_t)box_w, 32u, (cuuint32_t)box_h, (cuuint32_t)box_d, 1u};
    return encode_tmap_5d(tm, base, gdim, gstr, box, "conv_wgrad", CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
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
    // 2 * SMs blocks over all (co, ci) tiles, rounded DOWN: the 2-D TMA kernels run two CTAs per SM, so the grid is one
    // full wave (a 19 x 16 grid on 296 slots left 8 blocks for a second wave and doubled the kernel time); the 3-D
    // kernels (one CTA per SM) run it as two waves.
    long long gx = 2LL * sms / tiles;
    if (gx > dm.stages) gx = dm.stages;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)tiles);
    const bool tma_ok = dm.W % 4 == 0 && dm.Wo % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
    if (tma_ok) {
        using T = WgTmaCfg<KD, KHW, STRIDE, DIL>;
        CUtensorMap tm_x, tm_tail, tm_dy;
        int rc = encode_wgrad_map(&tm_x, x, dm.B, dm.Cin, dm.D, dm.H, dm.W, kWgPitch, T::BOX_H, KD);
        if (rc != CMFB200_OK) return rc;
        rc = encode_wgrad_map(&tm_tail, x, dm.B, dm.Cin, dm.D, dm.H, dm.W, 4, T::BOX_H, KD);
        if (rc != CMFB200_OK) return rc;
        rc = encode_wgrad_map(&tm_dy, dy, dm.B, dm.Cout, dm.Do, dm.Ho, dm.Wo, G::P, 1, 1);
        if (rc != CMFB200_OK) return rc;
        auto tkern = conv_wgrad_tma_kernel<KD, KHW, STRIDE, DIL>;
        CMF_CUDA(cudaFuncSetAttribute(tkern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
        tkern<<<grid, kWgThreads, T::SMEM_BYTES, st>>>(tm_x, tm_tail, tm_dy, dw, dm);
        CMF_LAUNCH_CHECK("conv_wgrad_tma_kernel");
        return CMFB200_OK;
    }
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
