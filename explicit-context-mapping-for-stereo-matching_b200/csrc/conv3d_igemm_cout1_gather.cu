// Classifier tail conv (3x3x3, 32 -> 1; classifN.2, cmf/models/cmfsm.py:624,629,634) of the bf16 aggregation as ONE small
// GEMM per input plane plus a 27-point gather, instead of a zero-padded 32-output-channel convolution.
//
//     P[pos, tap] = sum_ci x[pos, ci] * w[tap, ci]              (tensor cores: M = positions of the halo'd plane tile,
//                                                                N = 27 taps padded to 32, K = 32; the activation tile
//                                                                is used UNSHIFTED, 4 MMAs per plane instead of 18)
//     out[d,h,w]  = sum_{kd,kh,kw} P_{plane d+kd-1}[(h+kh, w+kw), (kd,kh,kw)]      (27 shared-memory reads per output)
//
// A persistent CTA walks a (h,w) tile column through depth like conv3d_igemm_kdstack.cu: TMA ring of plane tiles
// (18 x 10 positions x 32 channels), two TMEM accumulator sets of 2 x 32 columns (the 180 tile positions are two
// M = 128 row blocks; the rows past 180 read whatever follows the tile in shared memory and are never used), the four
// epilogue warps copy P into a 4-slot shared-memory ring ([tap][position], conflict-free) and, once planes d-1, d, d+1
// have landed, each of their 128 threads gathers one output of the 16 x 8 tile.
// Status: correct (tests/test_kernels_gpu.py) and 4.5x fewer MMAs, but measured at the SAME 76 us per launch as the
// zero-padded convolution: one plane per ~1000 clk, bound by the serial per-plane chain of the four epilogue warps
// (tcgen05.ld -> 54 stores -> barrier -> 27 loads).  Kept opt-in (CMF_B200_COUT1_GATHER=1) as the starting point for a
// version with eight epilogue warps / two planes per barrier.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {
constexpr int kGW = 10, kGH = 18;           // halo'd plane tile
constexpr int kGPos = kGW * kGH;            // 180 positions
constexpr int kGPitch = 192;                // P row pitch (positions) per tap
constexpr int kGTaps = 27;
constexpr int kGNS = 4;                     // TMA stages
constexpr int kGChunk = kGPos * 16;         // 2880 bytes: one 8-channel chunk of a plane tile
constexpr int kGABytes = 4 * kGChunk;       // Cin = 32
constexpr int kGWBytes = 4 * 32 * 16;       // [Cin/8][32 taps][8] bf16
constexpr int kGPBytes = kGTaps * kGPitch * 4;
constexpr int kGSmem = kGNS * kGABytes + kGWBytes + 4 * kGPBytes + 1024 + 1024;
static_assert(kGABytes % 128 == 0, "TMA destinations must stay 128-byte aligned");
static_assert(kGWBytes >= 255 * 16 + 16 - kGChunk, "the rows past the tile must stay inside the allocation");
}

__device__ __forceinline__ void cg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kIgThreads, 1)
    conv3d_igemm_cout1_gather_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wtap,
                                     float* __restrict__ y, int D, int H, int W, int tiles_w, int tiles_h,
                                     long long total_planes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // [NS][A_BYTES]
    uint8_t* sW = smem + kGNS * kGABytes;                 // tap weights (also absorbs the over-read of the last stage)
    float* sP = reinterpret_cast<float*>(sW + kGWBytes);  // [4 slots][27][192]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sP) + 4 * kGPBytes);
    uint64_t* barW = bars;
    uint64_t* fullA = bars + 1;          // [NS]
    uint64_t* emptyA = fullA + kGNS;     // [NS]
    uint64_t* tmemFull = emptyA + kGNS;  // [2]
    uint64_t* tmemEmpty = tmemFull + 2;  // [2] (128 epilogue threads arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemEmpty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(barW, 1);
        for (int i = 0; i < kGNS; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmemFull + i, 1);
            mbar_init(tmemEmpty + i, 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(barW, kGWBytes);
            bulk_g2s(sW, wtap, kGWBytes, barW);
            int g = 0;
            KdWalk walk((int)blockIdx.x, (int)gridDim.x, total_planes, D, tiles_w, tiles_h);
            KdUnit u;
            while (walk.next(u)) {
                for (int p = u.pl0; p <= u.pl1; ++p, ++g) {
                    const int s = g % kGNS;
                    if (g >= kGNS) mbar_wait(emptyA + s, ((g / kGNS) - 1) & 1);
                    mbar_arrive_expect_tx(fullA + s, kGABytes);
                    tma_load_5d(sA + s * kGABytes, &tmap_x, fullA + s, (u.tx * 8 - 1) * 8, u.ty * 16 - 1, p, 0, u.b);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(128), b_hi = umma_desc_hi(128);  // 8 consecutive positions = 128 contiguous bytes
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), 32 * 16);
        mbar_wait(barW, 0);
        tc_fence_after();
        int g = 0;
        KdWalk walk((int)blockIdx.x, (int)gridDim.x, total_planes, D, tiles_w, tiles_h);
        KdUnit u;
        while (walk.next(u)) {
            for (int p = u.pl0; p <= u.pl1; ++p, ++g) {
                const int s = g % kGNS, buf = g & 1;
                mbar_wait(fullA + s, (g / kGNS) & 1);
                if (g >= 2) mbar_wait(tmemEmpty + buf, ((g >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + s * kGABytes, kGChunk);
                const uint32_t dcol = tmem_base + buf * 64;
                if (elect_one()) {
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            const uint64_t ad = umma_desc_at(a_lo, a_hi, mt * 2048 + 2 * kc * kGChunk);
                            const uint64_t bd = umma_desc_at(w_lo, b_hi, 2 * kc * (32 * 16));
                            umma_bf16(dcol + mt * 32, ad, bd, idesc, kc != 0 ? 1u : 0u);
                        }
                    umma_commit(emptyA + s);
                    umma_commit(tmemFull + buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..5: land P planes in shared memory, gather the outputs
        const int quad = warp & 3;
        const int et = threadIdx.x - 64;           // 0..127 = output (h, w) of the tile
        const int oh_l = et >> 3, ow_l = et & 7;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        int g_base = 0, landed = 0;
        KdWalk walk((int)blockIdx.x, (int)gridDim.x, total_planes, D, tiles_w, tiles_h);
        KdUnit u;
        while (walk.next(u)) {
            const int h = u.ty * 16 + oh_l, w = u.tx * 8 + ow_l;
            const bool ok = (h < H) && (w < W);
#pragma unroll 1
            for (int d = u.d0; d < u.d1; ++d) {
                const int need = g_base + ((d + 1 < D ? d + 1 : D - 1) - u.pl0);
                while (landed <= need) {
                    const int buf = landed & 1;
                    float* P = sP + (size_t)(landed & 3) * (kGTaps * kGPitch);
                    mbar_wait(tmemFull + buf, (landed >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tlane + buf * 64 + mt * 32, v);
                        const int row = mt * 128 + quad * 32 + lane;
                        if (row < kGPos) {
#pragma unroll
                            for (int t = 0; t < kGTaps; ++t) P[t * kGPitch + row] = __uint_as_float(v[t]);
                        }
                    }
                    tc_fence_before();
                    cg_mbar_arrive(tmemEmpty + buf);
                    ++landed;
                    asm volatile("bar.sync 1, 128;" ::: "memory");  // the plane is visible to all four warps
                }
                float acc = 0.f;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int p = d + kd - 1;
                    if (p < 0 || p >= D) continue;  // warp-uniform: zero padding along depth
                    const float* P = sP + (size_t)((g_base + p - u.pl0) & 3) * (kGTaps * kGPitch);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            acc += P[((kd * 3 + kh) * 3 + kw) * kGPitch + (oh_l + kh) * kGW + ow_l + kw];
                }
                if (ok) y[(((size_t)u.b * D + d) * H + h) * W + w] = acc;
            }
            g_base += u.pl1 - u.pl0 + 1;
            // a slot is rewritten four landings later; every thread has passed at least one barrier in between
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_conv3d_igemm_cout1_gather_bf16_fwd(const void* x_c8, const void* tap_weights, float* y, int B,
                                                          int Cin, int D, int H, int W, void* stream) {
    CMF_REQUIRE(x_c8 && tap_weights && y, "conv3d_igemm_cout1_gather_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_igemm_cout1_gather_bf16_fwd: non-positive dimension");
    CMF_REQUIRE(Cin == 32, "conv3d_igemm_cout1_gather_bf16_fwd: unsupported Cin=%d (supported: 32)", Cin);
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, 4, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)4 * D * H * W * 16};
    const cuuint32_t box[5] = {kGW * 8, kGH, 1, 4, 1};
    if (int rc = encode_tmap_5d(&tmap, x_c8, gdim, gstr, box, "conv3d_igemm_cout1_gather")) return rc;
    CMF_CUDA(cudaFuncSetAttribute(conv3d_igemm_cout1_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem));
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16);
    const long long total_planes = (long long)tiles_w * tiles_h * B * D;
    const long long grid = total_planes < sms ? total_planes : sms;
    conv3d_igemm_cout1_gather_kernel<<<(unsigned)grid, kIgThreads, kGSmem, (cudaStream_t)stream>>>(
        tmap, reinterpret_cast<const __nv_bfloat16*>(tap_weights), y, D, H, W, tiles_w, tiles_h, total_planes);
    CMF_LAUNCH_CHECK("conv3d_igemm_cout1_gather_kernel");
    return CMFB200_OK;
}
