// K2 (throughput path) -- "kd-stacked" persistent implicit GEMM for the stride-1 3x3x3 layers: 32->32, 64->32 and
// 64->64 (as two independent groups of 32 output channels, each on half of the CTAs).
//
// A per-tap schedule issues one M128 x N32 MMA per (tap, k-step): with both operands in shared memory such an
// MMA reads 4 KB of A + 1 KB of B through the 128 B/clk port = 40 clk for 16 clk of tensor work (the 40 % ceiling
// measured there).  Here the three depth taps share one A read:
//
//     P_kd[d'] = conv2d_{kh,kw}( in[d'], W[kd] )            one MMA chain per INPUT plane d', N = 3 x 32 = 96
//     out[d]   = P_0[d-1] + P_1[d] + P_2[d+1]               three TMEM column blocks added in the epilogue
//
// so an input plane tile (18 x 10 voxels with halo, no depth halo at all) is loaded once, used by 9 x Cin/16 MMAs of
// N = 96 (4 KB of A + 3 KB of B = 56 clk for 48 clk of tensor work) and dropped.  A CTA walks a (h, w) tile column
// through depth: TMA ring of NS plane tiles, TMEM ring of four 96-column plane accumulators (the epilogue of out[d]
// reads planes d-1, d, d+1 while the MMA warp works on d+2), weights resident in shared memory in the stacked
// [kh,kw][Cin/8][kd*32+co][8] order (assembled from the ordinary per-tap packing by 512-byte bulk copies).
// Same C8 layout, same operands and the same fp32 accumulation as the other igemm kernels; only the order of the 27
// tap partial sums differs (3 chains of 9 taps, then 2 adds).
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {
constexpr int kKW = 10, kKH = 18;   // 8 x 16 output tile + halo
constexpr int kN = 96;              // 3 depth taps x 32 output channels
constexpr int kBufCols = 128;       // TMEM column pitch of one plane accumulator
}

template <int CIN, int NS>
struct KdCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int CHUNK_BYTES = kKH * kKW * 16;  // 2880
    static constexpr int A_BYTES = NC * CHUNK_BYTES;
    static constexpr int WCHUNK_BYTES = kN * 16;        // one 8-channel K chunk of the stacked B tile
    static constexpr int W_BYTES = 9 * NC * WCHUNK_BYTES;
    static constexpr int SMEM_BYTES = NS * A_BYTES + W_BYTES + 1024 + 4 * 32 * 2 * 8 + 1024;
    static_assert(A_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__device__ __forceinline__ void kd_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// COUT1: classifier tail (Cin -> 1): the weights arrive zero-padded to 32 output channels, only column 0 of every
// accumulator block is read back and written as fp32 [B][D][H][W] (no GroupNorm follows).
template <int CIN, int NS, bool COUT1>
__global__ void __launch_bounds__(kIgThreads, 1)
    conv3d_igemm_kdstack_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                                void* __restrict__ y_out, double* __restrict__ gn_sums, int D, int H, int W,
                                int tiles_w, int tiles_h, long long total_planes, int cout_total, int ctas_per_group) {
    using G = KdCfg<CIN, NS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                    // [NS][A_BYTES]
    uint8_t* sW = smem + NS * G::A_BYTES;  // stacked weights
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + G::W_BYTES);
    uint64_t* barW = bars;
    uint64_t* fullA = bars + 1;          // [NS]
    uint64_t* emptyA = fullA + NS;       // [NS]
    uint64_t* tmemFull = emptyA + NS;    // [4]
    uint64_t* tmemEmpty = tmemFull + 4;  // [4] (128 epilogue threads arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemEmpty + 4);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][32][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / ctas_per_group, rank = blockIdx.x - group * ctas_per_group;  // group = 32 couts

    if (threadIdx.x == 0) {
        mbar_init(barW, 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(tmemFull + i, 1);
            mbar_init(tmemEmpty + i, 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: stacked weights once, then one plane tile per ring slot
            mbar_arrive_expect_tx(barW, G::W_BYTES);
            for (int tap = 0; tap < 27; ++tap) {
                const int kd = tap / 9, khkw = tap % 9;
                for (int c = 0; c < G::NC; ++c)
                    bulk_g2s(sW + ((khkw * G::NC + c) * 3 + kd) * 512,
                             wpk + ((size_t)tap * G::NC + c) * cout_total * 8 + group * 256, 512, barW);
            }
            int g = 0;
            KdWalk walk(rank, ctas_per_group, total_planes, D, tiles_w, tiles_h);
            KdUnit u;
            while (walk.next(u)) {
                for (int p = u.pl0; p <= u.pl1; ++p, ++g) {
                    const int s = g % NS;
                    if (g >= NS) mbar_wait(emptyA + s, ((g / NS) - 1) & 1);
                    mbar_arrive_expect_tx(fullA + s, G::A_BYTES);
                    tma_load_5d(sA + s * G::A_BYTES, &tmap_x, fullA + s, (u.tx * 8 - 1) * 8, u.ty * 16 - 1, p, 0, u.b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues; immediates for every descriptor offset)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(kKW * 16), b_hi = umma_desc_hi(128);
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), G::WCHUNK_BYTES);
        mbar_wait(barW, 0);
        tc_fence_after();
        int g = 0;
        KdWalk walk(rank, ctas_per_group, total_planes, D, tiles_w, tiles_h);
        KdUnit u;
        while (walk.next(u)) {
            for (int p = u.pl0; p <= u.pl1; ++p, ++g) {
                const int s = g % NS, buf = g & 3;
                mbar_wait(fullA + s, (g / NS) & 1);
                if (g >= 4) mbar_wait(tmemEmpty + buf, ((g >> 2) - 1) & 1);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + s * G::A_BYTES, G::CHUNK_BYTES);
                const uint32_t dcol = tmem_base + buf * kBufCols;
                if (elect_one()) {
#pragma unroll
                    for (int khkw = 0; khkw < 9; ++khkw) {
                        const int kh = khkw / 3, kw = khkw % 3;
#pragma unroll
                        for (int kc = 0; kc < CIN / 16; ++kc) {
                            const uint64_t ad = umma_desc_at(a_lo, a_hi, (kh * kKW + kw) * 16 + 2 * kc * G::CHUNK_BYTES);
                            const uint64_t bd = umma_desc_at(w_lo, b_hi, (khkw * G::NC + 2 * kc) * G::WCHUNK_BYTES);
                            umma_bf16(dcol, ad, bd, idesc, (khkw | kc) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(emptyA + s);
                    umma_commit(tmemFull + buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..5: out[d] = P_0[d-1] + P_1[d] + P_2[d+1]
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const size_t plane = (size_t)H * W;
        const int et = threadIdx.x - 64;  // 0..127
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        double tot_s = 0.0, tot_q = 0.0;  // lane l: channel l of the current sample
        int cur_b = -1;
        auto flush = [&](int b) {
            sred[(quad * 32 + lane) * 2 + 0] = tot_s;
            sred[(quad * 32 + lane) * 2 + 1] = tot_q;
            tot_s = tot_q = 0.0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et < 64) {
                const int c = et >> 1, which = et & 1;
                double a = 0.0;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) a += sred[(qd * 32 + c) * 2 + which];
                atomicAdd(gn_sums + ((size_t)b * cout_total + group * 32 + c) * 2 + which, a);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        int g_base = 0, acquired = 0;
        KdWalk walk(rank, ctas_per_group, total_planes, D, tiles_w, tiles_h);
        KdUnit u;
        while (walk.next(u)) {
            const int h = u.ty * 16 + (row >> 3), w = u.tx * 8 + (row & 7);
            const bool hw_ok = (h < H) && (w < W);
            if (gn_sums != nullptr && u.b != cur_b) {
                if (cur_b >= 0) flush(cur_b);
                cur_b = u.b;
            }
            float s[32], ss[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) s[c] = ss[c] = 0.f;
            int since = 0;
#pragma unroll 1
            for (int d = u.d0; d < u.d1; ++d) {
                const int need = g_base + ((d + 1 < D ? d + 1 : D - 1) - u.pl0);
                while (acquired <= need) {
                    mbar_wait(tmemFull + (acquired & 3), (acquired >> 2) & 1);
                    ++acquired;
                }
                tc_fence_after();
                if constexpr (COUT1) {
                    uint32_t r1, r0 = 0, r2 = 0;  // 0 = +0.0f
                    tmem_ld_32x32b_x1_issue(tlane + ((g_base + d - u.pl0) & 3) * kBufCols + 32, r1);
                    if (d - 1 >= 0) tmem_ld_32x32b_x1_issue(tlane + ((g_base + d - 1 - u.pl0) & 3) * kBufCols + 0, r0);
                    if (d + 1 < D) tmem_ld_32x32b_x1_issue(tlane + ((g_base + d + 1 - u.pl0) & 3) * kBufCols + 64, r2);
                    tmem_ld_wait();
                    const float o0 = (__uint_as_float(r1) + __uint_as_float(r0)) + __uint_as_float(r2);
                    if (d - 1 >= u.pl0) {
                        tc_fence_before();
                        kd_mbar_arrive(tmemEmpty + ((g_base + d - 1 - u.pl0) & 3));
                    }
                    if (hw_ok) reinterpret_cast<float*>(y_out)[((size_t)u.b * D + d) * plane + (size_t)h * W + w] = o0;
                    continue;
                }
                __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(y_out);
                float o[32];
                {
                    // the three plane blocks are requested back to back and awaited once (one TMEM round trip per output
                    // plane instead of three: the epilogue chain, not the MMAs, bounded this kernel)
                    uint32_t v1[32], v0[32], v2[32];
                    const bool has0 = d - 1 >= 0, has2 = d + 1 < D;  // warp-uniform
                    tmem_ld_32x32b_x32_issue(tlane + ((g_base + d - u.pl0) & 3) * kBufCols + 32, v1);  // kd = 1
                    if (has0) tmem_ld_32x32b_x32_issue(tlane + ((g_base + d - 1 - u.pl0) & 3) * kBufCols + 0, v0);
                    if (has2) tmem_ld_32x32b_x32_issue(tlane + ((g_base + d + 1 - u.pl0) & 3) * kBufCols + 64, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float a = __uint_as_float(v1[c]);
                        if (has0) a += __uint_as_float(v0[c]);
                        if (has2) a += __uint_as_float(v2[c]);
                        o[c] = a;
                    }
                }
                if (d - 1 >= u.pl0) {  // plane d-1 has no reader left
                    tc_fence_before();
                    kd_mbar_arrive(tmemEmpty + ((g_base + d - 1 - u.pl0) & 3));
                }
                if (hw_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            pk[e] = __floats2bfloat162_rn(o[j * 8 + 2 * e], o[j * 8 + 2 * e + 1]);
                            const float r0 = __low2float(pk[e]), r1 = __high2float(pk[e]);
                            s[j * 8 + 2 * e] += r0;
                            ss[j * 8 + 2 * e] = fmaf(r0, r0, ss[j * 8 + 2 * e]);
                            s[j * 8 + 2 * e + 1] += r1;
                            ss[j * 8 + 2 * e + 1] = fmaf(r1, r1, ss[j * 8 + 2 * e + 1]);
                        }
                        __nv_bfloat16* dst = y + ((((size_t)u.b * (cout_total >> 3) + group * 4 + j) * D + d) * plane + (size_t)h * W + w) * 8;
                        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(pk);
                    }
                }
                if (gn_sums != nullptr && (++since == 4 || d == u.d1 - 1)) {  // fp32 partials cover at most 4 planes
                    tot_s += (double)warp_transpose_sum32(s, lane);
                    tot_q += (double)warp_transpose_sum32(ss, lane);
#pragma unroll
                    for (int c = 0; c < 32; ++c) s[c] = ss[c] = 0.f;
                    since = 0;
                }
            }
            // the last output plane and (when it exists) the trailing halo plane are done as well
            tc_fence_before();
            kd_mbar_arrive(tmemEmpty + ((g_base + u.d1 - 1 - u.pl0) & 3));
            if (u.pl1 == u.d1) kd_mbar_arrive(tmemEmpty + ((g_base + u.d1 - u.pl0) & 3));
            g_base += u.pl1 - u.pl0 + 1;
        }
        if (gn_sums != nullptr && cur_b >= 0) flush(cur_b);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

template <int CIN, int NS, bool COUT1>
static int launch_kdstack(const void* x, const void* wpk, void* y, double* gn, int B, int Cout, int D, int H, int W,
                          cudaStream_t st) {
    using G = KdCfg<CIN, NS>;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)G::NC * D * H * W * 16};
    const cuuint32_t box[5] = {kKW * 8, kKH, 1, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "conv3d_igemm_kdstack")) return rc;
    auto kern = conv3d_igemm_kdstack_kernel<CIN, NS, COUT1>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16);
    const int groups = Cout / 32;  // every CTA keeps the stacked weights of ONE group of 32 output channels resident
    const long long total_planes = (long long)tiles_w * tiles_h * B * D;
    long long per_group = sms / groups;
    if (per_group < 1) per_group = 1;
    if (per_group > total_planes) per_group = total_planes;  // every CTA gets at least one plane
    kern<<<(unsigned)(per_group * groups), kIgThreads, G::SMEM_BYTES, st>>>(
        tmap, reinterpret_cast<const __nv_bfloat16*>(wpk), y, gn, D, H, W, tiles_w, tiles_h, total_planes, Cout,
        (int)per_group);
    CMF_LAUNCH_CHECK("conv3d_igemm_kdstack_kernel");
    return CMFB200_OK;
}

// used by cmfb200_conv3d_igemm_bf16_fwd (c8_bf16_ops.cu): 32->32, 64->32 and (as two groups of 32 output channels) 64->64
int conv3d_igemm_kdstack_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout, int D,
                                  int H, int W, cudaStream_t st) {
    if (Cin == 32) return launch_kdstack<32, 6, false>(x, wpk, y, gn, B, Cout, D, H, W, st);
    if (Cin == 64) return launch_kdstack<64, 4, false>(x, wpk, y, gn, B, Cout, D, H, W, st);
    CMF_REQUIRE(false, "conv3d_igemm_kdstack: unsupported Cin=%d", Cin);
}

}  // namespace cmfb200

extern "C" int cmfb200_conv3d_igemm_cout1_bf16_fwd(const void* x_c8, const void* packed_w32, float* y, int B, int Cin,
                                                   int D, int H, int W, void* stream) {
    using namespace cmfb200;
    CMF_REQUIRE(x_c8 && packed_w32 && y, "conv3d_igemm_cout1_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_igemm_cout1_bf16_fwd: non-positive dimension");
    CMF_REQUIRE(Cin == 32, "conv3d_igemm_cout1_bf16_fwd: unsupported Cin=%d (supported: 32)", Cin);
    return launch_kdstack<32, 6, true>(x_c8, packed_w32, y, nullptr, B, 32, D, H, W, (cudaStream_t)stream);
}
