// K2 (throughput path) -- PERSISTENT version of the stride-1 3x3x3 implicit GEMM of conv3d_igemm.cu (same C8
// layout, same descriptors, same math).  What changes is the schedule:
//   * one CTA per SM loops over output tiles (static round-robin);
//   * the 27 packed weight taps stay RESIDENT in shared memory for the CTA's lifetime when they fit
//     (32->32: 54 KB, 64->32: 108 KB); 64->64 (216 KB) streams them through a deep ring instead;
//   * the halo'd activation block is double-buffered where shared memory allows, so the TMA load of tile i+1
//     overlaps the MMAs of tile i;
//   * TMEM holds TWO accumulator sets: the epilogue warps drain tile i (tcgen05.ld -> bf16 store + GroupNorm
//     partial sums) while the MMA thread already issues tile i+1.
// The non-persistent kernel measured 18.8 % tensor-pipe activity: per CTA the weight ring was latency-bound and
// load / MMA / epilogue were serialised (profiles/README.md).  With operands in shared memory a 128xN MMA with
// N=32 is bounded by the 128 B/clk shared-memory read port (4 KB of A + 1 KB of B per K=16 step = 40 clk vs 16 clk
// of tensor time), i.e. ~40 % pipe activity is the SS-mode ceiling for these Cout=32 layers.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

constexpr int kPW = 10, kPH = 18;  // 8 x 16 output tile + 1-voxel halo on both sides

template <int CIN, int COUT, int BD, int NA, bool RESIDENT, int NSW>
struct PCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int PD = BD + 2;
    static constexpr int CHUNK_BYTES = PD * kPH * kPW * 16;
    static constexpr int A_BYTES = NC * CHUNK_BYTES;
    static constexpr int TAP_BYTES = CIN * COUT * 2;
    static constexpr int W_BYTES = (RESIDENT ? 27 : NSW) * TAP_BYTES;
    static constexpr int ACC_COLS = BD * COUT;  // one accumulator set
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = NA * A_BYTES + W_BYTES + 1024 + 4 * COUT * 2 * 8 + 1024;
    static_assert(A_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit in TMEM");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN, int COUT, int BD, int NA, bool RESIDENT, int NSW>
__global__ void __launch_bounds__(kIgThreads, 1)
    conv3d_igemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                                   __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int D, int H, int W,
                                   int tiles_w, int tiles_h, int tiles_d, int total_tiles) {
    using G = PCfg<CIN, COUT, BD, NA, RESIDENT, NSW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                      // [NA][A_BYTES]
    uint8_t* sW = smem + NA * G::A_BYTES;    // resident taps or ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + G::W_BYTES);
    uint64_t* barW = bars;                   // resident weights landed
    uint64_t* fullA = bars + 1;              // [NA]
    uint64_t* emptyA = fullA + NA;           // [NA]
    uint64_t* tmemFull = emptyA + NA;        // [2]
    uint64_t* tmemEmpty = tmemFull + 2;      // [2]  (128 epilogue threads arrive)
    uint64_t* fullW = tmemEmpty + 2;         // [NSW] ring (unused when RESIDENT)
    uint64_t* emptyW = fullW + NSW;          // [NSW]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyW + NSW);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][COUT][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_sample = tiles_w * tiles_h * tiles_d;

    if (threadIdx.x == 0) {
        mbar_init(barW, 1);
        for (int i = 0; i < NA; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmemFull + i, 1);
            mbar_init(tmemEmpty + i, 128);
        }
        for (int i = 0; i < NSW; ++i) {
            mbar_init(fullW + i, 1);
            mbar_init(emptyW + i, 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
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
            // ===== TMA producer
            if constexpr (RESIDENT) {
                mbar_arrive_expect_tx(barW, 27 * G::TAP_BYTES);
                for (int tap = 0; tap < 27; ++tap)
                    bulk_g2s(sW + tap * G::TAP_BYTES, wpk + (size_t)tap * CIN * COUT, G::TAP_BYTES, barW);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int b = tile / tiles_per_sample;
                int r = tile - b * tiles_per_sample;
                const int tz = r / (tiles_w * tiles_h);
                r -= tz * tiles_w * tiles_h;
                const int ty = r / tiles_w, tx = r - ty * tiles_w;
                const int buf = it % NA;
                if (it >= NA) mbar_wait(emptyA + buf, ((it / NA) - 1) & 1);
                mbar_arrive_expect_tx(fullA + buf, G::A_BYTES);
                tma_load_5d(sA + buf * G::A_BYTES, &tmap_x, fullA + buf, (tx * 8 - 1) * 8, ty * 16 - 1, tz * BD - 1, 0, b);
                if constexpr (!RESIDENT) {
                    for (int tap = 0; tap < 27; ++tap) {
                        const int g = it * 27 + tap;
                        const int s = g % NSW;
                        if (g >= NSW) mbar_wait(emptyW + s, ((g / NSW) - 1) & 1);
                        mbar_arrive_expect_tx(fullW + s, G::TAP_BYTES);
                        bulk_g2s(sW + s * G::TAP_BYTES, wpk + (size_t)tap * CIN * COUT, G::TAP_BYTES, fullW + s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the schedule (uniform control flow), one elected lane issues
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(kPW * 16), b_hi = umma_desc_hi(128);
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), COUT * 16);
        if constexpr (RESIDENT) {
            mbar_wait(barW, 0);
            tc_fence_after();
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it % NA, acc = it & 1;
            mbar_wait(fullA + buf, (it / NA) & 1);
            if (it >= 2) mbar_wait(tmemEmpty + acc, ((it >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + buf * G::A_BYTES, G::CHUNK_BYTES);
            const uint32_t d0 = tmem_base + acc * G::ACC_COLS;
            if constexpr (RESIDENT) {
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 27; ++tap) {
                        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
#pragma unroll
                        for (int mt = 0; mt < BD; ++mt) {
#pragma unroll
                            for (int kc = 0; kc < CIN / 16; ++kc) {
                                const uint64_t ad = umma_desc_at(a_lo, a_hi, ((((mt + kd) * kPH + kh) * kPW) + kw) * 16 + 2 * kc * G::CHUNK_BYTES);
                                const uint64_t bd = umma_desc_at(w_lo, b_hi, tap * G::TAP_BYTES + 2 * kc * (COUT * 16));
                                umma_bf16(d0 + mt * COUT, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(emptyA + buf);    // activation buffer may be refilled
                    umma_commit(tmemFull + acc);  // accumulators of this tile are final
                }
                __syncwarp();
            } else {
#pragma unroll 1
                for (int tap = 0; tap < 27; ++tap) {
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    const int g = it * 27 + tap;
                    const int s = g % NSW;
                    mbar_wait(fullW + s, (g / NSW) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t arow = ((((kd * kPH) + kh) * kPW) + kw) * 16;
#pragma unroll
                        for (int mt = 0; mt < BD; ++mt) {
#pragma unroll
                            for (int kc = 0; kc < CIN / 16; ++kc) {
                                const uint64_t ad = umma_desc_at(a_lo, a_hi, arow + mt * kPH * kPW * 16 + 2 * kc * G::CHUNK_BYTES);
                                const uint64_t bd = umma_desc_at(w_lo, b_hi, s * G::TAP_BYTES + 2 * kc * (COUT * 16));
                                umma_bf16(d0 + mt * COUT, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(emptyW + s);
                        if (tap == 26) {
                            umma_commit(emptyA + buf);
                            umma_commit(tmemFull + acc);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue warps 2..5
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const size_t plane = (size_t)H * W;
        const int et = threadIdx.x - 64;  // 0..127
        // GroupNorm statistics: lane l of a warp keeps the running double total of column (half*32 + l) over all
        // tiles of the current sample; they are combined across the four warps and flushed with one atomic per
        // channel when the sample changes / at the end (not per tile).
        double tot_s[COUT / 32], tot_q[COUT / 32];
#pragma unroll
        for (int i = 0; i < COUT / 32; ++i) tot_s[i] = tot_q[i] = 0.0;
        int cur_b = -1;
        auto flush = [&](int b) {
#pragma unroll
            for (int i = 0; i < COUT / 32; ++i) {
                sred[(quad * COUT + i * 32 + lane) * 2 + 0] = tot_s[i];
                sred[(quad * COUT + i * 32 + lane) * 2 + 1] = tot_q[i];
                tot_s[i] = tot_q[i] = 0.0;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < COUT * 2; i += 128) {
                const int c = i >> 1, which = i & 1;
                double a = 0.0;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) a += sred[(qd * COUT + c) * 2 + which];
                atomicAdd(gn_sums + ((size_t)b * COUT + c) * 2 + which, a);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int b = tile / tiles_per_sample;
            int r = tile - b * tiles_per_sample;
            const int tz = r / (tiles_w * tiles_h);
            r -= tz * tiles_w * tiles_h;
            const int ty = r / tiles_w, tx = r - ty * tiles_w;
            const int h = ty * 16 + (row >> 3), w = tx * 8 + (row & 7), d0 = tz * BD;
            const bool hw_ok = (h < H) && (w < W);
            const int acc = it & 1;
            if (gn_sums != nullptr && b != cur_b) {
                if (cur_b >= 0) flush(cur_b);
                cur_b = b;
            }
            mbar_wait(tmemFull + acc, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
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
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * G::ACC_COLS + mt * COUT + half * 32, v);
                    const int d = d0 + mt;
                    if (hw_ok && d < D) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 p[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                p[e] = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]));
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
                if (half == COUT / 32 - 1) {  // all TMEM reads of this tile are done: hand the accumulators back
                    tc_fence_before();
                    mbar_arrive(tmemEmpty + acc);
                }
                if (gn_sums != nullptr) {
                    tot_s[half] += (double)warp_transpose_sum32(s, lane);
                    tot_q[half] += (double)warp_transpose_sum32(ss, lane);
                }
            }
        }
        if (gn_sums != nullptr && cur_b >= 0) flush(cur_b);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(G::TMEM_COLS)
                     : "memory");
    }
}

template <int CIN, int COUT, int BD, int NA, bool RESIDENT, int NSW>
static int launch_persistent(const void* x, const void* wpk, void* y, double* gn, int B, int D, int H, int W,
                             cudaStream_t st) {
    using G = PCfg<CIN, COUT, BD, NA, RESIDENT, NSW>;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)G::NC * D * H * W * 16};
    const cuuint32_t box[5] = {kPW * 8, kPH, (cuuint32_t)G::PD, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "conv3d_igemm_persistent")) return rc;
    auto kern = conv3d_igemm_persistent_kernel<CIN, COUT, BD, NA, RESIDENT, NSW>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16), tiles_d = (int)cdiv(D, BD);
    const long long total = (long long)tiles_w * tiles_h * tiles_d * B;
    CMF_REQUIRE(total < (1ll << 31), "conv3d_igemm_persistent: too many tiles");
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(total < sms ? total : sms);
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(wpk),
                                                  reinterpret_cast<__nv_bfloat16*>(y), gn, D, H, W, tiles_w, tiles_h,
                                                  tiles_d, (int)total);
    CMF_LAUNCH_CHECK("conv3d_igemm_persistent_kernel");
    return CMFB200_OK;
}

int conv3d_igemm_kdstack_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout, int D,
                                  int H, int W, cudaStream_t st);  // conv3d_igemm_kdstack.cu

// dispatcher used by cmfb200_conv3d_igemm_bf16_fwd (conv3d_igemm.cu)
int conv3d_igemm_persistent_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                     int D, int H, int W, cudaStream_t st) {
    static const bool tap_schedule = getenv("CMFB200_IGEMM_PER_TAP") != nullptr;  // A/B switch: one N=32 MMA per tap
    if (!tap_schedule && (Cout == 32 || (Cin == 64 && Cout == 64)))
        return conv3d_igemm_kdstack_dispatch(x, wpk, y, gn, B, Cin, Cout, D, H, W, st);
    if (Cin == 32 && Cout == 32) return launch_persistent<32, 32, 4, 2, true, 1>(x, wpk, y, gn, B, D, H, W, st);
    if (Cin == 64 && Cout == 32) return launch_persistent<64, 32, 2, 1, true, 1>(x, wpk, y, gn, B, D, H, W, st);
    if (Cin == 64 && Cout == 64) return launch_persistent<64, 64, 2, 1, false, 12>(x, wpk, y, gn, B, D, H, W, st);
    CMF_REQUIRE(false, "conv3d_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->32, 64->32, 64->64", Cin,
                Cout);
}

}  // namespace cmfb200
