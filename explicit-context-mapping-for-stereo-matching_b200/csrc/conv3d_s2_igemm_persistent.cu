// K2 (throughput path) -- PERSISTENT stride-2 3x3x3 conv (hourglass.conv1 / conv3, cmf/models/cmfsm.py:244-254) on
// the parity-split C8/bf16 input.  Decomposition (see c8_s2_entry.cu): input
// index i = 2*o + k - 1, so tap k=1 reads parity-0 inputs at o and taps k=0 / k=2 read parity-1 inputs at o-1 / o;
// the eight parity sub-volumes of a tile are eight TMA boxes feeding 1, 2, 4 or 8 taps each into one accumulator.
// What changes is the schedule (the one-tile-per-CTA kernel re-streamed all 27 weight taps per tile and paid the CTA
// start-up -- TMEM allocation, barrier init, first-load latency -- once per tile):
//   * one CTA per SM walks output tiles; the 27 taps of one group of GC output channels stay resident in shared
//     memory (32->64: one group of 64; 64->64: two groups of 32 on disjoint halves of the grid);
//   * the parity boxes of consecutive tiles flow through one ring, so the loads of tile i+1 overlap the MMAs of i;
//   * two TMEM accumulator sets: the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {
constexpr int kPW2 = 9, kPH2 = 17;  // 8 x 16 output tile + one halo voxel towards -h / -w (and -d)

// taps that read input parity p along one axis: p=0 -> {k=1}; p=1 -> {k=0, k=2}
__host__ __device__ constexpr int p2_ntaps(int p) { return p ? 2 : 1; }
__host__ __device__ constexpr int p2_tap(int p, int j) { return p ? (j == 0 ? 0 : 2) : 1; }
// position of the tap's source inside the box (box origin = output index - 1): k=0 -> o-1 -> 0 ; k=1,2 -> o -> 1
__host__ __device__ constexpr int p2_shift(int k) { return k == 0 ? 0 : 1; }
}

template <int CIN, int GC, int BD, int NSA>
struct S2pCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int PD = BD + 1;
    static constexpr int CHUNK_BYTES = PD * kPH2 * kPW2 * 16;
    static constexpr int STAGE_BYTES = NC * CHUNK_BYTES;
    static constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) & ~127;
    static constexpr int TAP_BYTES = CIN * GC * 2;
    static constexpr int W_BYTES = 27 * TAP_BYTES;
    static constexpr int ACC_COLS = BD * GC;
    static constexpr int SMEM_BYTES = NSA * STAGE_STRIDE + W_BYTES + 1024 + 4 * GC * 2 * 8 + 1024;
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit in TMEM");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__device__ __forceinline__ void s2p_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN, int GC, int BD, int NSA>
__global__ void __launch_bounds__(kIgThreads, 1)
    conv3d_s2_igemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_xs, const __nv_bfloat16* __restrict__ wpk,
                                      __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int Do, int Ho, int Wo,
                                      int tiles_w, int tiles_h, int tiles_d, int total_tiles, int cout_total,
                                      int ctas_per_group) {
    using G = S2pCfg<CIN, GC, BD, NSA>;
    constexpr int TMEM_COLS = (2 * G::ACC_COLS <= 64) ? 64 : (2 * G::ACC_COLS <= 128) ? 128 : (2 * G::ACC_COLS <= 256) ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                              // [NSA][STAGE_STRIDE]
    uint8_t* sW = smem + NSA * G::STAGE_STRIDE;      // [27][CIN/8][GC][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + G::W_BYTES);
    uint64_t* barW = bars;
    uint64_t* fullA = bars + 1;           // [NSA]
    uint64_t* emptyA = fullA + NSA;       // [NSA]
    uint64_t* tmemFull = emptyA + NSA;    // [2]
    uint64_t* tmemEmpty = tmemFull + 2;   // [2] (128 epilogue threads arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemEmpty + 2);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][GC][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / ctas_per_group, rank = blockIdx.x - group * ctas_per_group;
    const int tiles_per_sample = tiles_w * tiles_h * tiles_d;

    if (threadIdx.x == 0) {
        mbar_init(barW, 1);
        for (int i = 0; i < NSA; ++i) {
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: weights of this channel group once, then 8 parity boxes per tile through the ring
            mbar_arrive_expect_tx(barW, G::W_BYTES);
            for (int tap = 0; tap < 27; ++tap)
                for (int c = 0; c < G::NC; ++c)
                    bulk_g2s(sW + tap * G::TAP_BYTES + c * GC * 16,
                             wpk + ((size_t)tap * G::NC + c) * cout_total * 8 + group * GC * 8, GC * 16, barW);
            int g = 0;
            for (int tile = rank; tile < total_tiles; tile += ctas_per_group) {
                const int b = tile / tiles_per_sample;
                int r = tile - b * tiles_per_sample;
                const int tz = r / (tiles_w * tiles_h);
                r -= tz * tiles_w * tiles_h;
                const int ty = r / tiles_w, tx = r - ty * tiles_w;
                for (int par = 0; par < 8; ++par, ++g) {
                    const int s = g % NSA;
                    if (g >= NSA) mbar_wait(emptyA + s, ((g / NSA) - 1) & 1);
                    mbar_arrive_expect_tx(fullA + s, G::STAGE_BYTES);
                    tma_load_5d(sA + s * G::STAGE_STRIDE, &tmap_xs, fullA + s, (tx * 8 - 1) * 8, ty * 16 - 1, tz * BD - 1,
                                par * G::NC, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues; the 27 taps are fully unrolled)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GC >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(kPW2 * 16), b_hi = umma_desc_hi(128);
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), GC * 16);
        mbar_wait(barW, 0);
        tc_fence_after();
        int g = 0, it = 0;
        for (int tile = rank; tile < total_tiles; tile += ctas_per_group, ++it) {
            const int acc = it & 1;
            if (it >= 2) mbar_wait(tmemEmpty + acc, ((it >> 1) - 1) & 1);
            const uint32_t dcol = tmem_base + acc * G::ACC_COLS;
#pragma unroll
            for (int par = 0; par < 8; ++par, ++g) {
                const int s = g % NSA;
                mbar_wait(fullA + s, (g / NSA) & 1);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + s * G::STAGE_STRIDE, G::CHUNK_BYTES);
                if (elect_one()) {
                    const int pd = par >> 2, ph = (par >> 1) & 1, pw = par & 1;
#pragma unroll
                    for (int jd = 0; jd < p2_ntaps(pd); ++jd)
#pragma unroll
                        for (int jh = 0; jh < p2_ntaps(ph); ++jh)
#pragma unroll
                            for (int jw = 0; jw < p2_ntaps(pw); ++jw) {
                                const int kd = p2_tap(pd, jd), kh = p2_tap(ph, jh), kw = p2_tap(pw, jw);
                                const int tap = kd * 9 + kh * 3 + kw;
                                const int arow = ((p2_shift(kd) * kPH2 + p2_shift(kh)) * kPW2 + p2_shift(kw)) * 16;
#pragma unroll
                                for (int mt = 0; mt < BD; ++mt)
#pragma unroll
                                    for (int kc = 0; kc < CIN / 16; ++kc) {
                                        const uint64_t ad = umma_desc_at(a_lo, a_hi, arow + mt * kPH2 * kPW2 * 16 + 2 * kc * G::CHUNK_BYTES);
                                        const uint64_t bd = umma_desc_at(w_lo, b_hi, tap * G::TAP_BYTES + 2 * kc * (GC * 16));
                                        // parity 0 (the centre tap) comes first: it initialises the accumulators
                                        umma_bf16(dcol + mt * GC, ad, bd, idesc, (par | kc) != 0 ? 1u : 0u);
                                    }
                            }
                    umma_commit(emptyA + s);
                    if (par == 7) umma_commit(tmemFull + acc);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..5
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const size_t plane = (size_t)Ho * Wo;
        const int et = threadIdx.x - 64;  // 0..127
        double tot_s[GC / 32], tot_q[GC / 32];
#pragma unroll
        for (int i = 0; i < GC / 32; ++i) tot_s[i] = tot_q[i] = 0.0;
        int cur_b = -1;
        auto flush = [&](int b) {
#pragma unroll
            for (int i = 0; i < GC / 32; ++i) {
                sred[(quad * GC + i * 32 + lane) * 2 + 0] = tot_s[i];
                sred[(quad * GC + i * 32 + lane) * 2 + 1] = tot_q[i];
                tot_s[i] = tot_q[i] = 0.0;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < GC * 2; i += 128) {
                const int c = i >> 1, which = i & 1;
                double a = 0.0;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) a += sred[(qd * GC + c) * 2 + which];
                atomicAdd(gn_sums + ((size_t)b * cout_total + group * GC + c) * 2 + which, a);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        int it = 0;
        for (int tile = rank; tile < total_tiles; tile += ctas_per_group, ++it) {
            const int b = tile / tiles_per_sample;
            int r = tile - b * tiles_per_sample;
            const int tz = r / (tiles_w * tiles_h);
            r -= tz * tiles_w * tiles_h;
            const int ty = r / tiles_w, tx = r - ty * tiles_w;
            const int h = ty * 16 + (row >> 3), w = tx * 8 + (row & 7), d0 = tz * BD;
            const bool hw_ok = (h < Ho) && (w < Wo);
            const int acc = it & 1;
            if (gn_sums != nullptr && b != cur_b) {
                if (cur_b >= 0) flush(cur_b);
                cur_b = b;
            }
            mbar_wait(tmemFull + acc, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < GC / 32; ++half) {
                float s[32], ss[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) s[c] = ss[c] = 0.f;
#pragma unroll 1
                for (int mt = 0; mt < BD; ++mt) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * G::ACC_COLS + mt * GC + half * 32, v);
                    const int d = d0 + mt;
                    if (hw_ok && d < Do) {
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
                            const int chunk = (group * GC) / 8 + half * 4 + j;
                            __nv_bfloat16* dst = y + ((((size_t)b * (cout_total >> 3) + chunk) * Do + d) * plane + (size_t)h * Wo + w) * 8;
                            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(p);
                        }
                    }
                }
                if (half == GC / 32 - 1) {  // all TMEM reads of this tile are done
                    tc_fence_before();
                    s2p_mbar_arrive(tmemEmpty + acc);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int CIN, int GC, int BD, int NSA>
static int launch_s2_persistent(const void* xs, const void* wpk, void* y, double* gn, int B, int Cout, int Do, int Ho,
                                int Wo, cudaStream_t st) {
    using G = S2pCfg<CIN, GC, BD, NSA>;
    CUtensorMap tmap;
    const cuuint64_t vol = (cuuint64_t)Do * Ho * Wo * 16;
    const cuuint64_t gdim[5] = {(cuuint64_t)Wo * 8, (cuuint64_t)Ho, (cuuint64_t)Do, (cuuint64_t)8 * G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)Wo * 16, (cuuint64_t)Ho * Wo * 16, vol, vol * 8 * G::NC};
    const cuuint32_t box[5] = {kPW2 * 8, kPH2, (cuuint32_t)G::PD, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, xs, gdim, gstr, box, "conv3d_s2_igemm_persistent")) return rc;
    auto kern = conv3d_s2_igemm_persistent_kernel<CIN, GC, BD, NSA>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles_w = (int)cdiv(Wo, 8), tiles_h = (int)cdiv(Ho, 16), tiles_d = (int)cdiv(Do, BD);
    const long long total = (long long)tiles_w * tiles_h * tiles_d * B;
    CMF_REQUIRE(total < (1ll << 31), "conv3d_s2_igemm_persistent: too many tiles");
    const int groups = Cout / GC;
    long long per_group = sms / groups;
    if (per_group < 1) per_group = 1;
    if (per_group > total) per_group = total;
    kern<<<(unsigned)(per_group * groups), kIgThreads, G::SMEM_BYTES, st>>>(
        tmap, reinterpret_cast<const __nv_bfloat16*>(wpk), reinterpret_cast<__nv_bfloat16*>(y), gn, Do, Ho, Wo, tiles_w,
        tiles_h, tiles_d, (int)total, Cout, (int)per_group);
    CMF_LAUNCH_CHECK("conv3d_s2_igemm_persistent_kernel");
    return CMFB200_OK;
}

// used by cmfb200_conv3d_s2_igemm_bf16_fwd (c8_s2_entry.cu)
int conv3d_s2_igemm_persistent_dispatch(const void* xs, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                        int Do, int Ho, int Wo, cudaStream_t st) {
    if (Cin == 32 && Cout == 64) return launch_s2_persistent<32, 64, 2, 3>(xs, wpk, y, gn, B, Cout, Do, Ho, Wo, st);
    if (Cin == 64 && Cout == 64) return launch_s2_persistent<64, 32, 1, 2>(xs, wpk, y, gn, B, Cout, Do, Ho, Wo, st);
    CMF_REQUIRE(false, "conv3d_s2_igemm_persistent: unsupported (Cin=%d, Cout=%d)", Cin, Cout);
}

}  // namespace cmfb200
