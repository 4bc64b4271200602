// K2 (throughput path) -- PERSISTENT transposed conv (k3 s2 p1 op1, hourglass.conv5/conv6, cmf/models/cmfsm.py:261-281)
// on the C8/bf16 layout.  Decomposition (see c8_s2_entry.cu): an output voxel of
// parity 0 along an axis takes tap k=1 of input i, parity 1 takes tap k=2 of input i and tap k=0 of input i+1.  What
// changes is the schedule (the one-tile-per-CTA kernel spent most of its time in CTA start-up and in re-streaming
// up to 110 KB of weights per tile):
//   * one CTA per SM walks tiles of 16 x 8 input positions x one input depth slice and produces ALL EIGHT output
//     parity classes of the tile (8 x 32 fp32 columns in TMEM, two such sets so the epilogue overlaps the next tile);
//   * the weights of one group of 32 output channels stay resident in shared memory (Cout = 64 runs as two groups on
//     disjoint halves of the grid); the depth taps kd=1 / kd=2 read the same activation tile, so they are stacked
//     into one N = 64 MMA (classes pd=0 / pd=1 are adjacent TMEM column blocks): 18 instead of 27 MMAs per k-step;
//   * halo'd activation blocks (box {9*8, 17, 2, Cin/8, 1}) are double-buffered.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {
constexpr int kDW = 9, kDH = 17;  // 8 x 16 input tile + one halo voxel towards +h / +w
}

template <int CIN>
struct DpCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int CHUNK_BYTES = 2 * kDH * kDW * 16;  // two depth planes
    static constexpr int A_BYTES = NC * CHUNK_BYTES;
    static constexpr int S_BYTES = 9 * NC * 1024;  // kd = 1,2 stacked: [khkw][Cin/8][2 x 32 couts][8]
    static constexpr int Z_BYTES = 9 * NC * 512;   // kd = 0:           [khkw][Cin/8][32 couts][8]
    static constexpr int SMEM_BYTES = 2 * A_BYTES + S_BYTES + Z_BYTES + 1024 + 4 * 32 * 2 * 8 + 1024;
    static_assert(A_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__device__ __forceinline__ void dp_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN>
__global__ void __launch_bounds__(kIgThreads, 1)
    deconv3d_igemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                                     __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int D, int H, int W,
                                     int tiles_w, int tiles_h, int total_tiles, int cout_total, int ctas_per_group) {
    using G = DpCfg<CIN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                   // [2][A_BYTES]
    uint8_t* sS = smem + 2 * G::A_BYTES;  // stacked kd=1,2 weights
    uint8_t* sZ = sS + G::S_BYTES;        // kd=0 weights
    uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + G::Z_BYTES);
    uint64_t* barW = bars;
    uint64_t* fullA = bars + 1;          // [2]
    uint64_t* emptyA = fullA + 2;        // [2]
    uint64_t* tmemFull = emptyA + 2;     // [2]
    uint64_t* tmemEmpty = tmemFull + 2;  // [2] (128 epilogue threads arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemEmpty + 2);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][32][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / ctas_per_group, rank = blockIdx.x - group * ctas_per_group;
    const int tiles_per_sample = tiles_w * tiles_h * D;

    if (threadIdx.x == 0) {
        mbar_init(barW, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
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
            // ===== TMA producer
            mbar_arrive_expect_tx(barW, G::S_BYTES + G::Z_BYTES);
            for (int tap = 0; tap < 27; ++tap) {
                const int kd = tap / 9, khkw = tap % 9;
                for (int c = 0; c < G::NC; ++c) {
                    const __nv_bfloat16* src = wpk + ((size_t)tap * G::NC + c) * cout_total * 8 + group * 256;
                    uint8_t* dst = (kd == 0) ? sZ + (khkw * G::NC + c) * 512 : sS + (khkw * G::NC + c) * 1024 + (kd - 1) * 512;
                    bulk_g2s(dst, src, 512, barW);
                }
            }
            int it = 0;
            for (int tile = rank; tile < total_tiles; tile += ctas_per_group, ++it) {
                const int b = tile / tiles_per_sample;
                int r = tile - b * tiles_per_sample;
                const int id = r / (tiles_w * tiles_h);
                r -= id * tiles_w * tiles_h;
                const int ty = r / tiles_w, tx = r - ty * tiles_w;
                const int buf = it & 1;
                if (it >= 2) mbar_wait(emptyA + buf, ((it >> 1) - 1) & 1);
                mbar_arrive_expect_tx(fullA + buf, G::A_BYTES);
                tma_load_5d(sA + buf * G::A_BYTES, &tmap_x, fullA + buf, tx * 8 * 8, ty * 16, id, 0, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged, one elected lane issues; immediates for every descriptor offset)
        constexpr uint32_t idescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idescZ = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(kDW * 16), b_hi = umma_desc_hi(128);
        const uint32_t s_lo = umma_desc_lo(smem_u32(sS), 1024), z_lo = umma_desc_lo(smem_u32(sZ), 512);
        mbar_wait(barW, 0);
        tc_fence_after();
        int it = 0;
        for (int tile = rank; tile < total_tiles; tile += ctas_per_group, ++it) {
            const int buf = it & 1;
            mbar_wait(fullA + buf, (it >> 1) & 1);
            if (it >= 2) mbar_wait(tmemEmpty + buf, ((it >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + buf * G::A_BYTES, G::CHUNK_BYTES);
            const uint32_t dcol = tmem_base + buf * 256;
            if (elect_one()) {
                uint32_t started = 0;  // bit per (ph,pw): accumulator pair already initialised (folds at compile time)
#pragma unroll
                for (int khkw = 0; khkw < 9; ++khkw) {  // kd = 1,2 (input plane id, classes pd = 0,1 side by side)
                    const int kh = khkw / 3, kw = khkw % 3;
                    const int sh = (kh == 0) ? 1 : 0, sw = (kw == 0) ? 1 : 0;
                    const int phpw = ((kh != 1) ? 2 : 0) | ((kw != 1) ? 1 : 0);
#pragma unroll
                    for (int kc = 0; kc < CIN / 16; ++kc) {
                        const uint64_t ad = umma_desc_at(a_lo, a_hi, (sh * kDW + sw) * 16 + 2 * kc * G::CHUNK_BYTES);
                        const uint64_t bd = umma_desc_at(s_lo, b_hi, (khkw * G::NC + 2 * kc) * 1024);
                        umma_bf16(dcol + phpw * 64, ad, bd, idescS, (((started >> phpw) & 1u) | (uint32_t)kc) ? 1u : 0u);
                    }
                    started |= 1u << phpw;
                }
#pragma unroll
                for (int khkw = 0; khkw < 9; ++khkw) {  // kd = 0 (input plane id+1, classes pd = 1)
                    const int kh = khkw / 3, kw = khkw % 3;
                    const int sh = (kh == 0) ? 1 : 0, sw = (kw == 0) ? 1 : 0;
                    const int phpw = ((kh != 1) ? 2 : 0) | ((kw != 1) ? 1 : 0);
#pragma unroll
                    for (int kc = 0; kc < CIN / 16; ++kc) {
                        const uint64_t ad = umma_desc_at(a_lo, a_hi, ((kDH + sh) * kDW + sw) * 16 + 2 * kc * G::CHUNK_BYTES);
                        const uint64_t bd = umma_desc_at(z_lo, b_hi, (khkw * G::NC + 2 * kc) * 512);
                        umma_bf16(dcol + phpw * 64 + 32, ad, bd, idescZ, 1u);
                    }
                }
                umma_commit(emptyA + buf);
                umma_commit(tmemFull + buf);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue warps 2..5
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..127
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
        const size_t oplane = (size_t)Ho * Wo;
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
        int it = 0;
        for (int tile = rank; tile < total_tiles; tile += ctas_per_group, ++it) {
            const int b = tile / tiles_per_sample;
            int r = tile - b * tiles_per_sample;
            const int id = r / (tiles_w * tiles_h);
            r -= id * tiles_w * tiles_h;
            const int ty = r / tiles_w, tx = r - ty * tiles_w;
            const int h = ty * 16 + (row >> 3), w = tx * 8 + (row & 7);
            const bool ok = (h < H) && (w < W);
            const int buf = it & 1;
            if (gn_sums != nullptr && b != cur_b) {
                if (cur_b >= 0) flush(cur_b);
                cur_b = b;
            }
            mbar_wait(tmemFull + buf, (it >> 1) & 1);
            tc_fence_after();
            float s[32], ss[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) s[c] = ss[c] = 0.f;
#pragma unroll 1
            for (int cls = 0; cls < 8; ++cls) {  // column block order: ((ph*2 + pw)*2 + pd)
                uint32_t v[32];
                tmem_ld_32x32b_x32(tlane + buf * 256 + cls * 32, v);
                if (ok) {
                    const int pd = cls & 1, pw = (cls >> 1) & 1, ph = cls >> 2;
                    const int od = 2 * id + pd, oh = 2 * h + ph, ow = 2 * w + pw;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            pk[e] = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]));
                            const float r0 = __low2float(pk[e]), r1 = __high2float(pk[e]);
                            s[j * 8 + 2 * e] += r0;
                            ss[j * 8 + 2 * e] = fmaf(r0, r0, ss[j * 8 + 2 * e]);
                            s[j * 8 + 2 * e + 1] += r1;
                            ss[j * 8 + 2 * e + 1] = fmaf(r1, r1, ss[j * 8 + 2 * e + 1]);
                        }
                        __nv_bfloat16* dst = y + ((((size_t)b * (cout_total >> 3) + group * 4 + j) * Do + od) * oplane +
                                                  (size_t)oh * Wo + ow) * 8;
                        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(pk);
                    }
                }
            }
            tc_fence_before();
            dp_mbar_arrive(tmemEmpty + buf);
            if (gn_sums != nullptr) {
                tot_s += (double)warp_transpose_sum32(s, lane);
                tot_q += (double)warp_transpose_sum32(ss, lane);
            }
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

// used by cmfb200_deconv3d_igemm_bf16_fwd (c8_s2_entry.cu); Cin = 64, Cout = 32 or 64
int deconv3d_igemm_persistent_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                       int D, int H, int W, cudaStream_t st) {
    CMF_REQUIRE(Cin == 64 && (Cout == 32 || Cout == 64), "deconv3d_igemm_persistent: unsupported (Cin=%d, Cout=%d)", Cin, Cout);
    using G = DpCfg<64>;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)G::NC * D * H * W * 16};
    const cuuint32_t box[5] = {kDW * 8, kDH, 2, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "deconv3d_igemm_persistent")) return rc;
    auto kern = deconv3d_igemm_persistent_kernel<64>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16);
    const long long total = (long long)tiles_w * tiles_h * D * B;
    CMF_REQUIRE(total < (1ll << 31), "deconv3d_igemm_persistent: too many tiles");
    const int groups = Cout / 32;
    long long per_group = sms / groups;
    if (per_group < 1) per_group = 1;
    if (per_group > total) per_group = total;
    kern<<<(unsigned)(per_group * groups), kIgThreads, G::SMEM_BYTES, st>>>(
        tmap, reinterpret_cast<const __nv_bfloat16*>(wpk), reinterpret_cast<__nv_bfloat16*>(y), gn, D, H, W, tiles_w,
        tiles_h, (int)total, Cout, (int)per_group);
    CMF_LAUNCH_CHECK("deconv3d_igemm_persistent_kernel");
    return CMFB200_OK;
}

}  // namespace cmfb200
