// K2 (throughput path), the two resolution-changing layers of the hourglass as tcgen05 implicit GEMMs on the
// C8/bf16 layout (see conv3d_igemm.cu for the layout / descriptor scheme):
//
//  * transposed conv k3 s2 p1 op1 (hourglass.conv5/conv6, cmf/models/cmfsm.py:261-281).  od = 2*id - 1 + kd, so an
//    output voxel of parity 0 along an axis takes tap k=1 of input i and parity 1 takes tap k=2 of input i and tap
//    k=0 of input i+1.  A CTA owns (one input depth slice, one depth parity pd) x 16 x 8 input positions and keeps
//    the FOUR (ph,pw) output-parity classes as four accumulator tiles in TMEM; every tap is one MMA chain into the
//    tile of its class with the A descriptor shifted by (sd,sh,sw) in {0,1}^3 on the halo'd input block
//    (box {9*8, 17, 2, C/8, 1}).  No zero insertion, 27 taps of work for 8 output voxels.
//
//  * conv k3 stride 2 (hourglass.conv1/conv3, :244-254).  Input index i = 2*o + k - 1: tap k=1 reads parity-0
//    inputs at o, taps k=0 / k=2 read parity-1 inputs at o-1 / o.  The producer layer writes a PARITY-SPLIT copy
//    [B][8 parities][C/8][D/2][H/2][W/2][8] (gn_apply_c8 `y_split`), which turns the strided gather into eight
//    dense sub-volumes: the kernel walks them as 8 pipeline stages (one TMA box each), each stage feeding the 1, 2,
//    4 or 8 taps that read that parity, all accumulating into the same TMEM tile.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

constexpr int kS2PW = 9, kS2PH = 17;  // 8 x 16 tile + one halo voxel on one side

// =====================================================================================================
// transposed convolution
// =====================================================================================================
template <int CIN, int COUT, int NS>
struct DeconvCfg {
    static constexpr int NC = CIN / 8;
    static constexpr int VOX = 2 * kS2PH * kS2PW;
    static constexpr int CHUNK_BYTES = VOX * 16;
    static constexpr int A_BYTES = NC * CHUNK_BYTES;
    static constexpr int TAP_BYTES = CIN * COUT * 2;
    static constexpr int TMEM_COLS = 4 * COUT;  // 128 or 256
    static constexpr int SMEM_BYTES = A_BYTES + NS * TAP_BYTES + 1024 + 4 * COUT * 2 * 8 + 1024;
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256, "TMEM allocation must be a power of two");
};

template <int CIN, int COUT, int NS>
__global__ void __launch_bounds__(kIgThreads, 2)
    deconv3d_igemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                               __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int D, int H, int W,
                               int tiles_w) {
    using G = DeconvCfg<CIN, COUT, NS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sW = smem + G::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + NS * G::TAP_BYTES);
    uint64_t* barA = bars;
    uint64_t* barD = bars + 1;
    uint64_t* full = bars + 2;
    uint64_t* empty = bars + 2 + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * NS);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [4][COUT][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * 8, h0 = tile_y * 16;
    const int id = blockIdx.y >> 1, pd = blockIdx.y & 1;
    const int b = blockIdx.z;
    const int nkd = pd ? 2 : 1;  // depth taps of this parity: pd=0 -> {k=1,+0}; pd=1 -> {k=2,+0}, {k=0,+1}
    const int ntap = nkd * 9;

    if (threadIdx.x == 0) {
        mbar_init(barA, 1);
        mbar_init(barD, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
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
            mbar_arrive_expect_tx(barA, G::A_BYTES);
            tma_load_5d(sA, &tmap_x, barA, w0 * 8, h0, id, 0, b);
            for (int t = 0; t < ntap; ++t) {
                const int s = t % NS;
                const int jd = t / 9, r = t % 9;
                const int kd = pd ? (jd == 0 ? 2 : 0) : 1;
                const int tap = kd * 9 + r;
                if (t >= NS) mbar_wait(empty + s, ((t / NS) - 1) & 1);
                mbar_arrive_expect_tx(full + s, G::TAP_BYTES);
                bulk_g2s(sW + s * G::TAP_BYTES, wpk + (size_t)tap * CIN * COUT, G::TAP_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp converged, one elected lane issues (see conv3d_igemm_persistent.cu)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_lo = umma_desc_lo(smem_u32(sA), G::CHUNK_BYTES), a_hi = umma_desc_hi(kS2PW * 16);
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), COUT * 16), b_hi = umma_desc_hi(128);
        uint32_t started = 0;  // bit per (ph,pw) class: accumulator already initialised
        mbar_wait(barA, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < ntap; ++t) {
            const int s = t % NS;
            const int jd = t / 9, kh = (t % 9) / 3, kw = t % 3;
            const int sd = (pd && jd == 1) ? 1 : 0;  // k=0 reads input i+1
            const int sh = (kh == 0) ? 1 : 0, sw = (kw == 0) ? 1 : 0;
            const int cls = ((kh != 1) ? 2 : 0) | ((kw != 1) ? 1 : 0);
            mbar_wait(full + s, (t / NS) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t arow = ((sd * kS2PH + sh) * kS2PW + sw) * 16;
#pragma unroll
                for (int kc = 0; kc < CIN / 16; ++kc) {
                    const uint64_t ad = umma_desc_at(a_lo, a_hi, arow + 2 * kc * G::CHUNK_BYTES);
                    const uint64_t bd = umma_desc_at(w_lo, b_hi, s * G::TAP_BYTES + 2 * kc * (COUT * 16));
                    umma_bf16(tmem_base + cls * COUT, ad, bd, idesc, (((started >> cls) & 1u) | (uint32_t)kc) ? 1u : 0u);
                }
                umma_commit(empty + s);
                if (t == ntap - 1) umma_commit(barD);
            }
            __syncwarp();
            started |= 1u << cls;
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int h = h0 + (row >> 3), w = w0 + (row & 7);
        const bool ok = (h < H) && (w < W);
        const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
        const size_t oplane = (size_t)Ho * Wo;
        const int od = 2 * id + pd;
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
            for (int cls = 0; cls < 4; ++cls) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + cls * COUT + half * 32, v);
                if (ok) {
                    const int oh = 2 * h + (cls >> 1), ow = 2 * w + (cls & 1);
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
                        __nv_bfloat16* dst =
                            y + ((((size_t)b * (COUT / 8) + chunk) * Do + od) * oplane + (size_t)oh * Wo + ow) * 8;
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

// =====================================================================================================
// stride-2 convolution on the parity-split input
// =====================================================================================================
template <int CIN, int COUT, int BD, int NSA, int NS>
struct S2Cfg {
    static constexpr int NC = CIN / 8;
    static constexpr int PD = BD + 1;
    static constexpr int VOX = PD * kS2PH * kS2PW;
    static constexpr int CHUNK_BYTES = VOX * 16;
    static constexpr int STAGE_BYTES = NC * CHUNK_BYTES;               // bytes moved by one TMA box
    static constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) & ~127;    // TMA shared-memory destinations: 128-byte aligned
    static constexpr int TAP_BYTES = CIN * COUT * 2;
    static constexpr int TMEM_COLS = (BD * COUT <= 64) ? 64 : (BD * COUT <= 128) ? 128 : (BD * COUT <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = NSA * STAGE_STRIDE + NS * TAP_BYTES + 1024 + 4 * COUT * 2 * 8 + 1024;
};

// taps that read input parity p along one axis: p=0 -> {k=1}; p=1 -> {k=0, k=2}
__device__ __forceinline__ int s2_ntaps(int p) { return p ? 2 : 1; }
__device__ __forceinline__ int s2_tap(int p, int j) { return p ? (j == 0 ? 0 : 2) : 1; }
// position of the tap's source inside the box (box origin = output index - 1): k=0 -> o-1 -> 0 ; k=1,2 -> o -> 1
__device__ __forceinline__ int s2_shift(int k) { return k == 0 ? 0 : 1; }

template <int CIN, int COUT, int BD, int NSA, int NS>
__global__ void __launch_bounds__(kIgThreads, (CIN == 32) ? 2 : 1)
    conv3d_s2_igemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_xs, const __nv_bfloat16* __restrict__ wpk,
                                __nv_bfloat16* __restrict__ y, double* __restrict__ gn_sums, int Do, int Ho, int Wo,
                                int tiles_w) {
    using G = S2Cfg<CIN, COUT, BD, NSA, NS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sW = smem + NSA * G::STAGE_STRIDE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + NS * G::TAP_BYTES);
    uint64_t* barD = bars;
    uint64_t* fullW = bars + 1;
    uint64_t* emptyW = fullW + NS;
    uint64_t* fullA = emptyW + NS;
    uint64_t* emptyA = fullA + NSA;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyA + NSA);
    double* sred = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x % tiles_w, tile_y = blockIdx.x / tiles_w;
    const int w0 = tile_x * 8, h0 = tile_y * 16, d0 = blockIdx.y * BD;
    const int b = blockIdx.z;

    if (threadIdx.x == 0) {
        mbar_init(barD, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(fullW + s, 1);
            mbar_init(emptyW + s, 1);
        }
        for (int s = 0; s < NSA; ++s) {
            mbar_init(fullA + s, 1);
            mbar_init(emptyA + s, 1);
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
            // ===== producer: 8 parity sub-volumes (A ring) interleaved with their weight taps (W ring)
            int t = 0;
            for (int par = 0; par < 8; ++par) {
                const int sa = par % NSA;
                if (par >= NSA) mbar_wait(emptyA + sa, ((par / NSA) - 1) & 1);
                mbar_arrive_expect_tx(fullA + sa, G::STAGE_BYTES);
                tma_load_5d(sA + sa * G::STAGE_STRIDE, &tmap_xs, fullA + sa, (w0 - 1) * 8, h0 - 1, d0 - 1, par * G::NC, b);
                const int pd = par >> 2, ph = (par >> 1) & 1, pw = par & 1;
                for (int jd = 0; jd < s2_ntaps(pd); ++jd)
                    for (int jh = 0; jh < s2_ntaps(ph); ++jh)
                        for (int jw = 0; jw < s2_ntaps(pw); ++jw, ++t) {
                            const int tap = s2_tap(pd, jd) * 9 + s2_tap(ph, jh) * 3 + s2_tap(pw, jw);
                            const int s = t % NS;
                            if (t >= NS) mbar_wait(emptyW + s, ((t / NS) - 1) & 1);
                            mbar_arrive_expect_tx(fullW + s, G::TAP_BYTES);
                            bulk_g2s(sW + s * G::TAP_BYTES, wpk + (size_t)tap * CIN * COUT, G::TAP_BYTES, fullW + s);
                        }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp converged, one elected lane issues
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_hi = umma_desc_hi(kS2PW * 16), b_hi = umma_desc_hi(128);
        const uint32_t w_lo = umma_desc_lo(smem_u32(sW), COUT * 16);
        int t = 0;
#pragma unroll 1
        for (int par = 0; par < 8; ++par) {
            const int sa = par % NSA;
            const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + sa * G::STAGE_STRIDE, G::CHUNK_BYTES);
            mbar_wait(fullA + sa, (par / NSA) & 1);
            tc_fence_after();
            const int pd = par >> 2, ph = (par >> 1) & 1, pw = par & 1;
            const int ntp = s2_ntaps(pd) * s2_ntaps(ph) * s2_ntaps(pw);
#pragma unroll 1
            for (int j = 0; j < ntp; ++j, ++t) {
                // same (jd, jh, jw) enumeration order as the producer
                const int jw = j % s2_ntaps(pw), jh = (j / s2_ntaps(pw)) % s2_ntaps(ph), jd = j / (s2_ntaps(pw) * s2_ntaps(ph));
                const int sd = s2_shift(s2_tap(pd, jd)), sh = s2_shift(s2_tap(ph, jh)), sw = s2_shift(s2_tap(pw, jw));
                const int s = t % NS;
                mbar_wait(fullW + s, (t / NS) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t arow = (((sd * kS2PH + sh) * kS2PW) + sw) * 16;
#pragma unroll
                    for (int mt = 0; mt < BD; ++mt) {
#pragma unroll
                        for (int kc = 0; kc < CIN / 16; ++kc) {
                            const uint64_t ad = umma_desc_at(a_lo, a_hi, arow + mt * kS2PH * kS2PW * 16 + 2 * kc * G::CHUNK_BYTES);
                            const uint64_t bd = umma_desc_at(w_lo, b_hi, s * G::TAP_BYTES + 2 * kc * (COUT * 16));
                            umma_bf16(tmem_base + mt * COUT, ad, bd, idesc, (t | kc) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(emptyW + s);
                    if (j == ntp - 1) umma_commit(emptyA + sa);  // this parity sub-volume may be overwritten
                    if (t == 26) umma_commit(barD);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int h = h0 + (row >> 3), w = w0 + (row & 7);
        const bool hw_ok = (h < Ho) && (w < Wo);
        const size_t plane = (size_t)Ho * Wo;
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
                        const int chunk = half * 4 + j;
                        __nv_bfloat16* dst = y + ((((size_t)b * (COUT / 8) + chunk) * Do + d) * plane + (size_t)h * Wo + w) * 8;
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

// C8 -> parity-split C8: [B][C/8][D][H][W][8] -> [B][8][C/8][D/2][H/2][W/2][8]   (D,H,W even)
__global__ void c8_parity_split_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int NC,
                                       int D, int H, int W) {
    const size_t spatial = (size_t)D * H * W;
    const size_t bc = blockIdx.y;  // b * NC + chunk
    const size_t b = bc / NC, chunk = bc % NC;
    const int D2 = D / 2, H2 = H / 2, W2 = W / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < spatial; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)(i / ((size_t)W * H));
        const int par = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
        const size_t dst = ((((b * 8 + par) * NC + chunk) * D2 + (d >> 1)) * H2 + (h >> 1)) * W2 + (w >> 1);
        *reinterpret_cast<uint4*>(y + dst * 8) = *reinterpret_cast<const uint4*>(x + (bc * spatial + i) * 8);
    }
}

template <int CIN, int COUT, int NS>
static int launch_deconv_igemm(const void* x, const void* wpk, void* y, double* gn, int B, int D, int H, int W,
                               cudaStream_t st) {
    using G = DeconvCfg<CIN, COUT, NS>;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                                (cuuint64_t)G::NC * D * H * W * 16};
    const cuuint32_t box[5] = {kS2PW * 8, kS2PH, 2, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "deconv3d_igemm")) return rc;
    auto kern = deconv3d_igemm_bf16_kernel<CIN, COUT, NS>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16);
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(2 * D), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "deconv3d_igemm: grid too large");
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(wpk),
                                                  reinterpret_cast<__nv_bfloat16*>(y), gn, D, H, W, tiles_w);
    CMF_LAUNCH_CHECK("deconv3d_igemm_bf16_kernel");
    return CMFB200_OK;
}

template <int CIN, int COUT, int BD, int NSA, int NS>
static int launch_s2_igemm(const void* xs, const void* wpk, void* y, double* gn, int B, int Do, int Ho, int Wo,
                           cudaStream_t st) {
    using G = S2Cfg<CIN, COUT, BD, NSA, NS>;
    static_assert(G::SMEM_BYTES <= 227 * 1024, "stride-2 igemm tile does not fit in shared memory");
    CUtensorMap tmap;
    const cuuint64_t vol = (cuuint64_t)Do * Ho * Wo * 16;
    const cuuint64_t gdim[5] = {(cuuint64_t)Wo * 8, (cuuint64_t)Ho, (cuuint64_t)Do, (cuuint64_t)8 * G::NC, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)Wo * 16, (cuuint64_t)Ho * Wo * 16, vol, vol * 8 * G::NC};
    const cuuint32_t box[5] = {kS2PW * 8, kS2PH, (cuuint32_t)G::PD, (cuuint32_t)G::NC, 1};
    if (int rc = encode_tmap_5d(&tmap, xs, gdim, gstr, box, "conv3d_s2_igemm")) return rc;
    auto kern = conv3d_s2_igemm_bf16_kernel<CIN, COUT, BD, NSA, NS>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(Wo, 8), tiles_h = (int)cdiv(Ho, 16);
    dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)cdiv(Do, BD), (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d_s2_igemm: grid too large");
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(wpk),
                                                  reinterpret_cast<__nv_bfloat16*>(y), gn, Do, Ho, Wo, tiles_w);
    CMF_LAUNCH_CHECK("conv3d_s2_igemm_bf16_kernel");
    return CMFB200_OK;
}

int conv3d_s2_igemm_persistent_dispatch(const void* xs, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                        int Do, int Ho, int Wo, cudaStream_t st);  // conv3d_s2_igemm_persistent.cu
int deconv3d_igemm_persistent_dispatch(const void* x, const void* wpk, void* y, double* gn, int B, int Cin, int Cout,
                                       int D, int H, int W, cudaStream_t st);  // deconv3d_igemm_persistent.cu

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_deconv3d_igemm_bf16_fwd(const void* x_c8, const void* packed_w, void* y_c8, double* gn_sums,
                                               int B, int Cin, int Cout, int D, int H, int W, void* stream) {
    CMF_REQUIRE(x_c8 && packed_w && y_c8, "deconv3d_igemm_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "deconv3d_igemm_bf16_fwd: non-positive dimension");
    cudaStream_t st = (cudaStream_t)stream;
    static const bool simple_schedule = getenv("CMFB200_IGEMM_SIMPLE") != nullptr;  // A/B switch: one tile per CTA
    if (!simple_schedule && Cin == 64 && (Cout == 32 || Cout == 64))
        return deconv3d_igemm_persistent_dispatch(x_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, D, H, W, st);
    if (Cin == 64 && Cout == 64) return launch_deconv_igemm<64, 64, 8>(x_c8, packed_w, y_c8, gn_sums, B, D, H, W, st);
    if (Cin == 64 && Cout == 32) return launch_deconv_igemm<64, 32, 16>(x_c8, packed_w, y_c8, gn_sums, B, D, H, W, st);
    CMF_REQUIRE(false, "deconv3d_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 64->64, 64->32", Cin, Cout);
}

extern "C" int cmfb200_conv3d_s2_igemm_bf16_fwd(const void* x_split_c8, const void* packed_w, void* y_c8,
                                                double* gn_sums, int B, int Cin, int Cout, int Do, int Ho, int Wo,
                                                void* stream) {
    CMF_REQUIRE(x_split_c8 && packed_w && y_c8, "conv3d_s2_igemm_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Do > 0 && Ho > 0 && Wo > 0, "conv3d_s2_igemm_bf16_fwd: non-positive dimension");
    cudaStream_t st = (cudaStream_t)stream;
    static const bool simple_schedule = getenv("CMFB200_IGEMM_SIMPLE") != nullptr;  // A/B switch: one tile per CTA
    if (!simple_schedule && Cout == 64 && (Cin == 32 || Cin == 64))
        return conv3d_s2_igemm_persistent_dispatch(x_split_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, Do, Ho, Wo, st);
    if (Cin == 32 && Cout == 64)
        return launch_s2_igemm<32, 64, 2, 2, 10>(x_split_c8, packed_w, y_c8, gn_sums, B, Do, Ho, Wo, st);
    if (Cin == 64 && Cout == 64)
        return launch_s2_igemm<64, 64, 2, 2, 10>(x_split_c8, packed_w, y_c8, gn_sums, B, Do, Ho, Wo, st);
    CMF_REQUIRE(false, "conv3d_s2_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->64, 64->64", Cin, Cout);
}

extern "C" int cmfb200_c8_parity_split(const void* x_c8, void* y_split_c8, int B, int C, int D, int H, int W,
                                       void* stream) {
    CMF_REQUIRE(x_c8 && y_split_c8, "c8_parity_split: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && D > 0 && H > 0 && W > 0, "c8_parity_split: bad shape");
    CMF_REQUIRE((D % 2 == 0) && (H % 2 == 0) && (W % 2 == 0), "c8_parity_split: D,H,W must be even (got %d,%d,%d)", D, H, W);
    CMF_REQUIRE((long long)B * (C / 8) <= 65535, "c8_parity_split: B*C/8 exceeds grid limit");
    const long long spatial = (long long)D * H * W;
    dim3 grid((unsigned)min((long long)kNumSMs * 8, cdiv(spatial, 256)), (unsigned)(B * (C / 8)));
    c8_parity_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_c8),
                                                                   reinterpret_cast<__nv_bfloat16*>(y_split_c8), C / 8,
                                                                   D, H, W);
    CMF_LAUNCH_CHECK("c8_parity_split_kernel");
    return CMFB200_OK;
}
