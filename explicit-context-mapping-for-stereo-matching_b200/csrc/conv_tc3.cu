// fp32-ACCURATE convolution on the 5th-generation tensor cores ("tc3": three-term bf16 split, tcgen05 / TMEM / TMA).
//
// Replaces the stride-1 nn.Conv2d layers of feature_extraction (cmf/models/cmfsm.py:126-236) and the stride-1
// nn.Conv3d layers of the aggregation network (:49-58, :240-303, :604-634) in the fp32 PARITY mode.  bf16/TF32
// operand rounding was measured at 0.06-0.9 px, so fp32 parity needs fp32-accurate products:
//
//     a = a0 + a1 + a2,  w = w0 + w1 + w2      (bf16 terms: a0 = bf16(a), a1 = bf16(a - a0), a2 = bf16(a - a0 - a1);
//                                               8 + 8 + 8 significand bits = the fp32 value EXACTLY)
//     a * w = a0 w0 + (a0 w1 + a1 w0) + (a0 w2 + a1 w1 + a2 w0) + O(2^-24 |a w|)
//
// Six bf16 x bf16 -> fp32 products per fp32 product, every one exact in the tensor core, accumulated in fp32 in TMEM.
// The accumulator rounds toward zero with a few guard bits (profiles/r02_probe_tc_rounding.txt: rms error of a
// 864-term dot product 2.5e-7 vs 2.1e-7 for an FFMA chain), so the dominant term a0 w0 gets its OWN accumulator and
// the five small terms (2^-8 and 2^-16 of the magnitude, their rounding errors scaled alike) share the others; the
// epilogue adds the accumulators in fp32 (smallest first).
//
// Layouts (2-D images are volumes with D = 1):
//   activations "C8S3": bf16 [B][C/8][3 terms][D][H][W][8]  -- an (8-channel group, term) plane is dense, a voxel's 8
//       channels are one 16-byte unit: every im2col row of every tap is a 16-byte unit at constant pitch = the
//       no-swizzle K-major UMMA canonical layout; ONE rank-5 TMA box {(8+2d)*8, 16+2d, 1, 6, 1} per (tile, 16-channel
//       K step) lands both 8-channel chunks x three terms with the halo, zero OOB fill = the conv padding;
//   raw conv output "C8F": fp32 [B][C/8][D][H][W][8] (+ per-(b,channel) double sum / sum of squares for GroupNorm);
//   weights: bf16 [K steps = KD * Cin/16][KH][KW][2 chunks][3 terms][Cout][8]: the three terms of one chunk are
//       adjacent 8-row groups, so "A-term x [w0|w1|w2]" is ONE MMA of N = 3 Cout ("term stacking": the A tile, 4 KB
//       per MMA through the 128 B/clk shared-memory port, is read 3x instead of 6x per tap):
//           Cout <= 64 :  a0 x [w0|w1|w2] -> cols [0,3C)   a1 x [w0|w1] -> cols [C,3C)   a2 x [w0] -> cols [2C,3C)
//           Cout = 128 :  six N = 128 MMAs, a0 w0 -> cols [0,C), the five small terms -> cols [C,2C)
// Work per CTA: T adjacent 16 x 8 output tiles (M = 128 each) x all Cout; a 3-D conv is the same loop with K steps
// running over (kd, Cin/16) and the TMA box taken from depth plane d + kd - 1.  Two rings: A (one K step of all T
// tiles per slot) and B (one (K step, kh) row of 3 taps per slot).  Warp roles: w0 TMA producer, w1 TMEM allocator +
// MMA issuer (one elected lane), w2-5 epilogue (tcgen05.ld -> fp32 adds -> C8F / NCHW store; GroupNorm partials in
// DOUBLE from the first element through a transposing warp reduction, one double atomic per channel per CTA).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {

template <int COUT, int T, int DIL, int KHW, bool STACK, int NA, int NB>
struct Tc3Cfg {
    static constexpr int HALO = (KHW == 3) ? DIL : 0;
    static constexpr int PW = 8 + 2 * HALO, PH = 16 + 2 * HALO;
    static constexpr int PLANE = PH * PW * 16;          // one (8-channel chunk, term) plane of a tile
    static constexpr int A_TILE = 6 * PLANE;            // 2 chunks x 3 terms
    static constexpr int A_STAGE = T * A_TILE;
    static constexpr int B_TAP = 6 * COUT * 16;         // one tap: 2 chunks x 3 terms x Cout rows x 16 B
    static constexpr int B_STAGE = KHW * B_TAP;         // one kh row
    static constexpr int ACC = (STACK ? 3 : 2) * COUT;  // TMEM columns per tile
    static constexpr int TMEM_NEED = T * ACC;
    static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128
                                     : TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int SRED_BYTES = 4 * COUT * 2 * 8;  // aliases the A ring after the last MMA
    static constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_STAGE + 1024 /*barriers + tmem slot*/ + 1024 /*align*/;
    static_assert(TMEM_NEED <= 512, "accumulators exceed TMEM");
    static_assert(A_TILE % 128 == 0 && B_STAGE % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(!STACK || 3 * COUT <= 256, "stacked N exceeds the UMMA limit");
    static_assert(SRED_BYTES <= NA * A_STAGE, "reduction scratch must fit in the A ring");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__host__ __device__ constexpr uint32_t tc3_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace

template <int COUT, int T, int DIL, int KHW, bool STACK, int NA, int NB, bool OUT_NCHW, int MINB>
__global__ void __launch_bounds__(kIgThreads, MINB)
    conv_tc3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __nv_bfloat16* __restrict__ wpk,
                    float* __restrict__ y, double* __restrict__ gn_sums, int D, int H, int W, int groups_w, int KD,
                    int KC, int row_off) {
    // H = OUTPUT rows; the input may carry extra (halo) rows: output row h reads input rows h + row_off - halo ..
    using G = Tc3Cfg<COUT, T, DIL, KHW, STACK, NA, NB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // [NA][T][6 planes]
    uint8_t* sB = smem + NA * G::A_STAGE;     // [NB][KHW taps][2][3][COUT][16 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NB * G::B_STAGE);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + NA;
    uint64_t* fullB = emptyA + NA;
    uint64_t* emptyB = fullB + NB;
    uint64_t* tmemFull = emptyB + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemFull + 1);
    double* sred = reinterpret_cast<double*>(sA);  // [4 quads][COUT][2], used after the last MMA has completed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x % groups_w, ty = blockIdx.x / groups_w;
    const int tx0 = gx * T, d = blockIdx.y, b = blockIdx.z;
    const int KS = KD * KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
        }
        for (int i = 0; i < NB; ++i) {
            mbar_init(fullB + i, 1);
            mbar_init(emptyB + i, 1);
        }
        mbar_init(tmemFull, 1);
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
            // ===== TMA producer: per K step the T activation tiles, then the KHW weight rows
            for (int ks = 0; ks < KS; ++ks) {
                const int kd = ks / KC, kc = ks - kd * KC;
                const int sa = ks % NA;
                if (ks >= NA) mbar_wait(emptyA + sa, ((ks / NA) - 1) & 1);
                mbar_arrive_expect_tx(fullA + sa, G::A_STAGE);
                const int dz = d + kd - (KD >> 1);
#pragma unroll
                for (int t = 0; t < T; ++t)
                    tma_load_5d(sA + sa * G::A_STAGE + t * G::A_TILE, &tmap_x, fullA + sa,
                                ((tx0 + t) * 8 - G::HALO) * 8, ty * 16 + row_off - G::HALO, dz, kc * 6, b);
                for (int kh = 0; kh < KHW; ++kh) {
                    const int gb = ks * KHW + kh, sb = gb % NB;
                    if (gb >= NB) mbar_wait(emptyB + sb, ((gb / NB) - 1) & 1);
                    mbar_arrive_expect_tx(fullB + sb, G::B_STAGE);
                    bulk_g2s(sB + sb * G::B_STAGE, reinterpret_cast<const uint8_t*>(wpk) + (size_t)gb * G::B_STAGE,
                             G::B_STAGE, fullB + sb);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp converged; one elected lane issues)
        const uint32_t a_hi = umma_desc_hi(G::PW * 16), b_hi = umma_desc_hi(128);
        for (int ks = 0; ks < KS; ++ks) {
            const int sa = ks % NA;
            mbar_wait(fullA + sa, (ks / NA) & 1);
            const uint32_t a_lo = umma_desc_lo(smem_u32(sA) + sa * G::A_STAGE, 3 * G::PLANE);
#pragma unroll 1
            for (int kh = 0; kh < KHW; ++kh) {
                const int gb = ks * KHW + kh, sb = gb % NB;
                mbar_wait(fullB + sb, (gb / NB) & 1);
                tc_fence_after();
                const uint32_t w_lo = umma_desc_lo(smem_u32(sB) + sb * G::B_STAGE, 3 * COUT * 16);
                const uint32_t accum = (ks | kh) != 0 ? 1u : 0u;  // 0 only for the very first tap row (zero-initialises the accumulators)
                if (elect_one()) {
#pragma unroll
                    for (int kw = 0; kw < KHW; ++kw) {
                        const uint32_t acc_first = (kw == 0) ? accum : 1u;
                        const uint32_t a_off = (KHW == 3) ? (uint32_t)((kh * DIL * G::PW + kw * DIL) * 16) : 0u;
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const uint32_t ab = t * G::A_TILE + a_off;
                            const uint32_t wb = kw * G::B_TAP;
                            const uint32_t dcol = tmem_base + t * G::ACC;
                            const uint64_t a0 = umma_desc_at(a_lo, a_hi, ab);
                            const uint64_t a1 = umma_desc_at(a_lo, a_hi, ab + G::PLANE);
                            const uint64_t a2 = umma_desc_at(a_lo, a_hi, ab + 2 * G::PLANE);
                            const uint64_t w0 = umma_desc_at(w_lo, b_hi, wb);
                            if constexpr (STACK) {
                                umma_bf16(dcol, a0, w0, tc3_idesc(3 * COUT), acc_first);
                                umma_bf16(dcol + COUT, a1, w0, tc3_idesc(2 * COUT), 1u);
                                umma_bf16(dcol + 2 * COUT, a2, w0, tc3_idesc(COUT), 1u);
                            } else {
                                const uint64_t w1 = umma_desc_at(w_lo, b_hi, wb + COUT * 16);
                                const uint64_t w2 = umma_desc_at(w_lo, b_hi, wb + 2 * COUT * 16);
                                umma_bf16(dcol, a0, w0, tc3_idesc(COUT), acc_first);
                                umma_bf16(dcol + COUT, a0, w1, tc3_idesc(COUT), acc_first);
                                umma_bf16(dcol + COUT, a1, w0, tc3_idesc(COUT), 1u);
                                umma_bf16(dcol + COUT, a0, w2, tc3_idesc(COUT), 1u);
                                umma_bf16(dcol + COUT, a1, w1, tc3_idesc(COUT), 1u);
                                umma_bf16(dcol + COUT, a2, w0, tc3_idesc(COUT), 1u);
                            }
                        }
                    }
                    umma_commit(emptyB + sb);
                    if (kh == KHW - 1) {
                        umma_commit(emptyA + sa);
                        if (ks == KS - 1) umma_commit(tmemFull);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..5
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..127
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const size_t plane = (size_t)H * W;
        double tot_s[COUT / 32], tot_q[COUT / 32];  // lane l: channel cb*32 + l
#pragma unroll
        for (int cb = 0; cb < COUT / 32; ++cb) tot_s[cb] = tot_q[cb] = 0.0;
        mbar_wait(tmemFull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const int h = ty * 16 + (row >> 3), w = (tx0 + t) * 8 + (row & 7);
            const bool ok = (h < H) && (w < W);
#pragma unroll 1
            for (int cb = 0; cb < COUT / 32; ++cb) {
                float o[32];
                {
                    uint32_t v0[32], v1[32];
                    const uint32_t base = tlane + t * G::ACC + cb * 32;
                    if constexpr (STACK) {
                        tmem_ld_32x32b_x32_issue(base + 2 * COUT, v0);
                        tmem_ld_32x32b_x32_issue(base + COUT, v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v0[c]) + __uint_as_float(v1[c]);
                        tmem_ld_32x32b_x32(base, v0);
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] += __uint_as_float(v0[c]);
                    } else {
                        tmem_ld_32x32b_x32_issue(base + COUT, v1);
                        tmem_ld_32x32b_x32_issue(base, v0);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v1[c]) + __uint_as_float(v0[c]);
                    }
                }
                if (ok) {
                    if constexpr (OUT_NCHW) {
                        float* dst = y + (((size_t)b * COUT + cb * 32) * D + d) * plane + (size_t)h * W + w;
#pragma unroll
                        for (int c = 0; c < 32; ++c) dst[(size_t)c * D * plane] = o[c];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float* dst = y + ((((size_t)b * (COUT / 8) + cb * 4 + j) * D + d) * plane + (size_t)h * W + w) * 8;
                            *reinterpret_cast<float4*>(dst) = make_float4(o[j * 8], o[j * 8 + 1], o[j * 8 + 2], o[j * 8 + 3]);
                            *reinterpret_cast<float4*>(dst + 4) =
                                make_float4(o[j * 8 + 4], o[j * 8 + 5], o[j * 8 + 6], o[j * 8 + 7]);
                        }
                    }
                }
                if (gn_sums != nullptr) {
                    // 32-position partials in fp32 (relative rounding ~3e-7 of a 32-term partial, random in sign),
                    // accumulated in double across tiles / CTAs: the statistics of the >= 1e4 values per channel these
                    // layers have keep ~1e-8 relative accuracy (the tiny SPP-branch GroupNorms, where partials in fp32
                    // were measured to hurt, run on the FFMA kernels with double-from-the-first-element sums)
                    float q[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        o[c] = ok ? o[c] : 0.f;
                        q[c] = o[c] * o[c];
                    }
                    tot_s[cb] += (double)warp_transpose_sum32(o, lane);
                    tot_q[cb] += (double)warp_transpose_sum32(q, lane);
                }
            }
        }
        if (gn_sums != nullptr) {
            // every MMA (= every read of the A ring) has completed: the ring is free to hold the reduction scratch
#pragma unroll
            for (int cb = 0; cb < COUT / 32; ++cb) {
                sred[((quad * COUT) + cb * 32 + lane) * 2 + 0] = tot_s[cb];
                sred[((quad * COUT) + cb * 32 + lane) * 2 + 1] = tot_q[cb];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < COUT * 2; i += 128) {
                const int c = i >> 1, which = i & 1;
                double a = 0.0;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) a += sred[(qd * COUT + c) * 2 + which];
                atomicAdd(gn_sums + ((size_t)b * COUT + c) * 2 + which, a);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(G::TMEM_COLS)
                     : "memory");
    }
}

// ---- weights: fp32 [Cout][Cin][KD][KH][KW] -> bf16 [KD*Cin/16][KH][KW][2][3][Cout][8] ---------------------------
__global__ void pack_tc3_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int Cout, int Cin,
                                       int KD, int KHW) {
    const int total = KD * KHW * KHW * Cin * Cout;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // i enumerates the (ks, kh, kw, k8, co, e) grid of ONE term plane
    int r = i;
    const int e = r % 8;
    r /= 8;
    const int co = r % Cout;
    r /= Cout;
    const int k8 = r % 2;
    r /= 2;
    const int kw = r % KHW;
    r /= KHW;
    const int kh = r % KHW;
    const int ks = r / KHW;
    const int KC = Cin / 16;
    const int kd = ks / KC, kc = ks % KC;
    const int ci = kc * 16 + k8 * 8 + e;
    const float v = w[((((size_t)co * Cin + ci) * KD + kd) * KHW + kh) * KHW + kw];
    const __nv_bfloat16 t0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(t0);
    const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
    const size_t base = ((((((size_t)ks * KHW + kh) * KHW + kw) * 2 + k8) * 3) * Cout + co) * 8 + e;
    p[base] = t0;
    p[base + (size_t)Cout * 8] = t1;
    p[base + (size_t)2 * Cout * 8] = t2;
}

// ---- GroupNorm apply of the tc3 pipeline -------------------------------------------------------------------------
// raw (C8F or NCHW fp32) -> y = GroupNorm(raw) (+residual) (ReLU), written as C8S3 (split into three bf16 terms)
// and/or NCHW fp32.  gn_sums == nullptr: no normalisation (pure layout conversion / split).
struct GnTc3Args {
    const float* raw;
    const double* sums;
    const float* gamma;
    const float* beta;
    const __nv_bfloat16* res_s3;
    const float* res_nchw;
    __nv_bfloat16* y_s3;
    __nv_bfloat16* y_split;  // parity-split C8S3 [B][8][C/8][3][D/2][H/2][W/2][8] (input of the stride-2 tensor-core conv)
    float* y_nchw;
    int C, cpg;
    long long spatial;
    float eps;
    int relu, raw_c8f;
    int pad, H, W;  // row bands: y_s3 / res_s3 carry `pad` extra rows above and below the H rows (0 = dense)
    int nchw_padded;  // row bands: y_nchw carries the same `pad` rows (the 32->1 tail reads its halo rows in place)
    // row bands, fused halo push: the first / last `push_rows` rows of the C8S3 result are ALSO stored into the neighbour
    // ranks' landing buffers (peer-mapped memory over NVLink), dense [B*C/8][3][D][push_rows][W][8]; either may be null
    __nv_bfloat16* push_up;
    __nv_bfloat16* push_dn;
    int push_rows;
};

__device__ __forceinline__ void split3(float v, __nv_bfloat16& t0, __nv_bfloat16& t1, __nv_bfloat16& t2) {
    t0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(t0);
    t1 = __float2bfloat16_rn(r1);
    t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
}

template <bool PUSH>
__global__ void __launch_bounds__(256, 4) gn_apply_tc3_kernel(const GnTc3Args a) {
    __shared__ float s_scale[8], s_shift[8];
    const int bg = blockIdx.y;  // b * C/8 + g
    const int NG = a.C / 8;
    const int b = bg / NG, g = bg - b * NG;
    if (threadIdx.x < 8) {
        const int c = g * 8 + threadIdx.x;
        float scale = 1.f, shift = 0.f;
        if (a.sums != nullptr) {
            const int g0 = (c / a.cpg) * a.cpg;
            double s = 0.0, ss = 0.0;
            for (int j = 0; j < a.cpg; ++j) {
                s += a.sums[2 * ((size_t)b * a.C + g0 + j)];
                ss += a.sums[2 * ((size_t)b * a.C + g0 + j) + 1];
            }
            const double n = (double)a.cpg * (double)a.spatial;
            const double mean = s / n;
            double var = ss / n - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const double rstd = rsqrt(var + (double)a.eps);
            scale = (float)(rstd * (double)a.gamma[c]);
            shift = (float)((double)a.beta[c] - mean * rstd * (double)a.gamma[c]);
        }
        s_scale[threadIdx.x] = scale;
        s_shift[threadIdx.x] = shift;
    }
    __syncthreads();
    // 32-bit position arithmetic (spatial < 2^31 is checked on the host); two positions per thread and iteration, every
    // load issued before the first use (the kernel is a pure HBM stream: bytes in flight are what matters)
    const unsigned S = (unsigned)a.spatial;
    const unsigned H = (unsigned)a.H, W = (unsigned)a.W, hw = H * W;
    const unsigned Hp = H + 2 * a.pad;
    const size_t Sp = a.pad ? (size_t)(S / H) * Hp : S;  // positions of a padded (8-channel group, term) plane
    const bool need_dhw = a.pad != 0 || a.y_split != nullptr;
    constexpr int U = 2;
    const unsigned step = gridDim.x * blockDim.x;
    for (unsigned p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < S; p0 += U * step) {
        float v[U][8], rn[U][8];
        uint4 rq[U][3];
        unsigned pos[U];
        size_t pp[U];
        bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pos[u] = p0 + u * step;
            on[u] = pos[u] < S;
            const unsigned p = on[u] ? pos[u] : 0u;
            pp[u] = p;
            if (a.pad) {
                const unsigned dz = p / hw, r = p - dz * hw;
                pp[u] = (size_t)dz * Hp * W + (size_t)a.pad * W + r;
            }
            if (a.raw_c8f) {
                const float4 lo = ld_streaming_f4(a.raw + ((size_t)bg * S + p) * 8);
                const float4 hi = ld_streaming_f4(a.raw + ((size_t)bg * S + p) * 8 + 4);
                v[u][0] = lo.x, v[u][1] = lo.y, v[u][2] = lo.z, v[u][3] = lo.w;
                v[u][4] = hi.x, v[u][5] = hi.y, v[u][6] = hi.z, v[u][7] = hi.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[u][e] = __ldg(a.raw + ((size_t)b * a.C + g * 8 + e) * S + p);
            }
            if (a.res_s3 != nullptr) {
                const size_t base = ((size_t)bg * 3 * Sp + pp[u]) * 8;
                rq[u][0] = *reinterpret_cast<const uint4*>(a.res_s3 + base);
                rq[u][1] = *reinterpret_cast<const uint4*>(a.res_s3 + base + Sp * 8);
                rq[u][2] = *reinterpret_cast<const uint4*>(a.res_s3 + base + 2 * Sp * 8);
            }
            if (a.res_nchw != nullptr) {
#pragma unroll
                for (int e = 0; e < 8; ++e) rn[u][e] = __ldg(a.res_nchw + ((size_t)b * a.C + g * 8 + e) * S + p);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!on[u]) continue;
            const unsigned p = pos[u];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] = fmaf(v[u][e], s_scale[e], s_shift[e]);
            if (a.res_s3 != nullptr) {
                const __nv_bfloat16* r0 = reinterpret_cast<const __nv_bfloat16*>(&rq[u][0]);
                const __nv_bfloat16* r1 = reinterpret_cast<const __nv_bfloat16*>(&rq[u][1]);
                const __nv_bfloat16* r2 = reinterpret_cast<const __nv_bfloat16*>(&rq[u][2]);
#pragma unroll
                for (int e = 0; e < 8; ++e)  // the three terms reconstruct the fp32 residual exactly
                    v[u][e] += (__bfloat162float(r0[e]) + __bfloat162float(r1[e])) + __bfloat162float(r2[e]);
            }
            if (a.res_nchw != nullptr) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[u][e] += rn[u][e];
            }
            if (a.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[u][e] = fmaxf(v[u][e], 0.f);
            }
            if (a.y_s3 != nullptr || a.y_split != nullptr) {
                __align__(16) __nv_bfloat16 t0[8], t1[8], t2[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3(v[u][e], t0[e], t1[e], t2[e]);
                if (a.y_s3 != nullptr) {
                    const size_t base = ((size_t)bg * 3 * Sp + pp[u]) * 8;
                    *reinterpret_cast<uint4*>(a.y_s3 + base) = *reinterpret_cast<const uint4*>(t0);
                    *reinterpret_cast<uint4*>(a.y_s3 + base + Sp * 8) = *reinterpret_cast<const uint4*>(t1);
                    *reinterpret_cast<uint4*>(a.y_s3 + base + 2 * Sp * 8) = *reinterpret_cast<const uint4*>(t2);
                    if (PUSH && a.push_rows > 0) {  // boundary rows go straight to the neighbours (stores over NVLink)
                        const unsigned dz = p / hw, r = p - dz * hw, h = r / W, w = r - h * W;
                        const unsigned pr = (unsigned)a.push_rows;
                        const size_t Ss = (size_t)(S / H) * pr;  // positions of a pushed (group, term) plane
                        if (a.push_up != nullptr && h < pr) {
                            const size_t q = ((size_t)bg * 3 * Ss + ((size_t)dz * pr + h) * W + w) * 8;
                            *reinterpret_cast<uint4*>(a.push_up + q) = *reinterpret_cast<const uint4*>(t0);
                            *reinterpret_cast<uint4*>(a.push_up + q + Ss * 8) = *reinterpret_cast<const uint4*>(t1);
                            *reinterpret_cast<uint4*>(a.push_up + q + 2 * Ss * 8) = *reinterpret_cast<const uint4*>(t2);
                        }
                        if (a.push_dn != nullptr && h + pr >= H) {
                            const size_t q = ((size_t)bg * 3 * Ss + ((size_t)dz * pr + (h + pr - H)) * W + w) * 8;
                            *reinterpret_cast<uint4*>(a.push_dn + q) = *reinterpret_cast<const uint4*>(t0);
                            *reinterpret_cast<uint4*>(a.push_dn + q + Ss * 8) = *reinterpret_cast<const uint4*>(t1);
                            *reinterpret_cast<uint4*>(a.push_dn + q + 2 * Ss * 8) = *reinterpret_cast<const uint4*>(t2);
                        }
                    }
                }
                if (a.y_split != nullptr) {
                    const unsigned dz = p / hw, r = p - dz * hw, h = r / W, w = r - h * W;
                    const unsigned q = ((dz & 1) << 2) | ((h & 1) << 1) | (w & 1);
                    const size_t Hc = (H >> 1) + 2 * a.pad;  // `pad` spare CELL rows above / below (row bands)
                    const size_t Sc = (size_t)(S >> 3) / (H >> 1) * Hc;
                    const size_t cell = ((size_t)(dz >> 1) * Hc + (h >> 1) + a.pad) * (W >> 1) + (w >> 1);
                    const size_t base = ((((size_t)b * 8 + q) * NG + g) * 3 * Sc + cell) * 8;
                    *reinterpret_cast<uint4*>(a.y_split + base) = *reinterpret_cast<const uint4*>(t0);
                    *reinterpret_cast<uint4*>(a.y_split + base + Sc * 8) = *reinterpret_cast<const uint4*>(t1);
                    *reinterpret_cast<uint4*>(a.y_split + base + 2 * Sc * 8) = *reinterpret_cast<const uint4*>(t2);
                }
            }
            if (a.y_nchw != nullptr) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    a.y_nchw[a.nchw_padded ? ((size_t)b * a.C + g * 8 + e) * Sp + pp[u] : ((size_t)b * a.C + g * 8 + e) * S + p] =
                        v[u][e];
            }
        }
    }
    (void)need_dhw;
}


// ---- K1 in C8S3: cost[b][chunk][term][d][y][x][8] --------------------------------------------------------------
// The concat cost volume (cmfsm.py:667-682) written directly in the layout conv_tc3 consumes: chunks 0..C/8-1 = left
// features masked by x >= d, the rest = right features shifted by d; the three bf16 terms of an element sum to the
// fp32 feature value exactly, so the volume is still the bit-exact artefact (tests compare the reconstruction with the
// oracle's volume).  Same scheme as cost_volume.cu / the C8 kernel: a CTA stages 2 rows of one (8-channel group, term)
// once in shared memory (right rows behind a zero prefix) and streams the D shifted / masked copies with 128-bit
// stores; for fixed (chunk, term, d) its rows are contiguous.
constexpr int kCvS3Rows = 2;
__global__ void __launch_bounds__(256) cost_volume_c8s3_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                               __nv_bfloat16* __restrict__ cost, int C, int h, int w,
                                                               int D, int DP, int pad) {
    extern __shared__ uint4 sv3[];  // [rows][DP + w]
    const int nc = C / 8;
    const int y0 = blockIdx.x * kCvS3Rows, chunk = blockIdx.y / 3, term = blockIdx.y % 3, b = blockIdx.z;
    const int rows = min(kCvS3Rows, h - y0);
    const bool right = chunk >= nc;
    const int c0 = (right ? chunk - nc : chunk) * 8;
    const size_t plane = (size_t)h * w;
    const int pitch = DP + w;
    const float* src = (right ? R : L) + ((size_t)b * C + c0) * plane + (size_t)y0 * w;
    for (int i = threadIdx.x; i < rows * w; i += 256) {  // rows are contiguous: i == r*w + x
        __align__(16) __nv_bfloat16 p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            __nv_bfloat16 t0, t1, t2;
            split3(__ldg(src + e * plane + i), t0, t1, t2);
            p[e] = term == 0 ? t0 : term == 1 ? t1 : t2;
        }
        sv3[(i / w) * pitch + DP + (i % w)] = *reinterpret_cast<const uint4*>(p);
    }
    for (int i = threadIdx.x; i < rows * DP; i += 256) sv3[(i / DP) * pitch + (i % DP)] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t oplane = (size_t)(h + 2 * pad) * w;  // row bands: `pad` extra rows above and below every depth plane
    uint4* out = reinterpret_cast<uint4*>(cost) + ((((size_t)b * 2 * nc + chunk) * 3 + term) * D) * oplane +
                 (size_t)(y0 + pad) * w;
    for (int d = warp; d < D; d += 8) {
        uint4* o = out + (size_t)d * oplane;
        for (int i = lane; i < rows * w; i += 32) {
            const int r = i / w, x = i - r * w;
            uint4 v = sv3[r * pitch + DP + (right ? x - d : x)];
            if (!right && x < d) v = make_uint4(0, 0, 0, 0);
            asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(o + i), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------
template <int COUT, int T, int DIL, int KHW, bool STACK, int NA, int NB, bool OUT_NCHW, int MINB = 1>
static int launch_tc3(const void* x, const void* wpk, float* y, double* gn, int B, int Cin, int D, int H_in, int W, int KD,
                      int row_off, int H, cudaStream_t st) {
    using G = Tc3Cfg<COUT, T, DIL, KHW, STACK, NA, NB>;
    CUtensorMap tmap;
    const cuuint64_t NJ = (cuuint64_t)3 * (Cin / 8);
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H_in, (cuuint64_t)D, NJ, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H_in * W * 16, (cuuint64_t)D * H_in * W * 16,
                                NJ * D * H_in * W * 16};
    const cuuint32_t box[5] = {(cuuint32_t)G::PW * 8, (cuuint32_t)G::PH, 1, 6, 1};
    if (int rc = encode_tmap_5d(&tmap, x, gdim, gstr, box, "conv_tc3")) return rc;
    static_assert(MINB == 1 || (MINB * G::SMEM_BYTES <= 227 * 1024 && MINB * G::TMEM_COLS <= 512), "co-residency");
    auto kern = conv_tc3_kernel<COUT, T, DIL, KHW, STACK, NA, NB, OUT_NCHW, MINB>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(W, 8), tiles_h = (int)cdiv(H, 16), groups_w = (int)cdiv(tiles_w, T);
    dim3 grid((unsigned)(groups_w * tiles_h), (unsigned)D, (unsigned)B);
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv_tc3: grid too large");
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, reinterpret_cast<const __nv_bfloat16*>(wpk), y, gn, D, H, W,
                                                  groups_w, KD, Cin / 16, row_off);
    CMF_LAUNCH_CHECK("conv_tc3_kernel");
    return CMFB200_OK;
}

template <bool OUT_NCHW>
static int dispatch_tc3(const void* x, const void* wpk, float* y, double* gn, int B, int Cin, int Cout, int D, int H, int W,
                        int KD, int KHW, int dil, int row_off, int H_out, cudaStream_t st) {
    if (KHW == 3 && dil == 1) {
        // 32 output channels: 2 tiles per CTA and TWO co-resident CTAs per SM (108 KB, 256 TMEM columns each): the
        // prologue / epilogue of one overlaps the MMA phase of the other (K is only 2-12 steps deep in these layers)
        if (Cout == 32) return launch_tc3<32, 2, 1, 3, true, 2, 4, OUT_NCHW, 2>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
        if (Cout == 64) return launch_tc3<64, 2, 1, 3, true, 2, 4, OUT_NCHW>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
        if (Cout == 128) return launch_tc3<128, 2, 1, 3, false, 2, 3, OUT_NCHW>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
    } else if (KHW == 3 && dil == 2) {
        if (Cout == 128) return launch_tc3<128, 2, 2, 3, false, 2, 3, OUT_NCHW>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
    } else if (KHW == 1) {
        if (Cout == 32) return launch_tc3<32, 4, 1, 1, true, 3, 3, OUT_NCHW>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
        if (Cout == 128) return launch_tc3<128, 2, 1, 1, false, 3, 3, OUT_NCHW>(x, wpk, y, gn, B, Cin, D, H, W, KD, row_off, H_out, st);
    }
    CMF_REQUIRE(false, "conv_tc3_fwd: unsupported (Cout=%d, k=%d, dilation=%d); supported: 3x3 d1 Cout 32/64/128, "
                       "3x3 d2 Cout 128, 1x1 Cout 32/128", Cout, KHW, dil);
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_pack_tc3_weight(const float* weight, void* packed, int Cout, int Cin, int KD, int KHW,
                                       void* stream) {
    CMF_REQUIRE(weight && packed, "pack_tc3_weight: null pointer");
    CMF_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cin > 0 && Cin % 16 == 0, "pack_tc3_weight: Cin %% 16 and Cout %% 8 required");
    CMF_REQUIRE((KD == 1 || KD == 3) && (KHW == 1 || KHW == 3), "pack_tc3_weight: kernel extents must be 1 or 3");
    const long long n = (long long)KD * KHW * KHW * Cin * Cout;
    pack_tc3_weight_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        weight, reinterpret_cast<__nv_bfloat16*>(packed), Cout, Cin, KD, KHW);
    CMF_LAUNCH_CHECK("pack_tc3_weight_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_conv_tc3_rows_fwd(const void* x_c8s3, const void* packed_w, float* y, double* gn_sums, int B,
                                         int Cin, int Cout, int D, int H, int W, int KD, int KHW, int dilation,
                                         int out_nchw, int row_off, int H_out, void* stream) {
    CMF_REQUIRE(x_c8s3 && packed_w && y, "conv_tc3_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && H_out > 0, "conv_tc3_fwd: non-positive dimension");
    CMF_REQUIRE(row_off >= 0 && row_off + H_out <= H + 64, "conv_tc3_rows_fwd: row window [%d, %d) outside the input (%d rows)",
                row_off, row_off + H_out, H);
    CMF_REQUIRE(Cin > 0 && Cin % 16 == 0, "conv_tc3_fwd: Cin=%d must be a multiple of 16", Cin);
    CMF_REQUIRE((KD == 1 || KD == 3), "conv_tc3_fwd: KD must be 1 or 3");
    CMF_REQUIRE((reinterpret_cast<uintptr_t>(x_c8s3) & 15) == 0, "conv_tc3_fwd: input must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (out_nchw)
        return dispatch_tc3<true>(x_c8s3, packed_w, y, gn_sums, B, Cin, Cout, D, H, W, KD, KHW, dilation, row_off, H_out, st);
    return dispatch_tc3<false>(x_c8s3, packed_w, y, gn_sums, B, Cin, Cout, D, H, W, KD, KHW, dilation, row_off, H_out, st);
}

extern "C" int cmfb200_conv_tc3_fwd(const void* x_c8s3, const void* packed_w, float* y, double* gn_sums, int B, int Cin,
                                    int Cout, int D, int H, int W, int KD, int KHW, int dilation, int out_nchw,
                                    void* stream) {
    return cmfb200_conv_tc3_rows_fwd(x_c8s3, packed_w, y, gn_sums, B, Cin, Cout, D, H, W, KD, KHW, dilation, out_nchw, 0, H,
                                     stream);
}

extern "C" int cmfb200_gn_apply_tc3_padded(const float* raw, int raw_is_c8f, const double* gn_sums, const float* gamma,
                                           const float* beta, const void* residual_c8s3, const float* residual_nchw,
                                           void* y_c8s3, float* y_nchw, int B, int C, int groups, long long spatial,
                                           float eps, int relu, int pad, int H, int W, void* y_split_c8s3, void* push_up,
                                           void* push_dn, int push_rows, int nchw_padded, void* stream) {
    CMF_REQUIRE(raw && (y_c8s3 || y_nchw || y_split_c8s3), "gn_apply_tc3: null pointer");
    CMF_REQUIRE(push_rows >= 0 && (push_rows == 0 || (y_c8s3 && H >= push_rows && W > 0)),
                "gn_apply_tc3: the halo push needs the C8S3 output and push_rows <= H");
    CMF_REQUIRE(y_split_c8s3 == nullptr || (H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && spatial % ((long long)H * W) == 0 &&
                                            (spatial / ((long long)H * W)) % 2 == 0),
                "gn_apply_tc3: the parity-split copy needs even D, H, W");
    CMF_REQUIRE(pad >= 0 && (pad == 0 || (H > 0 && W > 0 && spatial % ((long long)H * W) == 0)),
                "gn_apply_tc3: padded rows need H, W with spatial a multiple of H*W");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && spatial > 0 && spatial < (1LL << 31), "gn_apply_tc3: bad shape");
    CMF_REQUIRE(gn_sums == nullptr || (gamma && beta && groups > 0 && C % groups == 0), "gn_apply_tc3: bad GroupNorm args");
    CMF_REQUIRE((long long)B * (C / 8) <= 65535, "gn_apply_tc3: B*C/8 exceeds the grid limit");
    GnTc3Args a;
    a.raw = raw, a.sums = gn_sums, a.gamma = gamma, a.beta = beta;
    a.res_s3 = reinterpret_cast<const __nv_bfloat16*>(residual_c8s3), a.res_nchw = residual_nchw;
    a.y_s3 = reinterpret_cast<__nv_bfloat16*>(y_c8s3), a.y_nchw = y_nchw;
    a.y_split = reinterpret_cast<__nv_bfloat16*>(y_split_c8s3);
    a.push_up = reinterpret_cast<__nv_bfloat16*>(push_up), a.push_dn = reinterpret_cast<__nv_bfloat16*>(push_dn);
    a.push_rows = (push_up || push_dn) ? push_rows : 0;
    a.C = C, a.cpg = gn_sums ? C / groups : 1, a.spatial = spatial, a.eps = eps, a.relu = relu, a.raw_c8f = raw_is_c8f;
    a.pad = pad, a.H = H, a.W = W, a.nchw_padded = (nchw_padded && pad > 0) ? 1 : 0;
    long long bx = cdiv(spatial, 256 * 4);  // 4 positions per thread, all loads of an iteration in flight together
    if (bx > 8192) bx = 8192;
    dim3 grid((unsigned)bx, (unsigned)(B * (C / 8)));
    if (a.push_rows > 0)
        gn_apply_tc3_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else
        gn_apply_tc3_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    CMF_LAUNCH_CHECK("gn_apply_tc3_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_gn_apply_tc3(const float* raw, int raw_is_c8f, const double* gn_sums, const float* gamma,
                                    const float* beta, const void* residual_c8s3, const float* residual_nchw,
                                    void* y_c8s3, float* y_nchw, int B, int C, int groups, long long spatial, float eps,
                                    int relu, void* stream) {
    return cmfb200_gn_apply_tc3_padded(raw, raw_is_c8f, gn_sums, gamma, beta, residual_c8s3, residual_nchw, y_c8s3, y_nchw, B,
                                       C, groups, spatial, eps, relu, 0, 0, 0, nullptr, nullptr, nullptr, 0, 0, stream);
}

extern "C" int cmfb200_cost_volume_concat_c8s3_padded(const float* L, const float* R, void* cost_c8s3, int B, int C, int h,
                                                      int w, int D, int pad, void* stream) {
    CMF_REQUIRE(L && R && cost_c8s3, "cost_volume_concat_c8s3: null pointer");
    CMF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && h > 0 && w > 0 && D > 0, "cost_volume_concat_c8s3: bad shape");
    CMF_REQUIRE(B <= 65535, "cost_volume_concat_c8s3: B exceeds grid limit");
    const int DP = (D + 3) & ~3;
    const size_t smem = (size_t)kCvS3Rows * (DP + w) * 16;
    CMF_REQUIRE(smem <= 200 * 1024, "cost_volume_concat_c8s3: row block does not fit in shared memory");
    if (smem > 48 * 1024)
        CMF_CUDA(cudaFuncSetAttribute(cost_volume_c8s3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)cdiv(h, kCvS3Rows), (unsigned)(3 * 2 * (C / 8)), (unsigned)B);
    cost_volume_c8s3_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(L, R, reinterpret_cast<__nv_bfloat16*>(cost_c8s3),
                                                                        C, h, w, D, DP, pad);
    CMF_LAUNCH_CHECK("cost_volume_c8s3_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_cost_volume_concat_c8s3(const float* L, const float* R, void* cost_c8s3, int B, int C, int h, int w,
                                               int D, void* stream) {
    return cmfb200_cost_volume_concat_c8s3_padded(L, R, cost_c8s3, B, C, h, w, D, 0, stream);
}
