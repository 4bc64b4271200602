// fp32-ACCURATE stride-2 convolution and transposed convolution of the hourglasses on the tensor cores (three-term
// bf16 split, same arithmetic and layouts as conv_tc3.cu): hourglass.conv1 / conv3 (Conv3d k3 s2 p1,
// cmf/models/cmfsm.py:244-254) and conv5 / conv6 (ConvTranspose3d k3 s2 p1 op1, :261-281) in the fp32 parity mode.
//
// Both are sums of SMALL stride-1 convolutions over a half-resolution cell grid:
//   * stride 2:  in index i = 2o - 1 + k.  Tap k = 1 reads parity-0 inputs of cell o; taps k = 0 / 2 read parity-1 inputs
//     of cells o - 1 / o.  With the input stored PARITY-SPLIT ([B][8 classes q][C/8][3 terms][D/2][H/2][W/2][8], written
//     by the producing GroupNorm apply) every class q is a dense sub-volume that contributes 1, 2, 4 or 8 taps with cell
//     offsets in {-1, 0}: K steps run over (q, depth tap, 16-channel chunk), each with its own 1..4 in-plane taps, all into
//     ONE accumulator set.  27 taps of work, no zero insertion, no strided gather.
//   * transposed: out index o' = 2i - 1 + k.  Output parity 0 takes tap k = 1 of cell i; parity 1 takes k = 2 of cell i and
//     k = 0 of cell i + 1.  Every output parity class p is its own small conv (1..8 taps, cell offsets in {0, +1}); all
//     eight run in ONE launch (blockIdx.z = b * 8 + p), the epilogue writes the class's voxels (2i + p) of the C8F output
//     and adds to the same GroupNorm sums.
// The kernel is conv_tc3's with the fixed 3x3(x3) tap loops replaced by a K-step TABLE (box origin, depth offset, channel
// chunk, batch-dimension index, number of in-plane taps and their offsets inside the 17 x 9 box, weight offset) built on
// the host and passed as a __grid_constant__ parameter; one ring stage = the T activation boxes + the <= 4 weight taps
// of a K step.  Weights are packed (host side, cmf_b200.ops.pack_tc3_s2_weight / pack_tc3_deconv_weight) in table order.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

namespace {

constexpr int kGPW = 9, kGPH = 17;             // 8 x 16 tile + one halo voxel on one side
constexpr int kGPlane = kGPH * kGPW * 16;      // one (8-channel chunk, term) plane of a box
constexpr int kGATile = (6 * kGPlane + 127) & ~127;  // 2 chunks x 3 terms, rounded up: TMA destinations are 128-byte aligned

struct KStep {
    short ox, oy, oz;        // box origin relative to the tile origin (voxels) / depth offset relative to d
    short j0;                // first (chunk, term) plane index = kc * 6
    short bsel;              // added to b * bmul in the batch dimension of the tensor map (parity class of the input)
    short nkh, nkw;          // in-plane taps of this step (1 or 2 each)
    short hoff[2], woff[2];  // tap offsets inside the box (voxels)
    unsigned b_off;          // byte offset of this step's taps in the packed weights
};
template <int MAXSTEPS>
struct KTable {
    int n;
    short od0, oh0, ow0, pad_;  // output voxel of compute position (d,h,w): (d*os+od0, h*os+oh0, w*os+ow0)
    KStep s[MAXSTEPS];
};
// NTAB tables served by ONE launch (blockIdx.z = b * NTAB + table): the eight output parity classes of the transposed
// conv run concurrently instead of as eight latency-bound launches
template <int NTAB, int MAXSTEPS>
struct KTableSet {
    int bmul;
    KTable<MAXSTEPS> t[NTAB];
};
struct OutMap {
    int Do, Ho, Wo, os;
};

template <int COUT, int T, int NS>
struct GCfg {
    static constexpr int B_TAP = 6 * COUT * 16;
    static constexpr int A_STAGE = T * kGATile;
    static constexpr int STAGE = A_STAGE + 4 * B_TAP;
    static constexpr int ACC = 3 * COUT;
    static constexpr int TMEM_NEED = T * ACC;
    static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = NS * STAGE + 1024 + 1024;
    static_assert(TMEM_NEED <= 512 && 3 * COUT <= 256, "accumulators / stacked N");
    static_assert(kGATile % 128 == 0 && STAGE % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(4 * COUT * 2 * 8 <= NS * STAGE, "reduction scratch");
    static_assert(SMEM_BYTES <= 227 * 1024, "configuration does not fit in shared memory");
};

__host__ __device__ constexpr uint32_t g_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace

template <int COUT, int T, int NS, int NTAB, int MAXSTEPS>
__global__ void __launch_bounds__(kIgThreads, 1)
    conv_tc3g_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ KTableSet<NTAB, MAXSTEPS> tabs,
                     const __nv_bfloat16* __restrict__ wpk, float* __restrict__ y, double* __restrict__ gn_sums, int Hc,
                     int Wc, int groups_w, const OutMap om, int row_off) {
    // row bands: the input carries `row_off` spare (halo) rows above the Hc rows this launch computes
    using G = GCfg<COUT, T, NS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * G::STAGE);
    uint64_t* full = bars;
    uint64_t* empty = full + NS;
    uint64_t* tmemFull = empty + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmemFull + 1);
    double* sred = reinterpret_cast<double*>(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x % groups_w, ty = blockIdx.x / groups_w;
    const int tx0 = gx * T, d = blockIdx.y, b = blockIdx.z / NTAB;
    const KTable<MAXSTEPS>& tab = tabs.t[blockIdx.z % NTAB];

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
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
            for (int ks = 0; ks < tab.n; ++ks) {
                const KStep& st = tab.s[ks];
                const int s = ks % NS;
                if (ks >= NS) mbar_wait(empty + s, ((ks / NS) - 1) & 1);
                const uint32_t wbytes = (uint32_t)(st.nkh * st.nkw) * G::B_TAP;
                mbar_arrive_expect_tx(full + s, T * 6 * kGPlane + wbytes);  // the boxes are dense; kGATile only pads the slots
                uint8_t* stage = smem + s * G::STAGE;
#pragma unroll
                for (int t = 0; t < T; ++t)
                    tma_load_5d(stage + t * kGATile, &tmap_x, full + s, ((tx0 + t) * 8 + st.ox) * 8, ty * 16 + st.oy + row_off,
                                d + st.oz, st.j0, b * tabs.bmul + st.bsel);
                bulk_g2s(stage + G::A_STAGE, reinterpret_cast<const uint8_t*>(wpk) + st.b_off, wbytes, full + s);
            }
        }
    } else if (warp == 1) {
        const uint32_t a_hi = umma_desc_hi(kGPW * 16), b_hi = umma_desc_hi(128);
        for (int ks = 0; ks < tab.n; ++ks) {
            const KStep& st = tab.s[ks];
            const int s = ks % NS;
            mbar_wait(full + s, (ks / NS) & 1);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(smem) + s * G::STAGE, 3 * kGPlane);
            const uint32_t w_lo = umma_desc_lo(smem_u32(smem) + s * G::STAGE + G::A_STAGE, 3 * COUT * 16);
            if (elect_one()) {
                int tap = 0;
                for (int kh = 0; kh < st.nkh; ++kh) {
                    for (int kw = 0; kw < st.nkw; ++kw, ++tap) {
                        const uint32_t accum = (ks | tap) != 0 ? 1u : 0u;
                        const uint32_t a_off = (uint32_t)((st.hoff[kh] * kGPW + st.woff[kw]) * 16);
                        const uint64_t w0 = umma_desc_at(w_lo, b_hi, tap * G::B_TAP);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const uint32_t ab = t * kGATile + a_off;
                            const uint32_t dcol = tmem_base + t * G::ACC;
                            umma_bf16(dcol, umma_desc_at(a_lo, a_hi, ab), w0, g_idesc(3 * COUT), accum);
                            umma_bf16(dcol + COUT, umma_desc_at(a_lo, a_hi, ab + kGPlane), w0, g_idesc(2 * COUT), 1u);
                            umma_bf16(dcol + 2 * COUT, umma_desc_at(a_lo, a_hi, ab + 2 * kGPlane), w0, g_idesc(COUT), 1u);
                        }
                    }
                }
                umma_commit(empty + s);
                if (ks == tab.n - 1) umma_commit(tmemFull);
            }
            __syncwarp();
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const size_t oplane = (size_t)om.Ho * om.Wo;
        double tot_s[COUT / 32], tot_q[COUT / 32];
#pragma unroll
        for (int cb = 0; cb < COUT / 32; ++cb) tot_s[cb] = tot_q[cb] = 0.0;
        mbar_wait(tmemFull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const int h = ty * 16 + (row >> 3), w = (tx0 + t) * 8 + (row & 7);
            const bool ok = (h < Hc) && (w < Wc);
            const size_t opos = ((size_t)(d * om.os + tab.od0)) * oplane + (size_t)(h * om.os + tab.oh0) * om.Wo + (w * om.os + tab.ow0);
#pragma unroll 1
            for (int cb = 0; cb < COUT / 32; ++cb) {
                float o[32];
                {
                    uint32_t v0[32], v1[32];
                    const uint32_t base = tlane + t * G::ACC + cb * 32;
                    tmem_ld_32x32b_x32_issue(base + 2 * COUT, v0);
                    tmem_ld_32x32b_x32_issue(base + COUT, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v0[c]) + __uint_as_float(v1[c]);
                    tmem_ld_32x32b_x32(base, v0);
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c] += __uint_as_float(v0[c]);
                }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float* dst = y + (((size_t)b * (COUT / 8) + cb * 4 + j) * om.Do * oplane + opos) * 8;
                        *reinterpret_cast<float4*>(dst) = make_float4(o[j * 8], o[j * 8 + 1], o[j * 8 + 2], o[j * 8 + 3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[j * 8 + 4], o[j * 8 + 5], o[j * 8 + 6], o[j * 8 + 7]);
                    }
                }
                if (gn_sums != nullptr) {
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

// ---- host: K-step tables (the weight packers in cmf_b200/ops.py walk the same order) -----------------------------------
// One axis of a class with parity bit `bit`: number of taps, and for tap index i the offset INSIDE the box.
//   stride 2   (box origin at cell -1): bit 0 -> 1 tap (k=1) at box index 1 ; bit 1 -> k=0 at index 0, k=2 at index 1
//   transposed (box origin at cell  0): bit 0 -> 1 tap (k=1) at box index 0 ; bit 1 -> k=2 at index 0, k=0 at index 1
static void build_table_s2(KTable<48>& t, int KC, int B_TAP) {
    t.n = 0, t.od0 = t.oh0 = t.ow0 = t.pad_ = 0;
    unsigned off = 0;
    for (int q = 0; q < 8; ++q) {
        const int qd = q >> 2, qh = (q >> 1) & 1, qw = q & 1;
        for (int kdi = 0; kdi < 1 + qd; ++kdi)
            for (int kc = 0; kc < KC; ++kc) {
                KStep& s = t.s[t.n++];
                s.ox = -1, s.oy = -1, s.oz = (short)(qd ? kdi - 1 : 0), s.j0 = (short)(kc * 6), s.bsel = (short)q;
                s.nkh = (short)(1 + qh), s.nkw = (short)(1 + qw);
                s.hoff[0] = (short)(qh ? 0 : 1), s.hoff[1] = 1, s.woff[0] = (short)(qw ? 0 : 1), s.woff[1] = 1;
                s.b_off = off;
                off += (unsigned)(s.nkh * s.nkw) * B_TAP;
            }
    }
}
static unsigned build_table_deconv(KTable<8>& t, int p, int KC, int B_TAP, unsigned off) {
    const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
    t.n = 0, t.od0 = (short)pd, t.oh0 = (short)ph, t.ow0 = (short)pw, t.pad_ = 0;
    for (int kdi = 0; kdi < 1 + pd; ++kdi)
        for (int kc = 0; kc < KC; ++kc) {
            KStep& s = t.s[t.n++];
            s.ox = 0, s.oy = 0, s.oz = (short)kdi, s.j0 = (short)(kc * 6), s.bsel = 0;
            s.nkh = (short)(1 + ph), s.nkw = (short)(1 + pw);
            s.hoff[0] = 0, s.hoff[1] = 1, s.woff[0] = 0, s.woff[1] = 1;
            s.b_off = off;
            off += (unsigned)(s.nkh * s.nkw) * B_TAP;
        }
    return off;
}

template <int COUT, int T, int NS, int NTAB, int MAXSTEPS>
static int launch_g(const CUtensorMap& tmap, const KTableSet<NTAB, MAXSTEPS>& tabs, const void* wpk, float* y, double* gn,
                    int B, int Dc, int Hc, int Wc, const OutMap& om, int row_off, cudaStream_t st) {
    using G = GCfg<COUT, T, NS>;
    auto kern = conv_tc3g_kernel<COUT, T, NS, NTAB, MAXSTEPS>;
    CMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int tiles_w = (int)cdiv(Wc, 8), tiles_h = (int)cdiv(Hc, 16), groups_w = (int)cdiv(tiles_w, T);
    dim3 grid((unsigned)(groups_w * tiles_h), (unsigned)Dc, (unsigned)(B * NTAB));
    CMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv_tc3g: grid too large");
    kern<<<grid, kIgThreads, G::SMEM_BYTES, st>>>(tmap, tabs, reinterpret_cast<const __nv_bfloat16*>(wpk), y, gn, Hc, Wc,
                                                  groups_w, om, row_off);
    CMF_LAUNCH_CHECK("conv_tc3g_kernel");
    return CMFB200_OK;
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_conv_tc3_s2_rows_fwd(const void* x_split_c8s3, const void* packed_w, float* y_c8f, double* gn_sums,
                                            int B, int Cin, int Cout, int Do, int Ho, int Wo, int pad, void* stream) {
    // pad: spare cell rows above and below the Ho cell rows of every parity class (row bands; 0 = dense)
    CMF_REQUIRE(x_split_c8s3 && packed_w && y_c8f, "conv_tc3_s2_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Do > 0 && Ho > 0 && Wo > 0 && pad >= 0, "conv_tc3_s2_fwd: bad dimension");
    CMF_REQUIRE(Cout == 64 && (Cin == 32 || Cin == 64), "conv_tc3_s2_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->64, 64->64",
                Cin, Cout);
    const cuuint64_t NJ = (cuuint64_t)3 * (Cin / 8);
    const cuuint64_t Hp = (cuuint64_t)Ho + 2 * pad;
    const cuuint64_t vol = (cuuint64_t)Do * Hp * Wo * 16;
    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)Wo * 8, Hp, (cuuint64_t)Do, NJ, (cuuint64_t)B * 8};
    const cuuint64_t gstr[4] = {(cuuint64_t)Wo * 16, Hp * Wo * 16, vol, vol * NJ};
    const cuuint32_t box[5] = {kGPW * 8, kGPH, 1, 6, 1};
    if (int rc = encode_tmap_5d(&tmap, x_split_c8s3, gdim, gstr, box, "conv_tc3_s2")) return rc;
    KTableSet<1, 48> tabs;
    tabs.bmul = 8;
    build_table_s2(tabs.t[0], Cin / 16, 6 * Cout * 16);
    const OutMap om = {Do, Ho, Wo, 1};
    return launch_g<64, 2, 3, 1, 48>(tmap, tabs, packed_w, y_c8f, gn_sums, B, Do, Ho, Wo, om, pad, (cudaStream_t)stream);
}

extern "C" int cmfb200_conv_tc3_s2_fwd(const void* x_split_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B,
                                       int Cin, int Cout, int Do, int Ho, int Wo, void* stream) {
    return cmfb200_conv_tc3_s2_rows_fwd(x_split_c8s3, packed_w, y_c8f, gn_sums, B, Cin, Cout, Do, Ho, Wo, 0, stream);
}

extern "C" int cmfb200_deconv_tc3_rows_fwd(const void* x_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B,
                                           int Cin, int Cout, int D, int H, int W, int pad, void* stream) {
    // pad: spare rows above and below the H rows of x (row bands: the row below the band holds the neighbour's first row)
    CMF_REQUIRE(x_c8s3 && packed_w && y_c8f, "deconv_tc3_fwd: null pointer");
    CMF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && pad >= 0, "deconv_tc3_fwd: bad dimension");
    CMF_REQUIRE(Cin == 64 && (Cout == 32 || Cout == 64), "deconv_tc3_fwd: unsupported (Cin=%d, Cout=%d); supported: 64->64, 64->32",
                Cin, Cout);
    const cuuint64_t NJ = (cuuint64_t)3 * (Cin / 8);
    CUtensorMap tmap;
    const cuuint64_t Hp = (cuuint64_t)H + 2 * pad;
    const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, Hp, (cuuint64_t)D, NJ, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, Hp * W * 16, (cuuint64_t)D * Hp * W * 16, NJ * D * Hp * W * 16};
    const cuuint32_t box[5] = {kGPW * 8, kGPH, 1, 6, 1};
    if (int rc = encode_tmap_5d(&tmap, x_c8s3, gdim, gstr, box, "deconv_tc3")) return rc;
    KTableSet<8, 8> tabs;  // the eight output parity classes, one launch
    tabs.bmul = 1;
    unsigned off = 0;
    for (int p = 0; p < 8; ++p) off = build_table_deconv(tabs.t[p], p, Cin / 16, 6 * Cout * 16, off);
    const OutMap om = {2 * D, 2 * H, 2 * W, 2};
    if (Cout == 64) return launch_g<64, 2, 3, 8, 8>(tmap, tabs, packed_w, y_c8f, gn_sums, B, D, H, W, om, pad, (cudaStream_t)stream);
    return launch_g<32, 4, 3, 8, 8>(tmap, tabs, packed_w, y_c8f, gn_sums, B, D, H, W, om, pad, (cudaStream_t)stream);
}

extern "C" int cmfb200_deconv_tc3_fwd(const void* x_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B, int Cin,
                                      int Cout, int D, int H, int W, void* stream) {
    return cmfb200_deconv_tc3_rows_fwd(x_c8s3, packed_w, y_c8f, gn_sums, B, Cin, Cout, D, H, W, 0, stream);
}
