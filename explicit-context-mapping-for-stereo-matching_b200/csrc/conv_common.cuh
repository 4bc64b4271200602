// Tile geometry + GroupNorm-statistics epilogue shared by the fp32 direct convolution kernels (2-D and 3-D).
#pragma once
#include "common.cuh"

namespace cmfb200 {

constexpr int kConvThreads = 256;
constexpr int kTW = 32;  // output voxels along w per CTA
constexpr int kVPT = 4;  // consecutive-w output voxels per thread

// TW = output voxels along w per CTA: 32, or 16 when that wastes fewer lanes (e.g. w = 240 = 15 x 16 = 7.5 x 32)
template <int COUT, int CPT, int TW = kTW>
struct ConvTile {
    static constexpr int NCG = COUT / CPT;               // channel groups
    static constexpr int NQ = kConvThreads / NCG;        // voxel quads per CTA
    static constexpr int ROWS = NQ / (TW / kVPT);        // (d,h) rows of TW voxels
    static constexpr int TD = (COUT == 1) ? 4 : 1;
    static constexpr int TH = ROWS / TD;
    static_assert(NCG * NQ == kConvThreads && TD * TH * (TW / kVPT) == NQ, "bad tile");
};

// epilogue helper: block-level reduction of per-thread channel sums into gn_sums[b][co][2] (double atomics).
// Sums and sums of squares are accumulated in DOUBLE from the first element on: the variance is formed as
// E[x^2]-mean^2, so any fp32 rounding of a partial sum is amplified by E[x^2]/var.  Measured on the SPP branch
// (GroupNorm over 2 pooled values per channel): fp32 4-element partials gave 6.5e-4 relative error there and
// 1 px disparity error end to end on random-init weights.
template <int COUT, int CPT>
__device__ __forceinline__ void gn_epilogue(const double (&s)[CPT], const double (&ss)[CPT], int cg, void* sred_raw,
                                            double* __restrict__ gn_sums, int b, int c_total = COUT, int c_base = 0) {
    using T = ConvTile<COUT, CPT>;
    constexpr int WPG = T::NQ / 32;  // warps per channel group
    double* sred = reinterpret_cast<double*>(sred_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // sred aliases the operand buffers: everyone must be done with them
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const double a = warp_sum(s[c]), q = warp_sum(ss[c]);
        if (lane == 0) {
            sred[(warp * CPT + c) * 2 + 0] = a;
            sred[(warp * CPT + c) * 2 + 1] = q;
        }
    }
    __syncthreads();
    if (threadIdx.x < COUT * 2) {
        const int co = threadIdx.x >> 1, which = threadIdx.x & 1;
        const int g = co / CPT, c = co % CPT;
        double acc = 0.0;
        for (int wgi = 0; wgi < WPG; ++wgi) acc += sred[((g * WPG + wgi) * CPT + c) * 2 + which];
        atomicAdd(gn_sums + ((size_t)b * c_total + c_base + co) * 2 + which, acc);
    }
    (void)cg;
}

// ---- cp.async (LDGSTS) helpers: 4-byte copies with zero fill, 16-byte copies -------------------------
__device__ __forceinline__ void cp_async_4_zfill(float* smem_dst, const float* gsrc, bool valid) {
    const int src_bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace cmfb200
