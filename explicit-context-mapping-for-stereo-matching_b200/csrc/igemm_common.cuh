// tcgen05 / TMEM / TMA primitives and the TMA descriptor encoder shared by the implicit-GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace cmfb200 {

constexpr int kIgThreads = 192;  // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: epilogue

// ---- tcgen05 / TMA primitives --------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=0 (no swizzle)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// split form: issue several loads back to back, then ONE wait (the three plane blocks of the depth-stacked epilogue)
__device__ __forceinline__ void tmem_ld_32x32b_x32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x1_issue(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v;
}

// elect.sync: exactly one lane of a converged warp gets `true` (warp-uniform control flow around tcgen05.mma lets
// ptxas keep descriptors in uniform registers instead of electing + R2UR-ing per instruction)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// descriptor with the constant fields in the high word; only the 14-bit start-address field (low word) moves
__device__ __forceinline__ uint64_t umma_desc_at(uint32_t lo_base, uint32_t hi, uint32_t byte_offset) {
    return ((uint64_t)hi << 32) | (uint64_t)(lo_base + (byte_offset >> 4));
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }

// Transposing warp reduction: every lane holds 32 partial values (one per column); afterwards lane l holds the sum of
// column l over all 32 lanes.  31 shuffles + 31 adds instead of 32 x 5 (a per-value butterfly), fp32.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float give = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, give, off);
        }
    }
    return v[0];
}

// ---- work partition of the depth-walking kernels (conv3d_igemm_kdstack.cu, conv3d_igemm_cout1_gather.cu) ----
struct KdUnit {
    int b, ty, tx, d0, d1, pl0, pl1;
};

// Work partition: the (tile column, depth) space of one output-channel group is linearised (column-major, depth
// fastest) and cut into equal contiguous ranges, one per CTA of the group; a range is walked as segments that stay
// inside one column.  A segment [d0,d1) needs the input planes d0-1 .. d1 (clipped to the volume).
struct KdWalk {
    long long pos, end;
    int D, tiles_w, per_sample;
    __device__ __forceinline__ KdWalk(int rank, int nranks, long long total, int D_, int tiles_w_, int tiles_h_)
        : pos(total * rank / nranks), end(total * (rank + 1) / nranks), D(D_), tiles_w(tiles_w_),
          per_sample(tiles_w_ * tiles_h_) {}
    __device__ __forceinline__ bool next(KdUnit& u) {
        if (pos >= end) return false;
        const int col = (int)(pos / D);
        u.d0 = (int)(pos - (long long)col * D);
        const long long left = end - pos;
        u.d1 = (left < (long long)(D - u.d0)) ? u.d0 + (int)left : D;
        pos += u.d1 - u.d0;
        u.b = col / per_sample;
        const int r = col - u.b * per_sample;
        u.ty = r / tiles_w;
        u.tx = r - u.ty * tiles_w;
        u.pl0 = u.d0 > 0 ? u.d0 - 1 : 0;
        u.pl1 = u.d1 < D ? u.d1 : D - 1;
        return true;
    }
};

// ---- host: cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda.so) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    // resolved through the runtime so that libcmfb200.so has no link-time dependency on libcuda.so
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}


// rank-5 tensor map (bf16 unless stated), no swizzle, zero OOB fill
static inline int encode_tmap_5d(CUtensorMap* tmap, const void* base, const cuuint64_t (&gdim)[5],
                                 const cuuint64_t (&gstr)[4], const cuuint32_t (&box)[5], const char* who,
                                 CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    EncodeTiledFn encode = get_encode_fn();
    CMF_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(tmap, dtype, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CMF_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
    return CMFB200_OK;
}

}  // namespace cmfb200
