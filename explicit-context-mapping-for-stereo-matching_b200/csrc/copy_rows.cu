// Pitched row-block copy for the row-band halo exchange (callers: cmf_b200/parallel.py fill_row_halo_ /
// _exchange).  The boundary rows of a band activation [..., D, rows, W, 8] are `height` = prod(outer dims) blocks of
// `width` = n_rows * W * 8 * elem contiguous bytes, `pitch` bytes apart; the mailbox side is dense.  ATen's strided
// copy_ moves them 2 bf16 elements per thread (344 GB/s measured on the 28 MB halo planes of the 2048x3072 volume,
// 5.4 ms per forward independent of the number of ranks); this kernel moves 16 bytes per thread per step.
// Source or destination may be a peer GPU's memory mapped over NVLink (symmetric-memory mailbox).
#include "common.cuh"

namespace cmfb200 {

__global__ void __launch_bounds__(256) copy_2d_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long dst_pitch,
                                                      long long src_pitch, long long width, long long total) {
    // all quantities in 16-byte units
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long k = i + u * stride;
            if (k < total) {
                const long long r = k / width, c = k - r * width;
                v[u] = src[r * src_pitch + c];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long k = i + u * stride;
            if (k < total) {
                const long long r = k / width, c = k - r * width;
                dst[r * dst_pitch + c] = v[u];
            }
        }
    }
}

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    const double* p[kMaxPeers];
};

// dst[i] = scale * (p0[i] + p1[i] + ... + p(n-1)[i]), summed in rank order on every rank (identical bits everywhere).
// Replaces copy_ + (n-1) add_ + mul launches per GroupNorm layer of a band forward.
__global__ void sum_peers_kernel(double* __restrict__ dst, PeerPtrs src, int n, int count, double scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double acc = src.p[0][i];
    for (int r = 1; r < n; ++r) acc += src.p[r][i];
    dst[i] = acc * scale;
}

}  // namespace cmfb200

using namespace cmfb200;

extern "C" int cmfb200_sum_peers_f64(double* dst, const void* const* srcs, int n, int count, double scale, void* stream) {
    CMF_REQUIRE(dst && srcs && n >= 1 && n <= kMaxPeers && count > 0, "sum_peers_f64: bad arguments (1..%d peers)", kMaxPeers);
    PeerPtrs pp;
    for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = r < n ? reinterpret_cast<const double*>(srcs[r]) : nullptr;
    for (int r = 0; r < n; ++r) CMF_REQUIRE(pp.p[r] != nullptr, "sum_peers_f64: null peer pointer %d", r);
    sum_peers_kernel<<<(unsigned)cdiv(count, 128), 128, 0, (cudaStream_t)stream>>>(dst, pp, n, count, scale);
    CMF_LAUNCH_CHECK("sum_peers_kernel");
    return CMFB200_OK;
}

extern "C" int cmfb200_copy_2d(void* dst, long long dst_pitch, const void* src, long long src_pitch, long long width,
                               long long height, void* stream) {
    CMF_REQUIRE(dst && src, "copy_2d: null pointer");
    CMF_REQUIRE(width > 0 && height > 0 && dst_pitch >= width && src_pitch >= width, "copy_2d: bad extent");
    CMF_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)dst_pitch |
                  (uintptr_t)src_pitch | (uintptr_t)width) & 15) == 0,
                "copy_2d: pointers, pitches and width must be multiples of 16 bytes");
    const long long w = width / 16, total = w * height;
    int dev = 0, sms = kNumSMs;
    CMF_CUDA(cudaGetDevice(&dev));
    CMF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long blocks = std::min<long long>(cdiv(total, 256 * 4), (long long)sms * 8);
    copy_2d_kernel<<<(unsigned)std::max<long long>(blocks, 1), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), dst_pitch / 16, src_pitch / 16, w, total);
    CMF_LAUNCH_CHECK("copy_2d_kernel");
    return CMFB200_OK;
}
