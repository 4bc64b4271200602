// Probe: which rank-5 fp32 tensor-map shapes the TMA unit accepts for the wgrad stage box (run one variant per process:
// an illegal instruction poisons the context).   nvcc -arch=sm_100a -o /tmp/probe tools/probe_tma_f32.cu && /tmp/probe V
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n_floats, int c0, int c1, int c2, int c3,
                      int c4) {
    extern __shared__ __align__(128) uint8_t raw[];
    float* buf = reinterpret_cast<float*>(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    __shared__ uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n_floats * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
            "%7}], [%2];" ::"r"((uint32_t)__cvta_generic_to_shared(buf)),
            "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
            : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar_a)
        : "memory");
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int B = 1, C = 32, D = 4, H = 10, W = 40;
    std::vector<float> h((size_t)B * C * D * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *x, *out;
    cudaMalloc(&x, h.size() * 4);
    cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    int c[5] = {-1, 0, -1, -1, 0};
    const cuuint64_t sW = 4, sH = W * 4, sD = (cuuint64_t)H * W * 4, sC = (cuuint64_t)D * H * W * 4, sB = sC * C;
    (void)sW;
    int bw = 36;
    if (variant == 3) bw = 32;
    if (variant == 0) {  // natural order (W, H, D, C, B)
        cuuint64_t gd[5] = {W, H, D, C, B}, gs[4] = {sH, sD, sC, sB};
        cuuint32_t bx[5] = {(cuuint32_t)bw, 3, 3, 32, 1};
        for (int i = 0; i < 5; ++i) gdim[i] = gd[i], box[i] = bx[i];
        for (int i = 0; i < 4; ++i) gstr[i] = gs[i];
        c[1] = -1, c[2] = -1, c[3] = 0;
    } else {  // (W, C, H, D, B)
        cuuint64_t gd[5] = {W, C, H, D, B}, gs[4] = {sC, sH, sD, sB};
        cuuint32_t bx[5] = {(cuuint32_t)bw, 32, 3, 3, 1};
        if (variant == 4) bx[2] = bx[3] = 1;
        if (variant == 5) bx[1] = 8;
        if (variant == 6) bx[1] = 16;
        for (int i = 0; i < 5; ++i) gdim[i] = gd[i], box[i] = bx[i];
        for (int i = 0; i < 4; ++i) gstr[i] = gs[i];
    }
    if (variant == 7) c[0] = 0;  // aligned start column
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode rc=%d box {%u,%u,%u,%u,%u}\n", variant, (int)r, box[0], box[1], box[2], box[3], box[4]);
    if (r != CUDA_SUCCESS) return 1;
    const int n = box[0] * box[1] * box[2] * box[3] * box[4];
    cudaMalloc(&out, n * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4 + 256);
    probe<<<1, 128, n * 4 + 256>>>(tm, out, n, c[0], c[1], c[2], c[3], c[4]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> o(n);
    cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    // check against the expected gather
    long bad = 0;
    for (int i4 = 0; i4 < (int)box[4]; ++i4)
        for (int i3 = 0; i3 < (int)box[3]; ++i3)
            for (int i2 = 0; i2 < (int)box[2]; ++i2)
                for (int i1 = 0; i1 < (int)box[1]; ++i1)
                    for (int i0 = 0; i0 < (int)box[0]; ++i0) {
                        const long g[5] = {c[0] + i0, c[1] + i1, c[2] + i2, c[3] + i3, c[4] + i4};
                        bool ok = true;
                        size_t off = 0;
                        for (int d = 0; d < 5; ++d) {
                            ok = ok && g[d] >= 0 && g[d] < (long)gdim[d];
                            off += (size_t)g[d] * (d == 0 ? 4 : gstr[d - 1]);
                        }
                        const float want = ok ? h[off / 4] : 0.f;
                        const float got = o[(((size_t)(i4 * box[3] + i3) * box[2] + i2) * box[1] + i1) * box[0] + i0];
                        if (want != got) ++bad;
                    }
    printf("variant %d mismatches: %ld of %d\n", variant, bad, n);
    return 0;
}
