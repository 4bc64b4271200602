// Entry points of the two resolution-changing layers of the hourglass on the C8/bf16 layout (the tcgen05
// kernels are conv3d_s2_igemm_persistent.cu and deconv3d_igemm_persistent.cu) and the parity-split copy:
//
//  * conv k3 stride 2 (hourglass.conv1/conv3, cmf/models/cmfsm.py:244-254).  Input index i = 2*o + k - 1: tap k=1
//    reads parity-0 inputs at o, taps k=0 / k=2 read parity-1 inputs at o-1 / o.  The producer layer writes a
//    PARITY-SPLIT copy [B][8 parities][C/8][D/2][H/2][W/2][8] (gn_apply_c8 `y_split`, or `c8_parity_split`
//    below), which turns the strided gather into eight dense sub-volumes.
//  * transposed conv k3 s2 p1 op1 (hourglass.conv5/conv6, :261-281): od = 2*id - 1 + kd, no zero insertion.
#include "common.cuh"
#include "igemm_common.cuh"

namespace cmfb200 {

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
    CMF_REQUIRE(Cin == 64 && (Cout == 32 || Cout == 64),
                "deconv3d_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 64->64, 64->32", Cin, Cout);
    return deconv3d_igemm_persistent_dispatch(x_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, D, H, W, st);
}

extern "C" int cmfb200_conv3d_s2_igemm_bf16_fwd(const void* x_split_c8, const void* packed_w, void* y_c8,
                                                double* gn_sums, int B, int Cin, int Cout, int Do, int Ho, int Wo,
                                                void* stream) {
    CMF_REQUIRE(x_split_c8 && packed_w && y_c8, "conv3d_s2_igemm_bf16_fwd: null pointer");
    CMF_REQUIRE(B > 0 && Do > 0 && Ho > 0 && Wo > 0, "conv3d_s2_igemm_bf16_fwd: non-positive dimension");
    cudaStream_t st = (cudaStream_t)stream;
    CMF_REQUIRE(Cout == 64 && (Cin == 32 || Cin == 64),
                "conv3d_s2_igemm_bf16_fwd: unsupported (Cin=%d, Cout=%d); supported: 32->64, 64->64", Cin, Cout);
    return conv3d_s2_igemm_persistent_dispatch(x_split_c8, packed_w, y_c8, gn_sums, B, Cin, Cout, Do, Ho, Wo, st);
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
