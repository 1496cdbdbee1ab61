// tc_wgrad.cu -- instantiations + launcher of the weight-gradient GEMM (tc_wgrad.cuh) and the fixed-order
// reduction of its split-K slices.
#include "tc_wgrad.cuh"

namespace finc {
namespace tc {

template <int BN, int NPASS, int EW>
static int launch_wgrad_t(const CUtensorMap& mapP, const CUtensorMap& mapQ, float* out, const WgradGeom& g,
                          cudaStream_t st) {
    using C = WCfg<BN, NPASS, EW>;
    auto kern = wgrad_kernel<BN, NPASS, EW>;
    constexpr int smem = C::kSmemBytes + 1024;
    static_assert(smem <= kSmemLimit, "shared memory budget");
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err != cudaSuccess) return (int)err;
        configured[dev] = true;
    }
    const int work = g.m_tiles * g.n_tiles * g.slices;
    const int grid = work < sm_count_cached() ? work : sm_count_cached();
    if (grid <= 0) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(C::kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kern, mapP, mapQ, out, g);
}

int launch_wgrad(int BN, int npass, const CUtensorMap& mapP, const CUtensorMap& mapQ, float* out, const WgradGeom& g,
                 cudaStream_t st) {
    if (npass == 1) {
        switch (BN) {
            case 64: return launch_wgrad_t<64, 1, 4>(mapP, mapQ, out, g, st);
            case 128: return launch_wgrad_t<128, 1, 8>(mapP, mapQ, out, g, st);
            case 160: return launch_wgrad_t<160, 1, 8>(mapP, mapQ, out, g, st);
            default: return FINC_E_UNSUPPORTED;
        }
    }
    switch (BN) {
        case 64: return launch_wgrad_t<64, 3, 4>(mapP, mapQ, out, g, st);
        case 128: return launch_wgrad_t<128, 3, 8>(mapP, mapQ, out, g, st);
        case 160: return launch_wgrad_t<160, 3, 8>(mapP, mapQ, out, g, st);
        default: return FINC_E_UNSUPPORTED;
    }
}

// out[m * ld_dst + n] (+)= sum_s partial[s][m][n] for m < M, n < N, slices added in index order
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, int M, int N, int ld_src,
                                    size_t slice_stride, int slices, int ld_dst, int accumulate) {
    const long total = (long)M * N;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % N), m = (int)(idx / N);
        float s = 0.f;
        for (int k = 0; k < slices; ++k) s += partial[k * slice_stride + (size_t)m * ld_src + n];
        float* d = dst + (size_t)m * ld_dst + n;
        *d = accumulate ? *d + s : s;
    }
}

int launch_wgrad_reduce(const float* partial, float* dst, int M, int N, int ld_src, size_t slice_stride, int slices,
                        int ld_dst, int accumulate, cudaStream_t st) {
    const long total = (long)M * N;
    if (total <= 0) return 0;
    long blocks = (total + 255) / 256;
    if (blocks > 2048) blocks = 2048;
    wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, dst, M, N, ld_src, slice_stride, slices, ld_dst,
                                                         accumulate);
    return (int)cudaGetLastError();
}

}  // namespace tc
}  // namespace finc
