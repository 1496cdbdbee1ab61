// tc_host.cuh -- host-side declarations shared by the tensor-core translation units.
#pragma once

#include "tc_igemm.cuh"

namespace finc {
namespace tc {

// launchers implemented by the template-instantiation units; return cudaError_t as int, or
// FINC_E_UNSUPPORTED when (BN, npass) is not instantiated
int launch_igemm_nhwc(int BN, int npass, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapOut,
                      const Geom& g, const EpiArgs& e, cudaStream_t st);
int launch_igemm_coupling(int BN, int npass, const CUtensorMap& mapA, const CUtensorMap& mapB, const Geom& g,
                          const EpiArgs& e, cudaStream_t st);

template <int BN, int NPASS, int EPI>
inline int launch_igemm_t(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapOut, const Geom& g,
                          const EpiArgs& e, cudaStream_t st) {
    using C = Cfg<BN, NPASS>;
    auto kern = igemm_kernel<BN, NPASS, EPI>;
    constexpr int smem = C::kSmemBytes + 1024;  // + slack for the manual 1024-byte alignment
    static_assert(smem <= kSmemLimit, "shared memory budget");
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err != cudaSuccess) return (int)err;
        configured[dev] = true;
    }
    const int tiles = g.tiles_w * g.tiles_h * g.tiles_n * g.n_tiles;
    const int grid = tiles < sm_count_cached() ? tiles : sm_count_cached();
    if (grid <= 0) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kern, mapA, mapB, mapOut, g, e);
}

}  // namespace tc
}  // namespace finc
