// tc_host.cuh -- host-side declarations shared by the tensor-core translation units.
#pragma once

#include "tc_igemm.cuh"

namespace finc {
namespace tc {

// launchers implemented by the template-instantiation units; return cudaError_t as int, or
// FINC_E_UNSUPPORTED when the variant is not instantiated.  `cluster` = CTAs sharing one weight tile.
int launch_igemm_nhwc(int BN, int npass, int cluster, const CUtensorMap& mapA, const CUtensorMap& mapB,
                      const CUtensorMap& mapOut, const Geom& g, const EpiArgs& e, cudaStream_t st);
int launch_igemm_rows(int BN, int npass, const CUtensorMap& mapA, const CUtensorMap& mapB, const Geom& g,
                      const EpiArgs& e, cudaStream_t st);

struct WgradGeom;
int launch_wgrad(int BN, int npass, const CUtensorMap& mapP, const CUtensorMap& mapQ, float* out, const WgradGeom& g,
                 cudaStream_t st);
int launch_wgrad_reduce(const float* partial, float* dst, int M, int N, int ld_src, size_t slice_stride, int slices,
                        int ld_dst, int accumulate, cudaStream_t st);

template <int BN, int NPASS, int EPI, int CL, int EW>
inline int launch_igemm_t(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapOut, const Geom& g,
                          const EpiArgs& e, cudaStream_t st) {
    using C = Cfg<BN, NPASS, EW>;
    auto kern = igemm_kernel<BN, NPASS, EPI, CL, EW>;
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
    const int m_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    const int cluster_tiles = (m_tiles + CL - 1) / CL * g.n_tiles;
    const int max_clusters = sm_count_cached() / CL;
    const int clusters = cluster_tiles < max_clusters ? cluster_tiles : max_clusters;
    if (clusters <= 0) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CL), 1, 1);
    cfg.blockDim = dim3(C::kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CL > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return (int)cudaLaunchKernelEx(&cfg, kern, mapA, mapB, mapOut, g, e);
}

}  // namespace tc
}  // namespace finc
