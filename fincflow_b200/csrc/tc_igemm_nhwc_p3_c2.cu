// explicit instantiations: channels-last epilogue, 3-pass TF32, clusters of 2
#include "tc_host.cuh"

namespace finc {
namespace tc {

int launch_igemm_nhwc_p3_c2(int BN, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapOut,
                              const Geom& g, const EpiArgs& e, cudaStream_t st) {
    switch (BN) {
        case 32: return launch_igemm_t<32, 3, EPI_NHWC, 2, 4>(mapA, mapB, mapOut, g, e, st);
        case 64: return launch_igemm_t<64, 3, EPI_NHWC, 2, 4>(mapA, mapB, mapOut, g, e, st);
        case 128: return launch_igemm_t<128, 3, EPI_NHWC, 2, 8>(mapA, mapB, mapOut, g, e, st);
        default: return FINC_E_UNSUPPORTED;
    }
}

}  // namespace tc
}  // namespace finc
