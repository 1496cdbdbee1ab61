// explicit instantiations: plain row-major epilogue (narrow / odd-width outputs), 3-pass TF32
#include "tc_host.cuh"

namespace finc {
namespace tc {

int launch_igemm_rows_p3(int BN, const CUtensorMap& mapA, const CUtensorMap& mapB, const Geom& g, const EpiArgs& e,
                         cudaStream_t st) {
    switch (BN) {
        case 16: return launch_igemm_t<16, 3, EPI_ROWS, 1, 4>(mapA, mapB, mapA, g, e, st);
        case 32: return launch_igemm_t<32, 3, EPI_ROWS, 1, 4>(mapA, mapB, mapA, g, e, st);
        case 64: return launch_igemm_t<64, 3, EPI_ROWS, 1, 4>(mapA, mapB, mapA, g, e, st);
        case 128: return launch_igemm_t<128, 3, EPI_ROWS, 1, 8>(mapA, mapB, mapA, g, e, st);
        case 144: return launch_igemm_t<144, 3, EPI_ROWS, 1, 8>(mapA, mapB, mapA, g, e, st);
        default: return FINC_E_UNSUPPORTED;
    }
}

}  // namespace tc
}  // namespace finc
