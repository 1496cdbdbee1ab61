// instantiations of the fused FInC convolution for C = 1; output blocks [1]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<1>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 1: return dispatch_wt<1, 1>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
