// instantiations of the fused FInC convolution for C = 3; output blocks [3, 1]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<3>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 3: return dispatch_wt<3, 3>(WT, KH, a, grid, threads, smem, st);
        case 1: return dispatch_wt<3, 1>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
