// instantiations of the fused FInC convolution for C = 6; output blocks [6, 3, 2]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<6>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 6: return dispatch_wt<6, 6>(WT, KH, a, grid, threads, smem, st);
        case 3: return dispatch_wt<6, 3>(WT, KH, a, grid, threads, smem, st);
        case 2: return dispatch_wt<6, 2>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
