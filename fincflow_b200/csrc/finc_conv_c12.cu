// instantiations of the fused FInC convolution for C = 12; output blocks [4, 2, 1]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<12>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 4: return dispatch_wt<12, 4>(WT, KH, a, grid, threads, smem, st);
        case 2: return dispatch_wt<12, 2>(WT, KH, a, grid, threads, smem, st);
        case 1: return dispatch_wt<12, 1>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
