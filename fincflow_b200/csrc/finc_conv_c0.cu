// instantiations of the fused FInC convolution for C = runtime (generic); output blocks [1, 2, 3, 4, 6]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<0>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 1: return dispatch_wt<0, 1>(WT, KH, a, grid, threads, smem, st);
        case 2: return dispatch_wt<0, 2>(WT, KH, a, grid, threads, smem, st);
        case 3: return dispatch_wt<0, 3>(WT, KH, a, grid, threads, smem, st);
        case 4: return dispatch_wt<0, 4>(WT, KH, a, grid, threads, smem, st);
        case 6: return dispatch_wt<0, 6>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
