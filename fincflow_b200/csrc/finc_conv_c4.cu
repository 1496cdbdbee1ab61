// instantiations of the fused FInC convolution for C = 4; output blocks [4, 2]
#include "finc_conv.cuh"
namespace finc {
namespace conv {
template <>
int dispatch_ob<4>(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (OB) {
        case 4: return dispatch_wt<4, 4>(WT, KH, a, grid, threads, smem, st);
        case 2: return dispatch_wt<4, 2>(WT, KH, a, grid, threads, smem, st);
        default: return FINC_E_UNSUPPORTED;
    }
}
}  // namespace conv
}  // namespace finc
