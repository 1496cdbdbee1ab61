// explicit instantiation of the register-window wavefront inverse for C = 6, 5x5 kernels
#include "finc_inverse_rw.cuh"
namespace finc {
namespace rw {
template int dispatch_ck<6, 5>(int, const RwArgs&, dim3, int, size_t, cudaStream_t);
}
}
