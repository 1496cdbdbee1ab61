// explicit instantiation of the register-window wavefront inverse for C = 3, 3x3 kernels
#include "finc_inverse_rw.cuh"
namespace finc {
namespace rw {
template int dispatch_ck<3, 3>(int, const RwArgs&, dim3, int, size_t, cudaStream_t);
}
}
