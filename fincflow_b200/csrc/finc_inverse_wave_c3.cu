// explicit instantiation of the specialised wavefront inverse for C = 3 (3x3 and 5x5, P = 1,2,4,8)
#include "finc_inverse_wave.cuh"
namespace finc {
namespace wave {
template int dispatch_c<3>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
}
}
