// Pre-included (-include) when compiling the REFERENCE's CUDA inverse extension unmodified under
// torch >= 2.x: its `AT_DISPATCH_FLOATING_TYPES(input.type(), ...)`
// (fastflow/utils/fastflow_cuda_inverse/cinc_cuda_kernel_level2.cu:117, cinc_cuda_kernel_level1.cu:113)
// needs the `::detail::scalar_type(DeprecatedTypeProperties)` overload that newer ATen dropped.
// Test infrastructure only (oracle/build_ref_cuda.py); no reference code is copied.
#pragma once
#include <torch/extension.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
