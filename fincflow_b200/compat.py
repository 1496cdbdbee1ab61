"""Drop-in objects for the reference's JIT-built pybind11 extensions.

The reference calls `cinc_cuda_level2.inverse(input, kernel, output)` with the four quadrants
already flipped to TL form (fastflow/fastflow.py:79-92) and `cinc_cuda_level1.inverse` with one
TL-form convolution (layers/conv.py:191-218).  These shims keep that exact call signature and
return convention (`[output]`, same storage) on top of finc_inverse_f32, so the reference's
own FastFlowUnit.reverse_level2 runs unchanged on the persistent wavefront kernel:

    import fincflow_b200.compat as compat
    fastflow.cinc_cuda_level2 = compat.cinc_cuda_level2      # instead of torch cpp_extension.load
"""
from __future__ import annotations

from . import _native


class _Level2:
    """pybind signature: inverse(input[B,4Cq,H,W], kernel[4Cq,Cq,kH,kW], output) -> [output]
    (utils/fastflow_cuda_inverse/cinc_cuda_level2.cpp:19-32)"""

    @staticmethod
    def inverse(input, kernel, output):
        # all four groups arrive in TL form -> orders = (TL, TL, TL, TL); `output` need not be zeroed
        _native.inverse(input, kernel, G=4, orders=0, out=output)
        return [output]


class _Level1:
    """pybind signature: inverse(input[B,C,H,W], kernel[C,C,kH,kW], output) -> [output]
    (utils/fastflow_cuda_inverse/cinc_cuda_level1.cpp:19-32)"""

    @staticmethod
    def inverse(input, kernel, output):
        _native.inverse(input, kernel, G=1, orders=0, out=output)
        return [output]


cinc_cuda_level2 = _Level2()
cinc_cuda_level1 = _Level1()
