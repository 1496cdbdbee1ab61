// finc_affine.cu -- per-pixel channel affine map y[n,:,p] = A x[n,:,p] + b, sm_100a.
//
// First "next" row of the scope table: the glue that follows every FastFlowUnit
// in the reference's flow step is ActNorm (layers/actnorm.py:14-52: (x - t) * exp(-log_s) per
// channel) and Conv1x1 (layers/conv1x1.py:18-43: z = W x per pixel).  Both are per-pixel affine
// maps over the channel axis, so their composition is ONE such map with A = W diag(exp(-log_s)),
// b = -A t (and the reverse is the map with A^-1, +t): one HBM round trip instead of two, no
// per-call torch.inverse.  The same kernel with A^T is the backward-data pass.
//
// HBM-bound streaming kernel (2*C flop per 8 bytes): a thread owns VEC consecutive pixels of one
// image, reads its C channel vectors with coalesced vector loads (NCHW: channel stride = H*W),
// keeps them in registers and produces the C outputs four at a time with the rows of A^T as
// 128-bit shared-memory broadcasts (one LDS.128 feeds 4*VEC FMAs).
#include "finc_common.cuh"

namespace finc {

namespace {

template <int VEC>
struct VecT;
template <>
struct VecT<4> { using type = float4; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float* out) {
    if constexpr (VEC == 4) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
        out[0] = v.x; out[1] = v.y;
    } else {
        out[0] = __ldcs(p);
    }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float* v) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else p[0] = v[0];
}

template <int C, int VEC>
__global__ void __launch_bounds__(256) affine1x1_kernel(const float* __restrict__ x, const float* __restrict__ A,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        long n_items, int hwv, long HW) {
    constexpr int CP = (C + 3) & ~3;
    __shared__ __align__(16) float At[C * CP];  // At[i][o] = A[o][i]
    __shared__ __align__(16) float bs[CP];
    for (int e = threadIdx.x; e < C * CP; e += blockDim.x) {
        const int i = e / CP, o = e - i * CP;
        At[e] = o < C ? __ldg(A + o * C + i) : 0.f;
    }
    for (int o = threadIdx.x; o < CP; o += blockDim.x) bs[o] = (bias != nullptr && o < C) ? __ldg(bias + o) : 0.f;
    __syncthreads();
    for (long item = (long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (long)gridDim.x * blockDim.x) {
        const long n = item / hwv;
        const long off = n * C * HW + (item - n * hwv) * VEC;
        float xv[C][VEC];
#pragma unroll
        for (int i = 0; i < C; ++i) ldv<VEC>(x + off + i * HW, xv[i]);
#pragma unroll
        for (int ob = 0; ob < CP / 4; ++ob) {
            float acc[4][VEC];
            const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * ob);
#pragma unroll
            for (int v = 0; v < VEC; ++v) { acc[0][v] = b4.x; acc[1][v] = b4.y; acc[2][v] = b4.z; acc[3][v] = b4.w; }
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(At + i * CP + 4 * ob);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    acc[0][v] = fmaf(a.x, xv[i][v], acc[0][v]);
                    acc[1][v] = fmaf(a.y, xv[i][v], acc[1][v]);
                    acc[2][v] = fmaf(a.z, xv[i][v], acc[2][v]);
                    acc[3][v] = fmaf(a.w, xv[i][v], acc[3][v]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * ob + k < C) stv<VEC>(y + off + (4 * ob + k) * HW, acc[k]);
        }
    }
}

// any C: one thread per output element (coalesced along the pixels)
__global__ void affine1x1_generic_kernel(const float* __restrict__ x, const float* __restrict__ A,
                                         const float* __restrict__ bias, float* __restrict__ y, long total, int C,
                                         long HW) {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long p = e % HW;
        const long no = e / HW;
        const int o = (int)(no % C);
        const long n = no / C;
        float acc = bias ? __ldg(bias + o) : 0.f;
        const float* xp = x + n * C * HW + p;
        for (int i = 0; i < C; ++i) acc = fmaf(__ldg(A + o * C + i), xp[i * HW], acc);
        y[e] = acc;
    }
}

template <int C, int VEC>
int launch_inst(const float* x, const float* A, const float* bias, float* y, int B, long HW, cudaStream_t st) {
    const long n_items = (long)B * (HW / VEC);
    const long blocks = (n_items + 255) / 256;
    const long cap = (long)sm_count_cached() * 8;
    affine1x1_kernel<C, VEC><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(x, A, bias, y, n_items, (int)(HW / VEC), HW);
    return (int)cudaGetLastError();
}

template <int C, int VMAX>
int dispatch_vec(int vec, const float* x, const float* A, const float* bias, float* y, int B, long HW, cudaStream_t st) {
    if constexpr (VMAX >= 4) { if (vec == 4) return launch_inst<C, 4>(x, A, bias, y, B, HW, st); }
    if constexpr (VMAX >= 2) { if (vec >= 2) return launch_inst<C, 2>(x, A, bias, y, B, HW, st); }
    return launch_inst<C, 1>(x, A, bias, y, B, HW, st);
}

}  // namespace

int launch_affine1x1(const float* x, const float* A, const float* bias, float* y, int B, int C, long HW, cudaStream_t st) {
    // widest vector the layout allows: rows start at multiples of HW floats from 16-byte aligned bases
    int vec = 1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y);
    if (HW % 4 == 0 && (al & 15) == 0) vec = 4;
    else if (HW % 2 == 0 && (al & 7) == 0) vec = 2;
    switch (C) {
        case 4: return dispatch_vec<4, 4>(vec, x, A, bias, y, B, HW, st);
        case 8: return dispatch_vec<8, 4>(vec, x, A, bias, y, B, HW, st);
        case 12: return dispatch_vec<12, 4>(vec, x, A, bias, y, B, HW, st);
        case 16: return dispatch_vec<16, 4>(vec, x, A, bias, y, B, HW, st);
        case 24: return dispatch_vec<24, 4>(vec, x, A, bias, y, B, HW, st);
        case 32: return dispatch_vec<32, 2>(vec, x, A, bias, y, B, HW, st);
        case 48: return dispatch_vec<48, 2>(vec, x, A, bias, y, B, HW, st);
        case 96: return dispatch_vec<96, 1>(vec, x, A, bias, y, B, HW, st);
        default: break;
    }
    const long total = (long)B * C * HW;
    const long blocks = (total + 255) / 256;
    const long cap = (long)sm_count_cached() * 16;
    affine1x1_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(x, A, bias, y, total, C, HW);
    return (int)cudaGetLastError();
}

}  // namespace finc
