// finc_wgrad.cu -- masked weight gradient of the FInC convolution, sm_100a.
//
//   dw[g][o][i][a][b] = sum_{n,h,w} dz[n,gC+o,h,w] * x[n,gC+i,h+r(a),w+c(b)]
//   dw[g][o][i][a*][b*] = 0 for i >= o                    (PaddedConv2d.reset_gradients)
//
// Replaces cuDNN wgrad + the post-hoc `grad * mask.to(device)` of the reference
// (layers/conv.py:98-99, train/experiment.py:16-18,250).  The result can be written
// straight into a flat gradient bucket (the NCCL all-reduce buffer).
//
// Design: a long reduction (B*H*W terms) into few outputs (C*C*kH*kW per group).
//   * grid = (X, G, Z): CTA (x, g, z) streams every X-th chunk of `ipc` images of group g
//     (x and dz tiles, TMA bulk copies, 2-stage mbarrier pipeline) and owns the output
//     slice z of the group.
//   * output-stationary: warp <-> (input channel i, block of OBW output channels o,
//     row-split slot); lane <-> (column w, row phase).  Each lane keeps
//     OBW*kH*kW accumulators in registers for the whole kernel: per pixel OBW + kH*kW
//     conflict-free shared loads feed OBW*kH*kW FMAs.
//   * reduction is deterministic: xor-shuffle tree inside the warp, fixed-order sums over
//     the row-split warps (shared) and over the X CTAs (global partials; the last CTA to
//     finish, found with one atomic ticket per (g,z), does the final pass, applies the
//     mask and writes dw).  No floating-point atomics anywhere.
#include "finc_common.cuh"

namespace finc {

namespace {

constexpr int kWarps = 12;
constexpr int kCounterBytes = 4096;

struct WgArgs {
    const float* dz;
    const float* x;
    float* dw;
    float* partial;
    unsigned* counters;
    Shape s;
    unsigned flags;
    int ipc;
    int tile_floats;
    int tile_stride;
    int bulk;
    int nobw;
    int jpc;   // jobs per CTA
    int rsw;   // row-split warps per job
    int X;
    int Z;
    int nchunks;
};

template <int OBW, int KH, int KW>
__global__ void __launch_bounds__(kWarps * 32, 1) wgrad_kernel(const WgArgs a) {
    constexpr int NACC = OBW * KH * KW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Shape& s = a.s;
    const int C = s.C, H = s.H, W = s.W, HW = H * W;
    const int half = a.ipc * a.tile_stride;  // floats of x (or dz) tiles per stage
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    float* red = stage0 + 4 * (size_t)half;
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + kWarps * NACC + (((kWarps * NACC) & 1) ? 1 : 0));
    __shared__ int s_last;

    const int g = blockIdx.y, zz = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int njobs = C * a.nobw;
    const int jl_cta = warp / a.rsw, rs = warp - jl_cta * a.rsw;
    const int job = zz * a.jpc + jl_cta;
    const bool job_on = jl_cta < a.jpc && job < njobs;
    const int ci = job_on ? job / a.nobw : 0;
    const int obw = job_on ? job - ci * a.nobw : 0;
    const int ord = order_of(s.orders, g);

    const int WL = W < 32 ? W : 32;
    const int HS = W <= 32 ? 32 / W : 1;
    const int hsl = lane / WL, jl = lane - hsl * WL;
    const bool lane_on = job_on && hsl < HS;
    const int ncb = (W + 31) / 32;

    auto issue = [&](int chunk, int st) {  // thread 0 only
        const int n0 = chunk * a.ipc;
        const int nt = min(a.ipc, s.B - n0);
        mbar_arrive_expect_tx(&bars[st], (uint32_t)(2 * nt * a.tile_floats * 4));
        float* xs = stage0 + (size_t)st * 2 * half;
        float* ds = xs + half;
        for (int t = 0; t < nt; ++t) {
            const long off = ((long)(n0 + t) * s.G + g) * a.tile_floats;
            bulk_g2s(xs + t * a.tile_stride, a.x + off, (uint32_t)(a.tile_floats * 4), &bars[st]);
            bulk_g2s(ds + t * a.tile_stride, a.dz + off, (uint32_t)(a.tile_floats * 4), &bars[st]);
        }
    };

    if (a.bulk && threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        if ((int)blockIdx.x < a.nchunks) issue(blockIdx.x, 0);
        if ((int)blockIdx.x + a.X < a.nchunks) issue(blockIdx.x + a.X, 1);
    }
    __syncthreads();

    float acc[OBW][KH][KW];
#pragma unroll
    for (int o = 0; o < OBW; ++o)
#pragma unroll
        for (int aa = 0; aa < KH; ++aa)
#pragma unroll
            for (int b = 0; b < KW; ++b) acc[o][aa][b] = 0.f;

    int kk = 0;
    for (int chunk = blockIdx.x; chunk < a.nchunks; chunk += a.X, ++kk) {
        const int st = kk & 1;
        const int n0 = chunk * a.ipc;
        const int nt = min(a.ipc, s.B - n0);
        float* xs = stage0 + (size_t)st * 2 * half;
        float* ds = xs + half;
        if (a.bulk) {
            mbar_wait(&bars[st], (uint32_t)((kk >> 1) & 1));
        } else {
            for (int t = 0; t < nt; ++t) {
                const long off = ((long)(n0 + t) * s.G + g) * a.tile_floats;
                for (int e = threadIdx.x; e < a.tile_floats; e += blockDim.x) {
                    xs[t * a.tile_stride + e] = a.x[off + e];
                    ds[t * a.tile_stride + e] = a.dz[off + e];
                }
            }
            __syncthreads();
        }
        if (lane_on) {
            const int R = nt * H;
            for (int r = rs * HS + hsl; r < R; r += a.rsw * HS) {
                const int t = r / H, h = r - t * H;
                const float* xt = xs + t * a.tile_stride + ci * HW;
                const float* dt = ds + t * a.tile_stride + (obw * OBW) * HW + h * W;
                for (int cb = 0; cb < ncb; ++cb) {
                    const int w = cb * 32 + jl;
                    if (w >= W) break;
                    float dzv[OBW];
#pragma unroll
                    for (int o = 0; o < OBW; ++o) dzv[o] = (obw * OBW + o < C) ? dt[o * HW + w] : 0.f;
#pragma unroll
                    for (int aa = 0; aa < KH; ++aa) {
                        const int hh = h + row_off(ord, aa, KH);
                        if (hh < 0 || hh >= H) continue;
#pragma unroll
                        for (int b = 0; b < KW; ++b) {
                            const int wc = w + col_off(ord, b, KW);
                            const float xv = (wc >= 0 && wc < W) ? xt[hh * W + wc] : 0.f;
#pragma unroll
                            for (int o = 0; o < OBW; ++o) acc[o][aa][b] = fmaf(dzv[o], xv, acc[o][aa][b]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this stage
        if (a.bulk && threadIdx.x == 0 && chunk + 2 * a.X < a.nchunks) issue(chunk + 2 * a.X, st);
    }

    // ---- deterministic reduction: lanes -> warp, row-split warps -> CTA, CTAs -> dw --------
#pragma unroll
    for (int o = 0; o < OBW; ++o)
#pragma unroll
        for (int aa = 0; aa < KH; ++aa)
#pragma unroll
            for (int b = 0; b < KW; ++b) {
                float v = acc[o][aa][b];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) red[warp * NACC + (o * KH + aa) * KW + b] = v;
            }
    __syncthreads();

    const long nout = (long)s.G * C * C * KH * KW;
    const int ca = corner_a(ord, KH), cbn = corner_b(ord, KW);
    const bool direct = a.X == 1;
    for (int e = threadIdx.x; e < a.jpc * NACC; e += blockDim.x) {
        const int jj = e / NACC, idx = e - jj * NACC;
        const int jb = zz * a.jpc + jj;
        if (jb >= njobs) continue;
        float v = 0.f;
        for (int r = 0; r < a.rsw; ++r) v += red[(jj * a.rsw + r) * NACC + idx];
        const int i = jb / a.nobw, ob = jb - i * a.nobw;
        const int o = ob * OBW + idx / (KH * KW);
        if (o >= C) continue;
        const int ab = idx % (KH * KW);
        const long out = (((long)g * C + o) * C + i) * KH * KW + ab;
        if (direct) {
            if (!(a.flags & FINC_FLAG_NO_MASK) && ab == ca * KW + cbn && i >= o) v = 0.f;
            if (a.flags & FINC_FLAG_ACCUMULATE) v += a.dw[out];
            a.dw[out] = v;
        } else {
            a.partial[(long)blockIdx.x * nout + out] = v;
        }
    }
    if (direct) return;

    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&a.counters[g * a.Z + zz], 1u);
        s_last = (ticket == (unsigned)(a.X - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = threadIdx.x; e < a.jpc * NACC; e += blockDim.x) {
        const int jj = e / NACC, idx = e - jj * NACC;
        const int jb = zz * a.jpc + jj;
        if (jb >= njobs) continue;
        const int i = jb / a.nobw, ob = jb - i * a.nobw;
        const int o = ob * OBW + idx / (KH * KW);
        if (o >= C) continue;
        const int ab = idx % (KH * KW);
        const long out = (((long)g * C + o) * C + i) * KH * KW + ab;
        float v = 0.f;
        for (int xx = 0; xx < a.X; ++xx) v += __ldcg(&a.partial[(long)xx * nout + out]);
        if (!(a.flags & FINC_FLAG_NO_MASK) && ab == ca * KW + cbn && i >= o) v = 0.f;
        if (a.flags & FINC_FLAG_ACCUMULATE) v += a.dw[out];
        a.dw[out] = v;
    }
    if (threadIdx.x == 0) a.counters[g * a.Z + zz] = 0u;  // leave the ticket clean
}

template <int OBW, int KH, int KW>
int launch_inst(const WgArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
    auto kern = wgrad_kernel<OBW, KH, KW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kWarps * 32, smem, st>>>(a);
    return (int)cudaGetLastError();
}

// output-channel block per warp job; 0 = shape not covered by the tiled kernel
int pick_obw(int C, int kH, int kW) {
    if (kH == 3 && kW == 3) {
        if (C <= 4) return C;
        if (C % 12 == 0) return 12;
        if (C % 6 == 0) return 6;
        if (C % 4 == 0) return 4;
        if (C % 3 == 0) return 3;
        return 4;
    }
    if (kH == 5 && kW == 5) {
        if (C <= 4) return C;
        if (C % 4 == 0) return 4;
        if (C % 3 == 0) return 3;
        return 4;
    }
    return 0;
}

struct Plan {
    int obw, nobw, jpc, rsw, Z, X, ipc, nchunks;
};

bool make_plan(const Shape& s, Plan* p) {
    p->obw = pick_obw(s.C, s.kH, s.kW);
    if (!p->obw) return false;
    const long tile_bytes = (long)s.C * s.H * s.W * 4;
    if (tile_bytes > 32 * 1024) return false;
    p->nobw = (s.C + p->obw - 1) / p->obw;
    const int njobs = s.C * p->nobw;
    p->jpc = njobs < kWarps ? njobs : kWarps;
    p->rsw = kWarps / p->jpc;
    p->Z = (njobs + p->jpc - 1) / p->jpc;
    if ((long)s.G * p->Z * 4 > kCounterBytes) return false;
    int ipc = (int)(16 * 1024 / tile_bytes);
    if (ipc < 1) ipc = 1;
    if (ipc > 16) ipc = 16;
    if (ipc > s.B) ipc = s.B;
    const int sms = sm_count_cached();
    int xmax = sms / (s.G * p->Z);
    if (xmax < 1) xmax = 1;
    // keep at least ~2 chunks per CTA when the batch allows it
    while (ipc > 1 && (s.B + ipc - 1) / ipc < 2 * xmax) ipc >>= 1;
    p->ipc = ipc;
    p->nchunks = (s.B + ipc - 1) / ipc;
    p->X = xmax < p->nchunks ? xmax : p->nchunks;
    return true;
}

}  // namespace

size_t wgrad_workspace_floats(const Shape& s) {
    Plan p;
    const long nout = (long)s.G * s.C * s.C * s.kH * s.kW;
    if (!make_plan(s, &p)) return kCounterBytes / 4;
    return kCounterBytes / 4 + (size_t)p.X * nout;
}

int launch_wgrad_fast(const float* dz, const float* x, float* dw, float* workspace, size_t ws_floats, const Shape& s,
                      unsigned flags, cudaStream_t st, bool* handled) {
    *handled = false;
    Plan p;
    if (!make_plan(s, &p)) return 0;
    const long nout = (long)s.G * s.C * s.C * s.kH * s.kW;
    if (ws_floats < kCounterBytes / 4 + (size_t)p.X * nout) return FINC_E_WORKSPACE;
    WgArgs a{};
    a.dz = dz; a.x = x; a.dw = dw; a.s = s; a.flags = flags;
    a.counters = reinterpret_cast<unsigned*>(workspace);
    a.partial = workspace + kCounterBytes / 4;
    a.ipc = p.ipc; a.nobw = p.nobw; a.jpc = p.jpc; a.rsw = p.rsw; a.X = p.X; a.Z = p.Z; a.nchunks = p.nchunks;
    a.tile_floats = s.C * s.H * s.W;
    a.tile_stride = (a.tile_floats + 3) & ~3;
    a.bulk = (a.tile_floats % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(dz) & 15) == 0);
    const int nacc = p.obw * s.kH * s.kW;
    const size_t smem = (size_t)4 * a.ipc * a.tile_stride * 4 + (size_t)(kWarps * nacc + 1) * 4 + 32;
    if (smem > max_optin_smem_cached()) return 0;
    if (a.X > 1) {
        cudaError_t e = cudaMemsetAsync(a.counters, 0, (size_t)s.G * a.Z * 4, st);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(a.X, s.G, a.Z);
    *handled = true;
    if (s.kH == 3) {
        switch (p.obw) {
            case 1: return launch_inst<1, 3, 3>(a, grid, smem, st);
            case 2: return launch_inst<2, 3, 3>(a, grid, smem, st);
            case 3: return launch_inst<3, 3, 3>(a, grid, smem, st);
            case 4: return launch_inst<4, 3, 3>(a, grid, smem, st);
            case 6: return launch_inst<6, 3, 3>(a, grid, smem, st);
            case 12: return launch_inst<12, 3, 3>(a, grid, smem, st);
        }
    } else {
        switch (p.obw) {
            case 1: return launch_inst<1, 5, 5>(a, grid, smem, st);
            case 2: return launch_inst<2, 5, 5>(a, grid, smem, st);
            case 3: return launch_inst<3, 5, 5>(a, grid, smem, st);
            case 4: return launch_inst<4, 5, 5>(a, grid, smem, st);
        }
    }
    *handled = false;
    return 0;
}

}  // namespace finc
