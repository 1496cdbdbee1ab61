// finc_inverse_wave.cuh -- specialised wavefront inverse for the shapes FInC flows use
// (C in {1,2,3,4,6,12,24}, 3x3 / 5x5 kernels).  Same contract and pipeline as
// finc_inverse.cu (warp-persistent workers, TMA bulk load -> in-place solve in shared
// memory -> TMA bulk store) with a lower-latency, higher-throughput inner loop:
//
//   * lane = (column j, part p), P = 2^k parts per pixel.  The (kH*kW-1) non-corner taps of
//     a pixel are split over its P lanes (split-K over taps); each lane accumulates the
//     contribution of its taps to ALL C output channels, then the P partial sums are
//     combined with log2(P) xor-shuffles.  A 4x4x12 tile therefore keeps 32 lanes busy
//     (4 columns x 8 parts) instead of 4, a 16x16x3 tile 32 instead of 16.
//   * the channel-triangular corner solve runs in registers, redundantly on the P lanes of
//     the pixel (no divergence); lane p writes the channels o == p (mod P).
//   * weights of the lane's taps and the corner tap live in REGISTERS when they fit
//     (C <= 6), loaded once per item; otherwise they are vector loads from the
//     sweep-ordered table in shared memory.
//   * an item is a STACK of T tiles (n..n+T-1, g) of one group swept as one tall image:
//     lane j solves stacked row (step - j), so the wavefront never drains between tiles
//     (a lone 16x16 tile keeps only 52% of the lane-steps busy, a stack of 4 keeps 81%).
//     Dependencies are cut at tile boundaries exactly like the zero padding does.
//   * kernel size and channel count are template parameters: tap offsets, bounds and the
//     triangular solve are fully unrolled; no integer division in the step loop.
#pragma once
#include <type_traits>

#include "finc_common.cuh"

namespace finc {
namespace wave {

constexpr int kMaxWarps = 12;  // 384 threads: up to 170 registers per lane (weights + both streams)

struct WaveArgs {
    const float* z;
    const float* w;
    float* x;
    Shape s;
    int T;  // stacked tiles per item
    int S;  // pipeline stages per warp
    int gsplit;
    int bulk;
    int tile_floats;
    int tile_stride;
    int wk_floats;
    int prepared;  // 1: a.w is a prepared sweep-ordered table (after the header)
    unsigned long long* dbg;
    long n_items;
};

template <int C>
struct Pad4 {
    static constexpr int value = C <= 2 ? C : ((C + 3) / 4) * 4;
    // floats between consecutive taps of the weight table: an odd number of 16-byte groups, so
    // the P lanes of a pixel (different taps, same row) hit disjoint banks with LDS.128
    static constexpr int tap_stride = (C <= 2) ? C * value : (((C * value / 4) % 2 == 1) ? C * value : C * value + 4);
};

__device__ __forceinline__ void bulk_wait_read_1w() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int C, int KH, int KW, int P>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) inverse_wave_kernel(const WaveArgs a) {
    constexpr int CPP = Pad4<C>::value;
    constexpr int TS = Pad4<C>::tap_stride;
    constexpr int NT = KH * KW - 1;          // non-corner taps
    constexpr bool CREG = (C <= 12);            // corner weights in registers (C(C-1)/2 of them are used)
    constexpr int CB = 32 / P;                  // columns per block

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* wk = reinterpret_cast<float*>(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Shape& s = a.s;
    const int H = s.H, W = s.W;
    const int HW = H * W;
    const int stage_floats = a.T * a.tile_stride;
    const int wk_pad = (a.wk_floats + 31) & ~31;
    float* bufs = wk + wk_pad + (size_t)warp * a.S * stage_floats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wk + wk_pad + (size_t)nwarps * a.S * stage_floats) + warp * a.S;

    const int g_fixed = a.gsplit ? (int)blockIdx.y : -1;
    const long gw = (long)warp * gridDim.x + blockIdx.x;
    const long gstride = (long)gridDim.x * nwarps;

    auto item_g = [&](long item) -> int { return a.gsplit ? g_fixed : (int)(item % s.G); };
    auto item_n0 = [&](long item) -> int { return (int)(a.gsplit ? item : item / s.G) * a.T; };
    auto issue_load = [&](long item, int st) {  // lane 0 only
        const int g = item_g(item), n0 = item_n0(item);
        const int nt = min(a.T, s.B - n0);
        mbar_arrive_expect_tx(&bars[st], (uint32_t)(nt * a.tile_floats * 4));
        for (int t = 0; t < nt; ++t)
            bulk_g2s(bufs + st * stage_floats + t * a.tile_stride, a.z + ((long)(n0 + t) * s.G + g) * a.tile_floats,
                     (uint32_t)(a.tile_floats * 4), &bars[st]);
    };

    if (threadIdx.x == 0) dbg_mark(a.dbg, 0);
    const int n_pre = a.S == 3 ? 2 : a.S;
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wk + wk_pad + (size_t)nwarps * a.S * stage_floats) + nwarps * a.S;
    if (a.bulk && lane == 0) {
        for (int st = 0; st < a.S; ++st) mbar_init(&bars[st], 1);
        if (warp == 0) mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (a.prepared) __syncthreads();  // wbar initialised before anyone waits on it
    pdl_wait();
    pdl_trigger();
    if (a.prepared && threadIdx.x == 0) {  // the whole weight table: one bulk copy
        const uint32_t wbytes = (uint32_t)a.wk_floats * 4;
        mbar_arrive_expect_tx(wbar, wbytes);
        bulk_g2s(wk, a.w + kPrepHeaderFloats + (a.gsplit ? (size_t)g_fixed * a.wk_floats : 0), wbytes, wbar);
    }
    if (a.bulk && lane == 0) {
        for (int st = 0; st < n_pre; ++st) {
            const long item = gw + st * gstride;
            if (item < a.n_items) issue_load(item, st);
        }
    }
    {
        // sweep-ordered weights: wk[gl][kh][kw][i][CPP] = Ws[g][o][i][a(kh)][b(kw)]
        // (padding lanes o >= C of a row are never read into a stored result)
        constexpr int per_g = C * C * KH * KW;
        const int ng = a.gsplit ? 1 : s.G;
        if (a.prepared) {
            mbar_wait(wbar, 0);
        } else {
        const float* wsrc = a.w + (a.gsplit ? (long)g_fixed * per_g : 0);
        stage_weights(wsrc, ng * per_g, [&](int e, float v) {
            const int gl = e / per_g;
            const int g = a.gsplit ? g_fixed : gl;
            int r = e - gl * per_g;
            const int b = r % KW;
            r /= KW;
            const int aa = r % KH;
            r /= KH;
            const int i = r % C, o = r / C;
            const int ord = order_of(s.orders, g);
            const int kh = (ord & 2) ? aa : KH - 1 - aa;
            const int kw = (ord & 1) ? b : KW - 1 - b;
            wk[((gl * KH + kh) * KW + kw) * TS + i * CPP + o] = v;
        });
        __syncthreads();
        }
        if (threadIdx.x == 0) dbg_mark(a.dbg, 1);
    }

    // lane = p * CB + jj: the part index is the slow one, so the lanes of a 128-bit shared-memory
    // phase read the same weight vector (see finc_inverse_rw.cuh)
    const int p = lane / CB;       // part
    const int jj = lane - p * CB;  // column slot
    const int ncb = (W + CB - 1) / CB;

    // Taps in sweep coordinates (kh, kw) != (0,0).  NEAR taps (0,1) and (1,0) read pixels of the
    // previous anti-diagonal: they are the critical path.  FAR taps (kh + kw >= 2) read pixels
    // that are at least two diagonals old, so the far part of the NEXT pixel is computed in the
    // shadow of the current pixel's near part / shuffle reduction / triangular solve.
    // With P <= 2 parts per pixel the two streams are split (PIPE); with more parts the near
    // taps would sit on 2 of the P lanes only, so all taps stay in one balanced phase.
    constexpr bool PIPE = P <= 2;
    constexpr int SC = C;                                  // input channels per slot (a slot is a tap)
    constexpr int NNEAR = PIPE ? 2 : NT;                   // slots on the critical path
    constexpr int NFAR = PIPE ? NT - 2 : 0;                // slots computed one step ahead
    constexpr int NNL = (NNEAR + P - 1) / P;               // per lane
    constexpr int NFL = (NFAR + P - 1) / P;
    constexpr int NFLA = NFL > 0 ? NFL : 1;
    constexpr bool WN_REG = (NNL * SC * C <= 40);
    constexpr bool WF_REG = WN_REG && ((NNL + NFL) * SC * C <= 48);
    int nkh[NNL], nkw[NNL], nch[NNL], fkh[NFLA], fkw[NFLA], fch[NFLA];
#pragma unroll
    for (int m = 0; m < NNL; ++m) {
        const int e = m * P + p;
        nch[m] = 0;
        if constexpr (PIPE) {                     // 0 -> (0,1), 1 -> (1,0)
            nkh[m] = e < NNEAR ? (e == 1 ? 1 : 0) : KH + H;   // slots past the end never pass the bounds test
            nkw[m] = (e < NNEAR && e == 0) ? 1 : 0;
        } else {                                  // all taps, row-major without (0,0)
            nkh[m] = e < NNEAR ? (e + 1) / KW : KH + H;
            nkw[m] = e < NNEAR ? (e + 1) % KW : 0;
        }
    }
#pragma unroll
    for (int m = 0; m < NFL; ++m) {
        const int e = m * P + p;
        fch[m] = 0;
        const int q = e < KW - 2 ? e + 1 : e + 2;  // skip q = 0 (0,1) and q = KW-1 (1,0)
        fkh[m] = (e < NFAR) ? (q + 1) / KW : KH + H;
        fkw[m] = (e < NFAR) ? (q + 1) % KW : 0;
    }

    long k = 0;
    for (long item = gw; item < a.n_items; item += gstride, ++k) {
        const int st = (int)(k % a.S);
        const int g = item_g(item), n0 = item_n0(item);
        const int nt = min(a.T, s.B - n0);
        float* buf = bufs + st * stage_floats;
        if (a.bulk) {
            mbar_wait(&bars[st], (uint32_t)((k / a.S) & 1));
            if (threadIdx.x == 0 && k == 0) dbg_mark(a.dbg, 2);
        } else {
            for (int t = 0; t < nt; ++t) {
                const float* src = a.z + ((long)(n0 + t) * s.G + g) * a.tile_floats;
                for (int e = lane; e < a.tile_floats; e += 32) buf[t * a.tile_stride + e] = src[e];
            }
            __syncwarp();
        }
        const int ord = order_of(s.orders, g);
        const bool bot = ord & 2, right = ord & 1;
        const float* wg = wk + (size_t)(a.gsplit ? 0 : g) * KH * KW * TS;

        // per-item slot tables: shared-memory offset, weight row, register weights
        int noff[NNL], foff[NFLA];
        int noffi[NNL][SC], foffi[NFLA][SC];  // + i*HW, hoisted out of the step loop
        const float* nwp[NNL];
        const float* fwp[NFLA];
        float wn[WN_REG ? NNL : 1][WN_REG ? SC : 1][WN_REG ? C : 1];
        float wf[WF_REG ? NFLA : 1][WF_REG ? SC : 1][WF_REG ? C : 1];
        float wc[CREG ? C : 1][CREG ? C : 1];
#pragma unroll
        for (int m = 0; m < NNL; ++m) {
            const bool real = m * P + p < NNEAR;
            noff[m] = (bot ? nkh[m] : -nkh[m]) * W + (right ? nkw[m] : -nkw[m]) + nch[m] * HW;
            nwp[m] = wg + (size_t)((real ? nkh[m] : 0) * KW + nkw[m]) * TS + nch[m] * CPP;
#pragma unroll
            for (int i = 0; i < SC; ++i) noffi[m][i] = noff[m] + i * HW;
            if constexpr (WN_REG) {
#pragma unroll
                for (int i = 0; i < SC; ++i)
#pragma unroll
                    for (int o = 0; o < C; ++o) wn[m][i][o] = real ? -nwp[m][i * CPP + o] : 0.f;  // sign folded in
            }
        }
#pragma unroll
        for (int m = 0; m < NFL; ++m) {
            const bool real = m * P + p < NFAR;
            foff[m] = (bot ? fkh[m] : -fkh[m]) * W + (right ? fkw[m] : -fkw[m]) + fch[m] * HW;
            fwp[m] = wg + (size_t)((real ? fkh[m] : 0) * KW + fkw[m]) * TS + fch[m] * CPP;
#pragma unroll
            for (int i = 0; i < SC; ++i) foffi[m][i] = foff[m] + i * HW;
            if constexpr (WF_REG) {
#pragma unroll
                for (int i = 0; i < SC; ++i)
#pragma unroll
                    for (int o = 0; o < C; ++o) wf[m][i][o] = real ? -fwp[m][i * CPP + o] : 0.f;  // sign folded in
            }
        }
        if constexpr (CREG) {
#pragma unroll
            for (int i = 0; i < C; ++i)
#pragma unroll
                for (int o = 0; o < C; ++o) wc[i][o] = wg[i * CPP + o];
        }

        // acc[o] -= xv * w[i][o] for channel i of a slot (register weights are stored negated,
        // shared-memory weights use the negated-multiplicand FMA: no extra instruction either way)
        auto fma_row = [&](float (&acc)[C], float xv, const float* wrow, const float* wreg) {
            if (wreg != nullptr) {
#pragma unroll
                for (int o = 0; o < C; ++o) acc[o] = fmaf(xv, wreg[o], acc[o]);
            } else if constexpr (CPP % 4 == 0) {
#pragma unroll
                for (int v = 0; v < CPP / 4; ++v) {
                    const float4 f = *reinterpret_cast<const float4*>(wrow + 4 * v);
                    if (4 * v + 0 < C) acc[4 * v + 0] = fmaf(-xv, f.x, acc[4 * v + 0]);
                    if (4 * v + 1 < C) acc[4 * v + 1] = fmaf(-xv, f.y, acc[4 * v + 1]);
                    if (4 * v + 2 < C) acc[4 * v + 2] = fmaf(-xv, f.z, acc[4 * v + 2]);
                    if (4 * v + 3 < C) acc[4 * v + 3] = fmaf(-xv, f.w, acc[4 * v + 3]);
                }
            } else {
#pragma unroll
                for (int o = 0; o < C; ++o) acc[o] = fmaf(-xv, wrow[o], acc[o]);
            }
        };

        const int R = nt * H;  // stacked rows
        for (int cb = 0; cb < ncb; ++cb) {
            const int ws = cb * CB + jj;  // sweep column
            const int nc = min(CB, W - cb * CB);
            const bool col_on = jj < nc;
            const int wst = right ? W - 1 - ws : ws;
            const int nsteps = R + nc - 1;
            // current pixel (finished this step) and next pixel (far part prefetched this step)
            bool act_c = false;
            int hs_c = 0;
            float* base_c = buf;
            float accf[C];
#pragma unroll
            for (int o = 0; o < C; ++o) accf[o] = 0.f;
            int hn = 0;        // row of the next pixel inside its tile
            float* xn = buf;   // tile of the next pixel
            for (int step = -1; step < nsteps; ++step) {
                // (1) loads of the current pixel's near slots: the only ones that wait for the
                //     previous diagonal
                float xa[NNL][SC];
#pragma unroll
                for (int m = 0; m < NNL; ++m) {
                    const bool valid = act_c && hs_c >= nkh[m] && ws >= nkw[m];
#pragma unroll
                    for (int i = 0; i < SC; ++i) xa[m][i] = valid ? base_c[noffi[m][i]] : 0.f;
                }
                // (2) loads for the far part of the NEXT pixel (two or more diagonals old)
                const int rn = step + 1 - jj;
                const bool act_n = col_on && rn >= 0 && rn < R;
                float* base_n = xn + (bot ? H - 1 - hn : hn) * W + wst;
                float accn[C];
#pragma unroll
                for (int o = 0; o < C; ++o) accn[o] = (act_n && p == 0) ? base_n[o * HW] : 0.f;  // z enters once
                float xb[NFLA][SC];
#pragma unroll
                for (int m = 0; m < NFL; ++m) {
                    const bool valid = act_n && hn >= fkh[m] && ws >= fkw[m];
#pragma unroll
                    for (int i = 0; i < SC; ++i) xb[m][i] = valid ? base_n[foffi[m][i]] : 0.f;
                }
                // (3) near FMAs, then the P partial sums of the pixel start their shuffle reduction
                float acc[C];
#pragma unroll
                for (int o = 0; o < C; ++o) acc[o] = accf[o];
#pragma unroll
                for (int m = 0; m < NNL; ++m)
#pragma unroll
                    for (int i = 0; i < SC; ++i)
                        fma_row(acc, xa[m][i], nwp[m] + i * CPP, WN_REG ? &wn[WN_REG ? m : 0][WN_REG ? i : 0][0] : nullptr);
                if constexpr (P > 1) {
#pragma unroll
                    for (int off = CB; off < 32; off <<= 1)
#pragma unroll
                        for (int o = 0; o < C; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
                }
                // (4) far FMAs of the next pixel: independent work that fills the latency above
#pragma unroll
                for (int m = 0; m < NFL; ++m)
#pragma unroll
                    for (int i = 0; i < SC; ++i)
                        fma_row(accn, xb[m][i], fwp[m] + i * CPP, WF_REG ? &wf[WF_REG ? m : 0][WF_REG ? i : 0][0] : nullptr);
                // (5) corner tap: x[o] = acc[o] - sum_{i<o} W[o,i,corner] x[i]; one lane stores
#pragma unroll
                for (int i = 0; i < C - 1; ++i)
#pragma unroll
                    for (int o = i + 1; o < C; ++o) {
                        if constexpr (CREG) acc[o] = fmaf(-acc[i], wc[i][o], acc[o]);
                        else acc[o] = fmaf(-acc[i], wg[i * CPP + o], acc[o]);
                    }
                if (act_c && p == 0) {
#pragma unroll
                    for (int o = 0; o < C; ++o) base_c[o * HW] = acc[o];
                }
                __syncwarp();
#pragma unroll
                for (int o = 0; o < C; ++o) accf[o] = accn[o];
                act_c = act_n;
                hs_c = hn;
                base_c = base_n;
                if (act_n && ++hn == H) { hn = 0; xn += a.tile_stride; }
            }
        }

        if (threadIdx.x == 0 && k == 0) dbg_mark(a.dbg, 3);
        if (a.bulk) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                for (int t = 0; t < nt; ++t)
                    bulk_s2g(a.x + ((long)(n0 + t) * s.G + g) * a.tile_floats, buf + t * a.tile_stride,
                             (uint32_t)(a.tile_floats * 4));
                bulk_commit();
                if (a.S == 3) {
                    bulk_wait_read_1w();
                    const long nxt = item + 2 * gstride;
                    if (nxt < a.n_items) issue_load(nxt, (int)((k + 2) % 3));
                } else {
                    bulk_wait_read_all();
                    const long nxt = item + (long)a.S * gstride;
                    if (nxt < a.n_items) issue_load(nxt, st);
                }
            }
        } else {
            for (int t = 0; t < nt; ++t) {
                float* dst = a.x + ((long)(n0 + t) * s.G + g) * a.tile_floats;
                for (int e = lane; e < a.tile_floats; e += 32) dst[e] = buf[t * a.tile_stride + e];
            }
            __syncwarp();
        }
    }
    if (a.bulk && lane == 0) bulk_wait_all();
    if (threadIdx.x == 0) dbg_mark(a.dbg, 4);
}

template <int C, int KH, int KW, int P>
int launch_inst(const WaveArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    auto kern = inverse_wave_kernel<C, KH, KW, P>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    return launch_kernel(kern, grid, threads, smem, st, a);
}

// parts-per-pixel variants instantiated per channel count: wide tiles (small P) only occur
// with few channels, so the heavily unrolled (large C, small P) combinations are left to the
// generic tiled kernel
template <int C, int KH, int KW>
int dispatch_p(int P, const WaveArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if constexpr (C <= 6) {
        if (P == 1) return launch_inst<C, KH, KW, 1>(a, grid, threads, smem, st);
    }
    if constexpr (C <= 12) {
        if (P == 2) return launch_inst<C, KH, KW, 2>(a, grid, threads, smem, st);
    }
    if (P == 4) return launch_inst<C, KH, KW, 4>(a, grid, threads, smem, st);
    if (P == 8) return launch_inst<C, KH, KW, 8>(a, grid, threads, smem, st);
    return FINC_E_UNSUPPORTED;
}

// per-channel-count dispatch; explicitly instantiated in finc_inverse_wave_c<N>.cu so the
// instantiations compile in parallel
template <int C>
int dispatch_c(int kH, int P, const WaveArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if (kH == 3) return dispatch_p<C, 3, 3>(P, a, grid, threads, smem, st);
    return dispatch_p<C, 5, 5>(P, a, grid, threads, smem, st);
}

}  // namespace wave
}  // namespace finc
