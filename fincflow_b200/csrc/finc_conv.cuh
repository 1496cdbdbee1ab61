// finc_conv.cuh -- fused FInC forward (+logdet) and backward-input convolution for sm_100a.
//
// One launch covers a whole [B, G*C, H, W] tensor: all G groups (the four padding corners
// of a FastFlowUnit), no F.pad copy, no chunk/cat copies (reference:
// fastflow/fastflow.py:31-50, layers/conv.py:102-107).
//
// Execution model: persistent CTAs, TMA producer / FMA consumers.
//   * a CHUNK is CH tiles (n0..n0+CH-1, g) of ONE group: CH 1-D TMA bulk copies
//     (cp.async.bulk) that complete on a `full` mbarrier.  Keeping a chunk inside one group
//     makes the padding corner -- hence every branch of the inner loop -- CTA-uniform.
//   * one producer warp (an elected lane) runs ahead through a ring of S stages, throttled
//     by `empty` mbarriers the consumer warps arrive on; no __syncthreads after start-up.
//   * the consumer threads split a chunk into sub-items (tile, output-channel block, row,
//     WT-wide strip); CH is chosen so that a chunk is about one sub-item per thread for every
//     tile shape from 48x4x4 to 12x32x32, and shrinks until every SM has a chunk when the
//     batch is small.  A lane register-blocks OB output channels x WT pixels; per (input
//     channel, kernel row) it loads the row strip with two vector loads and the OB weights
//     of each tap with one broadcast vector load from wk[g][cin][a][b][cout-block][OBP].
//   * the inner loop is branch-free: rows and halos that fall outside the image are
//     redirected (pointer select, hoisted out of the channel loop) to a zero strip in shared
//     memory; with the channel count a template parameter all shared-memory offsets of the
//     weight table are immediates.
//   * results leave as 128-bit global stores; logdet (H*W*sum log|diag|) is written by the
//     otherwise idle producer warp of CTA 0.
//
// Backward-input is the same kernel: dx = conv(dz) with the weights transposed (o<->i),
// both kernel axes reversed and the opposite padding corner -- folded into the weight
// staging; the inner loop is identical.
#pragma once
#include "finc_common.cuh"

namespace finc {
namespace conv {

constexpr int kMaxConsumerWarps = 16;
constexpr int kFrontPad = 32;  // floats in front of the first stage: zero strip + halo slack

struct ConvArgs {
    const float* x;
    const float* w;
    float* y;
    float* logdet;    // nullable; [B], written by CTA (0,0) (forward only)
    int logdet_acc;   // 1: logdet[n] += value
    Shape s;
    int transpose;
    int CH;           // tiles per chunk
    int S;            // pipeline stages
    int nob;          // output-channel blocks
    int gsplit;       // 1: blockIdx.y selects the group and only its weights are staged
    int bulk;         // 1: TMA bulk copies usable (alignment / size)
    int prepared;     // 1: a.w is a prepared table [G][wk_per_g] (after the header): one bulk copy
    int tile_floats;  // C*H*W
    int wk_floats;    // weight floats in smem (all groups or one)
    int nstrip;       // W / WT
    int RB;           // output rows per sub-item (1, or 2 for the row-blocked variant)
    int HR;           // row blocks per tile = ceil(H / RB)
    unsigned m_nstrip, m_h, m_nob;  // magic multipliers for division by nstrip, H, nob
    unsigned long long* dbg;  // debug timestamps (nullptr = off)
    int nblk;         // image blocks = ceil(B / CH)
    long n_chunks;    // nblk * G (or nblk when gsplit)
};

template <int OB>
struct ObPad {
    static constexpr int value = OB <= 2 ? OB : ((OB + 3) / 4) * 4;
};

// n / d for n, d < 65536 via one multiply-high; m = ceil(2^32 / d), m == 0 encodes d == 1
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned m) { return m ? __umulhi(n, m) : n; }

// N consecutive floats from an address aligned to min(16, 4*N) bytes (N in {1,2,4})
template <int N>
__device__ __forceinline__ void ldn(const float* p, float* out) {
    if constexpr (N == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else if constexpr (N == 2) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        out[0] = v.x; out[1] = v.y;
    } else {
        out[0] = p[0];
    }
}

// halo of HALO floats; vector width limited by the strip width WT (alignment follows WT)
template <int WT, int HALO>
__device__ __forceinline__ void ld_halo(const float* p, float* out) {
    if constexpr (HALO == 0) {
    } else if constexpr (WT == 4 && HALO == 4) { ldn<4>(p, out); }
    else if constexpr (WT >= 2 && HALO % 2 == 0) {
#pragma unroll
        for (int q = 0; q < HALO; q += 2) ldn<2>(p + q, out + q);
    } else {
#pragma unroll
        for (int q = 0; q < HALO; ++q) out[q] = p[q];
    }
}

// one sub-item: OB output channels x WT pixels of rows h .. h+RB-1.  RIGHT = padded on the right.
// RB = 2 (row-blocked variant for the HBM-bound small-channel levels at large batch): the KH+1 input
// row strips and every weight vector are loaded once for two output rows, and the index arithmetic
// of a sub-item is amortised over twice the FMAs.
template <int CT, int OB, int WT, int KH, int KW, bool RIGHT, int RB>
__device__ __forceinline__ void conv_sub(const float* __restrict__ xt, const float* __restrict__ zrow,
                                         const float* __restrict__ wg, float* __restrict__ yt, int C, int H, int W,
                                         int HW, int h, int w0, int r0, int ob, int nob) {
    constexpr int OBP = ObPad<OB>::value;
    constexpr int HALO = KW - 1;
    constexpr int NR = KH + RB - 1;   // input rows feeding the RB output rows
    const int nobp = (CT > 0 ? (CT + OB - 1) / OB : nob) * OBP;
    const float* midp[NR];
    const float* halop[NR];
    int mstr[NR], hstr[NR];
#pragma unroll
    for (int ar = 0; ar < NR; ++ar) {
        const int hh = h + r0 + ar;
        const bool ok = hh >= 0 && hh < H;
        const float* row = xt + hh * W + w0;
        const bool hok = ok && (RIGHT ? (w0 + WT < W) : (w0 > 0));
        midp[ar] = ok ? row : zrow;
        mstr[ar] = ok ? HW : 0;
        halop[ar] = hok ? (RIGHT ? row + WT : row - HALO) : zrow;
        hstr[ar] = hok ? HW : 0;
    }
    float acc[RB][OB][WT];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int o = 0; o < OB; ++o)
#pragma unroll
            for (int q = 0; q < WT; ++q) acc[rb][o][q] = 0.f;
    const float* wb = wg + ob * OBP;

    // the vector halo (redirected as a whole to the zero strip) needs the halo to fit in the
    // neighbouring strip; narrow strips fall back to per-element column predicates
    constexpr bool VHALO = HALO <= WT;
    bool colok[VHALO ? 1 : HALO];
    if constexpr (!VHALO) {
#pragma unroll
        for (int q = 0; q < HALO; ++q) {
            const int col = RIGHT ? w0 + WT + q : w0 - HALO + q;
            colok[q] = col >= 0 && col < W;
        }
    }
    auto load_strip = [&](int cin, int ar, float (&xs)[WT + HALO]) {
        if constexpr (VHALO) {
            if constexpr (RIGHT) {
                ldn<WT>(midp[ar] + cin * mstr[ar], xs);
                ld_halo<WT, HALO>(halop[ar] + cin * hstr[ar], xs + WT);
            } else {
                ld_halo<WT, HALO>(halop[ar] + cin * hstr[ar], xs);
                ldn<WT>(midp[ar] + cin * mstr[ar], xs + HALO);
            }
        } else {
            const float* mp = midp[ar] + cin * mstr[ar];  // zero strip when the row is outside
            ldn<WT>(mp, xs + (RIGHT ? 0 : HALO));
#pragma unroll
            for (int q = 0; q < HALO; ++q) {
                const float v = mp[RIGHT ? WT + q : q - HALO];
                xs[RIGHT ? WT + q : q] = colok[q] ? v : 0.f;
            }
        }
    };
    auto load_w = [&](const float* wp, int bp, float (&wv)[OB]) {
        if constexpr (OBP % 4 == 0) {
#pragma unroll
            for (int v = 0; v < OBP / 4; ++v) {
                const float4 f = *reinterpret_cast<const float4*>(wp + bp * nobp + 4 * v);
                if (4 * v + 0 < OB) wv[4 * v + 0] = f.x;
                if (4 * v + 1 < OB) wv[4 * v + 1] = f.y;
                if (4 * v + 2 < OB) wv[4 * v + 2] = f.z;
                if (4 * v + 3 < OB) wv[4 * v + 3] = f.w;
            }
        } else if constexpr (OBP == 2) {
            const float2 f = *reinterpret_cast<const float2*>(wp + bp * nobp);
            wv[0] = f.x; wv[1] = f.y;
        } else {
            wv[0] = wp[bp * nobp];
        }
    };
    auto body = [&](int cin) {
        if constexpr (RB == 1) {
#pragma unroll
            for (int ap = 0; ap < KH; ++ap) {
                float xs[WT + HALO];
                load_strip(cin, ap, xs);
                const float* wp = wb + (size_t)((cin * KH + ap) * KW) * nobp;
#pragma unroll
                for (int bp = 0; bp < KW; ++bp) {
                    float wv[OB];
                    load_w(wp, bp, wv);
#pragma unroll
                    for (int o = 0; o < OB; ++o)
#pragma unroll
                        for (int q = 0; q < WT; ++q) acc[0][o][q] = fmaf(wv[o], xs[q + bp], acc[0][o][q]);
                }
            }
        } else {
            float xs[NR][WT + HALO];
#pragma unroll
            for (int ar = 0; ar < NR; ++ar) load_strip(cin, ar, xs[ar]);
#pragma unroll
            for (int ap = 0; ap < KH; ++ap) {
                const float* wp = wb + (size_t)((cin * KH + ap) * KW) * nobp;
#pragma unroll
                for (int bp = 0; bp < KW; ++bp) {
                    float wv[OB];
                    load_w(wp, bp, wv);
#pragma unroll
                    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
                        for (int o = 0; o < OB; ++o)
#pragma unroll
                            for (int q = 0; q < WT; ++q) acc[rb][o][q] = fmaf(wv[o], xs[ap + rb][q + bp], acc[rb][o][q]);
                }
            }
        }
    };
    // full unrolling only while the straight-line code of both padding variants stays inside the
    // instruction cache (C = 6, OB = 6 fully unrolled: 41 % of the stall samples were instruction fetch)
    if constexpr (CT > 0 && CT * KH * KW * OB * WT * RB <= 1000) {
#pragma unroll
        for (int cin = 0; cin < CT; ++cin) body(cin);
    } else {
        const int Cn = CT > 0 ? CT : C;
#pragma unroll 2
        for (int cin = 0; cin < Cn; ++cin) body(cin);
    }
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
        if (RB > 1 && h + rb >= H) break;
#pragma unroll
        for (int o = 0; o < OB; ++o) {
            const int oc = ob * OB + o;
            if (oc < C) {
                float* yp = yt + oc * HW + rb * W;
                if constexpr (WT == 4) {
                    *reinterpret_cast<float4*>(yp) = make_float4(acc[rb][o][0], acc[rb][o][1], acc[rb][o][2], acc[rb][o][3]);
                } else if constexpr (WT == 2) {
                    *reinterpret_cast<float2*>(yp) = make_float2(acc[rb][o][0], acc[rb][o][1]);
                } else {
                    yp[0] = acc[rb][o][0];
                }
            }
        }
    }
}

__device__ __forceinline__ int nob_check(int nob_arg, int ct, int ob) { return ct > 0 ? (ct + ob - 1) / ob : nob_arg; }

template <int CT, int OB, int WT, int KH, int KW, int RB = 1>
__global__ void __launch_bounds__((kMaxConsumerWarps + 1) * 32, 1) conv_cta_kernel(const ConvArgs a) {
    constexpr int OBP = ObPad<OB>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* wk = reinterpret_cast<float*>(smem_raw);
    const Shape& s = a.s;
    const int C = CT > 0 ? CT : s.C;
    const int H = s.H, W = s.W;
    const int HW = H * W;
    const int nthreads_c = blockDim.x - 32;  // consumer threads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = warp == (nthreads_c >> 5);
    const int stage_floats = a.CH * a.tile_floats;
    const int wk_pad = (a.wk_floats + 31) & ~31;
    float* zrow = wk + wk_pad + 8;  // 8 zero floats at [8,16) of the front pad, 16-byte aligned
    float* bufs = wk + wk_pad + kFrontPad;
    uint64_t* full = reinterpret_cast<uint64_t*>(bufs + (size_t)a.S * stage_floats + 8);  // +8: halo slack behind
    uint64_t* empty = full + a.S;
    uint64_t* wbar = empty + a.S;  // prepared weight table landed
    const int g_fixed = a.gsplit ? (int)blockIdx.y : -1;

    auto chunk_g = [&](long chunk) -> int { return a.gsplit ? g_fixed : (int)(chunk % s.G); };
    auto chunk_n0 = [&](long chunk) -> int { return (int)(a.gsplit ? chunk : chunk / s.G) * a.CH; };
    auto issue_load = [&](long chunk, int st) {  // the whole producer warp: copies are issued by 32 lanes
        const int g = chunk_g(chunk), n0 = chunk_n0(chunk);
        const int nt = min(a.CH, s.B - n0);
        if (lane == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)(nt * a.tile_floats * 4));
        __syncwarp();
        for (int t = lane; t < nt; t += 32)
            bulk_g2s(bufs + (size_t)st * stage_floats + t * a.tile_floats,
                     a.x + ((long)(n0 + t) * s.G + g) * a.tile_floats, (uint32_t)(a.tile_floats * 4), &full[st]);
    };

    if (threadIdx.x == 0) dbg_mark(a.dbg, 0);
    // ---- start-up: barriers and first loads, then the weight staging (overlaps the loads) ----
    if (a.bulk && is_producer && lane == 0) {
        for (int st = 0; st < a.S; ++st) {
            mbar_init(&full[st], 1);
            mbar_init(&empty[st], nthreads_c >> 5);
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    pdl_wait();     // predecessor (producer of x / last writer of y) has completed
    pdl_trigger();  // successor may start its own prologue
    if (a.bulk && is_producer) {
        if (a.prepared && lane == 0) {  // the whole weight table: one bulk copy
            // a table is only valid for the plan it was prepared for (kind, output block, blocks): refuse others
            if (blockIdx.x == 0 && blockIdx.y == 0 &&
                (__ldg(a.w) != 1179208259.f || (int)__ldg(a.w + 1) != a.transpose || (int)__ldg(a.w + 2) != OB ||
                 (int)__ldg(a.w + 3) != nob_check(a.nob, CT, OB)))
                __trap();
            const uint32_t wbytes = (uint32_t)a.wk_floats * 4;
            mbar_arrive_expect_tx(wbar, wbytes);
            bulk_g2s(wk, a.w + kPrepHeaderFloats + (a.gsplit ? (size_t)g_fixed * a.wk_floats : 0), wbytes, wbar);
        }
        for (int st = 0; st < a.S; ++st) {
            const long chunk = blockIdx.x + (long)st * gridDim.x;
            if (chunk < a.n_chunks) issue_load(chunk, st);
        }
    }
    {
        // wk[gl][cin][a'][b'][ob][OBP]; padding slots (cout >= C) stay uninitialised: they only
        // feed accumulators that are never stored
        constexpr int kk = KH * KW;
        const int per_g = C * C * kk;
        const int ng = a.gsplit ? 1 : s.G;
        if (threadIdx.x < kFrontPad) wk[wk_pad + threadIdx.x] = 0.f;
        if (!a.prepared) {
        const float* wsrc = a.w + (a.gsplit ? (long)g_fixed * per_g : 0);
        stage_weights(wsrc, ng * per_g, [&](int e, float v) {
            const int gl = e / per_g;
            int r = e - gl * per_g;
            const int b = r % KW;
            r /= KW;
            const int aa = r % KH;
            r /= KH;
            const int i = r % C, o = r / C;
            int cin, cout, ap, bp;
            if (!a.transpose) { cin = i; cout = o; ap = aa; bp = b; }
            else { cin = o; cout = i; ap = KH - 1 - aa; bp = KW - 1 - b; }
            wk[((((gl * C + cin) * KH + ap) * KW + bp) * a.nob + cout / OB) * OBP + cout % OB] = v;
        });
        }
        __syncthreads();
        if (threadIdx.x == 0) dbg_mark(a.dbg, 1);
    }

    if (is_producer) {
        // ---- producer warp: keep the ring full; CTA 0 also writes logdet -------------------------
        if (a.logdet != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
            float ld;
            if (a.prepared) {
                ld = __ldg(a.w + 4);  // computed once by finc_prepare_weights_f32
            } else {
                ld = 0.f;
                for (int e = lane; e < s.G * C; e += 32) {
                    const int g = e / C, o = e - g * C;
                    const int ord = order_of(s.orders, g);
                    ld += logf(fabsf(__ldg(a.w + (((long)g * C + o) * C + o) * KH * KW + corner_a(ord, KH) * KW + corner_b(ord, KW))));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, off);
                ld *= (float)H * (float)W;
            }
            for (int n = lane; n < s.B; n += 32) a.logdet[n] = a.logdet_acc ? a.logdet[n] + ld : ld;
        }
        if (a.bulk) {
            long k = a.S;
            for (long chunk = blockIdx.x + (long)a.S * gridDim.x; chunk < a.n_chunks; chunk += gridDim.x, ++k) {
                const int st = (int)(k % a.S);
                mbar_wait(&empty[st], (uint32_t)(((k / a.S) - 1) & 1));  // consumers released use #(k/S - 1)
                issue_load(chunk, st);
            }
        }
        return;
    }

    // ---- consumers --------------------------------------------------------------------------------
    const int nob = CT > 0 ? (CT + OB - 1) / OB : a.nob;
    const int sub_per_tile = nob * a.HR * a.nstrip;
    if (a.prepared) mbar_wait(wbar, 0);
    long k = 0;
    for (long chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x, ++k) {
        const int st = (int)(k % a.S);
        const int g = chunk_g(chunk), n0 = chunk_n0(chunk);
        const int nt = min(a.CH, s.B - n0);
        float* buf = bufs + (size_t)st * stage_floats;
        if (a.bulk) {
            mbar_wait(&full[st], (uint32_t)((k / a.S) & 1));
            if (threadIdx.x == 0 && k == 0) dbg_mark(a.dbg, 2);
        } else {
            // unaligned tensors: cooperative copy by the consumer threads
            asm volatile("bar.sync 1, %0;" ::"r"(nthreads_c) : "memory");  // previous chunk fully consumed
            for (int t = 0; t < nt; ++t) {
                const float* src = a.x + ((long)(n0 + t) * s.G + g) * a.tile_floats;
                for (int e = threadIdx.x; e < a.tile_floats; e += nthreads_c) buf[t * a.tile_floats + e] = src[e];
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nthreads_c) : "memory");
        }
        const int ord = order_of(s.orders, g) ^ (a.transpose ? 3 : 0);
        const int r0 = (ord & 2) ? 0 : -(KH - 1);
        const bool right = (ord & 1) != 0;
        const float* wg = wk + (size_t)(a.gsplit ? 0 : g) * C * KH * KW * nob * OBP;
        const int nsub = nt * sub_per_tile;
        for (int sub = threadIdx.x; sub < nsub; sub += nthreads_c) {
            unsigned r = (unsigned)sub;
            unsigned q = fastdiv(r, a.m_nstrip);
            const int strip = (int)(r - q * a.nstrip);
            r = q;
            q = fastdiv(r, a.m_h);
            const int h = (int)(r - q * a.HR) * RB;
            r = q;
            q = CT > 0 ? r / (unsigned)nob : fastdiv(r, a.m_nob);
            const int ob = (int)(r - q * nob);
            const int t = (int)q;
            const int w0 = strip * WT;
            const float* xt = buf + t * a.tile_floats;
            float* yt = a.y + ((long)(n0 + t) * s.G + g) * a.tile_floats + h * W + w0;
            if (right) conv_sub<CT, OB, WT, KH, KW, true, RB>(xt, zrow, wg, yt, C, H, W, HW, h, w0, r0, ob, nob);
            else conv_sub<CT, OB, WT, KH, KW, false, RB>(xt, zrow, wg, yt, C, H, W, HW, h, w0, r0, ob, nob);
        }
        if (threadIdx.x == 0) dbg_mark(a.dbg, 3);
        if (a.bulk) {
            __syncwarp();
            if (lane == 0) {  // release the stage: one arrival per consumer warp
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
            }
        }
    }
}

// the row-blocked variant exists for the shapes the host planner may select it for (conv_rb_supported)
constexpr bool rb2_instantiated(int CT, int OB, int WT, int KH) { return CT >= 1 && CT <= 3 && OB == CT && WT == 4 && KH == 3; }

template <int CT, int OB, int WT, int KH, int KW>
int launch_inst(const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if constexpr (rb2_instantiated(CT, OB, WT, KH)) {
        if (a.RB == 2) {
            auto kern2 = conv_cta_kernel<CT, OB, WT, KH, KW, 2>;
            cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return (int)e2;
            return launch_kernel(kern2, grid, threads, smem, st, a);
        }
    }
    if (a.RB != 1) return FINC_E_UNSUPPORTED;
    auto kern = conv_cta_kernel<CT, OB, WT, KH, KW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    return launch_kernel(kern, grid, threads, smem, st, a);
}

template <int CT, int OB, int WT>
int dispatch_k(int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if (KH == 3) return launch_inst<CT, OB, WT, 3, 3>(a, grid, threads, smem, st);
    if (KH == 5) return launch_inst<CT, OB, WT, 5, 5>(a, grid, threads, smem, st);
    if constexpr (CT == 0) {
        if (KH == 2) return launch_inst<CT, OB, WT, 2, 2>(a, grid, threads, smem, st);
    }
    return FINC_E_UNSUPPORTED;
}

template <int CT, int OB>
int dispatch_wt(int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    switch (WT) {
        case 4: return dispatch_k<CT, OB, 4>(KH, a, grid, threads, smem, st);
        case 2: return dispatch_k<CT, OB, 2>(KH, a, grid, threads, smem, st);
        default:
            if constexpr (CT == 0) return dispatch_k<CT, OB, 1>(KH, a, grid, threads, smem, st);
            return FINC_E_UNSUPPORTED;
    }
}

// (CT, OB) pairs instantiated; defined in finc_conv_c<N>.cu so they compile in parallel
template <int CT>
int dispatch_ob(int OB, int WT, int KH, const ConvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st);

}  // namespace conv
}  // namespace finc
