// finc_chain.cu -- a CHAIN of FInC units (optionally each followed by its ActNorm o Conv1x1 affine map) in ONE
// launch: the tile of an image never leaves shared memory between units.
//
// Why: at training batch sizes every FInC launch moves 1.6-6.3 MB (< 1 us at HBM speed) and is launch / latency
// bound ("the only cure is fewer, fatter launches (fusion with the affine glue / persistent level
// kernels)").  The four groups of a unit act on disjoint channel quarters and a unit maps an image to an image,
// so a block of images can be carried through any number of consecutive units by one CTA:
//
//     cur = x[block]                                   (one TMA bulk load: the block is contiguous)
//     for unit u in order:   nxt = FInC_u(cur)         (conv_sub of finc_conv.cuh, shared memory -> shared memory)
//                            [nxt = A_u nxt + b_u]     (the Glow glue of fastflow_cifar_multi_gpu.py:238-256)
//                            y[u][block] = nxt         (coalesced 128-bit stores by all threads; ping-pong buffers)
//
// Uses: (1) FastFlowStep inference: FastFlowUnit + ActNorm + Conv1x1 = ONE launch (n_units = 1, affine given;
// reference: fastflow/fastflow.py:31-50, layers/actnorm.py:14-52, layers/conv1x1.py:18-43);
// (2) stacks of consecutive units (FincStack): the forward pass of a level (all activations are still written,
// the backward pass needs them) and its backward-data chain (transposed weights, units in reverse order) are
// one launch each instead of one per unit.
//
// Thread mapping: a sub-item is (tile, output-channel block, row, WT-wide strip) as in conv_cta_kernel; the
// threads of the CTA are split into G warp-aligned segments, one per group, so that the padding corner (and with
// it every branch of conv_sub) is warp-uniform.  The accumulation order of an output element (input channel,
// kernel row, kernel column) is the one of conv_cta_kernel: results are bit-identical to the per-unit launches.
#include <algorithm>
#include <cstdlib>

#include "finc_conv.cuh"

namespace finc {
namespace chain {

using conv::ObPad;

struct ChainArgs {
    const float* x;       // [B, G*C, H, W]: input of the first unit of the chain
    const float* w;       // raw weights, unit u at w + u * w_stride  ([G*C, C, 3, 3] each)
    long w_stride;
    float* y;             // unit u writes y + u * y_stride (y_stride == 0: only the last unit of the chain writes)
    long y_stride;
    const float* A;       // optional affine map after every unit: A + u * GC*GC, bias + u * GC (nullptr = none)
    const float* bias;
    float* logdet;        // optional [B]: sum over the chain's units of H*W*sum log|diag| (forward chains)
    int logdet_acc;
    Shape s;
    int n_units, u_first, u_step;
    int transpose;        // backward-data: transposed weights, flipped taps, opposite corner
    int IPB;              // images per block
    int seg;              // threads per group segment (multiple of 32)
    int nob, nstrip;
    int tile_floats;      // C*H*W
    int n_blocks;
    int wk_floats;        // staged weight table of one unit (all groups)
};

// raw weights per thread kept in registers between the prefetch and the table write
template <int CT>
struct Prefetch {
    static constexpr int value = CT <= 3 ? 3 : CT == 6 ? 11 : 24;
};

template <int CT, int OB, int WT>
__global__ void __launch_bounds__(512, 1) chain_kernel(const ChainArgs a) {
    constexpr int KH = 3, KW = 3, OBP = ObPad<OB>::value, PF = Prefetch<CT>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Shape& s = a.s;
    constexpr int C = CT;
    const int H = s.H, W = s.W, HW = H * W, G = s.G, GC = G * C;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int blk_floats = a.IPB * G * a.tile_floats;
    const bool affine = a.A != nullptr;

    // ---- shared memory: [weight table][affine][zero strip / front pad][buf0][pad][buf1][pad][barrier]
    float* wk = reinterpret_cast<float*>(smem_raw);
    float* Asm = wk + ((a.wk_floats + 31) & ~31);
    float* front = Asm + (affine ? ((GC * GC + GC + 31) & ~31) : 0);
    float* zrow = front + 8;                    // 8 zero floats, 16-byte aligned
    float* buf0 = front + conv::kFrontPad;
    float* buf1 = buf0 + ((blk_floats + 31) & ~31) + 32;   // 32 floats of halo slack between and behind the buffers
    uint64_t* in_bar = reinterpret_cast<uint64_t*>(buf1 + ((blk_floats + 31) & ~31) + 32);

    if (tid == 0) {
        mbar_init(in_bar, 1);
        fence_mbar_init();
    }
    if (tid < conv::kFrontPad) front[tid] = 0.f;

    // ---- weight staging: raw [g][o][i][a][b] -> wk[g][cin][a'][b'][ob][OBP].  The destination of element e is the
    // same for every unit: computed once; the values of the NEXT unit are fetched into registers before the
    // current unit is computed and written into the table after it (the L2 round trip hides behind the FMAs).
    const int per_g_raw = C * C * KH * KW;
    const int n_raw = G * per_g_raw;
    auto dst_of = [&](int e) -> int {
        const int gl = e / per_g_raw;
        int r = e - gl * per_g_raw;
        const int b = r % KW;
        r /= KW;
        const int aa = r % KH;
        r /= KH;
        const int i = r % C, o = r / C;
        int cin, cout, ap, bp;
        if (!a.transpose) { cin = i; cout = o; ap = aa; bp = b; }
        else { cin = o; cout = i; ap = KH - 1 - aa; bp = KW - 1 - b; }
        return ((((gl * C + cin) * KH + ap) * KW + bp) * a.nob + cout / OB) * OBP + cout % OB;
    };
    int wdst[PF];
#pragma unroll
    for (int q = 0; q < PF; ++q) {
        const int e = tid + q * nthr;
        wdst[q] = e < n_raw ? dst_of(e) : -1;
    }
    float wv[PF];
    auto fetch = [&](int u) {
        const float* wu = a.w + (long)u * a.w_stride;
#pragma unroll
        for (int q = 0; q < PF; ++q)
            if (wdst[q] >= 0) wv[q] = __ldg(wu + tid + q * nthr);
    };
    auto write_table = [&](int u) {
#pragma unroll
        for (int q = 0; q < PF; ++q)
            if (wdst[q] >= 0) wk[wdst[q]] = wv[q];
        const float* wu = a.w + (long)u * a.w_stride;
        for (int base = PF * nthr; base < n_raw; base += 8 * nthr) {   // large C: the rest, eight loads in flight
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = base + q * nthr + tid;
                if (e < n_raw) v[q] = __ldg(wu + e);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = base + q * nthr + tid;
                if (e < n_raw) wk[dst_of(e)] = v[q];
            }
        }
        if (affine) {
            const float* Au = a.A + (long)u * GC * GC;
            const float* bu = a.bias + (long)u * GC;
            for (int e = tid; e < GC * GC; e += nthr) Asm[e] = __ldg(Au + e);
            for (int e = tid; e < GC; e += nthr) Asm[GC * GC + e] = __ldg(bu + e);
        }
    };
    __syncthreads();
    pdl_wait();
    pdl_trigger();

    // logdet of the whole chain: one warp of CTA 0
    if (a.logdet != nullptr && blockIdx.x == 0 && tid < 32) {
        float ld = 0.f;
        for (int e = tid; e < a.n_units * GC; e += 32) {
            const int j = e / GC, r = e - j * GC;
            const int g = r / C, o = r - g * C;
            const int ord = order_of(s.orders, g);
            const float* wu = a.w + (long)(a.u_first + j * a.u_step) * a.w_stride;
            ld += logf(fabsf(__ldg(wu + (((long)g * C + o) * C + o) * KH * KW + corner_a(ord, KH) * KW + corner_b(ord, KW))));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, off);
        ld *= (float)H * (float)W;
        for (int n = tid; n < s.B; n += 32) a.logdet[n] = a.logdet_acc ? a.logdet[n] + ld : ld;
    }

    const int g_mine = tid / a.seg;                 // group segment of this thread (warp-uniform)
    const int lane_in_seg = tid - g_mine * a.seg;
    const int sub_per_tile = a.nob * H * a.nstrip;
    const int ord = g_mine < G ? (order_of(s.orders, g_mine) ^ (a.transpose ? 3 : 0)) : 0;
    const int r0 = (ord & 2) ? 0 : -(KH - 1);
    const bool right = (ord & 1) != 0;
    const float* wg = wk + (size_t)g_mine * C * KH * KW * a.nob * OBP;

    uint32_t in_phase = 0;
    bool table_ready = false;
    for (int blk = blockIdx.x; blk < a.n_blocks; blk += gridDim.x) {
        const int n0 = blk * a.IPB;
        const int nt = min(a.IPB, s.B - n0);
        const uint32_t blk_bytes = (uint32_t)(nt * G * a.tile_floats) * 4;
        float* cur = buf0;
        float* nxt = buf1;
        if (tid == 0) {
            mbar_arrive_expect_tx(in_bar, blk_bytes);
            bulk_g2s(cur, a.x + (long)n0 * G * a.tile_floats, blk_bytes, in_bar);
        }
        if (!table_ready) {   // first block (a one-unit chain keeps its table for all blocks)
            fetch(a.u_first);
            write_table(a.u_first);
            table_ready = true;
            __syncthreads();
        }
        for (int j = 0; j < a.n_units; ++j) {
            const int u = a.u_first + j * a.u_step;
            const bool more_blocks = blk + (int)gridDim.x < a.n_blocks;
            const int u_next = j + 1 < a.n_units ? u + a.u_step : (more_blocks && a.n_units > 1 ? a.u_first : -1);
            if (u_next >= 0) fetch(u_next);
            if (j == 0) {
                mbar_wait(in_bar, in_phase);
                in_phase ^= 1;
            }
            // ---- FInC unit u: cur -> nxt ---------------------------------------------------------------------
            if (g_mine < G) {
                const int nsub_g = nt * sub_per_tile;
                for (int sub = lane_in_seg; sub < nsub_g; sub += a.seg) {
                    int r = sub;
                    const int strip = r % a.nstrip;
                    r /= a.nstrip;
                    const int h = r % H;
                    r /= H;
                    const int ob = r % a.nob;
                    const int t = r / a.nob;             // image inside the block
                    const int w0 = strip * WT;
                    const float* xt = cur + (t * G + g_mine) * a.tile_floats;
                    float* yt = nxt + (t * G + g_mine) * a.tile_floats + h * W + w0;
                    if (right) conv::conv_sub<CT, OB, WT, KH, KW, true, 1>(xt, zrow, wg, yt, C, H, W, HW, h, w0, r0, ob, a.nob);
                    else conv::conv_sub<CT, OB, WT, KH, KW, false, 1>(xt, zrow, wg, yt, C, H, W, HW, h, w0, r0, ob, a.nob);
                }
            }
            float* res = nxt;
            if (affine) {
                // ---- res[c', p] = sum_c A[c', c] nxt[c, p] + b[c']  -> cur (its content is consumed) ------------
                __syncthreads();
                const int items = nt * HW * ((GC + 3) / 4);
                for (int it = tid; it < items; it += nthr) {
                    const int p = it % HW;
                    int r = it / HW;
                    const int t = r % nt;
                    const int o4 = (r / nt) * 4;
                    const float* src = nxt + (size_t)t * GC * HW + p;
                    float acc[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = (o4 + q < GC) ? Asm[GC * GC + o4 + q] : 0.f;
                    for (int c = 0; c < GC; ++c) {
                        const float v = src[c * HW];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (o4 + q < GC) acc[q] = fmaf(Asm[(o4 + q) * GC + c], v, acc[q]);
                    }
                    float* dst = cur + (size_t)t * GC * HW + p;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (o4 + q < GC) dst[(o4 + q) * HW] = acc[q];
                }
                res = cur;
            }
            // ---- store the block: plain coalesced 128-bit stores by all threads (fire and forget; a TMA bulk store
            // per unit put its shared-memory read latency on the critical path of a 16-unit chain) ------------------
            __syncthreads();   // (A) every thread is done with the table and with cur; res is complete
            const bool last = j + 1 == a.n_units;
            if (a.y_stride != 0 || last) {
                float4* dst = reinterpret_cast<float4*>(a.y + (long)u * a.y_stride + (long)n0 * G * a.tile_floats);
                const float4* src = reinterpret_cast<const float4*>(res);
                const int n4 = nt * G * a.tile_floats / 4;
                for (int e = tid; e < n4; e += nthr) dst[e] = src[e];
            }
            if (u_next >= 0) write_table(u_next);
            if (!affine) {   // ping-pong
                float* tmp = cur;
                cur = nxt;
                nxt = tmp;
            }
            __syncthreads();   // (B) table of the next unit ready; res has been read
        }
    }
}

struct Plan {
    int OB, WT, IPB, seg, threads, nob, nstrip;
    size_t smem;
};

static bool make_plan(const Shape& s, bool affine, Plan& p) {
    const int C = s.C;
    if (s.kH != 3 || s.kW != 3 || s.G < 1 || s.G > 4) return false;
    if (!(C == 1 || C == 2 || C == 3 || C == 6 || C == 12 || C == 24)) return false;
    // small output-channel blocks: a block of images is all the parallelism a CTA has, and at 4x4 / 8x8 tiles the
    // per-unit critical path (one warp's FMA chain), not the FMA count, sets the time
    p.OB = C <= 3 ? C : (C == 6 ? 2 : (C == 12 ? 1 : 4));
    p.WT = s.W % 4 == 0 ? 4 : (s.W % 2 == 0 ? 2 : 1);
    p.nob = (C + p.OB - 1) / p.OB;
    p.nstrip = s.W / p.WT;
    const int tile = C * s.H * s.W;
    if ((s.G * tile) % 4 != 0) return false;   // 16-byte TMA bulk granularity of an image
    const int sub_per_tile = p.nob * s.H * p.nstrip;
    const size_t max_smem = max_optin_smem_cached();
    const int OBP = p.OB <= 2 ? p.OB : ((p.OB + 3) / 4) * 4;   // ObPad<OB>::value
    const int wk = s.G * C * 9 * p.nob * OBP;
    const int GC = s.G * C;
    auto smem_for = [&](int ipb) {
        const size_t blk = ((size_t)ipb * s.G * tile + 31) & ~(size_t)31;
        return (size_t)4 * (((wk + 31) & ~31) + (affine ? ((GC * GC + GC + 31) & ~31) : 0) + conv::kFrontPad + 2 * (blk + 32)) + 64 + 128;
    };
    double best = -1.0;
    int best_ipb = 0;
    const int sms = sm_count_cached();
    for (int ipb = 1; ipb <= 32 && ipb <= s.B; ++ipb) {
        if (smem_for(ipb) > max_smem) break;
        if (ipb > 1 && smem_for(ipb) > 100 * 1024) break;   // keep two CTAs per SM possible
        const int nsub = ipb * sub_per_tile;
        const int seg = nsub >= 128 ? 128 : (nsub + 31) / 32 * 32;
        const double util = (double)nsub / ((nsub + seg - 1) / seg * seg);
        const int blocks = (s.B + ipb - 1) / ipb;
        const double fill = blocks >= sms ? 1.0 : (double)blocks / sms;
        const double score = util * fill * (seg >= 64 ? 1.0 : 0.9);
        if (score > best + 1e-9) { best = score; best_ipb = ipb; }
    }
    if (best_ipb == 0) return false;
    if (const char* e = getenv("FINC_CHAIN_IPB")) {   // experiment knob
        const int v = atoi(e);
        if (v >= 1 && v <= s.B && smem_for(v) <= max_smem) best_ipb = v;
    }
    p.IPB = best_ipb;
    const int nsub = p.IPB * sub_per_tile;
    p.seg = nsub >= 128 ? 128 : (nsub + 31) / 32 * 32;
    p.threads = p.seg * s.G;
    if (affine && p.threads < 128) p.threads = 128;
    // enough threads that the register prefetch covers a unit's weights (the extra warps only stage and copy)
    const int pf = C <= 3 ? 3 : C == 6 ? 11 : 24;   // Prefetch<CT>::value
    const int need = ((s.G * C * C * 9 + pf - 1) / pf + 31) / 32 * 32;
    if (p.threads < need) p.threads = need < 512 ? need : 512;
    p.smem = smem_for(p.IPB);
    return true;
}

template <int CT, int OB>
static int launch_ct(const ChainArgs& a, const Plan& p, dim3 grid, cudaStream_t st) {
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) return (int)e;
        return launch_kernel(kern, grid, p.threads, p.smem, st, a);
    };
    switch (p.WT) {
        case 4: return go(chain_kernel<CT, OB, 4>);
        case 2: return go(chain_kernel<CT, OB, 2>);
        default: return go(chain_kernel<CT, OB, 1>);
    }
}

}  // namespace chain
}  // namespace finc

using namespace finc;

extern "C" {

int finc_chain_supported(int G, int C, int H, int W, int kH, int kW, int with_affine) {
    Shape s{1 << 20, G, C, H, W, kH, kW, 0};
    chain::Plan p;
    return chain::make_plan(s, with_affine != 0, p) ? 1 : 0;
}

int finc_chain_f32(const float* x, const float* w, long w_stride, float* y, long y_stride, const float* A,
                   const float* bias, float* logdet, int B, int G, int C, int H, int W, int kH, int kW,
                   unsigned orders, int n_units, int u_first, int u_step, unsigned flags, void* stream) {
    if (!x || !w || !y || B < 0 || n_units < 1 || (A == nullptr) != (bias == nullptr)) return FINC_E_BADARG;
    if (x == y && y_stride == 0) return FINC_E_BADARG;
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) & 15) return FINC_E_UNSUPPORTED;
    if (((size_t)w_stride * 4) % 16 != 0 || ((size_t)y_stride * 4) % 16 != 0) return FINC_E_UNSUPPORTED;
    Shape s{B, G, C, H, W, kH, kW, orders};
    chain::Plan p;
    if (!chain::make_plan(s, A != nullptr, p)) return FINC_E_UNSUPPORTED;
    if (B == 0) return FINC_OK;
    chain::ChainArgs a{};
    a.x = x; a.w = w; a.w_stride = w_stride; a.y = y; a.y_stride = y_stride; a.A = A; a.bias = bias;
    a.logdet = logdet; a.logdet_acc = (flags & FINC_FLAG_LOGDET_ACCUMULATE) ? 1 : 0;
    a.s = s; a.n_units = n_units; a.u_first = u_first; a.u_step = u_step;
    a.transpose = (flags & FINC_FLAG_CHAIN_TRANSPOSE) ? 1 : 0;
    a.IPB = p.IPB; a.seg = p.seg; a.nob = p.nob; a.nstrip = p.nstrip;
    a.tile_floats = C * H * W;
    a.n_blocks = (B + p.IPB - 1) / p.IPB;
    const int OBP = p.OB <= 2 ? p.OB : ((p.OB + 3) / 4) * 4;
    a.wk_floats = G * C * 9 * p.nob * OBP;
    const int sms = sm_count_cached();
    const int per_sm = p.smem <= 100 * 1024 && p.threads <= 512 ? 2 : 1;
    dim3 grid((unsigned)std::min(a.n_blocks, sms * per_sm));
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 1: return chain::launch_ct<1, 1>(a, p, grid, st);
        case 2: return chain::launch_ct<2, 2>(a, p, grid, st);
        case 3: return chain::launch_ct<3, 3>(a, p, grid, st);
        case 6: return chain::launch_ct<6, 2>(a, p, grid, st);
        case 12: return chain::launch_ct<12, 1>(a, p, grid, st);
        case 24: return chain::launch_ct<24, 4>(a, p, grid, st);
        default: return FINC_E_UNSUPPORTED;
    }
}

}  // extern "C"
