// tc_wgrad.cuh -- weight-gradient GEMM on the tensor cores: a reduction over PIXELS
//
//   D[m, n] = sum_p P[p, m] * Q[p, n]        P: [np, ldP], Q: [np, ldQ] channels-last fp32 activations
//
// (dW of a 1x1 / im2col-form convolution: P = gradient at the layer output, Q = the layer input.)
// Same machinery as tc_igemm.cuh (3xTF32 split, two-level accumulation, A from TMEM, warp-uniform
// single-lane issue) with the operands arranged for a reduction over the slow memory axis:
//   * A = P^T (M = 128 channels x K = 32 pixels per k-block): the transform warps read the pixel-major
//     tile from shared memory column-wise (thread = channel; a warp reads 128 contiguous bytes per pixel)
//     and store hi / lo into TMEM, lane = channel, column = pixel -- the transpose costs nothing.
//   * B = Q^T (N x K) must be a K-major shared-memory operand (rows = channels, 32 pixels = 128 bytes per
//     row, 128-byte swizzle).  Both operands are activations, so the transform warps split the Q tile
//     anyway: each thread reads one channel COLUMN of the raw pixel-major tile (a warp reads 128 contiguous
//     bytes per pixel), and after a named barrier writes that channel's hi / lo ROWS in the swizzled layout
//     over the raw tile -- transpose and split in one pass.  (An MN-major descriptor over the raw tile
//     returned zeros: kind::tf32 wants the 32-byte-atom swizzle for MN-major operands, measured with
//     tools/wgrad_debug.py.)
//   * split-K over the pixels: CTA (tile, slice) accumulates its pixel range and writes one partial tile;
//     a small kernel adds the slices in a fixed order (deterministic, no floating-point atomics).
#pragma once

#include "tc_igemm.cuh"

namespace finc {
namespace tc {

struct WgradGeom {
    int np;            // pixels (reduction length)
    int m_tiles;       // ceil(M / 128)
    int n_tiles;       // N / BN
    int slices;        // split-K factor
    int kb_per_slice;  // k-blocks (32 pixels) per slice
    int ld_out;        // row stride of one partial tile matrix [m_tiles * 128, ld_out]
    size_t slice_stride;  // floats between slices of the workspace
};

template <int BN, int NPASS, int EW>
struct WCfg {
    static constexpr int kThreads = 192 + 32 * EW;
    static constexpr int kEpiThreads = 32 * EW;
    static constexpr int kBBytes = BN * kBK * 4;
    static constexpr int kParts = NPASS == 3 ? 2 : 1;
    static constexpr int kStageBytes = kABytes + kParts * kBBytes;   // P tile + Q tile (raw -> hi) [+ lo]
    static constexpr int kAvail = kSmemLimit - 2048 - 1024;
    static constexpr int kStagesRaw = kAvail / kStageBytes;
    static constexpr int kACols = kParts * kBK;
    static constexpr int kStagesTmem = (512 - 2 * BN) / kACols;
    static constexpr int kStagesCap = kStagesRaw < kStagesTmem ? kStagesRaw : kStagesTmem;
    static constexpr int kStages = kStagesCap > 4 ? 4 : kStagesCap;
    static constexpr int kSmemBytes = kStages * kStageBytes + 2048;
    static constexpr int kFlush = NPASS == 3 ? 2 : 4;
    static constexpr int kNPartRaw = (512 - kStages * kACols) / BN;
    static constexpr int kNPart = kNPartRaw > 4 ? 4 : kNPartRaw;
    static constexpr int kAColBase = kNPart * BN;
    static constexpr int kUsedCols = kAColBase + kStages * kACols;
    static constexpr int kTmemCols = kUsedCols <= 32 ? 32 : kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128 : kUsedCols <= 256 ? 256 : 512;
    static constexpr int kCols0 = EW == 8 ? (BN / 2 + 15) / 16 * 16 : BN;
    static constexpr int kBatch = kCols0 <= 64 ? kCols0 : kCols0 % 32 == 0 ? 32 : 16;
    static_assert(kStages >= 2 && kNPart >= 2 && kUsedCols <= 512, "TMEM / pipeline budget");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256 && kCols0 <= 128, "N in 32-channel boxes");
};

template <int BN, int NPASS, int EW>
__global__ void __launch_bounds__(WCfg<BN, NPASS, EW>::kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ, float* __restrict__ out,
             const WgradGeom g) {
    using C = WCfg<BN, NPASS, EW>;
    constexpr int S = C::kStages;
    constexpr int NP = C::kNPart;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * C::kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* xf = bars + 2 * S;
    uint64_t* part_full = bars + 3 * S;
    uint64_t* part_empty = part_full + NP;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_empty + NP);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int n_work = g.m_tiles * g.n_tiles * g.slices;

    auto p_raw = [&](int s) { return smem + s * C::kStageBytes; };                       // [32 pixels][128 ch], no swizzle
    auto q_hi = [&](int s) { return smem + s * C::kStageBytes + kABytes; };              // raw [32 px][BN ch] -> hi [BN][32 px] swizzled
    auto q_lo = [&](int s) { return smem + s * C::kStageBytes + kABytes + C::kBBytes; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapP);
        tma_prefetch_desc(&mapQ);
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&xf[s], kXfThreads);
        }
        for (int i = 0; i < NP; ++i) {
            mbar_init(&part_full[i], 1);
            mbar_init(&part_empty[i], C::kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // work item -> (m tile, n tile, pixel slice); the slice index is the slow one so that concurrently
    // running CTAs share the activation tiles through L2
    auto decode = [&](int w, int& mt, int& nt, int& sl) {
        nt = w % g.n_tiles;
        mt = (w / g.n_tiles) % g.m_tiles;
        sl = w / (g.n_tiles * g.m_tiles);
    };

    if (warp == 0) {
        uint32_t it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int mt, nt, sl;
            decode(w, mt, nt, sl);
            for (int kb = 0; kb < g.kb_per_slice; ++kb, ++it) {
                const int s = it % S;
                const int p0 = (sl * g.kb_per_slice + kb) * kBK;
                mbar_wait_long(&empty[s], ((it / S) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], kABytes + C::kBBytes);
                    tma_load_2d(p_raw(s), &mapP, &full[s], mt * kBM, p0);
                    tma_load_2d(q_hi(s), &mapQ, &full[s], nt * BN, p0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
        // bookkeeping once per group of kFlush k-blocks (see tc_igemm.cuh: every satisfied mbarrier wait in the
        // issuing thread costs ~100 cycles that the shallow tensor-core queue cannot hide)
        uint32_t it = 0, gq = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            for (int kb0 = 0; kb0 < g.kb_per_slice; kb0 += C::kFlush, ++gq) {
                const int nk = g.kb_per_slice - kb0 < C::kFlush ? g.kb_per_slice - kb0 : C::kFlush;
                const uint32_t p = gq % NP;
                const uint32_t d_tmem = tmem_base + p * BN;
                uint32_t ready = 0;
                for (uint32_t spins = 0;; ++spins) {
                    uint32_t m = mbar_try_wait(&part_empty[p], ((gq / NP) & 1) ^ 1) ? 1u : 0u;
                    m |= mbar_try_wait(&xf[it % S], (it / S) & 1) ? 2u : 0u;   // both operands split (implies the TMA bytes landed)
#pragma unroll
                    for (int j = 1; j < C::kFlush; ++j)
                        if (j < nk) m |= mbar_test_wait(&xf[(it + j) % S], ((it + j) / S) & 1) ? (2u << j) : 0u;
                    ready = __reduce_and_sync(0xffffffffu, m);
                    if ((ready & 3u) == 3u) break;
                    if (spins > (1u << 26)) __trap();
                }
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < C::kFlush; ++j) {
                    if (j >= nk) break;
                    const int s = (it + j) % S;
                    if (j > 0 && !(ready & (2u << j))) {
                        mbar_wait_long(&xf[s], ((it + j) / S) & 1);
                        tc_fence_after();
                    }
                    const uint64_t db_hi = umma_desc_k_sw128(q_hi(s)), db_lo = umma_desc_k_sw128(q_lo(s));
                    const uint32_t ta_hi = tmem_base + C::kAColBase + s * C::kACols, ta_lo = ta_hi + kBK;
                    if (elect_one()) {
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
#pragma unroll
                            for (int k = 0; k < kBK / kUmmaK; ++k) {
                                const uint64_t adv = (uint64_t)(k * kUmmaK * 4 >> 4);   // 8 pixels = 32 bytes inside the swizzle row
                                const uint32_t acc = (j != 0 || pass != 0 || k != 0) ? 1u : 0u;
                                const uint32_t ta = ((NPASS == 3 && pass == 0) ? ta_lo : ta_hi) + k * kUmmaK;
                                const uint64_t db = (NPASS == 3 && pass == 1) ? db_lo : db_hi;
                                umma_tf32_ts(d_tmem, ta, db + adv, idesc, acc);
                            }
                        }
                        umma_commit(&empty[s]);
                        if (j == nk - 1) umma_commit(&part_full[p]);
                    }
                    __syncwarp();
                }
                it += nk;
            }
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        const int ch = q * 32 + lane;        // channel of P handled by this thread = TMEM lane
        const int t = threadIdx.x - 64;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + C::kAColBase;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            for (int kb = 0; kb < g.kb_per_slice; ++kb, ++it) {
                const int s = it % S;
                mbar_wait_long(&full[s], (it / S) & 1);
                // (a) P tile, transposed on the fly: column `ch` of the [32 pixels][128 channels] tile
                const float* src = reinterpret_cast<const float*>(p_raw(s)) + ch;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float v = src[k * kBM];
                    const float h = tf32_rn(v);
                    hi[k] = __float_as_uint(h);
                    if (NPASS == 3) lo[k] = __float_as_uint(v - h);
                }
                tmem_st_x32(lane_addr + s * C::kACols, hi);
                if (NPASS == 3) tmem_st_x32(lane_addr + s * C::kACols + kBK, lo);
                // (b) Q tile: column n of the raw [32 pixels][BN channels] tile -> rows n of the K-major hi / lo tiles
                constexpr int kColsPerThread = (BN + kXfThreads - 1) / kXfThreads;
                float qv[kColsPerThread][32];
                const float* qraw = reinterpret_cast<const float*>(q_hi(s));
#pragma unroll
                for (int c = 0; c < kColsPerThread; ++c) {
                    const int n = c * kXfThreads + t;
                    if (n < BN) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) qv[c][k] = qraw[k * BN + n];
                    }
                }
                named_bar_sync(2, kXfThreads);   // every column has been read: the raw tile may be overwritten
#pragma unroll
                for (int c = 0; c < kColsPerThread; ++c) {
                    const int n = c * kXfThreads + t;
                    if (n < BN) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 h, l;
                            h.x = tf32_rn(qv[c][4 * j + 0]); h.y = tf32_rn(qv[c][4 * j + 1]);
                            h.z = tf32_rn(qv[c][4 * j + 2]); h.w = tf32_rn(qv[c][4 * j + 3]);
                            l.x = qv[c][4 * j + 0] - h.x; l.y = qv[c][4 * j + 1] - h.y;
                            l.z = qv[c][4 * j + 2] - h.z; l.w = qv[c][4 * j + 3] - h.w;
                            const int off = n * 128 + ((j ^ (n & 7)) << 4);   // 128-byte swizzle: chunk ^ (row % 8)
                            *reinterpret_cast<float4*>(q_hi(s) + off) = h;
                            if (NPASS == 3) *reinterpret_cast<float4*>(q_lo(s) + off) = l;
                        }
                    }
                }
                fence_proxy_async_smem();
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&xf[s]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;       // output row = channel of P
        const int hsel = (warp - 6) >> 2;
        const int col0 = hsel ? C::kCols0 : 0;
        const int ncols = EW == 8 ? (hsel ? BN - C::kCols0 : C::kCols0) : BN;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + col0;
        const int groups = (g.kb_per_slice + C::kFlush - 1) / C::kFlush;
        uint32_t gq = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int mt, nt, sl;
            decode(w, mt, nt, sl);
            float acc[C::kCols0];
#pragma unroll
            for (int i = 0; i < C::kCols0; ++i) acc[i] = 0.f;
            for (int gi = 0; gi < groups; ++gi, ++gq) {
                const uint32_t p = gq % NP;
                mbar_wait_long(&part_full[p], (gq / NP) & 1);
                tc_fence_after();
                const uint32_t taddr = lane_addr + p * BN;
#pragma unroll
                for (int b0 = 0; b0 < C::kCols0; b0 += C::kBatch) {
                    uint32_t v[C::kBatch / 16][16];
#pragma unroll
                    for (int c = 0; c < C::kBatch / 16; ++c)
                        if (b0 + c * 16 < ncols) tmem_ld_x16(taddr + b0 + c * 16, v[c]);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < C::kBatch / 16; ++c)
                        if (b0 + c * 16 < ncols) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[b0 + c * 16 + i] += __uint_as_float(v[c][i]);
                        }
                }
                tc_fence_before();
                mbar_arrive(&part_empty[p]);
            }
            float4* dst = reinterpret_cast<float4*>(out + sl * g.slice_stride + (size_t)(mt * kBM + row) * g.ld_out + nt * BN + col0);
#pragma unroll
            for (int i = 0; i < C::kCols0 / 4; ++i)
                if (i * 4 < ncols) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace tc
}  // namespace finc
