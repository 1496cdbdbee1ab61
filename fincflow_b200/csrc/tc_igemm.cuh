// tc_igemm.cuh -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[pixel, n] = sum_{tap, c} X[pixel + tap, c] * Wt[tap][n][c]        (1x1 or 3x3 / pad 1)
//
// for channels-last fp32 activations X [B, H, W, Cin] and weights pre-arranged as K-major rows
// [part][tap][n][c].  The contraction runs on tcgen05.mma kind::tf32 with accumulators in TMEM.
// fp32 parity with an fp32 reference ("not TF32") comes from the 3xTF32
// split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = tf32(a), a_lo = tf32(a - a_hi)):
// weights are split once on the device (`part` 0 = hi, 1 = lo); activation tiles are split in
// shared memory by four transform warps between the TMA load and the MMA.  NPASS = 1 is the plain
// single-pass TF32 product (PyTorch's default conv precision) for callers that ask for it.
//
// Warp roles of one persistent CTA (320 threads, one CTA per SM):
//   warp 0      TMA producer: 4-D box loads of the activation tile (out-of-image taps are zero
//               filled by the TMA unit = the conv padding) + 2-D loads of the weight tile
//   warp 1      TMEM allocation, single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 2-5   hi/lo split of the activation tile in shared memory (NPASS = 3 only)
//   warps 6-9   epilogue: tcgen05.ld -> bias / ReLU / coupling math -> TMA store or NCHW stores
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once

#include "tc_common.cuh"

namespace finc {
namespace tc {

constexpr int kBM = 128;           // pixels per tile = TMEM lanes
constexpr int kBK = 32;            // fp32 per k-block: one 128-byte swizzle row
constexpr int kUmmaK = 8;          // K of one tcgen05.mma kind::tf32
constexpr int kABytes = kBM * kBK * 4;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 128;
constexpr int kXfThreads = 128;
constexpr int kStagingBytes = kBM * 128;  // one 32-column output chunk of a tile
constexpr int kSmemLimit = 232448;        // 227 KB opt-in maximum per CTA

enum Epi { EPI_NHWC = 0, EPI_COUPLING = 1 };

struct Geom {
    int W, H, B;                      // image
    int wb, hb, nb;                   // pixel box of one tile, wb * hb * nb == 128
    int tiles_w, tiles_h, tiles_n;
    int taps;                         // 1 (1x1) or 9 (3x3, pad 1)
    int kb_per_tap;                   // Cin_pad / 32
    int n_tiles;                      // Npad / BN
    int n_rows;                       // Npad: weight rows per (part, tap)
};

struct EpiArgs {
    const float* bias;    // [Npad] (zero padded)
    const float* scale;   // [Npad] multiplier applied after the bias (coupling: exp(3 logs)), or null
    int relu;
    // EPI_NHWC with `mask`: out *= (mask[pixel, n] > 0)   (ReLU backward); channels-last, ld = Npad
    const float* mask;
    // EPI_COUPLING (layers/coupling.py:73-99): x, y NCHW [B, C, H, W]; rowsum [B*H*W] per-pixel sum of log_s
    const float* x;
    float* y;
    float* rowsum;
    int C;
    int reverse;
};

template <int BN, int NPASS>
struct Cfg {
    static constexpr int kBBytes = BN * kBK * 4;
    static constexpr int kParts = NPASS == 3 ? 2 : 1;
    static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
    static constexpr int kAvail = kSmemLimit - 2 * kStagingBytes - 2048;
    static constexpr int kStagesRaw = kAvail / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 2048;
    static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    static_assert(kStages >= 2, "pipeline needs two stages");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128");
};

struct TileCoord {
    int w0, h0, n0, ncol0;
};
__device__ __forceinline__ TileCoord tile_coord(const Geom& g, int tile, int BN) {
    const int nt = tile % g.n_tiles;
    int m = tile / g.n_tiles;
    TileCoord t;
    t.ncol0 = nt * BN;
    t.w0 = (m % g.tiles_w) * g.wb;
    m /= g.tiles_w;
    t.h0 = (m % g.tiles_h) * g.hb;
    t.n0 = (m / g.tiles_h) * g.nb;
    return t;
}

template <int BN, int NPASS, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapOut, const Geom g, const EpiArgs e) {
    using C = Cfg<BN, NPASS>;
    constexpr int S = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + S * C::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
    uint64_t* full = bars;             // [S]  TMA bytes landed
    uint64_t* empty = bars + S;        // [S]  MMAs that read the stage have completed
    uint64_t* xf = bars + 2 * S;       // [S]  activation tile split into hi / lo
    uint64_t* acc_full = bars + 3 * S; // [2]  accumulator stage complete
    uint64_t* acc_empty = acc_full + 2;  // [2]  accumulator stage drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_total = g.tiles_w * g.tiles_h * g.tiles_n * g.n_tiles;
    const int kblocks = g.taps * g.kb_per_tap;

    auto a_hi = [&](int s) { return smem + s * C::kStageBytes; };
    auto a_lo = [&](int s) { return smem + s * C::kStageBytes + kABytes; };
    auto b_hi = [&](int s) { return smem + s * C::kStageBytes + C::kParts * kABytes; };
    auto b_lo = [&](int s) { return smem + s * C::kStageBytes + C::kParts * kABytes + C::kBBytes; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (EPI == EPI_NHWC) tma_prefetch_desc(&mapOut);
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&xf[s], kXfThreads);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything above overlaps the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
                const TileCoord t = tile_coord(g, tile, BN);
                for (int tap = 0; tap < g.taps; ++tap) {
                    const int dy = g.taps == 9 ? tap / 3 - 1 : 0, dx = g.taps == 9 ? tap % 3 - 1 : 0;
                    for (int kb = 0; kb < g.kb_per_tap; ++kb, ++it) {
                        const int s = it % S;
                        mbar_wait_long(&empty[s], ((it / S) & 1) ^ 1);
                        mbar_arrive_expect_tx(&full[s], kABytes + C::kParts * C::kBBytes);
                        tma_load_4d(a_hi(s), &mapA, &full[s], kb * kBK, t.w0 + dx, t.h0 + dy, t.n0);
                        tma_load_2d(b_hi(s), &mapB, &full[s], kb * kBK, tap * g.n_rows + t.ncol0);
                        if (NPASS == 3)
                            tma_load_2d(b_lo(s), &mapB, &full[s], kb * kBK, (g.taps + tap) * g.n_rows + t.ncol0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            uint32_t it = 0, j = 0;
            for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++j) {
                const uint32_t as = j & 1;
                mbar_wait_long(&acc_empty[as], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kblk = 0; kblk < kblocks; ++kblk, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    mbar_wait_long(&full[s], ph);
                    if (NPASS == 3) mbar_wait_long(&xf[s], ph);
                    tc_fence_after();
                    const uint64_t da_hi = umma_desc_k_sw128(a_hi(s)), db_hi = umma_desc_k_sw128(b_hi(s));
                    const uint64_t da_lo = umma_desc_k_sw128(a_lo(s)), db_lo = umma_desc_k_sw128(b_lo(s));
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint64_t adv = (uint64_t)(k * kUmmaK * 4 >> 4);  // 32 bytes per K step inside the swizzle row
                        if (NPASS == 3) {
                            // small terms first, the dominant product last
                            umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, (kblk | k) != 0);
                            umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
                            umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, 1);
                        } else {
                            umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, (kblk | k) != 0);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else if (warp < 6) {
        // ===================== hi / lo split of the activation tile =====================
        if (NPASS == 3) {
            const int t = threadIdx.x - 64;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
                for (int kblk = 0; kblk < kblocks; ++kblk, ++it) {
                    const int s = it % S;
                    mbar_wait_long(&full[s], (it / S) & 1);
                    float4* hi = reinterpret_cast<float4*>(a_hi(s));
                    float4* lo = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
                    for (int i = 0; i < kABytes / 16 / kXfThreads; ++i) {
                        const int idx = i * kXfThreads + t;   // elementwise: the swizzle does not matter
                        const float4 v = hi[idx];
                        float4 h, l;
                        h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                        hi[idx] = h;
                        lo[idx] = l;
                    }
                    fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
                    mbar_arrive(&xf[s]);
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;       // accumulator row = pixel inside the tile
        const int te = threadIdx.x - 192;
        uint32_t j = 0, cc = 0;
        for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++j) {
            const TileCoord t = tile_coord(g, tile, BN);
            const uint32_t as = j & 1;
            mbar_wait_long(&acc_full[as], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
            if (EPI == EPI_NHWC) {
                // pixel of this row (the ReLU-backward mask needs it; the TMA store clips by itself)
                const int pw = t.w0 + row % g.wb, phh = t.h0 + (row / g.wb) % g.hb, pn = t.n0 + row / (g.wb * g.hb);
                const bool valid = pw < g.W && phh < g.H && pn < g.B;
                const size_t pix = ((size_t)pn * g.H + phh) * g.W + pw;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c, ++cc) {
                    uint32_t v[2][16];
                    tmem_ld_x16(taddr + c * 32, v[0]);
                    tmem_ld_x16(taddr + c * 32 + 16, v[1]);
                    tmem_wait_ld();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        mbar_arrive(&acc_empty[as]);
                    }
                    uint8_t* buf = staging + (cc & 1) * kStagingBytes;
                    if (te == 0) bulk_wait_read<1>();   // the store that last read `buf` has finished reading
                    named_bar_sync(1, kEpiThreads);
                    const int n0 = t.ncol0 + c * 32;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + jj);
                        float4 o;
                        o.x = __uint_as_float(v[jj >> 2][(jj & 3) * 4 + 0]) + bv.x;
                        o.y = __uint_as_float(v[jj >> 2][(jj & 3) * 4 + 1]) + bv.y;
                        o.z = __uint_as_float(v[jj >> 2][(jj & 3) * 4 + 2]) + bv.z;
                        o.w = __uint_as_float(v[jj >> 2][(jj & 3) * 4 + 3]) + bv.w;
                        if (e.relu) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        if (e.mask != nullptr) {
                            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid) m = __ldg(reinterpret_cast<const float4*>(e.mask + pix * g.n_rows + n0) + jj);
                            o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f;
                            o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
                        }
                        // 128-byte swizzle of the staging tile: 16-byte chunk index ^ (row % 8)
                        *reinterpret_cast<float4*>(buf + row * 128 + ((jj ^ (row & 7)) << 4)) = o;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, kEpiThreads);
                    if (te == 0) {
                        tma_store_4d(&mapOut, buf, n0, t.w0, t.h0, t.n0);
                        bulk_commit();
                    }
                }
            } else {
                // ---- affine-coupling epilogue (layers/coupling.py:73-99) ----
                float acc[BN];
#pragma unroll
                for (int c = 0; c < BN / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld_x16(taddr + c * 16, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[c * 16 + i] = __uint_as_float(v[i]);
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[as]);
                const int pw = t.w0 + row % g.wb, phh = t.h0 + (row / g.wb) % g.hb, pn = t.n0 + row / (g.wb * g.hb);
                if (pw < g.W && phh < g.H && pn < g.B) {
                    const int half = e.C / 2;
                    const size_t plane = (size_t)g.H * g.W;
                    const size_t base = (size_t)pn * e.C * plane + (size_t)phh * g.W + pw;
                    float rs = 0.f;
#pragma unroll
                    for (int jc = 0; jc < BN / 2; ++jc) {
                        if (jc < half) {
                            const float hs = (acc[2 * jc] + __ldg(e.bias + 2 * jc)) * __ldg(e.scale + 2 * jc);
                            const float tt = (acc[2 * jc + 1] + __ldg(e.bias + 2 * jc + 1)) * __ldg(e.scale + 2 * jc + 1);
                            const float log_s = 2.0f * tanhf(hs * 0.5f);
                            const size_t o2 = base + (size_t)(half + jc) * plane;
                            const float x2 = e.x[o2];
                            e.y[o2] = e.reverse ? (x2 - tt) * expf(-log_s) : x2 * expf(log_s) + tt;
                            rs += log_s;
                            if (e.y != e.x) {
                                const size_t o1 = base + (size_t)jc * plane;
                                e.y[o1] = e.x[o1];
                            }
                        }
                    }
                    if (e.rowsum != nullptr) e.rowsum[((size_t)pn * g.H + phh) * g.W + pw] = rs;
                }
            }
        }
        if (EPI == EPI_NHWC && te == 0) bulk_wait_all();
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace tc
}  // namespace finc
