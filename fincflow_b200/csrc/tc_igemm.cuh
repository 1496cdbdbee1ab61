// tc_igemm.cuh -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[pixel, n] = sum_{tap, c} X[pixel + tap, c] * Wt[tap][n][c]        (1x1 or 3x3 / pad 1)
//
// for channels-last fp32 activations X [B, H, W, Cin] and weights pre-arranged as K-major rows
// [part][tap][n][c].  The products run on tcgen05.mma kind::tf32 with accumulators in TMEM.
//
// fp32 parity ("not TF32") and speed rest on four design points, each
// measured on B200 (profiles/r2_tc_*.md):
//  (1) 3xTF32 split products  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (a_hi = tf32(a),
//      a_lo = a - a_hi): weights are split once on the device (`part` 0 = hi, 1 = lo); activation
//      tiles are split on the fly by four transform warps.  NPASS = 1 is the plain single-pass TF32
//      product (PyTorch's default conv precision) for callers that ask for it.
//  (2) two-level accumulation.  The tensor core adds into its fp32 accumulator with truncation, so
//      the error of a long chain grows linearly with K (3.4e-6 at K = 512, 2.6e-5 at K = 4608 where
//      cuDNN's fp32 is at 7e-7).  The MMAs of one GROUP (kFlush k-blocks of 32) accumulate into a
//      fresh TMEM partial and the epilogue warps add the partials into fp32 registers with
//      round-to-nearest: 1e-7 .. 2e-7 at every K.
//  (3) the A operand (activations) comes from TENSOR MEMORY.  With both operands in shared memory
//      the shared-memory data pipe was saturated (tensor-core operand fetch 47 % + the split warps'
//      loads / stores 40 % + TMA writes) at 26 % tensor-pipe utilisation.  The split warps read
//      every activation element anyway, so they write hi / lo straight into TMEM (tcgen05.st) and
//      the tensor core fetches only the weight tile from shared memory.
//  (4) thread-block clusters: the CL CTAs of a cluster work on CL different pixel tiles against the
//      SAME weight tile; each loads 1/CL of it and TMA-multicasts it to all (a 128x128 tile pulls
//      48 KB per k-block from L2 -- the 1x1 512 -> 512 layer ran at the L2 bandwidth limit).
//
// Warp roles of one persistent CTA (192 + 32 EW threads, one CTA per SM):
//   warp 0      TMA producer: 4-D box loads of the activation tile (out-of-image taps are zero
//               filled by the TMA unit = the conv padding) + 2-D (multicast) loads of the weight tile
//   warp 1      TMEM allocation, single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 2-5   activation tile shared memory -> registers -> hi / lo split -> TMEM (tcgen05.st)
//   warps 6..   EW = 4 or 8 epilogue warps: accumulate partials (tcgen05.ld + FADD), then bias /
//               ReLU / mask and TMA store (channels-last) or plain row stores; with EW = 8 two warps
//               share a TMEM lane quarter and split the columns (wide N with <= 80 sums per thread)
#pragma once

#include "tc_common.cuh"

namespace finc {
namespace tc {

constexpr int kBM = 128;           // pixels per tile = TMEM lanes
constexpr int kBK = 32;            // fp32 per k-block: one 128-byte swizzle row
constexpr int kUmmaK = 8;          // K of one tcgen05.mma kind::tf32
constexpr int kABytes = kBM * kBK * 4;
constexpr int kXfThreads = 128;
constexpr int kStagingBytes = kBM * 128;  // one 32-column output chunk of a tile
constexpr int kSmemLimit = 232448;        // 227 KB opt-in maximum per CTA

enum Epi { EPI_NHWC = 0, EPI_ROWS = 2 };

struct Geom {
    int W, H, B;                      // image
    int wb, hb, nb;                   // pixel box of one tile, wb * hb * nb == 128
    int tiles_w, tiles_h, tiles_n;
    int taps;                         // 1 (1x1) or 9 (3x3, pad 1)
    int kb_per_tap;                   // Cin_pad / 32
    int n_tiles;                      // Npad / BN
    int n_rows;                       // Npad: weight rows per (part, tap)
    int group_n;                      // block-diagonal GEMM (0 = off): output columns [q*group_n, (q+1)*group_n) read
                                      // activation channels q*group_n + k  (one launch for all groups of the dense inverse)
};

struct EpiArgs {
    const float* bias;    // EPI_NHWC: [Npad] (zero padded), or null
    int relu;
    // EPI_NHWC with `mask`: out *= (mask[pixel, n] > 0)   (ReLU backward); channels-last, ld = Npad
    const float* mask;
    // output: y[pixel * ld_out + n]; EPI_NHWC applies bias / ReLU / mask, EPI_ROWS stores the raw sums
    float* y;
    int ld_out;
};

template <int BN, int NPASS, int EW>
struct Cfg {
    static constexpr int kThreads = 192 + 32 * EW;
    static constexpr int kEpiThreads = 32 * EW;
    static constexpr int kBBytes = BN * kBK * 4;
    static constexpr int kParts = NPASS == 3 ? 2 : 1;
    static constexpr int kStageBytes = kABytes + kParts * kBBytes;   // raw activation tile + weight tile(s)
    static constexpr int kAvail = kSmemLimit - 2 * kStagingBytes - 2048 - 1024;
    static constexpr int kStagesRaw = kAvail / kStageBytes;
    static constexpr int kACols = kParts * kBK;                        // TMEM columns of one activation stage
    static constexpr int kStagesTmem = (512 - 2 * BN) / kACols;        // leave room for two partials
    static constexpr int kStagesCap = kStagesRaw < kStagesTmem ? kStagesRaw : kStagesTmem;
    static constexpr int kStages = kStagesCap > 4 ? 4 : kStagesCap;
    static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStagingBytes + 2048;
    // two-level accumulation
    static constexpr int kFlush = NPASS == 3 ? 2 : 4;                  // k-blocks per TMEM partial (2: 24 MMAs, 1.5k cycles -- hides the epilogue round trip behind a ring of two)
    static constexpr int kNPartRaw = (512 - kStages * kACols) / BN;
    static constexpr int kNPart = kNPartRaw > 4 ? 4 : kNPartRaw;       // ring depth
    static constexpr int kAColBase = kNPart * BN;                      // activation stages sit behind the partials
    static constexpr int kUsedCols = kAColBase + kStages * kACols;
    static constexpr int kTmemCols = kUsedCols <= 32 ? 32 : kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128 : kUsedCols <= 256 ? 256 : 512;
    // columns per epilogue thread: EW = 8 splits the tile's columns between two warps of a lane quarter
    static constexpr int kCols0 = EW == 8 ? (BN / 2 + 15) / 16 * 16 : BN;
    static constexpr int kBatch = kCols0 <= 64 ? kCols0 : kCols0 % 32 == 0 ? 32 : 16;   // columns per tcgen05.wait::ld
    static_assert(EW == 4 || EW == 8, "epilogue warps");
    static_assert(kStages >= 2, "pipeline needs two stages");
    static_assert(kNPart >= 2, "partial ring needs two buffers");
    static_assert(kUsedCols <= 512, "TMEM columns");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256 && kCols0 <= 128, "UMMA N for M = 128; <= 128 sums per thread");
};

struct TileCoord {
    int w0, h0, n0, ncol0;
};
// m beyond the last pixel tile (cluster padding) lands outside the tensor: loads read zeros, stores are clipped
__device__ __forceinline__ TileCoord tile_coord(const Geom& g, int m, int nt, int BN) {
    TileCoord t;
    t.ncol0 = nt * BN;
    t.w0 = (m % g.tiles_w) * g.wb;
    m /= g.tiles_w;
    t.h0 = (m % g.tiles_h) * g.hb;
    t.n0 = (m / g.tiles_h) * g.nb;
    return t;
}

template <int BN, int NPASS, int EPI, int CL, int EW>
__global__ void __launch_bounds__(Cfg<BN, NPASS, EW>::kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapOut, const Geom g, const EpiArgs e) {
    using C = Cfg<BN, NPASS, EW>;
    constexpr int S = C::kStages;
    constexpr int NP = C::kNPart;
    constexpr int BNS = BN / CL;          // weight rows this CTA loads (and multicasts)
    static_assert(BN % CL == 0 && BNS % 8 == 0, "weight slice must keep the 1024-byte swizzle phase");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + S * C::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
    uint64_t* full = bars;               // [S]  TMA bytes landed
    uint64_t* empty = bars + S;          // [S]  MMAs (of every CTA of the cluster) that read the stage have completed
    uint64_t* xf = bars + 2 * S;         // [S]  activation tile split into TMEM
    uint64_t* part_full = bars + 3 * S;  // [NP] partial accumulator complete
    uint64_t* part_empty = part_full + NP;  // [NP] partial added into the running sums
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_empty + NP);

    // warp index through a shuffle: the compiler then KNOWS the role branches are warp-uniform and keeps
    // loop counters / descriptors of the single-thread roles in uniform registers (measured: with
    // `threadIdx.x >> 5` every tcgen05.mma was preceded by ELECT + R2UR.BROADCAST sequences, ~100 issue
    // cycles per MMA, and the MMA warp -- not the tensor core -- was the bottleneck)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int m_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    const int cluster_tiles = (m_tiles + CL - 1) / CL * g.n_tiles;
    const int kblocks = g.taps * g.kb_per_tap;
    constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1);

    auto a_raw = [&](int s) { return smem + s * C::kStageBytes; };
    auto b_hi = [&](int s) { return smem + s * C::kStageBytes + kABytes; };
    auto b_lo = [&](int s) { return smem + s * C::kStageBytes + kABytes + C::kBBytes; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (EPI == EPI_NHWC) tma_prefetch_desc(&mapOut);
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);
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
    if (CL > 1) cluster_sync(); else __syncthreads();   // peers' barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything above overlaps the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
        {
            uint32_t it = 0;
            for (int ct = cluster_id; ct < cluster_tiles; ct += n_clusters) {
                const TileCoord t = tile_coord(g, (ct / g.n_tiles) * CL + (int)rank, ct % g.n_tiles, BN);
                for (int tap = 0; tap < g.taps; ++tap) {
                    const int dy = g.taps == 9 ? tap / 3 - 1 : 0, dx = g.taps == 9 ? tap % 3 - 1 : 0;
                    for (int kb = 0; kb < g.kb_per_tap; ++kb, ++it) {
                        const int s = it % S;
                        mbar_wait_long(&empty[s], ((it / S) & 1) ^ 1);   // stage free in EVERY CTA of the cluster
                        const int wrow = tap * g.n_rows + t.ncol0 + (int)rank * BNS;
                        const int c0 = kb * kBK + (g.group_n ? t.ncol0 / g.group_n * g.group_n : 0);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&full[s], kABytes + C::kParts * C::kBBytes);
                            tma_load_4d(a_raw(s), &mapA, &full[s], c0, t.w0 + dx, t.h0 + dy, t.n0);
                            if (CL == 1) {
                                tma_load_2d(b_hi(s), &mapB, &full[s], kb * kBK, wrow);
                                if (NPASS == 3) tma_load_2d(b_lo(s), &mapB, &full[s], kb * kBK, g.taps * g.n_rows + wrow);
                            } else {
                                tma_load_2d_mc(b_hi(s) + rank * BNS * 128, &mapB, &full[s], kb * kBK, wrow, kAllCtas);
                                if (NPASS == 3)
                                    tma_load_2d_mc(b_lo(s) + rank * BNS * 128, &mapB, &full[s], kb * kBK,
                                                   g.taps * g.n_rows + wrow, kAllCtas);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        {
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            // The tensor-core queue is shallow: tools/mma_issue_bench.cu shows that a pause of the issuing thread
            // is hidden only up to ~100 cycles, and that ONE satisfied mbarrier wait costs ~100 cycles of latency
            // (three waits + fence per k-block, as this loop was first written: 82 instead of 64 cycles per MMA).
            // So the bookkeeping is done once per GROUP (the kFlush k-blocks of one TMEM partial): one polling
            // loop probes the partial's `empty` barrier and the split barriers of all its k-blocks together (the
            // probes overlap), and only a k-block that was not ready at that moment is waited for separately.
            // xf[s] is arrived on by threads that have seen full[s], so it covers the weight tile as well.
            uint32_t it = 0, gq = 0;   // k-block counter, group (= partial) counter
            for (int ct = cluster_id; ct < cluster_tiles; ct += n_clusters) {
                for (int kb0 = 0; kb0 < kblocks; kb0 += C::kFlush, ++gq) {
                    const int nk = kblocks - kb0 < C::kFlush ? kblocks - kb0 : C::kFlush;
                    const uint32_t p = gq % NP;
                    const uint32_t d_tmem = tmem_base + p * BN;
                    uint32_t ready = 0;
                    for (uint32_t spins = 0;; ++spins) {
                        uint32_t m = mbar_try_wait(&part_empty[p], ((gq / NP) & 1) ^ 1) ? 1u : 0u;
                        m |= mbar_try_wait(&xf[it % S], (it / S) & 1) ? 2u : 0u;
#pragma unroll
                        for (int j = 1; j < C::kFlush; ++j)
                            if (j < nk) m |= mbar_test_wait(&xf[(it + j) % S], ((it + j) / S) & 1) ? (2u << j) : 0u;
                        ready = __reduce_and_sync(0xffffffffu, m);   // warp-uniform by construction
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
                        const uint64_t db_hi = umma_desc_k_sw128(b_hi(s)), db_lo = umma_desc_k_sw128(b_lo(s));
                        const uint32_t ta_hi = tmem_base + C::kAColBase + s * C::kACols, ta_lo = ta_hi + kBK;
                        if (elect_one()) {
                            // small terms first (a_lo*b_hi, a_hi*b_lo), the dominant product last
#pragma unroll
                            for (int pass = 0; pass < NPASS; ++pass) {
#pragma unroll
                                for (int k = 0; k < kBK / kUmmaK; ++k) {
                                    const uint64_t adv = (uint64_t)(k * kUmmaK * 4 >> 4);  // 32 bytes per K step inside the swizzle row
                                    const uint32_t acc = (j != 0 || pass != 0 || k != 0) ? 1u : 0u;
                                    const uint32_t ta = ((NPASS == 3 && pass == 0) ? ta_lo : ta_hi) + k * kUmmaK;
                                    const uint64_t db = (NPASS == 3 && pass == 1) ? db_lo : db_hi;
                                    umma_tf32_ts(d_tmem, ta, db + adv, idesc, acc);
                                }
                            }
                            if (CL == 1) umma_commit(&empty[s]); else umma_commit_mc(&empty[s], kAllCtas);
                            if (j == nk - 1) umma_commit(&part_full[p]);
                        }
                        __syncwarp();
                    }
                    it += nk;
                }
            }
        }
    } else if (warp < 6) {
        // ===================== activation tile: shared memory -> hi / lo -> TMEM =====================
        // thread = one tile row (pixel): its 128 bytes sit in 8 swizzled 16-byte chunks; eight consecutive
        // rows read eight different chunk slots, so the loads are bank-conflict free
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + C::kAColBase;
        uint32_t it = 0;
        for (int ct = cluster_id; ct < cluster_tiles; ct += n_clusters) {
            for (int kblk = 0; kblk < kblocks; ++kblk, ++it) {
                const int s = it % S;
                mbar_wait_long(&full[s], (it / S) & 1);   // also: the MMAs that read TMEM stage s have completed
                const uint8_t* src = a_raw(s) + row * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(src + ((j ^ (row & 7)) << 4));
                    const float h0 = tf32_rn(v.x), h1 = tf32_rn(v.y), h2 = tf32_rn(v.z), h3 = tf32_rn(v.w);
                    hi[4 * j + 0] = __float_as_uint(h0); hi[4 * j + 1] = __float_as_uint(h1);
                    hi[4 * j + 2] = __float_as_uint(h2); hi[4 * j + 3] = __float_as_uint(h3);
                    if (NPASS == 3) {
                        lo[4 * j + 0] = __float_as_uint(v.x - h0); lo[4 * j + 1] = __float_as_uint(v.y - h1);
                        lo[4 * j + 2] = __float_as_uint(v.z - h2); lo[4 * j + 3] = __float_as_uint(v.w - h3);
                    }
                }
                tmem_st_x32(lane_addr + s * C::kACols, hi);
                if (NPASS == 3) tmem_st_x32(lane_addr + s * C::kACols + kBK, lo);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&xf[s]);
            }
        }
    } else {
        // ===================== accumulate partials + epilogue =====================
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;       // accumulator row = pixel inside the tile
        const int te = threadIdx.x - 192;
        const int hsel = (warp - 6) >> 2;    // EW = 8: which half of the columns
        const int col0 = hsel ? C::kCols0 : 0;
        const int ncols = EW == 8 ? (hsel ? BN - C::kCols0 : C::kCols0) : BN;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + col0;
        const int groups = (kblocks + C::kFlush - 1) / C::kFlush;
        uint32_t gq = 0, cc = 0;
        for (int ct = cluster_id; ct < cluster_tiles; ct += n_clusters) {
            const TileCoord t = tile_coord(g, (ct / g.n_tiles) * CL + (int)rank, ct % g.n_tiles, BN);
            float acc[C::kCols0];
#pragma unroll
            for (int i = 0; i < C::kCols0; ++i) acc[i] = 0.f;
            for (int gi = 0; gi < groups; ++gi, ++gq) {
                const uint32_t p = gq % NP;
                mbar_wait_long(&part_full[p], (gq / NP) & 1);
                tc_fence_after();
                const uint32_t taddr = lane_addr + p * BN;
                // all loads of a batch in flight before the single wait: one TMEM round trip per batch
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
            const int pw = t.w0 + row % g.wb, phh = t.h0 + (row / g.wb) % g.hb, pn = t.n0 + row / (g.wb * g.hb);
            const bool valid = pw < g.W && phh < g.H && pn < g.B;
            if constexpr (EPI == EPI_NHWC) {
                // each column half (EW = 8: two halves, one staging buffer and one named barrier each;
                // EW = 4: one half, two staging buffers) streams its 32-column chunks through TMA stores
                // (measured against plain per-row stores: the write-bound K = 64 layer takes 40 us vs 60 us)
                static_assert(C::kCols0 % 32 == 0 && BN % 32 == 0, "channels-last epilogue: 32-column chunks");
                const size_t pix = ((size_t)pn * g.H + phh) * g.W + pw;
                const bool leader = (te & 127) == 0;
#pragma unroll
                for (int c = 0; c < C::kCols0 / 32; ++c, ++cc) {
                    uint8_t* buf = staging + (EW == 8 ? hsel : (int)(cc & 1)) * kStagingBytes;
                    if (leader) {   // the store that last read `buf` has finished reading
                        if (EW == 8) bulk_wait_read<0>(); else bulk_wait_read<1>();
                    }
                    named_bar_sync(1 + hsel, 128);
                    const int n0 = t.ncol0 + col0 + c * 32;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const float4 bv = e.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(e.bias + n0) + jj)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 o;
                        o.x = acc[c * 32 + jj * 4 + 0] + bv.x;
                        o.y = acc[c * 32 + jj * 4 + 1] + bv.y;
                        o.z = acc[c * 32 + jj * 4 + 2] + bv.z;
                        o.w = acc[c * 32 + jj * 4 + 3] + bv.w;
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
                    named_bar_sync(1 + hsel, 128);
                    if (leader) {
                        tma_store_4d(&mapOut, buf, n0, t.w0, t.h0, t.n0);
                        bulk_commit();
                    }
                }
            } else if (valid) {
                float4* dst = reinterpret_cast<float4*>(e.y + (((size_t)pn * g.H + phh) * g.W + pw) * e.ld_out + t.ncol0 + col0);
#pragma unroll
                for (int i = 0; i < C::kCols0 / 4; ++i)
                    if (i * 4 < ncols) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
        }
        if (EPI == EPI_NHWC && (te & 127) == 0) bulk_wait_all();
    }
    pdl_trigger();
    tc_fence_before();
    if (CL > 1) cluster_sync(); else __syncthreads();   // no CTA leaves while a peer may still multicast into it
    if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace tc
}  // namespace finc
