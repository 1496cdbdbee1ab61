// finc_inverse.cu -- persistent anti-diagonal wavefront inverse (sampling direction), sm_100a.
//
// Replaces the reference's (H+W-1)*Cq kernel launches, each followed by
// cudaDeviceSynchronize, plus 6 flips + 3 cats + zeros_like per FastFlowUnit
// (fastflow/fastflow.py:78-100, utils/fastflow_cuda_inverse/cinc_cuda_kernel_level2.cu:98-132)
// by ONE launch per unit:
//
//   * every warp is a persistent worker; an item is T = 32/W tiles (n..n+T-1, g) of one
//     group, moved HBM -> shared by 1-D TMA bulk copies on a per-warp mbarrier ring, solved
//     IN PLACE in shared memory, and written back with TMA bulk stores (cp.async.bulk
//     shared->global); load of item k+1/k+2 and store of item k-1 overlap the solve of k.
//   * lanes own columns.  The sweep is skewed: at step s lane j solves sweep-row s-j of its
//     sweep-column j, so all pixels of anti-diagonal s are solved in parallel and every
//     dependency (earlier diagonals) is already in shared memory.  Diagonals are
//     separated by __syncwarp() only -- no launch, no CTA barrier, no grid sync.
//   * TR/BL/BR corners are index maps (sweep coordinates -> stored coordinates), not flips.
//   * per pixel the C-channel triangular system of the corner tap is solved in registers
//     after all other taps were accumulated for all C outputs at once
//     (one shared load of x feeds C FMAs; weights are broadcast vector loads from the
//     sweep-ordered table wk[g][kh][kw][i][CPP]).
//
// Arithmetic is fp32 FMA; the accumulation order differs from the reference's in-place
// `-=` chain only in association (1e-7..1e-6 relative).
#include "finc_common.cuh"

namespace finc {

namespace {

constexpr int kMaxWarps = 16;

struct InvArgs {
    const float* z;
    const float* w;
    float* x;
    Shape s;
    int T;            // tiles per item (lanes / columns)
    int S;            // stages (1, 2 or 3)
    int WL;           // lanes per tile = min(W, 32)
    int gsplit;
    int bulk;
    int tile_floats;
    int tile_stride;
    int wk_floats;
    long n_items;
};

template <int CP>
struct CpPad {
    static constexpr int value = CP <= 2 ? CP : ((CP + 3) / 4) * 4;
};

__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int CP>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) inverse_warp_kernel(const InvArgs a) {
    constexpr int CPP = CpPad<CP>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* wk = reinterpret_cast<float*>(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Shape& s = a.s;
    const int C = s.C, H = s.H, W = s.W, kH = s.kH, kW = s.kW;
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

    const int n_pre = a.S == 3 ? 2 : a.S;  // loads in flight before the first solve
    if (a.bulk && lane == 0) {
        for (int st = 0; st < a.S; ++st) mbar_init(&bars[st], 1);
        fence_mbar_init();
        for (int st = 0; st < n_pre; ++st) {
            const long item = gw + st * gstride;
            if (item < a.n_items) issue_load(item, st);
        }
    }
    {
        // sweep-ordered weights: wk[gl][kh][kw][i][CPP] = Ws[g][o][i][a(kh)][b(kw)]
        for (int e = threadIdx.x; e < a.wk_floats; e += blockDim.x) wk[e] = 0.f;
        __syncthreads();
        const int kk = kH * kW;
        const int per_g = C * C * kk;
        const int ng = a.gsplit ? 1 : s.G;
        for (int e = threadIdx.x; e < ng * per_g; e += blockDim.x) {
            const int gl = e / per_g;
            const int g = a.gsplit ? g_fixed : gl;
            int r = e - gl * per_g;
            const int b = r % kW;
            r /= kW;
            const int aa = r % kH;
            r /= kH;
            const int i = r % C, o = r / C;
            const int ord = order_of(s.orders, g);
            const int kh = (ord & 2) ? aa : kH - 1 - aa;
            const int kw = (ord & 1) ? b : kW - 1 - b;
            wk[(((gl * kH + kh) * kW + kw) * C + i) * CPP + o] = a.w[(long)g * per_g + (e - gl * per_g)];
        }
        __syncthreads();
    }

    const int tl = lane / a.WL;       // tile slot of this lane
    const int jl = lane - tl * a.WL;  // column slot inside the 32-column block
    const int ncb = (W + 31) / 32;

    long k = 0;
    for (long item = gw; item < a.n_items; item += gstride, ++k) {
        const int st = (int)(k % a.S);
        const int g = item_g(item), n0 = item_n0(item);
        const int nt = min(a.T, s.B - n0);
        float* buf = bufs + st * stage_floats;
        if (a.bulk) {
            mbar_wait(&bars[st], (uint32_t)((k / a.S) & 1));
        } else {
            for (int t = 0; t < nt; ++t) {
                const float* src = a.z + ((long)(n0 + t) * s.G + g) * a.tile_floats;
                for (int e = lane; e < a.tile_floats; e += 32) buf[t * a.tile_stride + e] = src[e];
            }
            __syncwarp();
        }
        const int ord = order_of(s.orders, g);
        const bool bot = ord & 2, right = ord & 1;
        const float* wg = wk + (size_t)(a.gsplit ? 0 : g) * kH * kW * C * CPP;
        float* xt = buf + tl * a.tile_stride;
        const bool lane_on = tl < nt;

        for (int cb = 0; cb < ncb; ++cb) {
            const int ws = cb * 32 + jl;              // sweep column of this lane
            const int nc = min(32, W - cb * 32);      // columns in this block
            const bool col_on = lane_on && jl < nc;
            const int wst = right ? W - 1 - ws : ws;  // stored column
            const int nsteps = H + nc - 1;
            for (int step = 0; step < nsteps; ++step) {
                const int hs = step - jl;
                if (col_on && hs >= 0 && hs < H) {
                    const int h = bot ? H - 1 - hs : hs;
                    const int pix = h * W + wst;
                    float acc[CP];
#pragma unroll
                    for (int o = 0; o < CP; ++o) acc[o] = (o < C) ? xt[o * HW + pix] : 0.f;
                    const int khmax = min(kH - 1, hs), kwmax = min(kW - 1, ws);
                    for (int kh = 0; kh <= khmax; ++kh) {
                        const int hh = bot ? h + kh : h - kh;
                        for (int kw = (kh == 0 ? 1 : 0); kw <= kwmax; ++kw) {
                            const int src = hh * W + (right ? wst + kw : wst - kw);
                            const float* wp = wg + (size_t)((kh * kW + kw) * C) * CPP;
                            for (int i = 0; i < C; ++i) {
                                const float xv = -xt[i * HW + src];
                                if constexpr (CPP % 4 == 0) {
#pragma unroll
                                    for (int v = 0; v < CPP / 4; ++v) {
                                        const float4 f = *reinterpret_cast<const float4*>(wp + i * CPP + 4 * v);
                                        if (4 * v + 0 < CP) acc[4 * v + 0] = fmaf(xv, f.x, acc[4 * v + 0]);
                                        if (4 * v + 1 < CP) acc[4 * v + 1] = fmaf(xv, f.y, acc[4 * v + 1]);
                                        if (4 * v + 2 < CP) acc[4 * v + 2] = fmaf(xv, f.z, acc[4 * v + 2]);
                                        if (4 * v + 3 < CP) acc[4 * v + 3] = fmaf(xv, f.w, acc[4 * v + 3]);
                                    }
                                } else {
#pragma unroll
                                    for (int o = 0; o < CP; ++o) acc[o] = fmaf(xv, wp[i * CPP + o], acc[o]);
                                }
                            }
                        }
                    }
                    // corner tap: x[o] = acc[o] - sum_{i<o} W[o,i,corner] x[i]; diagonal/upper never read
#pragma unroll
                    for (int i = 0; i < CP - 1; ++i) {
                        if (i + 1 < C) {  // warp-uniform; rows i >= C of the table do not exist
#pragma unroll
                            for (int o = i + 1; o < CP; ++o) acc[o] = fmaf(-acc[i], wg[i * CPP + o], acc[o]);
                        }
                    }
#pragma unroll
                    for (int o = 0; o < CP; ++o)
                        if (o < C) xt[o * HW + pix] = acc[o];
                }
                __syncwarp();
            }
        }

        // ---- write back + pipeline bookkeeping ---------------------------------------------
        if (a.bulk) {
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                for (int t = 0; t < nt; ++t)
                    bulk_s2g(a.x + ((long)(n0 + t) * s.G + g) * a.tile_floats, buf + t * a.tile_stride,
                             (uint32_t)(a.tile_floats * 4));
                bulk_commit();
                if (a.S == 3) {
                    // stage of item k-1 has been read out; refill it with item k+2
                    bulk_wait_read_1();
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
    if (a.bulk && lane == 0) bulk_wait_all();  // smem must stay valid until the stores have read it
}

template <int CP>
int launch_inst(const InvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    auto kern = inverse_warp_kernel<CP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, threads, smem, st>>>(a);
    return (int)cudaGetLastError();
}

int pick_cp(int C) {
    const int opts[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
    for (int o : opts)
        if (C <= o) return o;
    return 0;
}

}  // namespace

int launch_inverse_fast(const float* z, const float* w, float* x, const Shape& s, cudaStream_t st, bool* handled) {
    *handled = false;
    const int CP = pick_cp(s.C);
    if (CP == 0) return 0;
    const long tile_floats_l = (long)s.C * s.H * s.W;
    if (tile_floats_l * 4 > 96 * 1024) return 0;
    InvArgs a{};
    a.z = z; a.w = w; a.x = x; a.s = s;
    a.tile_floats = (int)tile_floats_l;
    const int CPP = CP <= 2 ? CP : ((CP + 3) / 4) * 4;
    const size_t wk_per_g = (size_t)s.kH * s.kW * s.C * CPP;
    const size_t smem_max = max_optin_smem_cached();
    const size_t budget = smem_max > 8192 ? smem_max - 4096 : 0;
    if (wk_per_g * s.G * 4 <= budget / 2) { a.gsplit = 0; a.wk_floats = (int)(wk_per_g * s.G); }
    else if (wk_per_g * 4 <= budget / 2) { a.gsplit = 1; a.wk_floats = (int)wk_per_g; }
    else return 0;
    a.bulk = (a.tile_floats % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    a.WL = s.W < 32 ? s.W : 32;
    a.T = 32 / a.WL;
    // shared-memory stride between the T tiles of a stage: congruent to WL mod 32 so that the
    // T column groups of a warp fall into disjoint banks
    {
        int stride = (a.tile_floats + 3) & ~3;
        if (a.T > 1 && a.WL % 4 == 0) {
            while (stride % 32 != a.WL % 32) stride += 4;
        }
        a.tile_stride = stride;
    }
    const int sms = sm_count_cached();
    const size_t wk_bytes = (size_t)((a.wk_floats + 31) & ~31) * 4;
    const size_t stage_bytes = (size_t)a.T * a.tile_stride * 4;
    const int nbT = (s.B + a.T - 1) / a.T;
    a.n_items = a.gsplit ? nbT : (long)nbT * s.G;
    long ctas = a.gsplit ? sms / s.G : sms;
    if (ctas < 1) ctas = 1;
    if (ctas > a.n_items) ctas = a.n_items;
    const long items_per_cta = (a.n_items + ctas - 1) / ctas;
    // stages: 3 when warps will see several items and memory allows >= 8 warps, else 2, else 1
    int nwarps = 0;
    for (int S : {3, 2, 1}) {
        const size_t per_warp = S * stage_bytes + S * 8;
        int nw = (int)((budget - wk_bytes) / per_warp);
        if (nw > kMaxWarps) nw = kMaxWarps;
        const int want = (int)(items_per_cta < kMaxWarps ? items_per_cta : kMaxWarps);
        if (nw >= (S == 1 ? 1 : (want < 8 ? want : 8))) { a.S = S; nwarps = nw; break; }
    }
    if (nwarps < 1) return 0;
    if (items_per_cta <= nwarps) a.S = a.S > 1 ? 1 : a.S;  // one item per warp: no pipelining needed
    while (nwarps > 1 && (long)(nwarps - 1) * ctas >= a.n_items) --nwarps;
    const size_t smem = wk_bytes + (size_t)nwarps * (a.S * stage_bytes + a.S * 8) + 16;
    dim3 grid((unsigned)ctas, a.gsplit ? s.G : 1, 1);
    *handled = true;
    switch (CP) {
        case 1: return launch_inst<1>(a, grid, nwarps * 32, smem, st);
        case 2: return launch_inst<2>(a, grid, nwarps * 32, smem, st);
        case 3: return launch_inst<3>(a, grid, nwarps * 32, smem, st);
        case 4: return launch_inst<4>(a, grid, nwarps * 32, smem, st);
        case 6: return launch_inst<6>(a, grid, nwarps * 32, smem, st);
        case 8: return launch_inst<8>(a, grid, nwarps * 32, smem, st);
        case 12: return launch_inst<12>(a, grid, nwarps * 32, smem, st);
        case 16: return launch_inst<16>(a, grid, nwarps * 32, smem, st);
        case 24: return launch_inst<24>(a, grid, nwarps * 32, smem, st);
        case 32: return launch_inst<32>(a, grid, nwarps * 32, smem, st);
        default: *handled = false; return 0;
    }
}

}  // namespace finc
