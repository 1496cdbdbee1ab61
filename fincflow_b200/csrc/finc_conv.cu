// finc_conv.cu -- host-side planning for the fused FInC forward / backward-input convolution
// (kernel: finc_conv.cuh; instantiations: finc_conv_c<N>.cu).
#include "finc_conv.cuh"

#include <cstdlib>

namespace finc {

using conv::ConvArgs;
using conv::kMaxConsumerWarps;

namespace {

// output channels per lane, from the blocks instantiated for this channel count
// (finc_conv_c<N>.cu).  Larger blocks are more FMA-efficient; smaller ones give a small batch
// enough sub-items to occupy every SM.
int pick_ob(int ct, int C, long rows_strips_per_tile, long tiles_per_cta) {
    int cands[5];
    int n = 0;
    switch (ct) {
        case 1: cands[n++] = 1; break;
        case 2: cands[n++] = 2; cands[n++] = 1; break;
        case 3: cands[n++] = 3; cands[n++] = 1; break;
        case 4: cands[n++] = 4; cands[n++] = 2; break;
        case 6: cands[n++] = 6; cands[n++] = 3; cands[n++] = 2; break;
        case 12: case 24: cands[n++] = 4; cands[n++] = 2; cands[n++] = 1; break;
        default:  // generic: any divisor among {6,4,3,2,1}; 4 with a padded last block for C > 6
            for (int ob : {6, 4, 3, 2, 1}) {
                if (ob > C) continue;
                if (C % ob != 0 && !(ob == 4 && C > 6)) continue;
                if (ob == 6 && C % 4 == 0) continue;
                cands[n++] = ob;
            }
    }
    int best = cands[n - 1];
    if (const char* e = getenv("FINC_CONV_OB")) {  // experiment knob: force an instantiated block size
        const int want = atoi(e);
        for (int i = 0; i < n; ++i)
            if (cands[i] == want) return want;
    }
    for (int i = 0; i < n; ++i) {
        const long subs = tiles_per_cta * ((C + cands[i] - 1) / cands[i]) * rows_strips_per_tile;
        if (subs >= 384 || tiles_per_cta >= 64) return cands[i];
    }
    return best;
}

unsigned magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + d - 1) / d); }

}  // namespace

// everything about the weight table that does not depend on pointers or the batch split
struct WeightPlan {
    int ct, OB, OBP, nob, gsplit, WTn;
    size_t wk_per_g;
};

static bool plan_weights(const Shape& s, WeightPlan* p) {
    if (!(s.kH == s.kW && (s.kH == 3 || s.kH == 5 || s.kH == 2))) return false;
    if ((long)s.C * s.H * s.W * 4 > 48 * 1024) return false;
    p->WTn = s.W % 4 == 0 ? 4 : (s.W % 2 == 0 ? 2 : 1);  // nominal strip width (aligned tensors)
    p->ct = 0;
    if ((s.kH == 3 || s.kH == 5) && p->WTn >= 2)
        for (int c : {1, 2, 3, 4, 6, 12, 24})
            if (s.C == c) p->ct = c;
    const int sms = sm_count_cached();
    const size_t smem_max = max_optin_smem_cached();
    const size_t budget = smem_max > 8192 ? smem_max - 4096 : 0;
    p->gsplit = ((size_t)s.G * s.C * s.C * s.kH * s.kW * 4 * 4 / 3 > budget / 2) ? 1 : 0;
    const long n_tiles = p->gsplit ? s.B : (long)s.B * s.G;
    long ctas_max = p->gsplit ? sms / s.G : sms;
    if (ctas_max < 1) ctas_max = 1;
    const long spread = (n_tiles + ctas_max - 1) / ctas_max;
    p->OB = pick_ob(p->ct, s.C, (long)s.H * (s.W / p->WTn), spread);
    p->OBP = p->OB <= 2 ? p->OB : ((p->OB + 3) / 4) * 4;
    p->nob = (s.C + p->OB - 1) / p->OB;
    p->wk_per_g = (size_t)s.C * s.kH * s.kW * p->nob * p->OBP;
    if ((p->gsplit ? p->wk_per_g : p->wk_per_g * s.G) * 4 > budget / 2) return false;
    return true;
}

size_t conv_prepared_floats(const Shape& s) {
    WeightPlan p;
    if (!plan_weights(s, &p)) return 0;
    if ((p.gsplit ? p.wk_per_g : p.wk_per_g * s.G) % 4 != 0) return 0;  // bulk copies move 16-byte multiples
    if (((long)s.C * s.H * s.W) % 4 != 0) return 0;                       // (same for the tiles)
    return kPrepHeaderFloats + (((p.wk_per_g * s.G) + 3) & ~(size_t)3);
}

namespace {
// one CTA per (unit, group): global -> global transposition into the kernel's table layout
__global__ void conv_prepare_kernel(const float* __restrict__ w, float* __restrict__ out, size_t w_stride,
                                    size_t out_stride, Shape s, int transpose, int OB, int OBP, int nob) {
    const int u = blockIdx.x, g = blockIdx.y;
    const int C = s.C, KH = s.kH, KW = s.kW;
    const int per_g = C * C * KH * KW;
    const float* src = w + (size_t)u * w_stride + (size_t)g * per_g;
    float* dst = out + (size_t)u * out_stride + kPrepHeaderFloats + (size_t)g * C * KH * KW * nob * OBP;
    for (int e = threadIdx.x; e < per_g; e += blockDim.x) {
        int r = e;
        const int b = r % KW;
        r /= KW;
        const int aa = r % KH;
        r /= KH;
        const int i = r % C, o = r / C;
        int cin, cout, ap, bp;
        if (!transpose) { cin = i; cout = o; ap = aa; bp = b; }
        else { cin = o; cout = i; ap = KH - 1 - aa; bp = KW - 1 - b; }
        dst[((((size_t)cin * KH + ap) * KW + bp) * nob + cout / OB) * OBP + cout % OB] = __ldg(src + e);
    }
    if (g == 0 && threadIdx.x < 32) {
        // header: magic, kind, OB, nob, logdet = H*W*sum_g sum_o log|diag corner tap| (forward tables)
        float ld = 0.f;
        const float* wu = w + (size_t)u * w_stride;
        for (int e = threadIdx.x; e < s.G * C; e += 32) {
            const int gg = e / C, o = e - gg * C;
            const int ord = order_of(s.orders, gg);
            ld += logf(fabsf(__ldg(wu + (((size_t)gg * C + o) * C + o) * KH * KW + corner_a(ord, KH) * KW + corner_b(ord, KW))));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, off);
        if (threadIdx.x == 0) {
            float* h = out + (size_t)u * out_stride;
            h[0] = 1179208259.f; h[1] = (float)transpose; h[2] = (float)OB; h[3] = (float)nob;
            h[4] = ld * (float)s.H * (float)s.W;
        }
    }
}
}  // namespace

int launch_conv_prepare(const float* w, float* out, int n_units, size_t w_stride, size_t out_stride, const Shape& s,
                        bool transpose, cudaStream_t st) {
    WeightPlan p;
    if (!plan_weights(s, &p)) return FINC_E_UNSUPPORTED;
    // padding slots of the table are never read into a stored result: no need to clear them
    conv_prepare_kernel<<<dim3(n_units, s.G), 256, 0, st>>>(w, out, w_stride, out_stride, s, transpose ? 1 : 0, p.OB, p.OBP, p.nob);
    return (int)cudaGetLastError();
}

int launch_conv_fast(const float* x, const float* w, float* y, float* logdet, bool logdet_acc, const Shape& s,
                     bool transpose, bool prepared, int sm_div, cudaStream_t st, bool* handled) {
    *handled = false;
    WeightPlan wp;
    if (!plan_weights(s, &wp)) return 0;
    ConvArgs a{};
    a.x = x; a.w = w; a.y = y; a.logdet = transpose ? nullptr : logdet; a.logdet_acc = logdet_acc ? 1 : 0;
    a.s = s; a.transpose = transpose ? 1 : 0; a.dbg = debug_ts_buffer();
    a.tile_floats = s.C * s.H * s.W;
    a.bulk = (a.tile_floats % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    int WT = 1;
    const bool y16 = (reinterpret_cast<uintptr_t>(y) & 15) == 0, y8 = (reinterpret_cast<uintptr_t>(y) & 7) == 0;
    if (s.W % 4 == 0 && y16) WT = 4;  // smem rows are 16-byte aligned because tile_floats % 4 == 0 when W % 4 == 0
    else if (s.W % 2 == 0 && y8 && (a.tile_floats % 2 == 0)) WT = 2;
    a.nstrip = s.W / WT;
    int ct = wp.ct;
    if (WT < 2) ct = 0;  // (unaligned output: generic-C kernel; same table layout)
    a.prepared = prepared ? 1 : 0;
    if (prepared && (!a.bulk || (reinterpret_cast<uintptr_t>(w) & 15) != 0 || conv_prepared_floats(s) == 0)) return FINC_E_UNSUPPORTED;
    const int sms = sm_count_cached();
    const size_t smem_max = max_optin_smem_cached();
    const size_t budget = smem_max > 8192 ? smem_max - 4096 : 0;
    a.gsplit = wp.gsplit;
    const long n_tiles = a.gsplit ? s.B : (long)s.B * s.G;
    long ctas_max = (a.gsplit ? sms / s.G : sms) / (sm_div > 1 ? sm_div : 1);  // (the weight-table layout does not depend on sm_div)
    if (ctas_max < 1) ctas_max = 1;
    // experiment knob: FINC_CONV_SPLIT=k launches k smaller CTAs per SM (co-resident) instead of one
    static const int cta_split = getenv("FINC_CONV_SPLIT") ? atoi(getenv("FINC_CONV_SPLIT")) : 1;
    int max_cw = kMaxConsumerWarps;
    if (cta_split > 1 && !a.gsplit) { ctas_max *= cta_split; max_cw = (kMaxConsumerWarps + 1) / cta_split - 1; }
    const long spread = (n_tiles + ctas_max - 1) / ctas_max;
    const int OB = wp.OB;
    a.nob = wp.nob;
    a.wk_floats = (int)(a.gsplit ? wp.wk_per_g : wp.wk_per_g * s.G);
    // two output rows per sub-item for the HBM-bound small-channel levels once the batch is large
    // enough to keep every consumer thread busy anyway (amortises strip loads, weight loads and index
    // arithmetic; FINC_CONV_RB=1/2 overrides)
    a.RB = 1;
    if (ct >= 1 && ct <= 3 && OB == ct && WT == 4 && s.kH == 3 && s.H >= 2 &&
        spread * a.nob * s.H * a.nstrip >= 4L * kMaxConsumerWarps * 32)
        a.RB = 2;
    {
        static const char* e = getenv("FINC_CONV_RB");
        if (e && (e[0] == '1' || (e[0] == '2' && ct >= 1 && ct <= 3 && OB == ct && WT == 4 && s.kH == 3))) a.RB = e[0] - '0';
    }
    a.HR = (s.H + a.RB - 1) / a.RB;
    const int sub_per_tile = a.nob * a.HR * a.nstrip;
    if ((long)sub_per_tile >= 65536) return 0;
    a.m_nstrip = magic(a.nstrip); a.m_h = magic(a.HR); a.m_nob = magic(a.nob);

    // chunk: about one sub-item per consumer thread (512), at most 48 KB, and small enough that
    // every SM gets a chunk when the batch is small
    int CH = (max_cw * 32) / sub_per_tile;  // (rounded down: one pass of the consumer threads per chunk)
    if (CH < 1) CH = 1;
    while (CH > 1 && (long)CH * a.tile_floats * 4 > 48 * 1024) --CH;
    if (CH > spread) CH = (int)spread;
    if (CH > s.B) CH = s.B;
    if (CH < 1) CH = 1;
    if ((long)CH * sub_per_tile >= 65536) return 0;
    a.CH = CH;
    a.nblk = (s.B + CH - 1) / CH;
    a.n_chunks = a.gsplit ? a.nblk : (long)a.nblk * s.G;
    long ctas = ctas_max < a.n_chunks ? ctas_max : a.n_chunks;
    const long chunks_per_cta = (a.n_chunks + ctas - 1) / ctas;
    const size_t wk_bytes = (size_t)(((a.wk_floats + 31) & ~31) + conv::kFrontPad) * 4;
    const size_t stage_bytes = (size_t)CH * a.tile_floats * 4;
    int S = (int)(chunks_per_cta < 3 ? chunks_per_cta : 3);
    while (S > 1 && wk_bytes + S * stage_bytes + 256 > budget) --S;
    if (wk_bytes + S * stage_bytes + 256 > budget) return 0;
    a.S = S;
    int cw = (CH * sub_per_tile + 31) / 32;
    if (cw > max_cw) cw = max_cw;
    if (cw < 1) cw = 1;
    const size_t smem = wk_bytes + S * stage_bytes + 32 + (2 * S + 1) * 8 + 64;
    dim3 grid((unsigned)ctas, a.gsplit ? s.G : 1, 1);
    const int threads = (cw + 1) * 32;
    int rc;
    switch (ct) {
        case 1: rc = conv::dispatch_ob<1>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 2: rc = conv::dispatch_ob<2>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 3: rc = conv::dispatch_ob<3>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 4: rc = conv::dispatch_ob<4>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 6: rc = conv::dispatch_ob<6>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 12: rc = conv::dispatch_ob<12>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        case 24: rc = conv::dispatch_ob<24>(OB, WT, s.kH, a, grid, threads, smem, st); break;
        default: rc = conv::dispatch_ob<0>(OB, WT, s.kH, a, grid, threads, smem, st); break;
    }
    if (rc == FINC_E_UNSUPPORTED) return prepared ? FINC_E_UNSUPPORTED : 0;  // not instantiated: generic kernels
    *handled = true;
    return rc;
}

}  // namespace finc
