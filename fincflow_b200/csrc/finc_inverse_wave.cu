// finc_inverse_wave.cu -- host-side planning for the specialised wavefront inverse
// (kernel: finc_inverse_wave.cuh; instantiations: finc_inverse_wave_c<N>.cu).
#include "finc_inverse_wave.cuh"

namespace finc {

namespace wave {
extern template int dispatch_c<1>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<2>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<3>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<4>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<6>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<12>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
extern template int dispatch_c<24>(int, int, const WaveArgs&, dim3, int, size_t, cudaStream_t);
}  // namespace wave

using wave::WaveArgs;
using wave::kMaxWarps;

namespace {
struct WavePlan { int CPP, TS, gsplit; size_t wk_per_g; };
bool plan_wave_weights(const Shape& s, WavePlan* p) {
    if (!((s.kH == 3 && s.kW == 3) || (s.kH == 5 && s.kW == 5))) return false;
    const int C = s.C;
    if (!(C == 1 || C == 2 || C == 3 || C == 4 || C == 6 || C == 12 || C == 24)) return false;
    if ((long)C * s.H * s.W * 4 > 64 * 1024) return false;
    p->CPP = C <= 2 ? C : ((C + 3) / 4) * 4;
    p->TS = (C <= 2) ? C * p->CPP : (((C * p->CPP / 4) % 2 == 1) ? C * p->CPP : C * p->CPP + 4);
    p->wk_per_g = (size_t)s.kH * s.kW * p->TS;
    const size_t smem_max = max_optin_smem_cached();
    const size_t budget = smem_max > 8192 ? smem_max - 4096 : 0;
    if (p->wk_per_g * s.G * 4 <= budget / 2) p->gsplit = 0;
    else if (p->wk_per_g * 4 <= budget / 2) p->gsplit = 1;
    else return false;
    return true;
}
// parts per pixel must be an instantiated combination (finc_inverse_wave.cuh dispatch_p)
bool wave_parts_ok(const Shape& s) {
    int P = 1;
    while (P < 8 && s.W * (P * 2) <= 32) P *= 2;
    if (P == 1 && s.C > 6) return false;
    if (P == 2 && s.C > 12) return false;
    return true;
}

__global__ void wave_prepare_kernel(const float* __restrict__ w, float* __restrict__ out, size_t w_stride,
                                    size_t out_stride, Shape s, int CPP, int TS) {
    const int u = blockIdx.x, g = blockIdx.y;
    const int C = s.C, KH = s.kH, KW = s.kW;
    const int per_g = C * C * KH * KW;
    const float* src = w + (size_t)u * w_stride + (size_t)g * per_g;
    float* dst = out + (size_t)u * out_stride + kPrepHeaderFloats + (size_t)g * KH * KW * TS;
    const int ord = order_of(s.orders, g);
    for (int e = threadIdx.x; e < per_g; e += blockDim.x) {
        int r = e;
        const int b = r % KW;
        r /= KW;
        const int aa = r % KH;
        r /= KH;
        const int i = r % C, o = r / C;
        const int kh = (ord & 2) ? aa : KH - 1 - aa;
        const int kw = (ord & 1) ? b : KW - 1 - b;
        dst[(kh * KW + kw) * TS + i * CPP + o] = __ldg(src + e);
    }
    if (g == 0 && threadIdx.x == 0) {
        float* h = out + (size_t)u * out_stride;
        h[0] = 1179208259.f; h[1] = 2.f; h[2] = (float)CPP; h[3] = (float)TS;
    }
}
}  // namespace

size_t wave_prepared_floats(const Shape& s) {
    WavePlan p;
    if (!plan_wave_weights(s, &p)) return 0;
    if (!wave_parts_ok(s) && !rw_shape_supported(s)) return 0;  // no kernel would take the table
    if ((p.gsplit ? p.wk_per_g : p.wk_per_g * s.G) % 4 != 0) return 0;  // bulk copies move 16-byte multiples
    if (((long)s.C * s.H * s.W) % 4 != 0) return 0;                       // (same for the tiles)
    return kPrepHeaderFloats + ((p.wk_per_g * s.G + 3) & ~(size_t)3);
}

int launch_wave_prepare(const float* w, float* out, int n_units, size_t w_stride, size_t out_stride, const Shape& s,
                        cudaStream_t st) {
    WavePlan p;
    if (!plan_wave_weights(s, &p)) return FINC_E_UNSUPPORTED;
    wave_prepare_kernel<<<dim3(n_units, s.G), 256, 0, st>>>(w, out, w_stride, out_stride, s, p.CPP, p.TS);
    return (int)cudaGetLastError();
}

int launch_inverse_wave(const float* z, const float* w, float* x, const Shape& s, bool prepared, cudaStream_t st,
                        bool* handled) {
    *handled = false;
    if (!((s.kH == 3 && s.kW == 3) || (s.kH == 5 && s.kW == 5))) return 0;
    const int C = s.C;
    if (!(C == 1 || C == 2 || C == 3 || C == 4 || C == 6 || C == 12 || C == 24)) return 0;
    const long tile_floats_l = (long)C * s.H * s.W;
    if (tile_floats_l * 4 > 64 * 1024) return 0;
    WaveArgs a{};
    a.z = z; a.w = w; a.x = x; a.s = s; a.dbg = debug_ts_buffer(); a.prepared = prepared ? 1 : 0;
    a.tile_floats = (int)tile_floats_l;
    a.tile_stride = (a.tile_floats + 3) & ~3;
    const int CPP = C <= 2 ? C : ((C + 3) / 4) * 4;
    const int tap_stride = (C <= 2) ? C * CPP : (((C * CPP / 4) % 2 == 1) ? C * CPP : C * CPP + 4);
    const size_t wk_per_g = (size_t)s.kH * s.kW * tap_stride;
    const size_t smem_max = max_optin_smem_cached();
    const size_t budget = smem_max > 8192 ? smem_max - 4096 : 0;
    if (wk_per_g * s.G * 4 <= budget / 2) { a.gsplit = 0; a.wk_floats = (int)(wk_per_g * s.G); }
    else if (wk_per_g * 4 <= budget / 2) { a.gsplit = 1; a.wk_floats = (int)wk_per_g; }
    else return 0;
    a.bulk = (a.tile_floats % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    // parts per pixel: fill the warp with (columns x parts)
    int P = 1;
    while (P < 8 && s.W * (P * 2) <= 32) P *= 2;
    const int sms = sm_count_cached();
    const size_t wk_bytes = (size_t)((a.wk_floats + 31) & ~31) * 4;
    const size_t avail = budget - wk_bytes;
    const long tiles = a.gsplit ? s.B : (long)s.B * s.G;
    const long max_warps = (long)(a.gsplit ? sms / s.G : sms) * kMaxWarps;
    // stack depth: 1 while the batch cannot even fill the warps, up to 8 for large batches
    int T = 1;
    while (T < 8 && tiles / (T * 2) >= max_warps * 2 && (size_t)(T * 2) * a.tile_stride * 4 * 16 <= avail) T *= 2;
    a.T = T;
    const size_t stage_bytes = (size_t)a.T * a.tile_stride * 4;
    const int nbT = (s.B + a.T - 1) / a.T;
    a.n_items = a.gsplit ? nbT : (long)nbT * s.G;
    long ctas = a.gsplit ? sms / s.G : sms;
    if (ctas < 1) ctas = 1;
    if (ctas > a.n_items) ctas = a.n_items;
    const long items_per_cta = (a.n_items + ctas - 1) / ctas;
    int nwarps = (int)(items_per_cta < kMaxWarps ? items_per_cta : kMaxWarps);
    a.S = 1;
    if ((size_t)nwarps * (stage_bytes + 8) > avail) nwarps = (int)(avail / (stage_bytes + 8));
    if (nwarps < 1) return 0;
    // a second stage only pays when warps loop over several items and memory is left
    if (items_per_cta > nwarps && (size_t)nwarps * 2 * (stage_bytes + 8) <= avail) a.S = 2;
    if (prepared && (!a.bulk || (reinterpret_cast<uintptr_t>(w) & 15) != 0 || wave_prepared_floats(s) == 0)) return FINC_E_UNSUPPORTED;
    const size_t smem = wk_bytes + (size_t)nwarps * a.S * (stage_bytes + 8) + 8 + 16;
    dim3 grid((unsigned)ctas, a.gsplit ? s.G : 1, 1);
    int rc;
    switch (C) {
        case 1: rc = wave::dispatch_c<1>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        case 2: rc = wave::dispatch_c<2>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        case 3: rc = wave::dispatch_c<3>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        case 4: rc = wave::dispatch_c<4>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        case 6: rc = wave::dispatch_c<6>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        case 12: rc = wave::dispatch_c<12>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
        default: rc = wave::dispatch_c<24>(s.kH, P, a, grid, nwarps * 32, smem, st); break;
    }
    if (rc == FINC_E_UNSUPPORTED) return prepared ? FINC_E_UNSUPPORTED : 0;  // combination not instantiated: generic tiled kernel
    *handled = true;
    return rc;
}

}  // namespace finc
