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

int launch_inverse_wave(const float* z, const float* w, float* x, const Shape& s, cudaStream_t st, bool* handled) {
    *handled = false;
    if (!((s.kH == 3 && s.kW == 3) || (s.kH == 5 && s.kW == 5))) return 0;
    const int C = s.C;
    if (!(C == 1 || C == 2 || C == 3 || C == 4 || C == 6 || C == 12 || C == 24)) return 0;
    const long tile_floats_l = (long)C * s.H * s.W;
    if (tile_floats_l * 4 > 64 * 1024) return 0;
    WaveArgs a{};
    a.z = z; a.w = w; a.x = x; a.s = s; a.dbg = debug_ts_buffer();
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
    const size_t smem = wk_bytes + (size_t)nwarps * a.S * (stage_bytes + 8) + 16;
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
    if (rc == FINC_E_UNSUPPORTED) return 0;  // combination not instantiated: generic tiled kernel
    *handled = true;
    return rc;
}

}  // namespace finc
