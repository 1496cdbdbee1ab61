// ffma_peak.cu -- measured fp32 FFMA throughput of the device (roofline denominator for the
// compute-bound FInC kernels, C >= 6).  Three variants: 3-register-operand FFMAs with 8 and 16
// independent chains per thread, and a variant whose multiplicand is shared by all chains
// (register reuse cache friendly, like the kernels' inner loops).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ffma_peak tools/ffma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, bool SHARED>
__global__ void __launch_bounds__(1024) ffma_kernel(float* out, float b0, float c0, int iters) {
    float a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = b0 + i * 1e-6f; }
    float c = c0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = SHARED ? fmaf(a[i], b[0], c) : fmaf(a[i], b[i], c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

template <int CH, bool SHARED>
void run(const char* name, int sms, int threads, int ctas_per_sm) {
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ffma_kernel<CH, SHARED><<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 1e-7f, iters);
    cudaEventRecord(e0);
    ffma_kernel<CH, SHARED><<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 1e-7f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * ctas_per_sm * threads * iters * 8.0 * CH;
    printf("%-34s threads/SM=%4d: %7.2f TFLOP/s  (%.3f FFMA/clk/SM-lane-equivalent at 1.965 GHz: %.1f lanes/SM)\n", name,
           threads * ctas_per_sm, 2.0 * fma / ms / 1e9, 0.0, fma / (ms * 1e-3) / 1.965e9 / sms);
    cudaFree(out);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs: %d\n", sms);
    run<8, false>("3-reg FFMA, 8 chains", sms, 1024, 1);
    run<16, false>("3-reg FFMA, 16 chains", sms, 1024, 1);
    run<8, true>("shared multiplicand, 8 chains", sms, 1024, 1);
    run<8, false>("3-reg FFMA, 8 chains", sms, 128, 1);
    run<8, true>("shared multiplicand, 8 chains", sms, 128, 1);
    run<16, false>("3-reg FFMA, 16 chains", sms, 256, 1);
    return 0;
}
